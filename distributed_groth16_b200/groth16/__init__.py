"""Mirror of the reference crate `groth16` (hot-path part only) + the circom artefact loaders."""
from . import circom, ext_wit, mpc, prove, qap, setup, verify  # noqa: F401
from .proving_key import PackedProvingKeyShare, ProvingKey  # noqa: F401
