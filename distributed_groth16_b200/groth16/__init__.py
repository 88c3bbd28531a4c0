"""Mirror of the reference crate `groth16` (hot-path part only)."""
from . import ext_wit, prove, qap  # noqa: F401
from .proving_key import ProvingKey  # noqa: F401
