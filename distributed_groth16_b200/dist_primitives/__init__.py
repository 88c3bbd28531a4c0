"""Mirror of the reference crate `dist-primitives` (hot-path part only)."""
from .dmsm import GroupElement, d_msm, d_msm_mpc, packexp_from_public, unpackexp  # noqa: F401
from .dfft import d_fft, d_fft_mpc, d_ifft, d_ifft_mpc, fft_in_place_rearrange  # noqa: F401
