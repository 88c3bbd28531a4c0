"""Mirror of the reference crate `secret-sharing` (packed Shamir sharing parameters) on the GPU kernels."""
from .pss import PackedSharingParams  # noqa: F401
