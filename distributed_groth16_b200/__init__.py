"""distributed_groth16_b200 -- B200-native Groth16 proving hot path (BN254) behind the reference's
`dist-primitives::{d_msm, d_fft, d_ifft}` and `groth16::{ext_wit::h, prove::{A,B,C}}` API.

(The task names the package `distributed-groth16_b200`; a hyphen is not importable, hence the
underscore.)  All compute happens in csrc/*.cu through include/b200zk.h; this Python layer only
mirrors the reference's operator interface for tests, bench.py and multi-GPU orchestration."""
from ._native import B200zkError  # noqa: F401
from .context import MpcNetError, MultiplexedStreamID, Net  # noqa: F401
