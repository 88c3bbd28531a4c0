"""Host-side logic of the mirrored operator layer (no GPU)."""
import numpy as np
import pytest

from distributed_groth16_b200 import MpcNetError
from distributed_groth16_b200.dist_primitives import d_fft, d_msm, fft_in_place_rearrange


def test_fft_in_place_rearrange_matches_reference_permutation(cref):
    for lg in (0, 1, 3, 6, 10):
        x = cref.fr_generate(lg, 1 << lg)
        assert (fft_in_place_rearrange(x) == cref.bitrev(x)).all()
        assert (fft_in_place_rearrange(fft_in_place_rearrange(x)) == x).all()


def test_primitives_need_a_net():
    x = np.zeros((8, 4), dtype=np.uint64)
    with pytest.raises(MpcNetError):
        d_fft(x, False, 1, False, 8, None, None)
    with pytest.raises(MpcNetError):
        d_msm(np.zeros((8, 8), dtype=np.uint64), x, None, None)
