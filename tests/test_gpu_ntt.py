"""d_fft / d_ifft parity: CUDA NTT (csrc/ntt.cu) vs the oracle, through the C ABI.

Mirrors the reference's own checks (dist-primitives/src/dfft/mod.rs:285-556, examples/dfft_test.rs,
examples/local_dfft_test.rs): d_fft == dom.fft, d_ifft == dom.ifft, d_fft(d_ifft(x, rearrange)) == x."""
import numpy as np
import pytest

from distributed_groth16_b200 import B200zkError
from distributed_groth16_b200.dist_primitives import d_fft, d_ifft, fft_in_place_rearrange

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 13, 16, 17])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(net, cref, log_n, inverse):
    x = cref.fr_generate(0xB2000003 + log_n, 1 << log_n)
    got = net.ntt(x, inverse=inverse)
    assert (got == cref.ntt(x, inverse=inverse)).all()


@pytest.mark.parametrize("log_n", [1, 6, 11])
@pytest.mark.parametrize("inverse", [False, True])
def test_coset_ntt_matches_oracle(net, cref, log_n, inverse):
    x = cref.fr_generate(77 + log_n, 1 << log_n)
    got = net.ntt(x, inverse=inverse, coset=True)
    assert (got == cref.ntt(x, inverse=inverse, coset=True)).all()


def test_reference_sequence_0_to_m(net, cref):
    """examples/dfft_test.rs:20-26: x = 0..m-1, m = 1024."""
    from oracle import layout
    m = 1024
    x = layout.fr_to_arr(range(m))
    share = fft_in_place_rearrange(x)                       # dfft_test.rs:28
    out = d_fft(share, False, 1, False, m, None, net)
    assert (out == cref.ntt(x)).all()


@pytest.mark.parametrize("m", [8, 64, 4096])
def test_d_ifft_then_d_fft_roundtrip_with_rearrange_and_pad(net, cref, m):
    """dfft/mod.rs tests: d_ifft(rearrange=true, pad) feeds d_fft directly."""
    x = cref.fr_generate(5 + m, m)
    coeff_br = d_ifft(fft_in_place_rearrange(x), True, 2, False, m, None, net)      # bit-reversed, padded to 2m
    assert coeff_br.shape[0] == 2 * m
    exp = np.zeros((2 * m, 4), dtype=np.uint64)
    exp[:m] = cref.ntt(x, inverse=True)
    assert (coeff_br == cref.bitrev(exp)).all()
    evals = d_fft(coeff_br, False, 1, False, 2 * m, None, net)
    assert (evals == cref.ntt(exp)).all()
    assert (evals[0::2] == x).all()                        # even slots of the 2m-domain are the m-domain


def test_size_mismatch_and_domain_errors(net, cref):
    from distributed_groth16_b200 import MpcNetError
    x = cref.fr_generate(1, 8)
    with pytest.raises(MpcNetError):
        d_fft(x, False, 1, False, 16, None, net)
    with pytest.raises(B200zkError):
        net.ntt(x[:6])


def test_roundtrip_2_22_device_resident(net):
    """BASELINE config 3 size: iNTT(NTT(x)) == x at 2^22 (size-independent property), device buffers."""
    import torch
    n = 1 << 22
    x = net.generate_fr(0xB2000003, n)
    net.use_torch_stream(0)
    y = net.ntt_dev(x)
    z = net.ntt_dev(y, inverse=True)
    torch.cuda.synchronize()
    assert torch.equal(x, z)
    assert not torch.equal(x, y)


def test_linearity_2_20(net, cref):
    """NTT(a + b) == NTT(a) + NTT(b) at 2^20, and a 2^20 spot check against the oracle."""
    import torch
    n = 1 << 20
    a = net.generate_fr(11, n)
    b = net.generate_fr(12, n)
    ah, bh = a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64)
    s = net.field_op(1, 1, ah, bh)
    fs = net.ntt(s)
    fa, fb = net.ntt(ah), net.ntt(bh)
    assert (net.field_op(1, 1, fa, fb) == fs).all()
    assert (fa == cref.ntt(ah)).all()


def test_config3_2_22_vs_oracle(net, cref):
    """BASELINE config 3: 2^22 elements, forward, inverse and coset variants against the CPU twin."""
    n = 1 << 22
    x = net.generate_fr(0xB2000003, n).cpu().numpy().view(np.uint64)
    assert (net.ntt(x) == cref.ntt(x)).all()
    assert (net.ntt(x, inverse=True) == cref.ntt(x, inverse=True)).all()
    assert (net.ntt(x, coset=True) == cref.ntt(x, coset=True)).all()


def test_roundtrip_2_24_and_2_26_device_resident(net):
    """Largest sizes BASELINE lists (2^26 elements = 2 GiB per vector): iNTT(NTT(x)) == x."""
    import torch
    for log_n in (24, 26):
        x = net.generate_fr(0xB2000000 + log_n, 1 << log_n)
        net.use_torch_stream(0)
        y = net.ntt_dev(x)
        z = net.ntt_dev(y, inverse=True)
        torch.cuda.synchronize()
        assert torch.equal(x, z) and not torch.equal(x[:1024], y[:1024])
        del x, y, z
        torch.cuda.empty_cache()
