"""The n-party `d_fft` / `d_ifft` protocol mirror (dist_primitives/dfft.py::d_fft_mpc, d_ifft_mpc) on a CPU stand-in for
`Net` whose field operations and transforms are computed by the oracle: the reference's own tests `d_fft_works` /
`d_ifft_works` (dist-primitives/src/dfft/mod.rs:285-556) re-stated for BN254, plus the rearrange + pad chaining that
`ext_wit::h` relies on (groth16/src/ext_wit.rs:34-52).  (On a GPU the same code runs on `Net.field_op` / `Net.ntt`,
themselves covered by tests/test_gpu_field.py and tests/test_gpu_pss.py.)"""
import numpy as np
import pytest

from distributed_groth16_b200.dist_primitives import d_fft_mpc, d_ifft_mpc, fft_in_place_rearrange
from distributed_groth16_b200.secret_sharing import PackedSharingParams


class OracleNet:
    """field_op / ntt with the host-buffer signatures of `Net`, in big-int arithmetic."""

    def field_op(self, field, op, a, b):
        from oracle import bn254 as o, layout
        assert field == 1
        x, y = layout.arr_to_fr(np.asarray(a).reshape(-1, 4)), layout.arr_to_fr(np.asarray(b).reshape(-1, 4))
        f = (lambda p, q: p * q % o.R, lambda p, q: (p + q) % o.R, lambda p, q: (p - q) % o.R)[op]
        return layout.fr_to_arr([f(p, q) for p, q in zip(x, y)])

    def fr_powers(self, base, scale, n):
        from oracle import bn254 as o, layout
        return layout.fr_to_arr([scale * pow(base, i, o.R) % o.R for i in range(n)])

    def ntt(self, data, inverse=False, coset=False, **_):
        from oracle import bn254 as o, layout
        v = layout.arr_to_fr(np.asarray(data).reshape(-1, 4))
        return layout.fr_to_arr(o.intt(v, coset=coset) if inverse else o.ntt(v, coset=coset))


def _share(x_rearranged, pp):
    """pcoeff[i] = pack(x[i], x[i + M/l], ...); party p holds pcoeff[i][p] for every i (dfft/mod.rs:307-318)."""
    n_chunks = x_rearranged.shape[0] // pp.l
    packed = [pp.pack_from_public(x_rearranged[i::n_chunks]) for i in range(n_chunks)]
    return [np.stack([packed[i][p] for i in range(n_chunks)]) for p in range(pp.n)]


def _open(shares, pp, degree2=False):
    """transpose(result).flat_map(unpack)"""
    n_chunks = shares[0].shape[0]
    out = []
    for i in range(n_chunks):
        col = np.stack([s[i] for s in shares])
        out.append(pp.unpack2(col) if degree2 else pp.unpack(col))
    return np.concatenate(out)


@pytest.mark.parametrize("l,m", [(2, 8), (2, 64), (4, 32), (1, 16)])
def test_d_fft_and_d_ifft_protocols_equal_the_plain_transforms(cref, l, m):
    from oracle import bn254 as o, layout
    net = OracleNet()
    pp = PackedSharingParams(l, net)
    x = cref.fr_generate(100 + m + l, m)
    xi = layout.arr_to_fr(x)
    shares = _share(fft_in_place_rearrange(x), pp)
    got = _open(d_fft_mpc(shares, False, 1, False, m, pp, net), pp)
    assert (got == layout.fr_to_arr(o.ntt(xi))).all()                       # d_fft_works
    got = _open(d_ifft_mpc(shares, False, 1, False, m, pp, net), pp)
    assert (got == layout.fr_to_arr(o.intt(xi))).all()                      # d_ifft_works
    with pytest.raises(Exception):
        d_fft_mpc(shares, False, 1, False, 2 * m, pp, net)                  # "Mismatch of size in FFT"


def test_rearrange_and_pad_chain_like_ext_wit_h(cref):
    """d_ifft(rearrange, pad = 2) hands back shares that are directly the input of a d_fft over the doubled domain
    (ext_wit.rs:34-52): opening the second transform gives fft_2m(ifft_m(x) || 0)."""
    from oracle import bn254 as o, layout
    net = OracleNet()
    l, m = 2, 16
    pp = PackedSharingParams(l, net)
    x = cref.fr_generate(7, m)
    shares = _share(fft_in_place_rearrange(x), pp)
    mid = d_ifft_mpc(shares, True, 2, False, m, pp, net)
    assert all(s.shape[0] == 2 * m // l for s in mid)
    got = _open(d_fft_mpc(mid, False, 1, False, 2 * m, pp, net), pp)
    coeffs = o.intt(layout.arr_to_fr(x)) + [0] * m
    assert (got == layout.fr_to_arr(o.ntt(coeffs))).all()


def test_degree2_input_is_reduced_by_the_king(cref):
    """degree2 = true: the parties hold share-wise products; the king unpacks with `unpack2` (dfft/mod.rs:207-211)."""
    from oracle import bn254 as o, layout
    net = OracleNet()
    l, m = 2, 8
    pp = PackedSharingParams(l, net)
    a, b = cref.fr_generate(1, m), cref.fr_generate(2, m)
    sa, sb = _share(fft_in_place_rearrange(a), pp), _share(fft_in_place_rearrange(b), pp)
    prod = [net.field_op(1, 0, p, q) for p, q in zip(sa, sb)]
    got = _open(d_fft_mpc(prod, False, 1, True, m, pp, net), pp)
    ab = [p * q % o.R for p, q in zip(layout.arr_to_fr(a), layout.arr_to_fr(b))]
    assert (got == layout.fr_to_arr(o.ntt(ab))).all()
