"""`d_fft` / `d_ifft` -- /root/reference/dist-primitives/src/dfft/mod.rs:17-95.

Reference contract (asserted by its tests, dfft/mod.rs:285-556): the caller bit-reverses x and
stride-packs it into shares (:307-318); the protocol returns shares of dom.fft(x) / dom.ifft(x)
in natural order, or -- `rearrange` -- already bit-reversed and stride-packed for the next
transform, after zero-extending to `pad * m` (:225-245).  Without secret sharing (l = 1) a
"share" is the vector itself, so: input = bit-reversed x, output = natural (or bit-reversed when
`rearrange`) transform, zero-padded to pad*m.  `degree2` selects the king's unpack variant and
has no effect without PSS."""
from __future__ import annotations

import numpy as np

from ..context import MpcNetError, MultiplexedStreamID, Net


def bitrev_indices(n: int) -> np.ndarray:
    """rev[i] = i with its log2(n) bits reversed (int64)."""
    lg = n.bit_length() - 1
    assert 1 << lg == n
    idx = np.arange(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for b in range(lg):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lg - 1 - b)
    return rev.astype(np.int64)


def fft_in_place_rearrange(data: np.ndarray) -> np.ndarray:
    """Bit-reversal permutation (dfft/mod.rs:258-271); host-side index shuffle, returns a copy."""
    data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
    n = data.shape[0]
    lg = n.bit_length() - 1
    assert 1 << lg == n
    idx = np.arange(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for b in range(lg):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lg - 1 - b)
    out = np.empty_like(data)
    out[rev.astype(np.int64)] = data
    return out


def _run(share, rearrange, pad, dom_size, net, sid, inverse):
    if net is None:
        raise MpcNetError("NotConnected", "d_fft needs a Net (GPU context)")
    x = np.ascontiguousarray(share, dtype=np.uint64).reshape(-1, 4)
    # debug_assert_eq!(share.len() * pp.l, dom.size())   (dfft/mod.rs:31-37), l = 1 here
    if dom_size is not None and x.shape[0] != int(dom_size):
        raise MpcNetError("BadInput", "Mismatch of size in FFT, %d, %d." % (x.shape[0], int(dom_size)))
    return net.ntt(x, inverse=inverse, coset=False, bitrev_in=True, bitrev_out=bool(rearrange), pad=int(pad),
                   sid=int(sid))


def d_fft(pcoeff_share, rearrange: bool, pad: int, degree2: bool, dom, pp=None, net: Net | None = None,
          sid: MultiplexedStreamID = MultiplexedStreamID.Zero) -> np.ndarray:
    """dom: domain size (int) or an object with `.size()`."""
    size = dom.size() if hasattr(dom, "size") and callable(dom.size) else dom
    return _run(pcoeff_share, rearrange, pad, size, net, sid, inverse=False)


def d_ifft(peval_share, rearrange: bool, pad: int, degree2: bool, dom, pp=None, net: Net | None = None,
           sid: MultiplexedStreamID = MultiplexedStreamID.Zero) -> np.ndarray:
    size = dom.size() if hasattr(dom, "size") and callable(dom.size) else dom
    return _run(peval_share, rearrange, pad, size, net, sid, inverse=True)


# ---------------------------------------------------------------------------------------------------------------------
# The n-party protocol itself: what `d_fft` / `d_ifft` do when the vector is packed-secret-shared over
# n = 4l parties, all parties simulated in this process the way the reference's LocalTestNet does
# (mpc-net/src/multi.rs:289-316).  Field arithmetic goes through the CUDA library (`Net.field_op`, the sharing's transforms).
# ---------------------------------------------------------------------------------------------------------------------
_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_R_MONT = (1 << 256) % _FR


def _mont_limbs(values) -> np.ndarray:
    out = np.zeros((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = (int(v) % _FR) * _R_MONT % _FR
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def _butterfly(net, x, y, factors):
    """(x + y f, x - y f) on (rows, 4) limb arrays."""
    yf = net.field_op(1, 0, y, factors)
    return net.field_op(1, 1, x, yf), net.field_op(1, 2, x, yf)


def _fft1_in_place(px: np.ndarray, dom_size: int, l: int, gen: int, net) -> np.ndarray:
    """`fft1_in_place` (dfft/mod.rs:98-140): the stages every party runs on its own share vector.  The twiddle of
    column k is factor_stride^(k+1) -- the reference starts `factor` at `factor_stride`, not 1 (its final
    `rotate_right(1)` in fft2 undoes the shift)."""
    px = px.copy()
    log_l, log_n = l.bit_length() - 1, dom_size.bit_length() - 1
    for i in range(log_n, log_l, -1):
        ps = dom_size >> i
        stride = pow(gen, 1 << (i - 1), _FR)
        nj = (1 << (i - 1)) // l
        f = net.fr_powers(stride, stride, ps)                    # stride^(k+1), k < ps
        v = px.reshape(nj, 2, ps, 4)
        x = np.ascontiguousarray(v[:, 0]).reshape(-1, 4)
        y = np.ascontiguousarray(v[:, 1]).reshape(-1, 4)
        s, d = _butterfly(net, x, y, np.tile(f, (nj, 1)))
        v[:, 0] = s.reshape(nj, ps, 4)
        v[:, 1] = d.reshape(nj, ps, 4)
    return px


def _fft2_in_place(s1: np.ndarray, dom_size: int, l: int, gen: int, net) -> np.ndarray:
    """`fft2_in_place` (dfft/mod.rs:142-182): the last log2(l) stages, run by the king on the unpacked values."""
    log_l = l.bit_length() - 1
    for i in range(log_l, 0, -1):
        ps = dom_size >> i
        stride = pow(gen, 1 << (i - 1), _FR)
        half = 1 << (i - 1)
        f = net.fr_powers(stride, stride, ps)
        v = s1.reshape(ps, half, 2, 4)                           # s1[k * 2^i + 2j + b]
        x = np.ascontiguousarray(v[:, :, 0]).reshape(-1, 4)
        y = np.ascontiguousarray(v[:, :, 1]).reshape(-1, 4)
        s, d = _butterfly(net, x, y, np.repeat(f, half, axis=0))
        s2 = np.empty_like(s1)
        s2[: ps * half] = s                                      # s2[k * 2^(i-1) + j]
        s2[ps * half:] = d                                       # s2[(k + ps) * 2^(i-1) + j]
        s1 = s2
    return np.roll(s1, 1, axis=0)                                # rotate_right(1)


def _fft2_with_rearrange_pad(shares, rearrange, pad, degree2, dom_size, pp, gen, net):
    """`fft2_with_rearrange_pad` (dfft/mod.rs:184-256): every party sends its vector to the king, who unpacks element by
    element, finishes the transform, pads / rearranges and deals fresh packed shares back."""
    mbyl = shares[0].shape[0]
    s1 = np.zeros((mbyl * pp.l, 4), dtype=np.uint64)
    for i in range(mbyl):
        col = np.stack([sh[i] for sh in shares])                 # transpose(all_shares)[i]
        s1[i * pp.l:(i + 1) * pp.l] = pp.unpack2(col) if degree2 else pp.unpack(col)
    s1 = _fft2_in_place(s1, dom_size, pp.l, gen, net)
    if pad > 1:
        s1 = np.concatenate([s1, np.zeros(((pad - 1) * s1.shape[0], 4), dtype=np.uint64)])
    n_out = s1.shape[0] // pp.l
    if rearrange:
        s1 = fft_in_place_rearrange(s1)
        packed = [pp.pack_from_public(s1[i::n_out]) for i in range(n_out)]
    else:
        packed = [pp.pack_from_public(s1[i * pp.l:(i + 1) * pp.l]) for i in range(n_out)]       # pack_vec
    return [np.stack([packed[i][p] for i in range(n_out)]) for p in range(pp.n)]


def _dom_gen(dom_size: int) -> int:
    return pow(5, (_FR - 1) // dom_size, _FR)


def d_fft_mpc(pcoeff_shares, rearrange: bool, pad: int, degree2: bool, dom, pp, net) -> list:
    """All pp.n parties of `d_fft` (dfft/mod.rs:17-54) in one process: pcoeff_shares[p] is party p's share vector
    (dom.size() / pp.l elements, Montgomery limbs); returns the n output share vectors."""
    size = dom.size() if hasattr(dom, "size") and callable(dom.size) else int(dom)
    shares = [np.ascontiguousarray(s, dtype=np.uint64).reshape(-1, 4) for s in pcoeff_shares]
    if len(shares) != pp.n or any(s.shape[0] * pp.l != size for s in shares):
        raise MpcNetError("BadInput", "Mismatch of size in FFT, %d, %d." % (shares[0].shape[0] * pp.l, size))
    gen = _dom_gen(size)
    shares = [_fft1_in_place(s, size, pp.l, gen, net) for s in shares]
    return _fft2_with_rearrange_pad(shares, rearrange, pad, degree2, size, pp, gen, net)


def d_ifft_mpc(peval_shares, rearrange: bool, pad: int, degree2: bool, dom, pp, net) -> list:
    """All pp.n parties of `d_ifft` (dfft/mod.rs:56-95): scale by 1/size, then the same two phases with the inverse root."""
    size = dom.size() if hasattr(dom, "size") and callable(dom.size) else int(dom)
    shares = [np.ascontiguousarray(s, dtype=np.uint64).reshape(-1, 4) for s in peval_shares]
    if len(shares) != pp.n or any(s.shape[0] * pp.l != size for s in shares):
        raise MpcNetError("BadInput", "Mismatch of size in IFFT, %d, %d." % (shares[0].shape[0] * pp.l, size))
    gen = pow(_dom_gen(size), -1, _FR)
    ninv = _mont_limbs([pow(size, -1, _FR)])
    shares = [net.field_op(1, 0, s, np.tile(ninv, (s.shape[0], 1))) for s in shares]
    shares = [_fft1_in_place(s, size, pp.l, gen, net) for s in shares]
    return _fft2_with_rearrange_pad(shares, rearrange, pad, degree2, size, pp, gen, net)
