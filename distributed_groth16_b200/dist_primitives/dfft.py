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
