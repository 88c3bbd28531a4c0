"""`ext_wit::h` -- /root/reference/groth16/src/ext_wit.rs:16-101.

Two implementations of the same function:
  * `h`              -- the fused device pipeline (3 batched iNTT(m) -> coefficient shift by w_2m^j
                        -> 3 batched NTT(m) -> p*q - w), i.e. CircomReduction's formulation
                        (ark-circom/src/circom/qap.rs:64-89), one C-ABI call.
  * `h_via_d_fft`    -- the literal transcription of ext_wit.rs:34-92 on top of d_ifft/d_fft
                        (iNTT(m, rearrange, pad 2) -> NTT(2m) -> take odd slots -> p*q - w); kept
                        as a differential check that both read the reference the same way."""
from __future__ import annotations

import numpy as np

from ..context import MultiplexedStreamID, Net
from ..dist_primitives.dfft import d_fft, d_ifft, fft_in_place_rearrange
from .qap import PackedQAPShare


def _natural(share: PackedQAPShare):
    if share.rearranged:
        return tuple(fft_in_place_rearrange(v) for v in (share.a, share.b, share.c))
    return share.a, share.b, share.c


def h(qap_share: PackedQAPShare, pp=None, net: Net | None = None) -> np.ndarray:
    a, b, c = _natural(qap_share)
    return net.h_circom(a, b, c)


def h_via_d_fft(qap_share: PackedQAPShare, pp=None, net: Net | None = None) -> np.ndarray:
    m = qap_share.domain.size()
    vecs = (qap_share.a, qap_share.b, qap_share.c)
    if not qap_share.rearranged:
        vecs = tuple(fft_in_place_rearrange(v) for v in vecs)
    sids = (MultiplexedStreamID.Zero, MultiplexedStreamID.One, MultiplexedStreamID.Two)
    coeff = [d_ifft(v, True, 2, False, m, pp, net, s) for v, s in zip(vecs, sids)]          # ext_wit.rs:34-42
    evals = [d_fft(v, False, 1, False, 2 * m, pp, net, s) for v, s in zip(coeff, sids)]     # ext_wit.rs:44-52
    # king: pick index i*l + t with l = 2, t = 1 (ext_wit.rs:74-76), truncate to m
    p, q, w = (e[1::2][:m] for e in evals)
    return net.field_op(1, 2, net.field_op(1, 0, p, q), w)                                  # ext_wit.rs:88-92
