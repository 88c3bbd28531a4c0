"""`d_msm` -- /root/reference/dist-primitives/src/dmsm/mod.rs:70-98.

Reference: every party runs `G::msm` on its packed shares (:82), sends one group element to the
king, the king unpacks in the exponent and sums (:87-97).  Mathematically the result is the plain
MSM of the public vectors (examples/dmsm_test.rs:49-64).  Here the bases/scalars are *length-sharded*
over the GPUs of the box instead of secret-shared: every rank runs the Pippenger kernels on its
slice and the star gather is replaced by one NCCL all-gather of the XYZZ partials (128 B for G1,
256 B for G2 per rank) followed by a local point sum on every rank."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..context import MpcNetError, MultiplexedStreamID, Net


@dataclass
class GroupElement:
    """A curve point in canonical affine Montgomery limbs (what `G::into_affine()` would hold)."""
    limbs: np.ndarray          # (8,) for G1, (16,) for G2, uint64
    infinity: bool
    g2: bool = False

    def __eq__(self, other):
        return (self.g2 == other.g2 and self.infinity == other.infinity
                and (self.infinity or bool((self.limbs == other.limbs).all())))


def d_msm(bases, scalars, pp=None, net: Net | None = None, sid: MultiplexedStreamID = MultiplexedStreamID.Zero,
          g2: bool | None = None) -> GroupElement:
    """sum_i scalars[i] * bases[i].

    bases: (n, 8) G1 or (n, 16) G2 affine Montgomery limbs (host numpy, or CUDA int64 tensor);
    scalars: (n, 4) Fr Montgomery limbs.  `pp` (PackedSharingParams) is accepted for signature
    parity and ignored: the single-box build has no secret sharing.  With an initialised
    torch.distributed world, `bases`/`scalars` are this rank's slice.
    Raises MpcNetError("Generic", str(min_len)) on a length mismatch, like `?` at dmsm/mod.rs:82."""
    if net is None:
        raise MpcNetError("NotConnected", "d_msm needs a Net (GPU context)")
    is_torch = hasattr(bases, "data_ptr")
    width = int(bases.shape[-1]) if hasattr(bases, "shape") and len(bases.shape) == 2 else None
    if g2 is None:
        g2 = (width == 16)
    if net.n_parties() == 1 and not is_torch:
        limbs, inf = net.msm(bases, scalars, g2=g2, sid=int(sid))
        return GroupElement(limbs, inf, g2)
    import torch
    if not is_torch:
        dev = torch.device("cuda", net.device)
        b = np.ascontiguousarray(bases, dtype=np.uint64).view(np.int64)
        s = np.ascontiguousarray(scalars, dtype=np.uint64).view(np.int64)
        bases = torch.from_numpy(b).to(dev)
        scalars = torch.from_numpy(s).to(dev)
    net.use_torch_stream(int(sid))
    part = net.msm_dev(bases, scalars, g2=g2, sid=int(sid))
    if net.n_parties() > 1:
        import torch.distributed as dist
        gathered = torch.empty((net.n_parties(), part.numel()), dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(gathered, part.reshape(1, -1))
        limbs, inf = net.sum_points_dev(gathered, net.n_parties(), g2=g2, sid=int(sid))
    else:
        limbs, inf = net.sum_points_dev(part, 1, g2=g2, sid=int(sid))
    return GroupElement(limbs, inf, g2)
