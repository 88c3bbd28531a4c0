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
        # the star's gather / unpackexp / sum / scatter (dmsm/mod.rs:87-97) as one kernel over peer memory
        xch = getattr(net, "_partial_exchange", None)
        if xch is None:
            from ..parallel import PartialExchange
            xch = net._partial_exchange = PartialExchange(net)
        res = xch.sum(part, g2=g2, sid=int(sid)).cpu().numpy().view(np.uint64)
        w = 16 if g2 else 8
        return GroupElement(res[:w].copy(), bool(res[w]), g2)
    limbs, inf = net.sum_points_dev(part, 1, g2=g2, sid=int(sid))
    return GroupElement(limbs, inf, g2)


# ---------------------------------------------------------------------------------------------------------------------
# The reference's MPC protocol on top of the GPU kernels: packed secret sharing "in the exponent".
# ---------------------------------------------------------------------------------------------------------------------
_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_GEN = 5


def _fr_limbs_mont(values):
    """canonical ints -> Montgomery limbs; integer arithmetic on the O(n^2), n <= 16, domain constants only."""
    rows = []
    for v in values:
        x = (v % _FR) * (1 << 256) % _FR
        rows.append([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])
    return np.array(rows, dtype=np.uint64)


def _group_dft(net, points, size: int, inverse: bool, coset: bool, g2: bool):
    """FFT / iFFT of group elements over a radix-2 domain of `size` (offset = generator when `coset`): what
    `domain.fft_in_place(&mut Vec<G>)` does in unpackexp / packexp (dmsm/mod.rs:14,38,45,55,58).  Row i of the DFT matrix
    is a `size`-term MSM, executed by the Pippenger kernels."""
    w = 16 if g2 else 8
    pts = np.zeros((size, w), dtype=np.uint64)
    p = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, w)
    pts[: min(size, p.shape[0])] = p[: min(size, p.shape[0])]
    omega = pow(_GEN, (_FR - 1) // size, _FR)
    if inverse:
        omega = pow(omega, -1, _FR)
    out = np.zeros((size, w), dtype=np.uint64)
    ninv = pow(size, -1, _FR)
    for i in range(size):
        if not inverse:
            sc = [pow(omega, i * j, _FR) * (pow(_GEN, j, _FR) if coset else 1) for j in range(size)]
        else:
            scale = ninv * (pow(_GEN, -i, _FR) if coset else 1)
            sc = [pow(omega, i * j, _FR) * scale for j in range(size)]
        limbs, inf = net.msm(pts, _fr_limbs_mont(sc), g2=g2)
        if not inf:
            out[i] = limbs
    return out


def packexp_from_public(secrets, pp, net: Net, g2: bool = False) -> np.ndarray:
    """dmsm/mod.rs:50-68: pack l group elements into n shares (iFFT over `secret`, FFT over `share`)."""
    c = _group_dft(net, secrets, pp.secret_size, inverse=True, coset=True, g2=g2)
    return _group_dft(net, c, pp.share_size, inverse=False, coset=False, g2=g2)


def unpackexp(shares, degree2: bool, pp, net: Net, g2: bool = False) -> np.ndarray:
    """dmsm/mod.rs:7-48: interpolate the n shares, evaluate on the secret (or secret2) coset, keep the secrets."""
    c = _group_dft(net, shares, pp.share_size, inverse=True, coset=False, g2=g2)
    if degree2:
        e = _group_dft(net, c, pp.secret2_size, inverse=False, coset=True, g2=g2)
        return e[: 2 * pp.l: 2]
    e = _group_dft(net, c, pp.secret_size, inverse=False, coset=True, g2=g2)
    return e[: pp.l]


def packexp_from_public_batch(points, pp, net: Net, g2: bool = False):
    """packexp_from_public of every l-point chunk of `points` in one kernel launch (b200zk_points_matmul_dev with the
    sharing's pack matrix): CUDA int64 (chunks * l, w) -> (chunks, n, w).  A trailing partial chunk is padded with the
    identity, as `cfg_chunks!` + `packexp_from_public`'s resize do (proving_key.rs:66-80, dmsm/mod.rs:54)."""
    import torch
    from .._native import c_vp
    w = 16 if g2 else 8
    k = int(points.shape[0])
    chunks = -(-k // pp.l)
    if chunks * pp.l != k:
        pad = torch.zeros((chunks * pp.l, w), dtype=torch.int64, device=points.device)
        pad[:k] = points
        points = pad
    m = net.to_device(pp.pack_matrix().reshape(-1, 4))
    out = torch.empty((chunks, pp.n, w), dtype=torch.int64, device=points.device)
    net.check(net._lib.b200zk_points_matmul_dev(net._h, 0, 1 if g2 else 0, c_vp(points.data_ptr()), chunks, pp.l,
                                                c_vp(m.data_ptr()), pp.n, c_vp(out.data_ptr())))
    net.sync(0)
    return out


def unpackexp_batch(shares, degree2: bool, pp, net: Net, g2: bool = False):
    """unpackexp of (chunks, n, w) share vectors in one launch -> (chunks, l, w)."""
    import torch
    from .._native import c_vp
    w = 16 if g2 else 8
    chunks = int(shares.shape[0])
    m = net.to_device(pp.unpack_matrix(degree2).reshape(-1, 4))
    out = torch.empty((chunks, pp.l, w), dtype=torch.int64, device=shares.device)
    net.check(net._lib.b200zk_points_matmul_dev(net._h, 0, 1 if g2 else 0, c_vp(shares.data_ptr()), chunks, pp.n,
                                                c_vp(m.data_ptr()), pp.l, c_vp(out.data_ptr())))
    net.sync(0)
    return out


def d_msm_mpc(bases_shares, scalar_shares, pp, net: Net, g2: bool = False) -> GroupElement:
    """The reference protocol itself, all n parties simulated on this GPU (what LocalTestNet does in-process,
    mpc-net/src/multi.rs:289-316): every party runs `G::msm` on its packed shares (dmsm/mod.rs:82); the king unpacks the
    degree-2 sharing in the exponent and sums the l secrets (dmsm/mod.rs:91-95).
    bases_shares[p], scalar_shares[p]: party p's share vectors."""
    w = 16 if g2 else 8
    c_shares = np.zeros((pp.n, w), dtype=np.uint64)
    for p in range(pp.n):
        limbs, inf = net.msm(bases_shares[p], scalar_shares[p], g2=g2)
        if not inf:
            c_shares[p] = limbs
    secrets = unpackexp(c_shares, True, pp, net, g2=g2)
    ones = _fr_limbs_mont([1] * secrets.shape[0])
    limbs, inf = net.msm(secrets, ones, g2=g2)
    return GroupElement(limbs, inf, g2)
