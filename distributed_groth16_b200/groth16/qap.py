"""`qap::qap` and its carriers -- /root/reference/groth16/src/qap.rs:17-91, groth16/src/lib.rs:11-35.

`qap(matrices, full_assignment, net)` keeps the reference's name and argument meaning: R1CS matrices A, B
(`ConstraintMatrices`, here CSR arrays resident in HBM) times the full assignment z -> the three QAP
evaluation vectors over the domain of size next_pow2(num_constraints + num_inputs) (qap.rs:53).  The sparse
mat-vec runs on the GPU (csrc/qap.cu); the host only reorders index arrays."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .._native import c_vp
from ..formats import coo_to_csr


@dataclass
class Radix2Domain:
    m: int

    def size(self) -> int:
        return self.m


@dataclass
class PackedQAPShare:
    num_inputs: int
    num_constraints: int
    a: np.ndarray      # (m, 4) Fr Montgomery limbs; bit-reversed order when `rearranged` (QAP::pss, qap.rs:151-152)
    b: np.ndarray
    c: np.ndarray
    domain: Radix2Domain
    rearranged: bool = False


class ConstraintMatrices:
    """ark_relations::r1cs::ConstraintMatrices (A and B only, as `qap()` uses them) in CSR form on the device."""

    def __init__(self, net, num_instance_variables: int, num_constraints: int, a_coo, b_coo, values_montgomery_depth: int):
        """a_coo / b_coo: (rows u32, cols u32, vals (nnz, 4) u64).  values_montgomery_depth: how many Montgomery
        reductions turn the stored words into Montgomery form: 1 for zkey coefficients (stored * R^2,
        ark-circom/src/zkey.rs:333-338), -1 for canonical values (r1cs), 0 if already Montgomery."""
        self.net = net
        self.num_instance_variables = int(num_instance_variables)
        self.num_constraints = int(num_constraints)
        self.csr = []
        for rows, cols, vals in (a_coo, b_coo):
            ptr, col, val = coo_to_csr(np.asarray(rows), np.asarray(cols), np.asarray(vals), self.num_constraints)
            d_val = net.to_device(val.reshape(-1, 4)) if val.size else net.to_device(np.zeros((1, 4), dtype=np.uint64))
            if values_montgomery_depth > 0:
                d_val = net.fr_convert(d_val, to_mont=False, times=values_montgomery_depth)
            elif values_montgomery_depth < 0:
                d_val = net.fr_convert(d_val, to_mont=True, times=-values_montgomery_depth)
            d_col = net.to_device(col.view(np.int32) if col.size else np.zeros(1, dtype=np.int32))
            self.csr.append((net.to_device(ptr.view(np.int32)), d_col, d_val))


@dataclass
class QAP:
    """groth16/src/qap.rs:17-29 with the vectors resident in HBM (CUDA int64 (m, 4) tensors)."""
    num_inputs: int
    num_constraints: int
    a: object
    b: object
    c: object
    domain: Radix2Domain


def qap_pss(q: QAP, pp) -> list:
    """`QAP::pss` (groth16/src/qap.rs:143-187): bit-reverse a, b, c, cut each into m/l strided chunks
    (x[i], x[i + m/l], ...), pack every chunk, and hand party p the p-th share of every chunk.
    Returns pp.n PackedQAPShare objects whose vectors are CUDA int64 (m/l, 4) tensors."""
    import torch
    from ..dist_primitives.dfft import bitrev_indices
    m = q.domain.size()
    assert m % pp.l == 0
    idx = torch.from_numpy(bitrev_indices(m)).to(q.a.device)

    def pack(x):
        xr = torch.empty_like(x)
        xr[idx] = x                                                   # fft_in_place_rearrange
        chunks = xr.reshape(pp.l, m // pp.l, 4).permute(1, 0, 2).contiguous()     # chunk i = x[i], x[i + m/l], ...
        return pp.pack_from_public_batch(chunks)                      # (m/l, n, 4)

    pa, pb, pc = pack(q.a), pack(q.b), pack(q.c)
    return [PackedQAPShare(q.num_inputs, q.num_constraints, pa[:, p].contiguous(), pb[:, p].contiguous(), pc[:, p].contiguous(),
                           q.domain, rearranged=True) for p in range(pp.n)]


def qap(matrices: ConstraintMatrices, full_assignment, net=None) -> QAP:
    """full_assignment: CUDA int64 (n_vars, 4) tensor, Montgomery form."""
    import torch
    net = net or matrices.net
    num_inputs, nc = matrices.num_instance_variables, matrices.num_constraints
    m = 1
    while m < nc + num_inputs:          # D::new(num_constraints + num_inputs)
        m <<= 1
    log_m = m.bit_length() - 1
    a = torch.empty((m, 4), dtype=torch.int64, device=full_assignment.device)
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    (ap, ac, av), (bp, bc, bv) = matrices.csr
    net.check(net._lib.b200zk_qap_dev(net._h, 0, c_vp(ap.data_ptr()), c_vp(ac.data_ptr()), c_vp(av.data_ptr()),
                                      c_vp(bp.data_ptr()), c_vp(bc.data_ptr()), c_vp(bv.data_ptr()), nc, num_inputs,
                                      c_vp(full_assignment.data_ptr()), log_m, c_vp(a.data_ptr()), c_vp(b.data_ptr()),
                                      c_vp(c.data_ptr())))
    return QAP(num_inputs, nc, a, b, c, Radix2Domain(m))
