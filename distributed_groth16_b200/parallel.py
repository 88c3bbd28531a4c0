"""Multi-GPU orchestration (one process per GPU, `torch.distributed`): the replacement of the
reference's king/client star (/root/reference/mpc-net/src/lib.rs:61-139, used by every `d_*` primitive
through dist-primitives/src/channel/mod.rs:8-56) for the single-box setting.

* MSM shards by index range; the only exchange is an all-gather of one XYZZ partial per rank.
* NTT is a four-step (Bailey) transform: local column NTTs with the w_N^(col*k1) twiddle fused into
  their last pass -> ONE all-to-all -> local row NTTs.  Distributed vectors live in the
  "column layout": for a length-N vector and `ncols` columns, rank g of P owns the columns
  n mod ncols in [g*ncols/P, (g+1)*ncols/P), each stored contiguously (N/ncols entries):
      local[c][r] = x[r*ncols + g*ncols/P + c].
  A transform with (rows, cols) maps layout(ncols = cols) -> layout(ncols = rows); chaining iNTT -> NTT
  (the h pipeline) therefore needs no redistribution: the second transform swaps rows and cols.

All field arithmetic is done by the CUDA library through a small backend object so that the host
logic (index maps, collectives) can be exercised with the CPU oracle under gloo in tests/."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._native import c_vp


# ---------------------------------------------------------------------------------------------
# layout helpers (host side, numpy) -- used by callers to scatter/gather whole vectors and by tests
# ---------------------------------------------------------------------------------------------
def to_column_layout(x: np.ndarray, ncols: int, world: int, rank: int) -> np.ndarray:
    """x: (N, 4) -> local (ncols/world, N/ncols, 4)."""
    n = x.shape[0]
    cg = ncols // world
    m = x.reshape(n // ncols, ncols, 4)
    return np.ascontiguousarray(m[:, rank * cg:(rank + 1) * cg].transpose(1, 0, 2))


def from_column_layout(parts, ncols: int) -> np.ndarray:
    """inverse of to_column_layout given every rank's local array (in rank order)."""
    world = len(parts)
    cg = ncols // world
    rows = parts[0].shape[1]
    out = np.empty((rows, ncols, 4), dtype=parts[0].dtype)
    for g, p in enumerate(parts):
        out[:, g * cg:(g + 1) * cg] = np.asarray(p).transpose(1, 0, 2)
    return out.reshape(rows * ncols, 4)


def split_log(log_n: int):
    """(log_rows, log_cols) of the four-step split: rows >= cols."""
    return (log_n + 1) // 2, log_n // 2


# ---------------------------------------------------------------------------------------------
# compute backend on the GPU (the only product backend)
# ---------------------------------------------------------------------------------------------
class GpuBackend:
    def __init__(self, net):
        self.net = net

    def batched_ntt_post(self, x, log_t, batch, inverse, log_base=0, shift=False, b0=0, alpha=0, beta=0, gamma=0,
                         post=True):
        import torch
        net = self.net
        out = torch.empty_like(x)
        if not post:
            net.check(net._lib.b200zk_ntt_fr_dev(net._h, 0, c_vp(x.data_ptr()), c_vp(out.data_ptr()), log_t, int(inverse),
                                                 0, batch))
        else:
            net.check(net._lib.b200zk_ntt_fr_batched_post_dev(net._h, 0, c_vp(x.data_ptr()), c_vp(out.data_ptr()), log_t,
                                                              batch, int(inverse), log_base, int(shift),
                                                              ctypes.c_uint64(b0), ctypes.c_uint64(alpha),
                                                              ctypes.c_uint64(beta), ctypes.c_uint64(gamma)))
        return out

    def mul_sub(self, a, b, c):
        import torch
        net = self.net
        out = torch.empty_like(a)
        net.check(net._lib.b200zk_fr_mul_sub_dev(net._h, 0, c_vp(a.data_ptr()), c_vp(b.data_ptr()), c_vp(c.data_ptr()),
                                                 c_vp(out.data_ptr()), a.numel() // 4))
        return out

    def msm_partial(self, bases, scalars, g2=False):
        return self.net.msm_dev(bases, scalars, g2=g2)

    def sum_points(self, xyzz, count, g2=False):
        return self.net.sum_points_dev(xyzz, count, g2=g2)


# ---------------------------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------------------------
def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def sharded_ntt(backend, local, log_rows: int, log_cols: int, inverse: bool = False, shift_log_m: int | None = None,
                group=None):
    """Four-step NTT of a length-2^(log_rows+log_cols) vector in layout(ncols = 2^log_cols).

    local: (cols/P, rows, 4) int64 tensor.  Returns (rows/P, cols, 4) = layout(ncols = 2^log_rows) of the
    transform.  shift_log_m: when set (inverse transforms of the h pipeline), output coefficient j is also
    multiplied by w_{2m}^j, m = 2^shift_log_m (the odd-coset shift of ext_wit::h / CircomReduction)."""
    import torch
    import torch.distributed as dist
    world, rank = _world(group)
    rows, cols = 1 << log_rows, 1 << log_cols
    assert rows % world == 0 and cols % world == 0, "rows and cols must be divisible by the number of GPUs"
    cg, rl = cols // world, rows // world
    assert tuple(local.shape) == (cg, rows, 4), (tuple(local.shape), (cg, rows, 4))
    log_n = log_rows + log_cols
    # 1+2: column transforms (size rows) with the twiddle w_N^((col0 + c) * k1) fused
    y = backend.batched_ntt_post(local.reshape(cg * rows, 4), log_rows, cg, inverse, log_base=log_n, shift=False,
                                 b0=rank * cg, alpha=1, beta=0, gamma=0)
    # 3: all-to-all transpose: destination g' receives the k1 in [g'*rl, (g'+1)*rl) of all my columns
    send = y.reshape(cg, world, rl, 4).permute(1, 0, 2, 3).contiguous()
    if world > 1:
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
    else:
        recv = send
    rows_in = recv.reshape(world, cg, rl, 4).permute(2, 0, 1, 3).contiguous()       # [k1 local][n2 global]
    # 4: row transforms (size cols); optional coefficient shift w_2m^(k1 + rows*k2)
    if shift_log_m is None:
        out = backend.batched_ntt_post(rows_in.reshape(rl * cols, 4), log_cols, rl, inverse, post=False)
    else:
        out = backend.batched_ntt_post(rows_in.reshape(rl * cols, 4), log_cols, rl, inverse, log_base=shift_log_m + 1,
                                       shift=True, b0=rank * rl, alpha=0, beta=1, gamma=rows)
    return out.reshape(rl, cols, 4)


def sharded_h(backend, a, b, c, log_m: int, group=None):
    """ext_wit::h (groth16/src/ext_wit.rs:16-101) on vectors in layout(ncols = 2^log_cols), (log_rows, log_cols) =
    split_log(log_m); returns h in the same layout.  One all-to-all per transform, six transforms (the 3+3 of
    ext_wit.rs:34-52, with the 2m-domain evaluation replaced by the coefficient shift of CircomReduction)."""
    log_rows, log_cols = split_log(log_m)
    ev = []
    for v in (a, b, c):
        coef = sharded_ntt(backend, v, log_rows, log_cols, inverse=True, shift_log_m=log_m, group=group)
        ev.append(sharded_ntt(backend, coef, log_cols, log_rows, inverse=False, group=group))
    return backend.mul_sub(ev[0], ev[1], ev[2])


def sharded_msm(backend, bases, scalars, g2: bool = False, group=None):
    """d_msm over length-sharded inputs: local Pippenger, all-gather of the XYZZ partials, local point sum
    (replaces send_to_king / unpackexp / recv_from_king, dist-primitives/src/dmsm/mod.rs:87-97)."""
    import torch
    import torch.distributed as dist
    world, _ = _world(group)
    part = backend.msm_partial(bases, scalars, g2)
    if world > 1:
        gathered = torch.empty((world, part.numel()), dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(gathered, part.reshape(1, -1), group=group)
    else:
        gathered = part.reshape(1, -1)
    return backend.sum_points(gathered, world, g2)


class PartialExchange:
    """d_msm's exchange without NCCL: one mailbox per rank (b200zk_peer_alloc + CUDA IPC, opened by every peer once), then per
    call ONE kernel per rank that stores its XYZZ partial into every peer's mailbox over NVLink, waits on sequence flags for
    everybody's partial, adds them up and normalises (`b200zk_msm_exchange_sum_dev`).  Replaces send_to_king / unpackexp /
    sum / recv_from_king of dist-primitives/src/dmsm/mod.rs:87-97 -- and round 1's all-gather + host-synchronising sum."""
    MAILBOX_BYTES = 8192

    def __init__(self, net, group=None):
        import torch
        import torch.distributed as dist
        self.net, self.group = net, group
        self.world, self.rank = _world(group)
        lib, h = net._lib, net._h
        ptr = c_vp()
        hb = (ctypes.c_uint8 * 64)()
        net.check(lib.b200zk_peer_alloc(h, self.MAILBOX_BYTES, ctypes.byref(ptr), hb))
        self.local = ptr.value
        self.boxes = [None] * self.world
        self.boxes[self.rank] = self.local
        if self.world > 1:
            dev = torch.device("cuda", net.device)
            mine = torch.tensor(list(bytes(hb)), dtype=torch.uint8, device=dev)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine.reshape(1, -1), group=group)     # also orders every rank's zero-fill before first use
            allh = allh.cpu().numpy()
            for g in range(self.world):
                if g != self.rank:
                    raw = (ctypes.c_uint8 * 64).from_buffer_copy(allh[g].tobytes())
                    p = c_vp()
                    net.check(lib.b200zk_peer_open(h, raw, ctypes.byref(p)))
                    self.boxes[g] = p.value
        self._ptrs = (c_vp * self.world)(*[c_vp(p) for p in self.boxes])
        self.seq = 0

    def sum(self, part, g2: bool = False, out=None, sid: int = 0):
        """part: this rank's XYZZ partial (CUDA int64 tensor).  Returns a CUDA tensor: affine limbs (8 / 16 words) + infinity flag,
        identical on every rank; nothing is synchronised with the host."""
        import torch
        w = 16 if g2 else 8
        if out is None:
            out = torch.empty(w + 1, dtype=torch.int64, device=part.device)
        self.seq += 1
        net = self.net
        net.check(net._lib.b200zk_msm_exchange_sum_dev(net._h, int(sid), 1 if g2 else 0, c_vp(part.data_ptr()), self._ptrs, self.world,
                                                       self.rank, ctypes.c_uint64(self.seq), c_vp(out.data_ptr())))
        return out

    def close(self):
        lib, h = self.net._lib, self.net._h
        for g in range(self.world):
            if g != self.rank and self.boxes[g]:
                lib.b200zk_peer_close(h, c_vp(self.boxes[g]))
        lib.b200zk_peer_free(h, c_vp(self.local))


# ---------------------------------------------------------------------------------------------
# multi-GPU prove (BASELINE config 5: MSM split + four-step NTT all-to-all)
# ---------------------------------------------------------------------------------------------
class ShardedProvingKey:
    """The slice of a proving key one rank keeps resident.

    a_query / b_g1_query / b_g2_query: rows [g*n_vars/P, (g+1)*n_vars/P) of the global arrays (rank 0's slice starts
    with index 0); l_query: the same fraction of its n_vars - n_inputs rows; h_query: in the column layout of
    `sharded_h` (local[c][r] = h_query[r*ncols + c0 + c], flattened), so that it lines up with the h this rank
    computes.  vk_points: the 56 limbs of b200zk_pk_upload (replicated)."""

    def __init__(self, net, a_query, b_g1_query, b_g2_query, l_query, h_query_cols, n_inputs, vk_points, tables=True):
        from .groth16.proving_key import ProvingKey
        self.net = net
        self.a_query, self.b_g1_query, self.b_g2_query = a_query, b_g1_query, b_g2_query
        self.l_query, self.h_query = l_query, h_query_cols.reshape(-1, 8)
        # fixed-base window tables of this rank's slices (b200zk_msm_table_*; same policy as b200zk_pk_precompute)
        self.tables = {}
        if tables and os.environ.get("B200ZK_PK_TABLES", "1") != "0":
            for name, g2 in (("a_query", False), ("b_g1_query", False), ("b_g2_query", True), ("l_query", False),
                             ("h_query", False)):
                q = getattr(self, name)
                n = int(q.shape[0])
                if n >= 64:
                    c = net.msm_table_auto_window(n)
                    self.tables[name] = (net.msm_table_build(q, c, g2=g2), c)
        # the device-side pk object is only used for query[0] / vk in the final assembly
        self.pk = ProvingKey.from_device(net, a_query[:1], b_g1_query[:1], b_g2_query[:1], a_query[:0], h_query_cols.reshape(-1, 8)[:1],
                                         1, vk_points)


def sharded_prove(net, spk: ShardedProvingKey, z_shard, z_aux_shard, a, b, c, log_m: int, r=None, s=None, group=None,
                  xch=None):
    """Groth16 proof with every vector sharded over the ranks of `group`.

    z_shard: this rank's rows of the full assignment (aligned with spk.a_query); z_aux_shard: its rows of
    z[n_inputs:] (aligned with spk.l_query); a, b, c: QAP evaluations in the column layout (see sharded_ntt).
    xch: a P2PExchange -> the h pipeline uses the fused NTT + NVLink exchange instead of NCCL all-to-all.
    Every rank returns the same 128 bytes."""
    import torch
    import torch.distributed as dist
    world, rank = _world(group)
    be = GpuBackend(net)
    zero = np.zeros(4, dtype=np.uint64)
    r = zero if r is None else np.ascontiguousarray(r, dtype=np.uint64)
    s = zero if s is None else np.ascontiguousarray(s, dtype=np.uint64)
    need_b1 = bool(r.any())
    if xch is not None:          # fused four-step: NTT column kernels store straight into the peers' buffers
        h = sharded_h_p2p(net, xch, a, b, c, log_m).reshape(-1, 4)
    else:
        h = sharded_h(be, a, b, c, log_m, group=group).reshape(-1, 4)
    dev = z_shard.device
    # slots: 0 A(G1) 1 L 2 H 3 B1 (each 16 words) then B2 (32 words)
    parts = torch.zeros(4 * 16 + 32, dtype=torch.int64, device=dev)
    def msm(name, scalars, out, g2=False):
        if name in spk.tables:
            table, c = spk.tables[name]
            net.msm_table_dev(table, scalars, c, out, g2=g2)
        else:
            net.msm_dev(getattr(spk, name), scalars, out, g2=g2)
    msm("a_query", z_shard, parts[0:16])
    msm("l_query", z_aux_shard, parts[16:32])
    msm("h_query", h, parts[32:48])
    if need_b1:
        msm("b_g1_query", z_shard, parts[48:64])
    msm("b_g2_query", z_shard, parts[64:96], g2=True)
    if world > 1:
        gathered = torch.empty((world, parts.numel()), dtype=parts.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, parts.reshape(1, -1), group=group)
    else:
        gathered = parts.reshape(1, -1)
    tot = torch.empty_like(parts)
    lib, hnd = net._lib, net._h
    for k in range(4):        # G1 slots: stride between ranks = 96 words = 6 G1-XYZZ points
        net.check(lib.b200zk_xyzz_sum_dev(hnd, 0, 0, c_vp(gathered.data_ptr() + k * 128), world, 6, c_vp(tot.data_ptr() + k * 128)))
    net.check(lib.b200zk_xyzz_sum_dev(hnd, 0, 1, c_vp(gathered.data_ptr() + 512), world, 3, c_vp(tot.data_ptr() + 512)))
    out = (ctypes.c_uint8 * 128)()
    base = tot.data_ptr()
    net.check(lib.b200zk_groth16_assemble_dev(hnd, spk.pk._h, c_vp(base), c_vp(base + 512), c_vp(base + 128), c_vp(base + 256),
                                              c_vp(base + 384) if need_b1 else None, c_vp(r.ctypes.data), c_vp(s.ctypes.data),
                                              0, out))
    return bytes(out)


# ---------------------------------------------------------------------------------------------
# fused compute + exchange: the four-step NTT with the all-to-all folded into the column kernel
# ---------------------------------------------------------------------------------------------
class P2PExchange:
    """Peer-mapped receive buffers for `sharded_ntt_p2p`.

    Every rank allocates two row-major receive buffers with b200zk_peer_alloc (cudaMalloc + CUDA IPC), the 64-byte
    handles are all-gathered once, and every rank maps every peer's buffers (NVLink peer access).  The column kernel
    of a transform then stores its outputs directly into the owners' buffers; two buffers alternate so that a rank
    still reading transform t's rows is never overwritten by a faster peer already in transform t+1 (the barrier of
    transform t+1 orders transform t+2's writes after those reads)."""

    def __init__(self, net, n_elems: int, group=None):
        import torch
        import torch.distributed as dist
        self.net, self.group = net, group
        self.world, self.rank = _world(group)
        self.n_elems = n_elems
        self.bytes = n_elems * 32
        lib, h = net._lib, net._h
        self.local = []
        handles = []
        for _ in range(2):
            ptr = c_vp()
            hb = (ctypes.c_uint8 * 64)()
            net.check(lib.b200zk_peer_alloc(h, self.bytes, ctypes.byref(ptr), hb))
            self.local.append(ptr.value)
            handles.append(bytes(hb))
        self.peers = [[None] * self.world for _ in range(2)]
        if self.world > 1:
            dev = torch.device("cuda", net.device)
            mine = torch.tensor(list(handles[0] + handles[1]), dtype=torch.uint8, device=dev)
            allh = torch.empty((self.world, 128), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine.reshape(1, -1), group=group)
            allh = allh.cpu().numpy()
            for g in range(self.world):
                for b in range(2):
                    if g == self.rank:
                        self.peers[b][g] = self.local[b]
                    else:
                        raw = (ctypes.c_uint8 * 64).from_buffer_copy(allh[g, 64 * b:64 * (b + 1)].tobytes())
                        ptr = c_vp()
                        net.check(lib.b200zk_peer_open(h, raw, ctypes.byref(ptr)))
                        self.peers[b][g] = ptr.value
        else:
            for b in range(2):
                self.peers[b][0] = self.local[b]
        self.turn = 0
        self._flag = None

    def recv_tensor(self, b: int, shape):
        """torch view of local receive buffer b (no copy)."""
        import torch

        class _Holder:
            pass
        hold = _Holder()
        n = int(np.prod(shape))
        hold.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (self.local[b], False), "version": 2}
        return torch.as_tensor(hold, device=torch.device("cuda", self.net.device)).reshape(shape)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            if self._flag is None:
                self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", self.net.device))
            dist.all_reduce(self._flag, group=self.group)      # stream-ordered: all ranks' column kernels have completed

    def close(self):
        lib, h = self.net._lib, self.net._h
        for b in range(2):
            for g in range(self.world):
                if g != self.rank and self.peers[b][g]:
                    lib.b200zk_peer_close(h, c_vp(self.peers[b][g]))
            lib.b200zk_peer_free(h, c_vp(self.local[b]))


def sharded_ntt_p2p(net, xch: P2PExchange, local, log_rows: int, log_cols: int, inverse: bool = False,
                    shift_log_m: int | None = None):
    """Same transform and layouts as `sharded_ntt`, with the exchange fused into the column kernel: no pack,
    no NCCL all-to-all, no unpack -- the kernel's stores are the transfer (b200zk_ntt_fr_fourstep_cols_p2p_dev)."""
    world, rank = xch.world, xch.rank
    rows, cols = 1 << log_rows, 1 << log_cols
    cg, rl = cols // world, rows // world
    assert tuple(local.shape) == (cg, rows, 4) and rl * cols <= xch.n_elems
    b = xch.turn
    xch.turn ^= 1
    ptrs = (c_vp * world)(*[c_vp(p) for p in xch.peers[b]])
    net.check(net._lib.b200zk_ntt_fr_fourstep_cols_p2p_dev(net._h, 0, c_vp(local.data_ptr()), ptrs, world, log_rows,
                                                           cg.bit_length() - 1, log_rows + log_cols,
                                                           ctypes.c_uint64(rank * cg), int(inverse)))
    xch.barrier()
    rows_in = xch.recv_tensor(b, (rl * cols, 4))
    be = GpuBackend(net)
    if shift_log_m is None:
        out = be.batched_ntt_post(rows_in, log_cols, rl, inverse, post=False)
    else:
        out = be.batched_ntt_post(rows_in, log_cols, rl, inverse, log_base=shift_log_m + 1, shift=True, b0=rank * rl,
                                  alpha=0, beta=1, gamma=rows)
    return out.reshape(rl, cols, 4)


def sharded_h_p2p(net, xch: P2PExchange, a, b, c, log_m: int):
    """`sharded_h` on the fused transforms."""
    log_rows, log_cols = split_log(log_m)
    ev = []
    for v in (a, b, c):
        coef = sharded_ntt_p2p(net, xch, v, log_rows, log_cols, inverse=True, shift_log_m=log_m)
        ev.append(sharded_ntt_p2p(net, xch, coef, log_cols, log_rows, inverse=False))
    return GpuBackend(net).mul_sub(ev[0], ev[1], ev[2])
