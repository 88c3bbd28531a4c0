"""Shared checks for the sharded (multi-rank) paths; used by the gloo CPU test (oracle backend) and by the
GPU worker (CUDA backend)."""
import numpy as np

from distributed_groth16_b200 import parallel as par


class OracleBackend:
    """CPU stand-in for parallel.GpuBackend built on the oracle: lets the host logic (index maps, collectives)
    run under gloo with no GPU.  TEST ONLY."""

    def __init__(self):
        from oracle import bn254 as o, cref, layout
        self.o, self.cref, self.L = o, cref, layout

    def batched_ntt_post(self, x, log_t, batch, inverse, log_base=0, shift=False, b0=0, alpha=0, beta=0, gamma=0, post=True):
        import torch
        o, L = self.o, self.L
        t = 1 << log_t
        arr = x.numpy().view(np.uint64).reshape(batch, t, 4)
        out = np.empty_like(arr)
        if post:
            base = o.fr_root_of_unity(1 << log_base)
            if inverse and not shift:
                base = pow(base, -1, o.R)
        for b in range(batch):
            v = self.cref.ntt(arr[b], inverse=inverse)
            if post:
                vals = L.arr_to_fr(v)
                vals = [val * pow(base, (b + b0) * (alpha * k + beta) + gamma * k, o.R) % o.R for k, val in enumerate(vals)]
                v = L.fr_to_arr(vals)
            out[b] = v
        return torch.from_numpy(out.reshape(batch * t, 4).view(np.int64))

    def mul_sub(self, a, b, c):
        import torch
        L, o = self.L, self.o
        va, vb, vc = (L.arr_to_fr(t.numpy().view(np.uint64).reshape(-1, 4)) for t in (a, b, c))
        out = L.fr_to_arr([(x * y - z) % o.R for x, y, z in zip(va, vb, vc)])
        return torch.from_numpy(out.view(np.int64)).reshape(a.shape)

    def msm_partial(self, bases, scalars, g2=False):
        import torch
        res, inf = self.cref.msm_g1(bases.numpy().view(np.uint64), scalars.numpy().view(np.uint64))
        xyzz = np.zeros(16, dtype=np.uint64)
        if not inf:
            one = self.L.fq_to_limbs(1)
            xyzz[:8] = res
            xyzz[8:12] = one
            xyzz[12:16] = one
        return torch.from_numpy(xyzz.view(np.int64))

    def sum_points(self, xyzz, count, g2=False):
        arr = xyzz.numpy().view(np.uint64).reshape(count, 16)
        pts = np.array([row[:8] if row[8:].any() else np.zeros(8, dtype=np.uint64) for row in arr])
        return self.cref.msm_g1(pts, self.L.fr_to_arr([1] * count))


def check_all(backend, to_dev, rank, world, log_m=6, msm_n=64):
    """Runs on every rank; returns a dict of booleans."""
    import torch
    from oracle import cref, layout
    res = {}
    m = 1 << log_m
    log_rows, log_cols = par.split_log(log_m)
    rows, cols = 1 << log_rows, 1 << log_cols
    x = cref.fr_generate(0xABC, m)
    loc = to_dev(par.to_column_layout(x, cols, world, rank))
    for inverse in (False, True):
        out = par.sharded_ntt(backend, loc, log_rows, log_cols, inverse=inverse)
        exp = par.to_column_layout(cref.ntt(x, inverse=inverse), rows, world, rank)
        res["ntt_inverse=%s" % inverse] = bool((out.cpu().numpy().view(np.uint64) == exp).all())
    a, b, c = (cref.fr_generate(s, m) for s in (11, 12, 13))
    la, lb, lc = (to_dev(par.to_column_layout(v, cols, world, rank)) for v in (a, b, c))
    h = par.sharded_h(backend, la, lb, lc, log_m)
    exp_h = par.to_column_layout(cref.h_circom(a, b, c), cols, world, rank)
    res["h"] = bool((h.cpu().numpy().view(np.uint64) == exp_h).all())
    bases = cref.g1_generate(77, msm_n * world)
    scalars = cref.fr_generate(78, msm_n * world)
    sl = slice(rank * msm_n, (rank + 1) * msm_n)
    got, inf = par.sharded_msm(backend, to_dev(bases[sl]), to_dev(scalars[sl]))
    exp, einf = cref.msm_g1(bases, scalars)
    res["msm"] = bool(inf == einf and (got == exp).all())
    return res


def check_sharded_prove(net, to_dev, rank, world, log_m=10, rs=(0, 0)):
    """GPU backend only: sharded_prove on `world` ranks vs the CPU twin's proof for the same global instance."""
    import torch
    from oracle import cref, layout
    m = 1 << log_m
    n_vars, n_inputs = m, 2
    log_rows, log_cols = par.split_log(log_m)
    cols = 1 << log_cols
    aq, b1, lq, hq = cref.g1_generate(201, n_vars), cref.g1_generate(202, n_vars), cref.g1_generate(204, n_vars - n_inputs), cref.g1_generate(205, m)
    b2 = cref.g2_generate(203, n_vars)
    vk1, vk2 = cref.g1_generate(206, 3), cref.g2_generate(207, 2)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    z = cref.fr_generate(208, n_vars)
    z[0] = layout.fr_to_arr([1])[0]
    a, b, c = (cref.fr_generate(sd, m) for sd in (209, 210, 211))
    r, s = layout.fr_to_arr([rs[0]])[0], layout.fr_to_arr([rs[1]])[0]
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), r, s)
    sl = slice(rank * n_vars // world, (rank + 1) * n_vars // world)
    n_aux = n_vars - n_inputs
    sl_aux = slice(rank * n_aux // world, (rank + 1) * n_aux // world)
    spk = par.ShardedProvingKey(net, to_dev(aq[sl]), to_dev(b1[sl]), to_dev(b2[sl]), to_dev(lq[sl_aux]),
                                to_dev(_cols_g1(hq, cols, world, rank)), n_inputs, vk)
    la, lb, lc = (to_dev(par.to_column_layout(v, cols, world, rank)) for v in (a, b, c))
    got = par.sharded_prove(net, spk, to_dev(z[sl]), to_dev(z[n_inputs:][sl_aux]), la, lb, lc, log_m, r, s)
    return got == exp


def _cols_g1(pts, ncols, world, rank):
    """column layout for an (N, 8) point array (same index map as to_column_layout)."""
    n = pts.shape[0]
    cg = ncols // world
    mview = pts.reshape(n // ncols, ncols, 8)
    return np.ascontiguousarray(mview[:, rank * cg:(rank + 1) * cg].transpose(1, 0, 2))


def check_p2p(net, to_dev, rank, world, log_m=10):
    """Fused four-step (P2P stores) == NCCL four-step == oracle."""
    from oracle import cref
    m = 1 << log_m
    log_rows, log_cols = par.split_log(log_m)
    rows, cols = 1 << log_rows, 1 << log_cols
    xch = par.P2PExchange(net, max(rows, cols) * max(rows, cols) // world)
    ok = True
    x = cref.fr_generate(0xF00D, m)
    loc = to_dev(par.to_column_layout(x, cols, world, rank))
    for inverse in (False, True):
        out = par.sharded_ntt_p2p(net, xch, loc, log_rows, log_cols, inverse=inverse)
        exp = par.to_column_layout(cref.ntt(x, inverse=inverse), rows, world, rank)
        ok = ok and bool((out.cpu().numpy().view(np.uint64) == exp).all())
    a, b, c = (cref.fr_generate(sd, m) for sd in (21, 22, 23))
    la, lb, lc = (to_dev(par.to_column_layout(v, cols, world, rank)) for v in (a, b, c))
    h = par.sharded_h_p2p(net, xch, la, lb, lc, log_m)
    ok = ok and bool((h.cpu().numpy().view(np.uint64) == par.to_column_layout(cref.h_circom(a, b, c), cols, world, rank)).all())
    import torch
    torch.cuda.synchronize()
    xch.close()
    return ok
