"""BASELINE config 5: Groth16 prove of a 2^log_m-constraint synthetic circuit sharded over all ranks
(MSM split + four-step NTT all-to-all).  Launch: torchrun --nproc-per-node P tools/prove_sharded_bench.py [log_m].
Rank 0 also proves the whole instance on its own GPU and checks that the 128 proof bytes are identical."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from distributed_groth16_b200 import Net, parallel as par  # noqa: E402
from distributed_groth16_b200._constants import FR_ONE_MONT  # noqa: E402
from distributed_groth16_b200.groth16 import ProvingKey, prove  # noqa: E402


def cols_layout(t, ncols, world, rank):
    """device tensor (N, w) -> column layout (ncols/world, N/ncols, w) of this rank."""
    n, w = t.shape
    cg = ncols // world
    return t.reshape(n // ncols, ncols, w)[:, rank * cg:(rank + 1) * cg].permute(1, 0, 2).contiguous()


def main():
    log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    net = Net(local)
    net.use_torch_stream(0)
    m = 1 << log_m
    n_vars, n_inputs = m, 2
    n_aux = n_vars - n_inputs
    log_rows, log_cols = par.split_log(log_m)
    cols = 1 << log_cols
    # global dummy instance (same seeds on every rank), then this rank's slices
    aq, b1 = net.generate_g1(301, n_vars), net.generate_g1(302, n_vars)
    b2 = net.generate_g2(303, n_vars)
    lq, hq = net.generate_g1(304, n_aux), net.generate_g1(305, m)
    vk = np.concatenate([net.generate_g1(306, 3).cpu().numpy().view(np.uint64).reshape(-1),
                         net.generate_g2(307, 2).cpu().numpy().view(np.uint64).reshape(-1)])
    z = net.generate_fr(308, n_vars)
    z[0] = torch.from_numpy(np.array(FR_ONE_MONT, dtype=np.uint64).view(np.int64)).to(z.device)
    a, b, c = (net.generate_fr(sd, m) for sd in (309, 310, 311))
    sl = slice(rank * n_vars // world, (rank + 1) * n_vars // world)
    sla = slice(rank * n_aux // world, (rank + 1) * n_aux // world)
    spk = par.ShardedProvingKey(net, aq[sl].contiguous(), b1[sl].contiguous(), b2[sl].contiguous(), lq[sla].contiguous(),
                                cols_layout(hq, cols, world, rank), n_inputs, vk)
    z_sh, zaux_sh = z[sl].contiguous(), z[n_inputs:][sla].contiguous()
    la, lb, lc = (cols_layout(v, cols, world, rank) for v in (a, b, c))
    rows = 1 << log_rows
    xch = par.P2PExchange(net, rows * rows // world)
    out = {"config": "Groth16 prove 2^%d synthetic, sharded x%d" % (log_m, world), "log_m": log_m, "world": world,
           "fixed_base_table_gb_per_rank": sum(t.numel() * 8 for t, _ in spk.tables.values()) / 2**30,
           "fixed_base_windows": {k: c for k, (_, c) in spk.tables.items()}}
    proofs = {}
    for mode, x in (("nccl_all_to_all", None), ("p2p_fused", xch)):
        times = []
        for it in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            proof = par.sharded_prove(net, spk, z_sh, zaux_sh, la, lb, lc, log_m, xch=x)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor(sorted(times[2:])[1], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_prove_sharded_" + mode] = float(t)
        proofs[mode] = proof
        # the h pipeline alone (6 transforms + pointwise), same two exchange mechanisms
        ht = []
        be = par.GpuBackend(net)
        for it in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if x is None:
                par.sharded_h(be, la, lb, lc, log_m)
            else:
                par.sharded_h_p2p(net, x, la, lb, lc, log_m)
            torch.cuda.synchronize()
            ht.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor(sorted(ht[1:])[1], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_h_sharded_" + mode] = float(t)
    assert proofs["nccl_all_to_all"] == proofs["p2p_fused"]
    proof = proofs["p2p_fused"]
    if rank == 0:
        pk = ProvingKey.from_device(net, aq, b1, b2, lq, hq, n_inputs, vk)
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            single = prove.create_proof_dev(pk, z, a, b, c)
            ts.append((time.perf_counter() - t0) * 1e3)
        out["ms_prove_single_gpu"] = sorted(ts)[1]
        out["single_gpu_table_gb"] = pk.table_bytes / 2**30
        out["bit_exact_sharded_vs_single_gpu"] = bool(single == proof)
        out["proof_hex"] = proof.hex()
        print("SHARDED_PROVE " + json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
