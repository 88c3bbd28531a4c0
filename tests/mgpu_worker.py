"""torchrun worker: the sharded paths on real GPUs (NCCL).  Launched by tests/test_gpu_multi.py."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from distributed_groth16_b200 import Net, parallel as par
    import mgpu_common as mc
    net = Net(local)
    net.use_torch_stream(0)
    dev = torch.device("cuda", local)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)
    res = {}
    for log_m in (6, 12, 17):
        r = mc.check_all(par.GpuBackend(net), to_dev, rank, world, log_m=log_m, msm_n=1 << 10)
        res.update({"%s@2^%d" % (k, log_m): v for k, v in r.items()})
    for log_m, rs in ((8, (0, 0)), (12, (5, 7))):
        res["sharded_prove@2^%d" % log_m] = bool(mc.check_sharded_prove(net, to_dev, rank, world, log_m=log_m, rs=rs))
    for log_m in (6, 13, 16):
        res["p2p_fused_ntt@2^%d" % log_m] = bool(mc.check_p2p(net, to_dev, rank, world, log_m=log_m))
    # d_msm over length-sharded inputs through the peer-mailbox exchange kernel (several calls: the sequence flags and the two
    # mailbox parities are exercised), G1 and G2, and prove::A with L != identity on every rank (the extra term is added once)
    from oracle import cref, layout
    from distributed_groth16_b200.dist_primitives import d_msm
    from distributed_groth16_b200.groth16 import prove
    ok = True
    for it, n_loc in enumerate((257, 1 << 12, 1 << 12, 33)):
        bases, scalars = cref.g1_generate(900 + it, n_loc * world), cref.fr_generate(950 + it, n_loc * world)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        got = d_msm(to_dev(bases[sl]), to_dev(scalars[sl]), None, net)
        exp, inf = cref.msm_g1(bases, scalars)
        ok = ok and got.infinity == bool(inf) and bool((got.limbs == exp).all())
    b2, s2 = cref.g2_generate(77, 300 * world), cref.fr_generate(78, 300 * world)
    got = d_msm(to_dev(b2[rank * 300:(rank + 1) * 300]), to_dev(s2[rank * 300:(rank + 1) * 300]), None, net, g2=True)
    exp, inf = cref.msm_g2(b2, s2)
    ok = ok and (not got.infinity) and bool((got.limbs == exp).all())
    res["d_msm_mailbox_exchange"] = bool(ok)
    S, a = cref.g1_generate(990, 64 * world), cref.fr_generate(991, 64 * world)
    L, N = cref.g1_generate(992, 2)
    r = layout.fr_to_arr([12345])[0]
    A = prove.A(L=L, N=N, r=r, pp=None, S=to_dev(S[rank * 64:(rank + 1) * 64]), a=to_dev(a[rank * 64:(rank + 1) * 64])).compute(net)
    one = layout.fr_to_arr([1])[0]
    expA, _ = cref.msm_g1(np.concatenate([S, L[None], N[None]]), np.concatenate([a, one[None], r[None]]))
    res["prove_A_multi_rank_L_added_once"] = bool((A.limbs == expA).all())
    flat = torch.tensor([int(all(res.values()))], device=dev)
    dist.all_reduce(flat, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps({"ok": bool(flat.item()), "world": world, "rank0": res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
