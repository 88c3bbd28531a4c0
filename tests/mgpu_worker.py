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
    flat = torch.tensor([int(all(res.values()))], device=dev)
    dist.all_reduce(flat, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps({"ok": bool(flat.item()), "world": world, "rank0": res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
