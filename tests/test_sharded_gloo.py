"""N > 1 host logic on CPU: world_size-2 gloo, compute done by the oracle backend (no GPU).
Covers the four-step index maps / all-to-all of parallel.sharded_ntt, the chained iNTT->NTT layout of
parallel.sharded_h and the all-gather + sum of parallel.sharded_msm."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mgpu_common as mc
        res = mc.check_all(mc.OracleBackend(), lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)), rank, world)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_paths_under_gloo(world, cref):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res in results:
        assert all(res.values()), (rank, res)


def test_column_layout_roundtrip(cref):
    from distributed_groth16_b200 import parallel as par
    x = cref.fr_generate(1, 64)
    for world in (1, 2, 4):
        parts = [par.to_column_layout(x, 8, world, g) for g in range(world)]
        assert (par.from_column_layout(parts, 8) == x).all()
