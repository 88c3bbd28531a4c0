"""Sharded MSM / four-step NTT / h: single-GPU degenerate case always, real multi-GPU (NCCL) when the box has >= 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_paths_world1(net):
    """world = 1 exercises the batched-post NTT kernels (column twiddle, coefficient shift) and the layout maps."""
    import torch
    sys.path.insert(0, HERE)
    import mgpu_common as mc
    from distributed_groth16_b200 import parallel as par
    net.use_torch_stream(0)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    for log_m in (2, 5, 10, 15):
        res = mc.check_all(par.GpuBackend(net), to_dev, 0, 1, log_m=log_m, msm_n=256)
        assert all(res.values()), (log_m, res)
    for log_m in (3, 8, 12):
        assert mc.check_p2p(net, to_dev, 0, 1, log_m=log_m)
    for log_m, rs in ((4, (0, 0)), (9, (3, 4))):
        assert mc.check_sharded_prove(net, to_dev, 0, 1, log_m=log_m, rs=rs)


def test_sharded_paths_nccl_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-4000:]
    res = json.loads(lines[-1][len("MGPU_RESULT "):])
    assert res["ok"], res
