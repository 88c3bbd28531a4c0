"""End-to-end MSM timing through the host-buffer C-ABI call (b200zk_msm_g1 / _g2) with pinned host memory -- what bench.py's
`e2e` key measures -- for A/B runs of the staging switches (B200ZK_MSM_PART_WEIGHTS, B200ZK_MSM_PARTS).  Development aid."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402

net = Net(0)
net.use_torch_stream(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for log_n in [int(x) for x in (sys.argv[1:] or ["20"])]:
    n = 1 << log_n
    bases, scalars = net.generate_g1(0xB2000002, n), net.generate_fr(0xB2000002, n)
    hb = torch.empty((n, 8), dtype=torch.int64).pin_memory()
    hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hb.copy_(bases); hs.copy_(scalars)
    hb_np, hs_np = hb.numpy().view(np.uint64), hs.numpy().view(np.uint64)
    ref = None
    for _ in range(3):
        ref = net.msm(hb_np, hs_np)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        r = net.msm(hb_np, hs_np)
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
        assert (r[0] == ref[0]).all()
    ts.sort()
    print("e2e G1 MSM 2^%d: %.3f ms median (min %.3f)  %.1f Mpairs/s   weights=%s parts=%s" % (
        log_n, ts[len(ts) // 2], ts[0], n / ts[len(ts) // 2] / 1e3, os.environ.get("B200ZK_MSM_PART_WEIGHTS", "default"),
        os.environ.get("B200ZK_MSM_PARTS", "-")))
