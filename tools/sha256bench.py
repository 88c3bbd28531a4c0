"""Config 4 alone (the reference's sha256 circuit, m = 2^15): GPU prove time through bench.measure_prove_sha256, plus the
small-MSM timings that dominate it.  Development aid; bench.py is the contract."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from distributed_groth16_b200 import Net  # noqa: E402

net = Net(0)
net.use_torch_stream(0)
r = bench.measure_prove_sha256(net, with_cpu="--cpu" in sys.argv)
print("sha256 prove: %.3f ms (min %.3f)%s" % (r["ms"], r["ms_min"], "  bytes == proof.bin: %s" % r.get("bytes_equal_reference_proof_bin") if "--cpu" in sys.argv else ""))
d = np.load(os.path.join(ROOT, "tests", "golden", "sha256_circuit.npz"))
z = net.fr_convert(net.to_device(d["witness"]), to_mont=True)
n = int(z.shape[0])
for g2 in (False, True):
    bases = net.generate_g2(5, n) if g2 else net.generate_g1(5, n)
    out = torch.empty(32 if g2 else 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        net.msm_dev(bases, z, out, g2=g2)
    evs = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.msm_dev(bases, z, out, g2=g2); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    print("MSM %s over the sha256 witness (n = %d, 0/1-heavy): %.3f ms" % ("G2" if g2 else "G1", n, ts[len(ts) // 2]))
