"""Groth16 prove timing on a synthetic instance with a dummy CRS (as groth16/examples/local_groth_bench.rs does):
m = n_vars = 2^log_m, n_inputs = 2.  Checks the 128 proof bytes against the CPU twin unless --no-check."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402
from distributed_groth16_b200.groth16 import ProvingKey, prove  # noqa: E402


def main():
    log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    check = "--no-check" not in sys.argv
    m = 1 << log_m
    n_vars, n_inputs = m, 2
    net = Net(0)
    net.use_torch_stream(0)
    aq = net.generate_g1(101, n_vars)
    b1 = net.generate_g1(102, n_vars)
    b2 = net.generate_g2(103, n_vars)
    lq = net.generate_g1(104, n_vars - n_inputs)
    hq = net.generate_g1(105, m)
    vk1 = net.generate_g1(106, 3).cpu().numpy().view(np.uint64)
    vk2 = net.generate_g2(107, 2).cpu().numpy().view(np.uint64)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    z = net.generate_fr(108, n_vars)
    one = np.array([12436184717236109307, 3962172157175319849, 7381016538464732718, 1011752739694698287], dtype=np.uint64)
    z[0] = torch.from_numpy(one.view(np.int64)).cuda()
    a, b, c = (net.generate_fr(s, m) for s in (109, 110, 111))
    pk = ProvingKey.from_device(net, aq, b1, b2, lq, hq, n_inputs, vk)
    times = []
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = prove.create_proof_dev(pk, z, a, b, c)
        times.append((time.perf_counter() - t0) * 1e3)
    net.profile(True)
    net.profile_reset()
    proof2 = prove.create_proof_dev(pk, z, a, b, c)
    rep = net.profile_report()
    net.profile(False)
    assert proof == proof2
    out = {"log_m": log_m, "prove_ms": sorted(times[1:])[len(times[1:]) // 2], "prove_ms_all": times,
           "kernel_ms": {k: round(v["ms"], 3) for k, v in rep.items()}, "launches": sum(v["launches"] for v in rep.values())}
    out["pk_table_gb"] = pk.table_bytes / 2**30
    print("prove 2^%d: %.2f ms (runs %s)  pk tables %.2f GB" % (log_m, out["prove_ms"], ["%.1f" % t for t in times], out["pk_table_gb"]))
    print("  kernels:", out["kernel_ms"])
    if check:
        from oracle import cref
        cref.build()
        ncores = os.cpu_count()
        h2 = lambda t: t.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        hh = cref.h_circom(h2(a), h2(b), h2(c), ncores)
        exp = cref.groth16_prove(h2(aq), h2(b1), h2(b2), h2(lq), h2(hq), vk, n_inputs, h2(z), hh, np.zeros(4, np.uint64),
                                 np.zeros(4, np.uint64), nthreads=ncores)
        dt = time.perf_counter() - t0
        out["cpu_ms"] = dt * 1e3
        out["cpu_cores"] = ncores
        out["bit_exact"] = bool(exp == proof)
        print("  CPU twin: %.0f ms on %d threads; bit-exact: %s" % (dt * 1e3, ncores, exp == proof))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/provebench_%d.json" % log_m, "w"), indent=1)


if __name__ == "__main__":
    main()
