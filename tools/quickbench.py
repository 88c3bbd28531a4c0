"""Quick device-side timing of the two hot kernels (development aid; bench.py is the contract)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    net = Net(0)
    net.use_torch_stream(0)
    out = {}
    for log_n in [int(x) for x in (sys.argv[1:] or ["16", "20", "22"])]:
        n = 1 << log_n
        bases = net.generate_g1(1, n)
        scalars = net.generate_fr(2, n)
        xy = torch.empty(16, dtype=torch.int64, device="cuda")
        net.profile(True)
        net.profile_reset()
        med, best = timed(lambda: net.msm_dev(bases, scalars, xy))
        rep = net.profile_report()
        net.profile(False)
        med2, best2 = timed(lambda: net.msm_dev(bases, scalars, xy))
        out["msm_g1_2^%d" % log_n] = dict(ms_median=med2, ms_min=best2, mpairs_s=n / med2 / 1e3,
                                          kernels={k: v["ms"] / max(v["launches"], 1) * (v["launches"] / 7.0)
                                                   for k, v in rep.items()})
        print("MSM G1 2^%d: %.3f ms (min %.3f)  %.1f Mpairs/s" % (log_n, med2, best2, n / med2 / 1e3))
        print("   per-call kernel ms:", {k: round(v["ms"] / 7.0, 4) for k, v in rep.items()})
        del bases, scalars
    for log_n in (16, 20, 22, 24):
        n = 1 << log_n
        x = net.generate_fr(3, n)
        y = torch.empty_like(x)
        med, best = timed(lambda: net.ntt_dev(x, y))
        out["ntt_2^%d" % log_n] = dict(ms_median=med, ms_min=best, melem_s=n / med / 1e3, gbs=64.0 * n / med / 1e6)
        print("NTT 2^%d: %.3f ms (min %.3f)  %.1f Melem/s  %.1f GB/s algorithmic" % (log_n, med, best, n / med / 1e3,
                                                                                      64.0 * n / med / 1e6))
        del x, y
    n = 1 << 16
    b2 = net.generate_g2(5, n)
    s2 = net.generate_fr(6, n)
    xy2 = torch.empty(32, dtype=torch.int64, device="cuda")
    med, best = timed(lambda: net.msm_dev(b2, s2, xy2, g2=True), iters=3, warm=1)
    print("MSM G2 2^16: %.3f ms  %.2f Mpairs/s" % (med, n / med / 1e3))
    out["msm_g2_2^16"] = dict(ms_median=med)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/quickbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
