"""Fixed-base table MSM (b200zk_msm_table_*) vs the generic MSM, device timing (development aid).
usage: python tools/tablebench.py [log_n] [c ...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402
from quickbench import timed  # noqa: E402


def main():
    net = Net(0)
    net.use_torch_stream(0)
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    cs = [int(x) for x in sys.argv[2:]] or [16, 18, 20, 21]
    n = 1 << log_n
    for g2 in (False, True):
        bases = net.generate_g2(5, n) if g2 else net.generate_g1(5, n)
        scalars = net.generate_fr(6, n)
        out = torch.empty(32 if g2 else 16, dtype=torch.int64, device="cuda")
        med, _ = timed(lambda: net.msm_dev(bases, scalars, out, g2=g2), iters=3, warm=1)
        print("%s 2^%d generic: %.3f ms" % ("G2" if g2 else "G1", log_n, med))
        for c in cs:
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            table = net.msm_table_build(bases, c, g2=g2)
            t1.record(); torch.cuda.synchronize()
            net.msm_table_dev(table, scalars, c, out, g2=g2)
            net.profile(True); net.profile_reset()
            timed(lambda: net.msm_table_dev(table, scalars, c, out, g2=g2), iters=3, warm=1)
            rep = net.profile_report(); net.profile(False)
            med, best = timed(lambda: net.msm_table_dev(table, scalars, c, out, g2=g2), iters=3, warm=1)
            print("%s 2^%d table c=%d (W=%d, %.2f GB, build %.0f ms): %.3f ms  %.1f Mpairs/s" % (
                "G2" if g2 else "G1", log_n, c, net.msm_table_windows(c), table.numel() * 8 / 2**30, t0.elapsed_time(t1), med,
                n / med / 1e3))
            print("   per-call kernel ms:", {k: round(v["ms"] / 4.0, 4) for k, v in rep.items()})
            del table


if __name__ == "__main__":
    main()
