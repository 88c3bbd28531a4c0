"""G2 MSM device timing with the per-kernel breakdown (development aid)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402
from quickbench import timed  # noqa: E402


def main():
    net = Net(0)
    net.use_torch_stream(0)
    for log_n in [int(x) for x in (sys.argv[1:] or ["20"])]:
        n = 1 << log_n
        b2 = net.generate_g2(5, n)
        s2 = net.generate_fr(6, n)
        xy2 = torch.empty(32, dtype=torch.int64, device="cuda")
        net.msm_dev(b2, s2, xy2, g2=True)
        net.profile(True)
        net.profile_reset()
        timed(lambda: net.msm_dev(b2, s2, xy2, g2=True), iters=3, warm=1)
        rep = net.profile_report()
        net.profile(False)
        med, best = timed(lambda: net.msm_dev(b2, s2, xy2, g2=True), iters=3, warm=1)
        print("MSM G2 2^%d: %.3f ms  %.2f Mpairs/s" % (log_n, med, n / med / 1e3))
        print("   per-call kernel ms:", {k: round(v["ms"] / 4.0, 4) for k, v in rep.items()})


if __name__ == "__main__":
    main()
