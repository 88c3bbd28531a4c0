"""NTT-only timing / profiling target (BASELINE config 3: 2^22 forward + inverse)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402

net = Net(0)
net.use_torch_stream(0)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
x = net.generate_fr(3, 1 << log_n)
y = torch.empty_like(x)
for _ in range(3):
    net.ntt_dev(x, y)
    net.ntt_dev(y, x, inverse=True)
def timed(fn, label):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%s 2^%d: %.3f ms  %.1f GB/s algorithmic (64 B/element)" % (label, log_n, ms, 64.0 * (1 << log_n) / ms / 1e6))


net.ntt_dev(x, y, coset=True)
net.ntt_dev(y, x, inverse=True, coset=True)
timed(lambda: net.ntt_dev(x, y), "NTT")
timed(lambda: net.ntt_dev(x, y, inverse=True), "iNTT")
timed(lambda: net.ntt_dev(x, y, coset=True), "coset NTT")
timed(lambda: net.ntt_dev(x, y, inverse=True, coset=True), "coset iNTT")
