"""Kernel timeline of one Groth16 prove (B200ZK_PROFILE_TIMELINE=1): which launches overlap, where the tail is.
Development aid; event pairs add launch gaps, so the total is a little longer than the unprofiled prove."""
import os
import sys

os.environ["B200ZK_PROFILE_TIMELINE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from distributed_groth16_b200 import Net  # noqa: E402
from distributed_groth16_b200._constants import FR_ONE_MONT  # noqa: E402
from distributed_groth16_b200.groth16 import ProvingKey, prove  # noqa: E402


def sha256():
    """the reference's sha256 circuit (config 4): every launch >= 5 us"""
    from distributed_groth16_b200.groth16 import circom, setup
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "sha256_circuit.npz"))
    n_wires, n_pub, n_cons = (int(x) for x in d["dims"])
    coo = lambda k: (d[k + "_rows"], d[k + "_cols"], d[k + "_vals"])
    net = Net(0)
    net.use_torch_stream(0)
    pk, vk, mats = setup.circuit_specific_setup(net, n_wires, n_pub + 1, n_cons, coo("a"), coo("b"), coo("c"),
                                                (0x1234567, 0x2345678, 0x3456789, 0x456789A, 0x56789AB))
    z = net.fr_convert(net.to_device(d["witness"]), to_mont=True)
    zero = np.zeros(4, dtype=np.uint64)
    for _ in range(3):
        circom.prove_from_matrices(pk, mats, z, zero, zero)
    net.profile(True)
    net.profile_reset()
    circom.prove_from_matrices(pk, mats, z, zero, zero)
    rep = net.profile_report()
    net.profile(False)
    tl = sorted(rep.pop("_timeline"), key=lambda t: t[1])
    print("profiled sha256 prove: %.2f ms, %d launches" % (max(t[2] for t in tl), len(tl)))
    for name, t0, t1 in tl:
        if t1 - t0 >= 0.005:
            print("%8.3f %8.3f  %7.3f  %s" % (t0, t1, t1 - t0, name))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sha256":
        return sha256()
    log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    m = 1 << log_m
    n_vars, n_inputs = m, 2
    net = Net(0)
    net.use_torch_stream(0)
    aq, b1 = net.generate_g1(101, n_vars), net.generate_g1(102, n_vars)
    b2 = net.generate_g2(103, n_vars)
    lq, hq = net.generate_g1(104, n_vars - n_inputs), net.generate_g1(105, m)
    vk = np.concatenate([net.generate_g1(106, 3).cpu().numpy().view(np.uint64).reshape(-1),
                         net.generate_g2(107, 2).cpu().numpy().view(np.uint64).reshape(-1)])
    z = net.generate_fr(108, n_vars)
    z[0] = torch.from_numpy(np.array(FR_ONE_MONT, dtype=np.uint64).view(np.int64)).to(z.device)
    a, b, c = (net.generate_fr(sd, m) for sd in (109, 110, 111))
    pk = ProvingKey.from_device(net, aq, b1, b2, lq, hq, n_inputs, vk)
    for _ in range(3):
        prove.create_proof_dev(pk, z, a, b, c)
    net.profile(True)
    net.profile_reset()
    prove.create_proof_dev(pk, z, a, b, c)
    rep = net.profile_report()
    net.profile(False)
    tl = sorted(rep.pop("_timeline"), key=lambda t: t[1])
    end = max(t[2] for t in tl)
    print("profiled prove: %.2f ms, %d launches" % (end, len(tl)))
    for name, t0, t1 in tl:
        if t1 - t0 >= 0.05:
            print("%8.3f %8.3f  %7.3f  %s" % (t0, t1, t1 - t0, name))


if __name__ == "__main__":
    main()
