"""csrc/pairing.cuh compiled for the host vs the oracle's optimal-ate pairing (oracle/bn254.py, itself pinned on the
snarkjs fixtures of the reference: tests/test_oracle_golden.py)."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tower_limbs(f):
    """oracle Fq12 (12 coefficients in w, w^6 = 9 + u) -> tower order c0.a c0.b c0.c c1.a c1.b c1.c, each (c0, c1), Montgomery."""
    from oracle import bn254 as o
    out = []
    for i in (0, 2, 4, 1, 3, 5):
        y = f[i + 6]
        x = (f[i] + 9 * y) % o.P
        for v in (x, y):
            m = o.fq_mont(v)
            out += [(m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
    return out


def test_pairing_header_matches_the_oracle(tmp_path):
    from oracle import bn254 as o, layout
    a, b = 0x1234567, 0xABCDEF0123
    P, Q = o.G1_GEN, o.G2_GEN
    aP, bQ = o.G1.mul(P, a), o.G2.mul(Q, b)
    cases = [(P, Q), (aP, bQ), (o.G1.neg(o.G1.mul(P, a * b % o.R)), Q), (None, Q)]
    # e(P,Q) * e(aP,bQ) * e(-abP,Q) * 1 != 1 in general: add e(-P, Q) so that the product is 1
    cases.append((o.G1.neg(P), Q))
    blob = struct.pack("<Q", len(cases))
    for Pt, Qt in cases:
        e = o.pairing(Pt, Qt)
        if Pt is not None:
            assert o.fq12_pow(e, o.R) == o.FQ12_ONE and e != o.FQ12_ONE
        p_arr = layout.g1_to_arr([Pt]).reshape(-1)
        q_arr = layout.g2_to_arr([Qt]).reshape(-1)
        blob += p_arr.astype("<u8").tobytes() + q_arr.astype("<u8").tobytes()
        blob += np.array(tower_limbs(e), dtype="<u8").tobytes()
    blob += struct.pack("<Q", 1)
    vec = tmp_path / "pairing_vectors.bin"
    vec.write_bytes(blob)
    exe = tmp_path / "pairing_host_test"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "pairing_host_test.cpp")])
    out = subprocess.run([str(exe), str(vec)], capture_output=True, text=True)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
