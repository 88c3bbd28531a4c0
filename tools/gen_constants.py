"""Generates distributed_groth16_b200/csrc/bn254_constants.inc from oracle/bn254.py.

Build-time tool (not imported by the product): the field constants are *derived* from the BN
parameter u in oracle/bn254.py rather than pasted; tests/test_build_and_abi.py re-derives them and
compares with the committed .inc file."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import bn254 as o  # noqa: E402


def limbs32(x, n=8):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def arr(name, vals):
    body = ", ".join("0x%08xu" % v for v in vals)
    return ("    B2_HD static constexpr uint32_t %s(int i) {\n"
            "        constexpr uint32_t t[%d] = {%s};\n"
            "        return t[i];\n    }\n" % (name, len(vals), body))


def params(name, mod, extra=""):
    s = "struct %s {\n" % name
    s += arr("mod", limbs32(mod))
    s += arr("r1", limbs32(o.MONT_R % mod))
    s += arr("r2", limbs32(o.MONT_R ** 2 % mod))
    s += arr("mod_m2", limbs32(mod - 2))
    # fp29.cuh: 1 in the 2^261 Montgomery domain (2^261 mod p) as 9 limbs of 29 bits
    one261 = (1 << 261) % mod
    s += arr("one261", [(one261 >> (29 * i)) & ((1 << 29) - 1) for i in range(9)])
    s += "    static constexpr uint32_t INV = 0x%08xu;   // -mod^{-1} mod 2^32\n" % o.mont_inv32(mod)
    s += extra
    s += "};\n"
    return s


def pairing_params():
    """Constants of the optimal-ate pairing (csrc/pairing.cuh): twist Frobenius factors xi^((p-1)/3), xi^((p-1)/2),
    p^2-Frobenius factors xi^(i (p^2-1)/6) of w^i, the Miller loop count 6u+2 and the hard exponent (p^4-p^2+1)/r."""
    p, r = o.P, o.R
    g12 = o.fq2_pow(o.XI, (p - 1) // 3)
    g13 = o.fq2_pow(o.XI, (p - 1) // 2)
    s = "struct PairingConst {\n"
    for name, v in (("tw_x", g12), ("tw_y", g13)):
        s += arr(name + "_c0", limbs32(o.fq_mont(v[0])))
        s += arr(name + "_c1", limbs32(o.fq_mont(v[1])))
    g2 = o.fq2_pow(o.XI, (p * p - 1) // 6)
    assert g2[1] == 0
    acc = 1
    for i in range(1, 6):
        acc = acc * g2[0] % p
        s += arr("frob2_%d" % i, limbs32(o.fq_mont(acc)))
    loop = o.ATE_LOOP
    assert loop.bit_length() == 65
    s += "    static constexpr uint64_t ATE_LOOP_LO = 0x%016xull;   // 6u + 2 = 2^64 + LO\n" % (loop & ((1 << 64) - 1))
    hard = (p ** 4 - p ** 2 + 1) // r
    assert (p ** 4 - p ** 2 + 1) % r == 0
    nl = (hard.bit_length() + 31) // 32
    s += arr("hard_exp", limbs32(hard, nl))
    s += "    static constexpr int HARD_EXP_BITS = %d;\n" % hard.bit_length()
    s += "};\n"
    return s


def glv_constants():
    """BN254 G1 endomorphism phi(x, y) = (beta x, y) = lambda P and the short lattice basis used to split a scalar
    k = k1 + k2 lambda with |k1|, |k2| < 2^127 (Gallant-Lambert-Vanstone).  Everything is derived, nothing pasted."""
    import math
    r, p = o.R, o.P
    lam = pow(o.FR_GENERATOR, (r - 1) // 3, r)
    assert lam != 1 and pow(lam, 3, r) == 1
    beta = None
    target = o.G1.mul(o.G1_GEN, lam)
    for b in range(2, 50):
        c = pow(b, (p - 1) // 3, p)
        if c != 1:
            for cand in (c, c * c % p):
                if ((cand * o.G1_GEN[0]) % p, o.G1_GEN[1]) == target:
                    beta = cand
            break
    assert beta is not None
    seq = [(r, 1, 0), (lam, 0, 1)]
    while seq[-1][0] != 0:
        q = seq[-2][0] // seq[-1][0]
        seq.append((seq[-2][0] - q * seq[-1][0], seq[-2][1] - q * seq[-1][1], seq[-2][2] - q * seq[-1][2]))
    sq = math.isqrt(r)
    i = max(j for j, (rem, _, _) in enumerate(seq) if rem >= sq)
    (r0, _, t0), (r1, _, t1), (r2, _, t2) = seq[i], seq[i + 1], seq[i + 2]
    a1, b1 = r1, -t1
    a2, b2 = (r0, -t0) if r0 * r0 + t0 * t0 <= r2 * r2 + t2 * t2 else (r2, -t2)
    assert (a1 + b1 * lam) % r == 0 and (a2 + b2 * lam) % r == 0
    # orientation used by the device code: a1 > 0, b1 < 0, a2 > 0, b2 > 0
    assert a1 > 0 and b1 < 0 and a2 > 0 and b2 > 0, (a1, b1, a2, b2)
    g1 = ((1 << 256) * b2) // r
    g2 = ((1 << 256) * (-b1)) // r
    # G2: lambda Q = (beta^2 x, y) for Q on the twist
    Q = o.G2.mul(o.G2_GEN, 0xB200)
    b2sq = beta * beta % o.P
    assert o.G2.mul(Q, lam) == ((Q[0][0] * b2sq % o.P, Q[0][1] * b2sq % o.P), Q[1])
    return dict(lam=lam, beta=beta, a1=a1, nb1=-b1, a2=a2, b2=b2, g1=g1, g2=g2)


def glv_params():
    g = glv_constants()
    s = "struct GlvParams {\n"
    s += arr("beta", limbs32(o.fq_mont(g["beta"])))          # Fq, Montgomery
    # on the twist E'(Fq2) the same lambda acts as (x, y) -> (beta^2 x, y) (checked in glv_constants)
    s += arr("beta_g2", limbs32(o.fq_mont(g["beta"] * g["beta"] % o.P)))
    s += arr("lambda_mont", limbs32(o.fr_mont(g["lam"])))    # Fr, Montgomery (tests)
    s += arr("a1", limbs32(g["a1"], 2))
    s += arr("nb1", limbs32(g["nb1"], 4))                    # -b1 > 0
    s += arr("a2", limbs32(g["a2"], 4))
    s += arr("b2", limbs32(g["b2"], 2))
    s += arr("g1", limbs32(g["g1"], 3))                      # floor(2^256 b2 / r)
    s += arr("g2", limbs32(g["g2"], 5))                      # floor(2^256 (-b1) / r)
    s += "};\n"
    return s


def render():
    out = "// GENERATED by tools/gen_constants.py from oracle/bn254.py -- do not edit.\n"
    out += params("FqParams", o.P)
    # Fr extras: 2-adic root of unity (2^28-th), its inverse, generator 5 and inverse -- Montgomery form
    w = o.fr_root_of_unity(1 << o.FR_TWO_ADICITY)
    extra = arr("root28", limbs32(o.fr_mont(w)))
    extra += arr("root28_inv", limbs32(o.fr_mont(pow(w, -1, o.R))))
    extra += arr("gen", limbs32(o.fr_mont(o.FR_GENERATOR)))
    extra += arr("gen_inv", limbs32(o.fr_mont(pow(o.FR_GENERATOR, -1, o.R))))
    out += params("FrParams", o.R, extra)
    # curve constants (Montgomery): b for G1, b' = 3/(9+u) for G2, generators
    out += glv_params()
    out += "struct CurveConst {\n"
    out += arr("g1_b", limbs32(o.fq_mont(3)))
    out += arr("g2_b_c0", limbs32(o.fq_mont(o.B_G2[0])))
    out += arr("g2_b_c1", limbs32(o.fq_mont(o.B_G2[1])))
    out += arr("g1_gen_x", limbs32(o.fq_mont(o.G1_GEN[0])))
    out += arr("g1_gen_y", limbs32(o.fq_mont(o.G1_GEN[1])))
    out += arr("g2_gen_x0", limbs32(o.fq_mont(o.G2_GEN[0][0])))
    out += arr("g2_gen_x1", limbs32(o.fq_mont(o.G2_GEN[0][1])))
    out += arr("g2_gen_y0", limbs32(o.fq_mont(o.G2_GEN[1][0])))
    out += arr("g2_gen_y1", limbs32(o.fq_mont(o.G2_GEN[1][1])))
    out += "};\n"
    out += pairing_params()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "distributed_groth16_b200", "csrc",
                        "bn254_constants.inc")
    with open(path, "w") as f:
        f.write(render())
    print("wrote", os.path.normpath(path))
    v = o.MONT_R % o.R
    limbs = [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    path2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "distributed_groth16_b200", "_constants.py")
    with open(path2, "w") as f:
        f.write('"""GENERATED by tools/gen_constants.py -- Montgomery one of BN254 Fr (R mod r) as 4 u64 limbs."""\n'
                'FR_ONE_MONT = %r\n' % (limbs,))
    print("wrote", os.path.normpath(path2))
