"""BN254 big-integer oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Pure-Python restatement of the arithmetic the reference gets from arkworks 0.4
(`ark-ff`, `ark-ec`, `ark-poly`, `ark-bn254`, `ark-serialize`, the zkHubHQ `ark-groth16`
fork) -- crates that are NOT vendored under /root/reference and are unpinned there
(`Cargo.lock` is git-ignored, /root/reference/.gitignore:2).  What is restated, and the
reference call site each piece serves:

  * Fq / Fr / Fq2 arithmetic, Montgomery (R = 2^256) encode/decode       -- every site below
  * G1 / G2 short-Weierstrass arithmetic, naive and Pippenger MSM         -- `G::msm`, dist-primitives/src/dmsm/mod.rs:82
  * radix-2 NTT / iNTT / coset NTT with arkworks `Radix2EvaluationDomain`
    conventions (natural order in/out, generator 5)                        -- dist-primitives/src/dfft/mod.rs:17-95,
                                                                              secret-sharing/src/pss.rs:41-48
  * `fft_in_place_rearrange` (bit-reversal permutation)                    -- dist-primitives/src/dfft/mod.rs:258-271
  * `CircomReduction::witness_map_from_matrices` (h polynomial)            -- ark-circom/src/circom/qap.rs:27-92
  * `qap()` (QAP evaluation vectors)                                       -- groth16/src/qap.rs:44-91
  * Groth16 prove with explicit r, s                                       -- groth16/src/prove.rs:21-136,
                                                                              groth16/examples/sha256.rs:152-169,208-212
  * ark-serialize compressed point / Proof encoding                        -- zk-cli/src/main.rs:130-136
  * optimal-ate pairing + Groth16 verify                                   -- groth16/examples/sha256.rs:229-254
  * snarkjs .zkey / .r1cs / .wtns readers                                  -- ark-circom/src/zkey.rs:53-387,
                                                                              ark-circom/src/circom/r1cs_reader.rs:54-249

Pinning (see tests/test_oracle_golden.py): Montgomery byte goldens of Fq one / G1 gen /
G2 gen (ark-circom/src/zkey.rs:417-455); the 128-byte golden proof
zk-cli/test-circuits/sha256/proof.bin decodes to exactly the coordinates printed in
zk-cli/README.md:82 (x, y-sign flags, Fq2 ordering); snarkjs proof/vk fixtures verify
under the pairing here (fixtures/million/*.json) and the pairing value itself matches snarkjs'
`vk_alphabeta_12` (= e(alpha, beta)^(2u(6u^2+3u+1)), ffjavascript's hard-part multiple); a proof produced by `groth16_prove`
from the snarkjs-made complex-circuit-10000-10000.zkey verifies against that zkey's vk.
MSM and NTT outputs have no absolute vector of their own in the reference (its tests are
differential against arkworks); they are pinned end to end: oracle/ark_rand.py regenerates the
reference's seeded setup and the prover built on this module's arithmetic (C++ twin) returns the
reference's committed proof.bin bit for bit (tests/test_oracle_reference_proof.py).
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Constants.  p, r are the BN254 base / scalar primes; both are re-derived from the BN
# parameter u in tests and checked against the headers of the reference's fixtures.
# --------------------------------------------------------------------------------------
BN_U = 4965661367192848881
P = 36 * BN_U**4 + 36 * BN_U**3 + 24 * BN_U**2 + 6 * BN_U + 1          # Fq modulus
R = 36 * BN_U**4 + 36 * BN_U**3 + 18 * BN_U**2 + 6 * BN_U + 1          # Fr modulus (group order)
assert P == 21888242871839275222246405745257275088696311157297823662689037894645226208583
assert R == 21888242871839275222246405745257275088548364400416034343698204186575808495617

MONT_BITS = 256
MONT_R = 1 << MONT_BITS
FR_GENERATOR = 5          # ark-bn254 Fr::GENERATOR
FR_TWO_ADICITY = 28
B_G1 = 3


def fr_mont(x: int) -> int:
    return (x * MONT_R) % R


def fr_unmont(x: int) -> int:
    return (x * pow(MONT_R, -1, R)) % R


def fq_mont(x: int) -> int:
    return (x * MONT_R) % P


def fq_unmont(x: int) -> int:
    return (x * pow(MONT_R, -1, P)) % P


def mont_inv64(mod: int) -> int:
    """-mod^{-1} mod 2^64 (the CIOS constant)."""
    return (-pow(mod, -1, 1 << 64)) % (1 << 64)


def mont_inv32(mod: int) -> int:
    return (-pow(mod, -1, 1 << 32)) % (1 << 32)


def to_limbs64(x: int, n: int = 4) -> List[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs64(l: Sequence[int]) -> int:
    v = 0
    for i, w in enumerate(l):
        v |= int(w) << (64 * i)
    return v


# --------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1)
# --------------------------------------------------------------------------------------
Fq2 = Tuple[int, int]
FQ2_ZERO: Fq2 = (0, 0)
FQ2_ONE: Fq2 = (1, 0)


def fq2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fq2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fq2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % P, (-a[1]) % P)


def fq2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fq2_sqr(a: Fq2) -> Fq2:
    return fq2_mul(a, a)


def fq2_scalar(a: Fq2, k: int) -> Fq2:
    return ((a[0] * k) % P, (a[1] * k) % P)


def fq2_inv(a: Fq2) -> Fq2:
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return ((a[0] * n) % P, (-a[1] * n) % P)


def fq2_conj(a: Fq2) -> Fq2:
    return (a[0], (-a[1]) % P)


def fq2_pow(a: Fq2, e: int) -> Fq2:
    res = FQ2_ONE
    base = a
    while e:
        if e & 1:
            res = fq2_mul(res, base)
        base = fq2_sqr(base)
        e >>= 1
    return res


def fq2_sqrt(a: Fq2) -> Optional[Fq2]:
    """Square root in Fq2 (p = 3 mod 4), complex method."""
    if a == FQ2_ZERO:
        return FQ2_ZERO
    a0, a1 = a
    if a1 == 0:
        s = pow(a0, (P + 1) // 4, P)
        if s * s % P == a0:
            return (s, 0)
        s = pow((-a0) % P, (P + 1) // 4, P)
        assert s * s % P == (-a0) % P
        return (0, s)
    norm = (a0 * a0 + a1 * a1) % P
    alpha = pow(norm, (P + 1) // 4, P)
    if alpha * alpha % P != norm:
        return None
    inv2 = pow(2, -1, P)
    delta = (a0 + alpha) * inv2 % P
    x0 = pow(delta, (P + 1) // 4, P)
    if x0 * x0 % P != delta:
        delta = (a0 - alpha) * inv2 % P
        x0 = pow(delta, (P + 1) // 4, P)
        if x0 * x0 % P != delta:
            return None
    x1 = a1 * pow(2 * x0, -1, P) % P
    res = (x0, x1)
    return res if fq2_sqr(res) == a else None


XI: Fq2 = (9, 1)                              # non-residue for the sextic twist
B_G2: Fq2 = fq2_scalar(fq2_inv(XI), 3)        # twist b' = 3/(9+u)

# --------------------------------------------------------------------------------------
# Curve arithmetic.  Points are affine tuples or None (= infinity).  Generic over a field
# described by a small ops table so G1 and G2 share the code.
# --------------------------------------------------------------------------------------


class _F1:
    zero = 0
    one = 1

    @staticmethod
    def add(a, b):
        return (a + b) % P

    @staticmethod
    def sub(a, b):
        return (a - b) % P

    @staticmethod
    def mul(a, b):
        return (a * b) % P

    @staticmethod
    def neg(a):
        return (-a) % P

    @staticmethod
    def inv(a):
        return pow(a, -1, P)


class _F2:
    zero = FQ2_ZERO
    one = FQ2_ONE
    add = staticmethod(fq2_add)
    sub = staticmethod(fq2_sub)
    mul = staticmethod(fq2_mul)
    neg = staticmethod(fq2_neg)
    inv = staticmethod(fq2_inv)


class Curve:
    """y^2 = x^3 + b over field F; Jacobian internals, affine in/out."""

    def __init__(self, F, b):
        self.F = F
        self.b = b

    def is_on_curve(self, pt) -> bool:
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), self.b)

    def neg(self, pt):
        if pt is None:
            return None
        return (pt[0], self.F.neg(pt[1]))

    # Jacobian (X, Y, Z); infinity = Z == zero
    def to_jac(self, pt):
        if pt is None:
            return (self.F.one, self.F.one, self.F.zero)
        return (pt[0], pt[1], self.F.one)

    def from_jac(self, J):
        F = self.F
        X, Y, Z = J
        if Z == F.zero:
            return None
        zi = F.inv(Z)
        zi2 = F.mul(zi, zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_double(self, J):
        F = self.F
        X, Y, Z = J
        if Z == F.zero:
            return J
        A = F.mul(X, X)
        B = F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.mul(E, E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jac_add(self, J1, J2):
        F = self.F
        if J1[2] == F.zero:
            return J2
        if J2[2] == F.zero:
            return J1
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        Z1Z1 = F.mul(Z1, Z1)
        Z2Z2 = F.mul(Z2, Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jac_double(J1)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def add(self, p1, p2):
        return self.from_jac(self.jac_add(self.to_jac(p1), self.to_jac(p2)))

    def jac_mul(self, J, k: int):
        F = self.F
        acc = (F.one, F.one, F.zero)
        if k == 0:
            return acc
        for bit in bin(k)[2:]:
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add(acc, J)
        return acc

    def mul(self, pt, k: int):
        """k is reduced modulo the group order R (scalars are Fr elements)."""
        k %= R
        return self.from_jac(self.jac_mul(self.to_jac(pt), k))

    def msm_naive(self, bases, scalars):
        """sum_i scalars[i]*bases[i]; the mathematical definition d_msm must match
        (dist-primitives/examples/dmsm_test.rs:49-64)."""
        acc = self.to_jac(None)
        for b, s in zip(bases, scalars):
            if b is None or s % R == 0:
                continue
            acc = self.jac_add(acc, self.jac_mul(self.to_jac(b), s % R))
        return self.from_jac(acc)

    def msm(self, bases, scalars, c: int = 8):
        """Unsigned-window Pippenger (python speed-up of msm_naive; same result)."""
        n = min(len(bases), len(scalars))
        if n < 16:
            return self.msm_naive(bases[:n], scalars[:n])
        F = self.F
        inf = (F.one, F.one, F.zero)
        jb = [self.to_jac(b) for b in bases[:n]]
        sc = [s % R for s in scalars[:n]]
        nwin = (254 + c - 1) // c
        total = inf
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jac_double(total)
            buckets = [inf] * (1 << c)
            sh = w * c
            mask = (1 << c) - 1
            for j, s in zip(jb, sc):
                d = (s >> sh) & mask
                if d and j[2] != F.zero:
                    buckets[d] = self.jac_add(buckets[d], j)
            run = inf
            acc = inf
            for d in range((1 << c) - 1, 0, -1):
                run = self.jac_add(run, buckets[d])
                acc = self.jac_add(acc, run)
            total = self.jac_add(total, acc)
        return self.from_jac(total)


G1 = Curve(_F1, B_G1)
G2 = Curve(_F2, B_G2)
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)   # decimal coordinates as printed in ark-circom/src/zkey.rs:466-486 (test data)

# --------------------------------------------------------------------------------------
# NTT over Fr with arkworks Radix2EvaluationDomain conventions.
# --------------------------------------------------------------------------------------


def fr_root_of_unity(n: int) -> int:
    """group_gen of Radix2EvaluationDomain::new(n): GENERATOR^((r-1)/n), n a power of two."""
    assert n & (n - 1) == 0 and n <= (1 << FR_TWO_ADICITY)
    return pow(FR_GENERATOR, (R - 1) // n, R)


def bit_reverse_permute(v: List[int]) -> List[int]:
    """Result of `fft_in_place_rearrange` (dist-primitives/src/dfft/mod.rs:258-271)."""
    n = len(v)
    lg = n.bit_length() - 1
    out = [0] * n
    for i in range(n):
        out[int(bin(i)[2:].zfill(lg)[::-1], 2) if lg else 0] = v[i]
    return out


def _ntt_core(a: List[int], omega: int) -> List[int]:
    n = len(a)
    a = bit_reverse_permute(a)
    length = 2
    while length <= n:
        wlen = pow(omega, n // length, R)
        half = length // 2
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * wlen % R
        for start in range(0, n, length):
            for j in range(half):
                u = a[start + j]
                v = a[start + j + half] * tw[j] % R
                a[start + j] = (u + v) % R
                a[start + j + half] = (u - v) % R
        length <<= 1
    return a


def ntt(a: Sequence[int], coset: bool = False) -> List[int]:
    """dom.fft(a): out[i] = sum_j a[j] w^{ij} (natural order); coset: a[j] *= g^j first."""
    n = len(a)
    a = [x % R for x in a]
    if coset:
        g = 1
        for j in range(n):
            a[j] = a[j] * g % R
            g = g * FR_GENERATOR % R
    if n == 1:
        return a
    return _ntt_core(a, fr_root_of_unity(n))


def intt(a: Sequence[int], coset: bool = False) -> List[int]:
    """dom.ifft(a); coset: result[j] *= g^{-j} afterwards."""
    n = len(a)
    a = [x % R for x in a]
    if n > 1:
        a = _ntt_core(a, pow(fr_root_of_unity(n), -1, R))
    ninv = pow(n, -1, R)
    a = [x * ninv % R for x in a]
    if coset:
        gi = pow(FR_GENERATOR, -1, R)
        g = 1
        for j in range(n):
            a[j] = a[j] * g % R
            g = g * gi % R
    return a


def dft_naive(a: Sequence[int], inverse: bool = False) -> List[int]:
    n = len(a)
    w = fr_root_of_unity(n)
    if inverse:
        w = pow(w, -1, R)
    out = []
    for i in range(n):
        wi = pow(w, i, R)
        acc, x = 0, 1
        for j in range(n):
            acc = (acc + a[j] * x) % R
            x = x * wi % R
        out.append(acc)
    if inverse:
        ninv = pow(n, -1, R)
        out = [v * ninv % R for v in out]
    return out


# --------------------------------------------------------------------------------------
# QAP + CircomReduction h + Groth16 prove
# --------------------------------------------------------------------------------------


def next_pow2(n: int) -> int:
    m = 1
    while m < n:
        m <<= 1
    return m


def qap(matrix_a, matrix_b, num_inputs: int, num_constraints: int, z: Sequence[int]):
    """groth16/src/qap.rs:44-91.  matrices: list (per constraint) of [(coeff, wire)]."""
    m = next_pow2(num_constraints + num_inputs)
    a = [0] * m
    b = [0] * m
    for i in range(num_constraints):
        a[i] = sum(c * z[w] for c, w in matrix_a[i]) % R
        b[i] = sum(c * z[w] for c, w in matrix_b[i]) % R
    for j in range(num_inputs):
        a[num_constraints + j] = z[j] % R
    c = [0] * m
    for i in range(num_constraints):
        c[i] = a[i] * b[i] % R
    return a, b, c


def h_circom(a: Sequence[int], b: Sequence[int], c: Sequence[int]) -> List[int]:
    """CircomReduction::witness_map_from_matrices after the mat-vec
    (ark-circom/src/circom/qap.rs:64-89) == what ext_wit::h reconstructs
    (groth16/src/ext_wit.rs:16-101): h_i = A(w2m^{2i+1}) B(..) - C(..)."""
    m = len(a)
    w2m = fr_root_of_unity(2 * m)

    def shift(v):
        co = intt(v)
        g = 1
        for j in range(m):
            co[j] = co[j] * g % R
            g = g * w2m % R
        return ntt(co)

    ea, eb, ec = shift(a), shift(b), shift(c)
    return [(x * y - w) % R for x, y, w in zip(ea, eb, ec)]


class ProvingKey:
    """Field order follows ark_groth16::ProvingKey as filled by ark-circom/src/zkey.rs:103-134."""

    def __init__(self):
        self.alpha_g1 = None
        self.beta_g1 = None
        self.beta_g2 = None
        self.gamma_g2 = None
        self.delta_g1 = None
        self.delta_g2 = None
        self.ic: list = []            # gamma_abc_g1
        self.a_query: list = []
        self.b_g1_query: list = []
        self.b_g2_query: list = []
        self.h_query: list = []
        self.l_query: list = []
        self.n_vars = 0
        self.n_public = 0
        self.domain_size = 0


def groth16_prove(pk: ProvingKey, z: Sequence[int], h: Sequence[int], r: int = 0, s: int = 0,
                  mirror_reference_bg1: bool = False):
    """Proof elements as computed by the reference:

      A = alpha + a_query[0] + r*delta_g1 + MSM(a_query[1..], z[1..])        prove.rs:21-46 + sha256.rs:208-209
      B = beta2 + b_g2_query[0] + s*delta_g2 + MSM_G2(b_g2_query[1..], z[1..])  prove.rs:62-85 + sha256.rs:210-212
      C = MSM(l_query, aux) + MSM(h_query, h) + s*A + r*B1 - r*s*delta_g1      (single-node arkworks formula)

    z[0] must be 1 (the constant wire).  Returns affine (A, B, C)."""
    n_inputs = pk.n_public + 1
    zz = [v % R for v in z]
    assert zz[0] == 1
    msm_a = G1.msm(pk.a_query[1:], zz[1:])
    A = G1.add(G1.add(pk.alpha_g1, pk.a_query[0]), msm_a)
    A = G1.add(A, G1.mul(pk.delta_g1, r))
    msm_b2 = G2.msm(pk.b_g2_query[1:], zz[1:])
    B = G2.add(G2.add(pk.beta_g2, pk.b_g2_query[0]), msm_b2)
    B = G2.add(B, G2.mul(pk.delta_g2, s))
    aux = zz[n_inputs:]
    l_acc = G1.msm(pk.l_query, aux)
    h_acc = G1.msm(pk.h_query[: len(h)], h)
    C = G1.add(l_acc, h_acc)
    C = G1.add(C, G1.mul(A, s))
    if r % R != 0 or mirror_reference_bg1:
        B1 = G1.add(G1.add(pk.beta_g1, pk.b_g1_query[0]), G1.msm(pk.b_g1_query[1:], zz[1:]))
        B1 = G1.add(B1, G1.mul(pk.delta_g1, s))
        C = G1.add(C, G1.mul(B1, r))
        C = G1.add(C, G1.neg(G1.mul(pk.delta_g1, (r * s) % R)))
    return A, B, C


# --------------------------------------------------------------------------------------
# ark-serialize compressed encoding (Compress::Yes)
# --------------------------------------------------------------------------------------
_FLAG_NEG = 0x80
_FLAG_INF = 0x40


def _fq_is_neg(y: int) -> bool:
    """arkworks: flag set when y > -y, i.e. y > (p-1)/2."""
    return y > (P - 1) // 2


def _fq2_is_neg(y: Fq2) -> bool:
    """Fq2 lexicographic order compares c1 first, then c0."""
    ny = fq2_neg(y)
    return (y[1], y[0]) > (ny[1], ny[0])


def g1_compress(pt) -> bytes:
    if pt is None:
        b = bytearray(32)
        b[31] |= _FLAG_INF
        return bytes(b)
    b = bytearray(pt[0].to_bytes(32, "little"))
    if _fq_is_neg(pt[1]):
        b[31] |= _FLAG_NEG
    return bytes(b)


def g2_compress(pt) -> bytes:
    if pt is None:
        b = bytearray(64)
        b[63] |= _FLAG_INF
        return bytes(b)
    b = bytearray(pt[0][0].to_bytes(32, "little") + pt[0][1].to_bytes(32, "little"))
    if _fq2_is_neg(pt[1]):
        b[63] |= _FLAG_NEG
    return bytes(b)


def g1_decompress(buf: bytes):
    assert len(buf) == 32
    flags = buf[31] & 0xC0
    if flags & _FLAG_INF:
        return None
    x = int.from_bytes(bytes(buf[:31]) + bytes([buf[31] & 0x3F]), "little")
    y2 = (x * x * x + B_G1) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise ValueError("x not on curve")
    if _fq_is_neg(y) != bool(flags & _FLAG_NEG):
        y = (-y) % P
    return (x, y)


def g2_decompress(buf: bytes):
    assert len(buf) == 64
    flags = buf[63] & 0xC0
    if flags & _FLAG_INF:
        return None
    c0 = int.from_bytes(buf[:32], "little")
    c1 = int.from_bytes(bytes(buf[32:63]) + bytes([buf[63] & 0x3F]), "little")
    x = (c0, c1)
    y2 = fq2_add(fq2_mul(fq2_sqr(x), x), B_G2)
    y = fq2_sqrt(y2)
    if y is None:
        raise ValueError("x not on twist")
    if _fq2_is_neg(y) != bool(flags & _FLAG_NEG):
        y = fq2_neg(y)
    return (x, y)


def proof_compress(A, B, C) -> bytes:
    """Proof<Bn254>::serialize_with_mode(Compress::Yes): A (32) || B (64) || C (32)."""
    return g1_compress(A) + g2_compress(B) + g1_compress(C)


def proof_decompress(buf: bytes):
    assert len(buf) == 128
    return g1_decompress(buf[:32]), g2_decompress(buf[32:96]), g1_decompress(buf[96:])


# --------------------------------------------------------------------------------------
# Pairing (optimal ate) -- only used to *verify* proofs in tests.
# Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)  (w^6 = 9 + u)
# --------------------------------------------------------------------------------------
Fq12 = Tuple[int, ...]
FQ12_ONE: Fq12 = (1,) + (0,) * 11


def fq12_mul(a: Fq12, b: Fq12) -> Fq12:
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):          # w^12 = 18 w^6 - 82
        v = t[k]
        if v:
            t[k - 6] += 18 * v
            t[k - 12] -= 82 * v
    return tuple(x % P for x in t[:12])


def fq12_pow(a: Fq12, e: int) -> Fq12:
    res = FQ12_ONE
    base = a
    while e:
        if e & 1:
            res = fq12_mul(res, base)
        base = fq12_mul(base, base)
        e >>= 1
    return res


def _embed_fq2(c: Fq2, k: int) -> List[int]:
    """(a + b u) * w^k with u = w^6 - 9  ->  polynomial coefficients."""
    out = [0] * 12
    a, b = c
    assert k < 6
    out[k] = (a - 9 * b) % P
    out[k + 6] = b % P
    return out


def _line(lmbda: Fq2, T, Pt) -> Fq12:
    """Line through untwisted T (slope lmbda on the twist) evaluated at P in G1:
    yP - lmbda*xP*w + (lmbda*xT - yT)*w^3."""
    xP, yP = Pt
    xT, yT = T
    c1 = _embed_fq2(fq2_scalar(fq2_neg(lmbda), xP), 1)
    c3 = _embed_fq2(fq2_sub(fq2_mul(lmbda, xT), yT), 3)
    out = [0] * 12
    out[0] = yP % P
    for i in range(12):
        out[i] = (out[i] + c1[i] + c3[i]) % P
    return tuple(out)


ATE_LOOP = 6 * BN_U + 2


def _frob_twist(Q):
    """pi(Q) expressed on the twist."""
    x, y = Q
    g2 = fq2_pow(XI, (P - 1) // 3)
    g3 = fq2_pow(XI, (P - 1) // 2)
    return (fq2_mul(fq2_conj(x), g2), fq2_mul(fq2_conj(y), g3))


def miller_loop(Pt, Q) -> Fq12:
    if Pt is None or Q is None:
        return FQ12_ONE
    T = Q
    f = FQ12_ONE
    bits = bin(ATE_LOOP)[3:]
    for bit in bits:
        # doubling step
        lam = fq2_mul(fq2_scalar(fq2_sqr(T[0]), 3), fq2_inv(fq2_scalar(T[1], 2)))
        l = _line(lam, T, Pt)
        x3 = fq2_sub(fq2_sqr(lam), fq2_scalar(T[0], 2))
        y3 = fq2_sub(fq2_mul(lam, fq2_sub(T[0], x3)), T[1])
        T = (x3, y3)
        f = fq12_mul(fq12_mul(f, f), l)
        if bit == "1":
            lam = fq2_mul(fq2_sub(Q[1], T[1]), fq2_inv(fq2_sub(Q[0], T[0])))
            l = _line(lam, T, Pt)
            x3 = fq2_sub(fq2_sub(fq2_sqr(lam), T[0]), Q[0])
            y3 = fq2_sub(fq2_mul(lam, fq2_sub(T[0], x3)), T[1])
            T = (x3, y3)
            f = fq12_mul(f, l)
    Q1 = _frob_twist(Q)
    Q2 = _frob_twist(Q1)
    nQ2 = (Q2[0], fq2_neg(Q2[1]))
    for Qi in (Q1, nQ2):
        lam = fq2_mul(fq2_sub(Qi[1], T[1]), fq2_inv(fq2_sub(Qi[0], T[0])))
        l = _line(lam, T, Pt)
        x3 = fq2_sub(fq2_sub(fq2_sqr(lam), T[0]), Qi[0])
        y3 = fq2_sub(fq2_mul(lam, fq2_sub(T[0], x3)), T[1])
        T = (x3, y3)
        f = fq12_mul(f, l)
    return f


def final_exponentiation(f: Fq12) -> Fq12:
    return fq12_pow(f, (P**12 - 1) // R)


def pairing(Pt, Q) -> Fq12:
    return final_exponentiation(miller_loop(Pt, Q))


def pairing_product_is_one(pairs) -> bool:
    f = FQ12_ONE
    for Pt, Q in pairs:
        f = fq12_mul(f, miller_loop(Pt, Q))
    return final_exponentiation(f) == FQ12_ONE


def groth16_verify(vk_alpha_g1, vk_beta_g2, vk_gamma_g2, vk_delta_g2, ic, public_inputs, A, B, C) -> bool:
    """e(A,B) == e(alpha,beta) e(sum ic_i x_i, gamma) e(C, delta)   (sha256.rs:229-254)."""
    acc = ic[0]
    for x, pt in zip(public_inputs, ic[1:]):
        acc = G1.add(acc, G1.mul(pt, x))
    return pairing_product_is_one([
        (A, B),
        (G1.neg(vk_alpha_g1), vk_beta_g2),
        (G1.neg(acc), vk_gamma_g2),
        (G1.neg(C), vk_delta_g2),
    ])


# --------------------------------------------------------------------------------------
# snarkjs binary formats
# --------------------------------------------------------------------------------------


def _sections(buf: bytes, magic: bytes):
    assert buf[:4] == magic, buf[:4]
    _version, nsec = struct.unpack_from("<II", buf, 4)
    off = 12
    secs = {}
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", buf, off)
        off += 12
        secs.setdefault(sid, []).append((off, ln))
        off += ln
    return secs


def _rd_fq_mont(buf, off) -> int:
    """zkey points are stored in Montgomery form (zkey.rs:340-345)."""
    return fq_unmont(int.from_bytes(buf[off:off + 32], "little"))


def _rd_g1(buf, off):
    x = int.from_bytes(buf[off:off + 32], "little")
    y = int.from_bytes(buf[off + 32:off + 64], "little")
    if x == 0 and y == 0:
        return None                       # zkey.rs:353-362
    return (fq_unmont(x), fq_unmont(y))


def _rd_g2(buf, off):
    v = [int.from_bytes(buf[off + 32 * i:off + 32 * i + 32], "little") for i in range(4)]
    if all(t == 0 for t in v):
        return None
    v = [fq_unmont(t) for t in v]
    return ((v[0], v[1]), (v[2], v[3]))


def read_zkey(buf: bytes):
    """Returns (ProvingKey, matrix_a, matrix_b, num_constraints).  Follows zkey.rs:53-218:
    sections 2 (header) 3 (IC) 4 (coeffs) 5 (A) 6 (B1) 7 (B2) 8 (L/C) 9 (H)."""
    secs = _sections(buf, b"zkey")
    off, _ = secs[2][0]
    n8q = struct.unpack_from("<I", buf, off)[0]
    off += 4
    q = int.from_bytes(buf[off:off + n8q], "little")
    off += n8q
    n8r = struct.unpack_from("<I", buf, off)[0]
    off += 4
    rr = int.from_bytes(buf[off:off + n8r], "little")
    off += n8r
    assert q == P and rr == R, "zkey is not BN254"
    n_vars, n_public, domain_size = struct.unpack_from("<III", buf, off)
    off += 12
    pk = ProvingKey()
    pk.n_vars, pk.n_public, pk.domain_size = n_vars, n_public, domain_size
    pk.alpha_g1 = _rd_g1(buf, off); off += 64
    pk.beta_g1 = _rd_g1(buf, off); off += 64
    pk.beta_g2 = _rd_g2(buf, off); off += 128
    pk.gamma_g2 = _rd_g2(buf, off); off += 128
    pk.delta_g1 = _rd_g1(buf, off); off += 64
    pk.delta_g2 = _rd_g2(buf, off); off += 128

    def g1_sec(sid, n):
        o, _ = secs[sid][0]
        return [_rd_g1(buf, o + 64 * i) for i in range(n)]

    pk.ic = g1_sec(3, n_public + 1)
    pk.a_query = g1_sec(5, n_vars)
    pk.b_g1_query = g1_sec(6, n_vars)
    o7, _ = secs[7][0]
    pk.b_g2_query = [_rd_g2(buf, o7 + 128 * i) for i in range(n_vars)]
    pk.l_query = g1_sec(8, n_vars - n_public - 1)
    pk.h_query = g1_sec(9, domain_size)

    # coefficients: stored * R^2, i.e. Montgomery form of the Montgomery form (zkey.rs:333-338)
    o4, _ = secs[4][0]
    ncoef = struct.unpack_from("<I", buf, o4)[0]
    o4 += 4
    mats = [[[] for _ in range(domain_size)] for _ in range(2)]
    max_c = 0
    rinv2 = pow(MONT_R, -2, R)
    for _ in range(ncoef):
        mi, ci, si = struct.unpack_from("<III", buf, o4)
        o4 += 12
        val = int.from_bytes(buf[o4:o4 + 32], "little") * rinv2 % R
        o4 += 32
        max_c = max(max_c, ci)
        mats[mi][ci].append((val, si))
    num_constraints = max_c - n_public
    return pk, mats[0][:num_constraints], mats[1][:num_constraints], num_constraints


def read_r1cs(buf: bytes):
    """circom .r1cs (r1cs_reader.rs:54-249): returns dict with n_wires, n_pub_out, n_pub_in,
    n_prv_in, n_constraints and constraints [(A, B, C)] each a list of (coeff, wire)."""
    secs = _sections(buf, b"r1cs")
    off, _ = secs[1][0]
    fs = struct.unpack_from("<I", buf, off)[0]
    off += 4
    prime = int.from_bytes(buf[off:off + fs], "little")
    off += fs
    assert prime == R, "r1cs prime is not BN254 Fr"      # r1cs_reader.rs:180-188
    n_wires, n_pub_out, n_pub_in, n_prv_in = struct.unpack_from("<IIII", buf, off)
    off += 16
    _n_labels = struct.unpack_from("<Q", buf, off)[0]
    off += 8
    n_constraints = struct.unpack_from("<I", buf, off)[0]
    off, _ = secs[2][0]
    cons = []
    for _ in range(n_constraints):
        abc = []
        for _k in range(3):
            nterm = struct.unpack_from("<I", buf, off)[0]
            off += 4
            lc = []
            for _t in range(nterm):
                w = struct.unpack_from("<I", buf, off)[0]
                off += 4
                lc.append((int.from_bytes(buf[off:off + fs], "little"), w))
                off += fs
            abc.append(lc)
        cons.append(tuple(abc))
    return dict(n_wires=n_wires, n_pub_out=n_pub_out, n_pub_in=n_pub_in, n_prv_in=n_prv_in,
                n_constraints=n_constraints, constraints=cons)


def read_wtns(buf: bytes) -> List[int]:
    secs = _sections(buf, b"wtns")
    off, _ = secs[1][0]
    n8 = struct.unpack_from("<I", buf, off)[0]
    off += 4
    prime = int.from_bytes(buf[off:off + n8], "little")
    off += n8
    assert prime == R
    n = struct.unpack_from("<I", buf, off)[0]
    off, _ = secs[2][0]
    return [int.from_bytes(buf[off + n8 * i:off + n8 * (i + 1)], "little") for i in range(n)]
