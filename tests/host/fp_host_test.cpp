// Host-side unit test of the device limb schedule in csrc/fp.cuh (compiled with g++; the PTX carry
// primitives run through their C emulation).  Reference: oracle/liboracle.so (orc_field_op).
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include "../../distributed_groth16_b200/csrc/fp.cuh"
#include "../../distributed_groth16_b200/csrc/ec.cuh"

extern "C" void orc_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out);
extern "C" void orc_fr_generate(uint64_t seed, size_t n, uint64_t* out);
extern "C" void orc_g1_generate(uint64_t seed, size_t n, uint64_t* out, int nthreads);
extern "C" void orc_g2_generate(uint64_t seed, size_t n, uint64_t* out, int nthreads);
extern "C" int orc_msm_g1_naive(const uint64_t* bases, const uint64_t* scalars, size_t n, uint64_t* out, int* inf);
extern "C" int orc_msm_g2_naive(const uint64_t* bases, const uint64_t* scalars, size_t n, uint64_t* out, int* inf);

using namespace b200zk;

static uint64_t rng_state = 0x1234567;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <class F, int FIELD>
static int test_field(const char* name) {
    int fails = 0;
    const int N = 20000;
    for (int it = 0; it < N; ++it) {
        uint64_t a[4], b[4];
        // random canonical-ish elements: use orc to_mont of random 256-bit reduced by generator
        uint64_t buf[8];
        orc_fr_generate(rnd(), 2, buf);
        memcpy(a, buf, 32); memcpy(b, buf + 4, 32);
        if (FIELD == 0) { a[3] &= 0x0FFFFFFFFFFFFFFFULL; b[3] &= 0x0FFFFFFFFFFFFFFFULL; }   // < 2^252 < q
        if (it == 0) { memset(a, 0, 32); }
        if (it == 1) { memset(b, 0, 32); }
        if (it == 2) { for (int i = 0; i < 8; ++i) { ((uint32_t*)a)[i] = F::zero().l[i]; } a[0] = 1; }
        if (it == 3) { F m1 = F::neg(F::one()); memcpy(a, m1.l, 32); memcpy(b, m1.l, 32); }
        F fa, fb; memcpy(fa.l, a, 32); memcpy(fb.l, b, 32);
        uint64_t exp[4];
        F r;
        r = F::mul(fa, fb); orc_field_op(FIELD, 0, a, b, exp); if (memcmp(r.l, exp, 32)) { fails++; if (fails < 5) printf("%s mul mismatch it=%d\n", name, it); }
        r = F::add(fa, fb); orc_field_op(FIELD, 1, a, b, exp); if (memcmp(r.l, exp, 32)) { fails++; if (fails < 5) printf("%s add mismatch it=%d\n", name, it); }
        r = F::sub(fa, fb); orc_field_op(FIELD, 2, a, b, exp); if (memcmp(r.l, exp, 32)) { fails++; if (fails < 5) printf("%s sub mismatch it=%d\n", name, it); }
        if (it < 2000) {
            r = F::inv(fa); orc_field_op(FIELD, 3, a, nullptr, exp);
            if (!fa.is_zero() && memcmp(r.l, exp, 32)) { fails++; printf("%s inv mismatch it=%d\n", name, it); }
            r = F::inv_fermat(fa);
            if (!fa.is_zero() && memcmp(r.l, exp, 32)) { fails++; printf("%s inv_fermat mismatch it=%d\n", name, it); }
            r = F::from_mont(fa); orc_field_op(FIELD, 5, a, nullptr, exp); if (memcmp(r.l, exp, 32)) { fails++; printf("%s from_mont mismatch\n", name); }
        }
    }
    printf("%s: %d iterations, %d failures\n", name, N, fails);
    return fails;
}

// sum_i s_i P_i with the device curve code (double-and-add through xyzz), compared with the oracle
template <class C>
static int test_curve(const char* name, int is_g2) {
    typedef typename C::F F;
    const int n = 24;
    const int PL = is_g2 ? 16 : 8;
    uint64_t* pts = (uint64_t*)malloc(n * PL * 8);
    uint64_t sc[4 * n];
    if (is_g2) orc_g2_generate(77, n, pts, 1); else orc_g1_generate(77, n, pts, 1);
    orc_fr_generate(99, n, sc);
    memset(sc + 4 * 3, 0, 32);                      // zero scalar
    memset(pts + PL * 5, 0, PL * 8);                // infinity base
    memcpy(pts + PL * 7, pts + PL * 6, PL * 8);     // duplicate base (exercises doubling when scalars equal)
    memcpy(sc + 4 * 7, sc + 4 * 6, 32);
    xyzz_t<F> acc = xyzz_t<F>::identity();
    for (int i = 0; i < n; ++i) {
        affine_t<F> p; memcpy(&p, pts + PL * i, PL * 8);
        Fr s; memcpy(s.l, sc + 4 * i, 32);
        Fr k = Fr::from_mont(s);
        xyzz_t<F> t = xyzz_t<F>::identity();
        for (int bit = 255; bit >= 0; --bit) {
            t = xyzz_t<F>::dbl(t);
            if ((k.l[bit >> 5] >> (bit & 31)) & 1) xyzz_t<F>::madd(t, p, false);
        }
        acc = xyzz_t<F>::add(acc, t);
    }
    // also exercise P + P through madd, P + (-P), and negated madd
    {
        affine_t<F> p; memcpy(&p, pts, PL * 8);
        xyzz_t<F> t = xyzz_t<F>::identity();
        xyzz_t<F>::madd(t, p, false); xyzz_t<F>::madd(t, p, false);      // 2P via equal-point path
        xyzz_t<F> u = xyzz_t<F>::identity();
        xyzz_t<F>::madd(u, p, false); u = xyzz_t<F>::dbl(u);
        affine_t<F> ta = xyzz_t<F>::to_affine(t), ua = xyzz_t<F>::to_affine(u);
        if (memcmp(&ta, &ua, sizeof(ta))) { printf("%s: madd doubling path mismatch\n", name); return 1; }
        xyzz_t<F>::madd(t, p, true); xyzz_t<F>::madd(t, p, true);        // back to identity
        if (!t.is_inf()) { printf("%s: P + (-P) path mismatch\n", name); return 1; }
        {   // latency-optimised (row-interleaved) group law == plain group law
            affine_t<F> q; memcpy(&q, pts + PL * 9, PL * 8);
            xyzz_t<F> v = xyzz_t<F>::dbl(xyzz_t<F>::from_affine(q));
            xyzz_t<F> s1 = xyzz_t<F>::add(u, v), s2 = xyzz_t<F>::add_ilp(u, v);
            xyzz_t<F> d1 = xyzz_t<F>::dbl(s1), d2 = xyzz_t<F>::dbl_ilp(s2);
            if (memcmp(&s1, &s2, sizeof(s1)) || memcmp(&d1, &d2, sizeof(d1))) { printf("%s: ilp group law mismatch\n", name); return 1; }
            xyzz_t<F> e1 = xyzz_t<F>::add_ilp(u, u), e2 = xyzz_t<F>::dbl(u);
            affine_t<F> ea = xyzz_t<F>::to_affine(e1), eb = xyzz_t<F>::to_affine(e2);
            if (memcmp(&ea, &eb, sizeof(ea))) { printf("%s: add_ilp equal-operands mismatch\n", name); return 1; }
        }
        xyzz_t<F> w = xyzz_t<F>::add(u, u);                               // add with equal operands
        xyzz_t<F> w2 = xyzz_t<F>::dbl(u);
        affine_t<F> wa = xyzz_t<F>::to_affine(w), w2a = xyzz_t<F>::to_affine(w2);
        if (memcmp(&wa, &w2a, sizeof(wa))) { printf("%s: add equal-operands mismatch\n", name); return 1; }
    }
    affine_t<F> res = xyzz_t<F>::to_affine(acc);
    uint64_t exp[16]; int inf = 0;
    if (is_g2) orc_msm_g2_naive(pts, sc, n, exp, &inf); else orc_msm_g1_naive(pts, sc, n, exp, &inf);
    int bad = memcmp(&res, exp, PL * 8) != 0;
    printf("%s: msm-by-double-and-add %s\n", name, bad ? "MISMATCH" : "ok");
    free(pts);
    return bad;
}

int main() {
    int fails = 0;
    fails += test_field<Fq, 0>("Fq");
    fails += test_field<Fr, 1>("Fr");
    fails += test_curve<G1Curve>("G1", 0);
    fails += test_curve<G2Curve>("G2", 1);
    printf(fails ? "FAILED\n" : "ALL OK\n");
    return fails ? 1 : 0;
}
