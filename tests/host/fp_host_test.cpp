// Host-side unit test of the device limb schedule in csrc/fp.cuh (compiled with g++; the PTX carry
// primitives run through their C emulation).  Reference: oracle/liboracle.so (orc_field_op).
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include "../../distributed_groth16_b200/csrc/fp.cuh"
#include "../../distributed_groth16_b200/csrc/ec.cuh"
#include "../../tools/experiments/ec29.cuh"
#include "../../tools/experiments/batch_affine.cuh"
#include "../../tools/experiments/fp52.cuh"
#include <cfenv>

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
        { F sq = F::sqr(fa), mm = F::mul(fa, fa); if (sq != mm) { fails++; if (fails < 5) printf("%s sqr mismatch it=%d\n", name, it); } }
        {   // unreduced product + separate reduction == the row-interleaved product
            uint32_t t[16]; F::mul_wide(t, fa.l, fb.l); F w; F::template redc<1>(w, t);
            if (w != F::mul(fa, fb)) { fails++; if (fails < 5) printf("%s mul_wide/redc mismatch it=%d\n", name, it); }
        }
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


// ---- fp29.cuh / ec29.cuh: the 9 x 29-bit carry-free field and the bucket group law built on it ----------------------
template <class P, int FIELD>
static int test_field29(const char* name) {
    typedef Fp<P> F;
    typedef Fp29<P> G;
    int fails = 0;
    const int N = 20000;
    for (int it = 0; it < N; ++it) {
        uint64_t buf[8];
        orc_fr_generate(rnd(), 2, buf);
        F a, b; memcpy(a.l, buf, 32); memcpy(b.l, buf + 4, 32);
        if (FIELD == 0) { a.l[7] &= 0x0FFFFFFFu; b.l[7] &= 0x0FFFFFFFu; }
        if (it == 0) a = F::zero();
        if (it == 1) b = F::zero();
        if (it == 2) { a = F::neg(F::one()); b = a; }                       // p - R: large canonical values
        if (it == 3) { a = F::zero(); a.l[0] = 1; a = F::neg(a); b = a; }   // p - 1, the largest canonical integer
        G ga = G::from_mont256(a), gb = G::from_mont256(b);
        F r;
        r = G::to_mont256(G::mul(ga, gb)); if (r != F::mul(a, b)) { if (fails++ < 5) printf("%s mul29 mismatch it=%d\n", name, it); }
        r = G::to_mont256(G::sqr(ga)); if (r != F::mul(a, a)) { if (fails++ < 5) printf("%s sqr29 mismatch it=%d\n", name, it); }
        r = G::to_mont256(G::norm(G::add_lazy(ga, gb))); if (r != F::add(a, b)) { if (fails++ < 5) printf("%s add29 mismatch it=%d\n", name, it); }
        {   // subtraction needs value(b) < K p: reduce the operands below 2 p first (x * 1)
            G ra = G::mul(ga, G::one()), rb = G::mul(gb, G::one());
            r = G::to_mont256(G::template sub<2>(ra, rb)); if (r != F::sub(a, b)) { if (fails++ < 5) printf("%s sub29 mismatch it=%d\n", name, it); }
            G lz = G::template sub_lazy<4, 31>(ra, G::add_lazy(rb, G::dbl_lazy(rb)));     // a - 3 b + 4 p  (2 + 2 * 2 ... uses < 3.6 p)
            r = G::to_mont256(G::norm(lz)); if (r != F::sub(a, F::add(b, F::dbl(b)))) { if (fails++ < 5) printf("%s sub_lazy<.,31> mismatch it=%d\n", name, it); }
            r = G::to_mont256(G::norm(G::template neg_lazy<2, 29>(rb))); if (r != F::neg(b)) { if (fails++ < 5) printf("%s neg29 mismatch it=%d\n", name, it); }
            // mixed-domain product: 2^261-domain x (plain re-limbed 2^256-domain) lands in the 2^256 domain
            r = G::canon256(G::mul(ga, G::relimb(b))); if (r != F::mul(a, b)) { if (fails++ < 5) printf("%s mixed mul mismatch it=%d\n", name, it); }
            bool z = G::mul(ra, rb).is_zero_mod_2p();
            if (z != (a.is_zero() || b.is_zero())) { if (fails++ < 5) printf("%s zero test mismatch it=%d\n", name, it); }
        }
    }
    printf("%s (29-bit limbs): %d iterations, %d failures\n", name, N, fails);
    return fails;
}

static int test_curve29() {
    const int n = 64;
    uint64_t* pts = (uint64_t*)malloc(n * 64);
    orc_g1_generate(4242, n, pts, 1);
    int fails = 0;
    for (int round = 0; round < 50; ++round) {
        xyzz_t<Fq> ref = xyzz_t<Fq>::identity();
        xyzz29_g1 acc = xyzz29_g1::identity();
        for (int step = 0; step < 40; ++step) {
            uint64_t r = rnd();
            affine_t<Fq> p; memcpy(&p, pts + 8 * (r % n), 64);
            bool neg = (r >> 20) & 1;
            if (step == 7) p = affine_t<Fq>::infinity();
            if (round % 5 == 1 && step == 1) { memcpy(&p, pts + 8 * 3, 64); neg = false; }      // P + P right after the first point
            if (round % 5 == 1 && step == 0) { memcpy(&p, pts + 8 * 3, 64); neg = false; }
            if (round % 5 == 2 && step == 0) { memcpy(&p, pts + 8 * 4, 64); neg = true; }       // -P then +P: identity, then keep going
            if (round % 5 == 2 && step == 1) { memcpy(&p, pts + 8 * 4, 64); neg = false; }
            if (round % 5 == 3 && step == 20) {   // acc == +-p deep inside a chain: add the current sum itself
                affine_t<Fq> cur = xyzz_t<Fq>::to_affine(ref);
                p = cur; neg = (round & 1) != 0;
            }
            xyzz_t<Fq>::madd(ref, p, neg);
            xyzz29_g1::madd(acc, p, neg);
            xyzz_t<Fq> got = xyzz29_g1::to_xyzz(acc);
            affine_t<Fq> ga = xyzz_t<Fq>::to_affine(got), ra = xyzz_t<Fq>::to_affine(ref);
            if (memcmp(&ga, &ra, sizeof(ga)) || got.is_inf() != ref.is_inf()) {
                if (fails++ < 5) printf("G1 (29-bit limbs): madd mismatch round %d step %d\n", round, step);
                break;
            }
        }
    }
    printf("G1 (29-bit limbs) bucket group law: %d failures\n", fails);
    free(pts);
    return fails;
}


// ---- batch_affine.cuh: affine additions around a shared inversion (Montgomery's trick) vs the XYZZ group law ----------
template <class C>
static int test_batch_affine(const char* name, int is_g2) {
    typedef typename C::F F;
    typedef batch_affine<F> BA;
    const int n = 48, SLOTS = 12, STEPS = 9;
    const int PL = is_g2 ? 16 : 8;
    uint64_t* raw = (uint64_t*)malloc(n * PL * 8);
    if (is_g2) orc_g2_generate(31337, n, raw, 1); else orc_g1_generate(31337, n, raw, 1);
    affine_t<F>* pts = (affine_t<F>*)raw;
    int fails = 0;
    for (int round = 0; round < 20; ++round) {
        affine_t<F> acc[SLOTS];
        xyzz_t<F> ref[SLOTS];
        for (int k = 0; k < SLOTS; ++k) { acc[k] = affine_t<F>::infinity(); ref[k] = xyzz_t<F>::identity(); }
        for (int step = 0; step < STEPS; ++step) {
            affine_t<F> pt[SLOTS];
            bool neg[SLOTS];
            for (int k = 0; k < SLOTS; ++k) {
                uint64_t r = rnd();
                pt[k] = pts[r % n];
                neg[k] = (r >> 33) & 1;
                if (k == 1 && step == 3) pt[k] = affine_t<F>::infinity();                               // infinity operand
                if (k == 2 && step >= 1 && step <= 2) { pt[k] = pts[5]; neg[k] = false; }                 // P + P  (doubling)
                if (k == 2 && step == 0) { pt[k] = pts[5]; neg[k] = false; }
                if (k == 3 && step == 0) { pt[k] = pts[6]; neg[k] = true; }                             // -P, then +P: cancels
                if (k == 3 && step == 1) { pt[k] = pts[6]; neg[k] = false; }
                if (k == 4 && step == 5) { pt[k] = acc[k]; neg[k] = (round & 1) != 0; }                  // acc +- acc mid-chain
                pt[k] = BA::signed_point(pt[k], neg[k]);
            }
            // forward: denominators, running product
            F d[SLOTS], pre[SLOTS], run = F::one();
            int cs[SLOTS];
            for (int k = 0; k < SLOTS; ++k) {
                cs[k] = BA::prepare(acc[k], pt[k], d[k]);
                pre[k] = run;
                if (BA::needs_inverse(cs[k])) run = F::mul(run, d[k]);
            }
            F inv = F::inv(run);
            for (int k = SLOTS - 1; k >= 0; --k) {
                F dinv = inv;
                if (BA::needs_inverse(cs[k])) { dinv = F::mul(inv, pre[k]); inv = F::mul(inv, d[k]); }
                BA::finish(cs[k], acc[k], pt[k], dinv);
            }
            for (int k = 0; k < SLOTS; ++k) {
                xyzz_t<F>::madd(ref[k], pt[k], false);
                affine_t<F> ra = xyzz_t<F>::to_affine(ref[k]);
                if (memcmp(&ra, &acc[k], sizeof(ra))) { if (fails++ < 5) printf("%s batch_affine mismatch round %d step %d slot %d\n", name, round, step, k); }
            }
        }
    }
    printf("%s batched affine additions: %d failures\n", name, fails);
    free(raw);
    return fails;
}

static int test_fq2_lazy_product() {
    // Fq2::mul (three unreduced products, two reductions) vs the schoolbook formula on Fq, including the largest operands
    int fails = 0;
    Fq big = Fq::zero(); big.l[0] = 1; big = Fq::neg(big);          // p - 1
    for (int it = 0; it < 5000; ++it) {
        uint64_t buf[16];
        orc_fr_generate(rnd(), 4, buf);
        Fq2 a, b;
        memcpy(a.c0.l, buf, 32); memcpy(a.c1.l, buf + 4, 32); memcpy(b.c0.l, buf + 8, 32); memcpy(b.c1.l, buf + 12, 32);
        a.c0.l[7] &= 0x0FFFFFFFu; a.c1.l[7] &= 0x0FFFFFFFu; b.c0.l[7] &= 0x0FFFFFFFu; b.c1.l[7] &= 0x0FFFFFFFu;
        if (it < 16) { if (it & 1) a.c0 = big; if (it & 2) a.c1 = big; if (it & 4) b.c0 = big; if (it & 8) b.c1 = big; }
        if (it == 16) { a = Fq2::zero(); }
        if (it == 17) { a.c0 = Fq::zero(); b.c1 = Fq::zero(); }
        Fq2 got = Fq2::mul(a, b), want;
        want.c0 = Fq::sub(Fq::mul(a.c0, b.c0), Fq::mul(a.c1, b.c1));
        want.c1 = Fq::add(Fq::mul(a.c0, b.c1), Fq::mul(a.c1, b.c0));
        if (!(got == want)) { if (fails++ < 5) printf("Fq2 lazy product mismatch it=%d\n", it); }
    }
    printf("Fq2 lazy-reduction product: %d failures\n", fails);
    return fails;
}

// ---- fp52.cuh: 5 x 52-bit limbs, products from pairs of round-towards-zero FMAs -------------------------------------------
template <class P, int FIELD>
static int test_field52(const char* name) {
    typedef Fp<P> F;
    typedef Fp52<P> G;
    int fails = 0;
    const int N = 20000;
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    for (int it = 0; it < N; ++it) {
        uint64_t buf[8];
        orc_fr_generate(rnd(), 2, buf);
        F a, b; memcpy(a.l, buf, 32); memcpy(b.l, buf + 4, 32);
        if (FIELD == 0) { a.l[7] &= 0x0FFFFFFFu; b.l[7] &= 0x0FFFFFFFu; }
        if (it == 0) a = F::zero();
        if (it == 1) b = F::zero();
        if (it == 2) { a = F::neg(F::one()); b = a; }
        if (it == 3) { a = F::zero(); a.l[0] = 1; a = F::neg(a); b = a; }   // p - 1, the largest canonical integer
        G ga = G::from_mont256(a), gb = G::from_mont256(b);
        F r = G::to_mont256(G::mul(ga, gb));
        if (r != F::mul(a, b)) { if (fails++ < 5) printf("%s mul52 mismatch it=%d\n", name, it); }
        // a chain: ((a b) a) b
        r = G::to_mont256(G::mul(G::mul(G::mul(ga, gb), ga), gb));
        if (r != F::mul(F::mul(F::mul(a, b), a), b)) { if (fails++ < 5) printf("%s mul52 chain mismatch it=%d\n", name, it); }
    }
    fesetround(old);
    printf("%s (52-bit limbs, FMA products): %d iterations, %d failures\n", name, N, fails);
    return fails;
}

int main() {
    int fails = 0;
    fails += test_field52<FqParams, 0>("Fq");
    fails += test_field52<FrParams, 1>("Fr");
    fails += test_field<Fq, 0>("Fq");
    fails += test_field<Fr, 1>("Fr");
    fails += test_curve<G1Curve>("G1", 0);
    fails += test_curve<G2Curve>("G2", 1);
    fails += test_field29<FqParams, 0>("Fq");
    fails += test_field29<FrParams, 1>("Fr");
    fails += test_fq2_lazy_product();
    fails += test_curve29();
    fails += test_batch_affine<G1Curve>("G1", 0);
    fails += test_batch_affine<G2Curve>("G2", 1);
    printf(fails ? "FAILED\n" : "ALL OK\n");
    return fails ? 1 : 0;
}
