// ec.cuh -- short-Weierstrass (a = 0) group arithmetic for BN254 G1 (over Fq) and G2 (over Fq2).
//
// Replaces arkworks' `short_weierstrass::{Affine,Projective}` (un-vendored) as used by `G::msm`
// (/root/reference/dist-primitives/src/dmsm/mod.rs:82) and the proof assembly
// (/root/reference/groth16/src/prove.rs:36-44,75-83,128-134).
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// a mixed add costs 8M+2S and needs no field inversion, which is what the bucket kernels want.
// Infinity: affine (0,0) (the zkey convention, ark-circom/src/zkey.rs:353-373); XYZZ with ZZ = 0.
// Results only ever leave the library as canonical affine coordinates, so they are bit-identical
// to arkworks' regardless of the internal coordinate system.
#pragma once
#include "fp.cuh"

namespace b200zk {

template <class F>
struct affine_t {
    F x, y;
    B2_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    B2_HD static affine_t infinity() { affine_t r; r.x = F::zero(); r.y = F::zero(); return r; }
};

template <class F>
struct xyzz_t {
    F x, y, zz, zzz;

    B2_HD static xyzz_t identity() {
        xyzz_t r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r;
    }
    B2_HD bool is_inf() const { return zz.is_zero(); }

    B2_HD static xyzz_t from_affine(const affine_t<F>& p) {
        if (p.is_inf()) return identity();
        xyzz_t r; r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one(); return r;
    }

    // 2*p for an affine p (mdbl-2008-s-1)
    B2_HD_NI static xyzz_t dbl_affine(const F& px, const F& py) {
        F U = F::dbl(py);
        F V = F::sqr(U);
        F W = F::mul(U, V);
        F S = F::mul(px, V);
        F X2 = F::sqr(px);
        F M = F::add(F::dbl(X2), X2);
        xyzz_t r;
        r.x = F::sub(F::sqr(M), F::dbl(S));
        r.y = F::sub(F::mul(M, F::sub(S, r.x)), F::mul(W, py));
        r.zz = V;
        r.zzz = W;
        return r;
    }

    // dbl-2008-s-1
    B2_HD_NI static xyzz_t dbl(const xyzz_t& p) {
        if (p.is_inf()) return p;
        F U = F::dbl(p.y);
        F V = F::sqr(U);
        F W = F::mul(U, V);
        F S = F::mul(p.x, V);
        F X2 = F::sqr(p.x);
        F M = F::add(F::dbl(X2), X2);
        xyzz_t r;
        r.x = F::sub(F::sqr(M), F::dbl(S));
        r.y = F::sub(F::mul(M, F::sub(S, r.x)), F::mul(W, p.y));
        r.zz = F::mul(V, p.zz);
        r.zzz = F::mul(W, p.zzz);
        return r;
    }

    // acc += (negate ? -p : p), p affine (madd-2008-s); handles p = inf, acc = inf, acc = +-p
    B2_HD static void madd(xyzz_t& acc, const affine_t<F>& p, bool negate) {
        if (p.is_inf()) return;
        F py = negate ? F::neg(p.y) : p.y;
        if (acc.is_inf()) {
            acc.x = p.x; acc.y = py; acc.zz = F::one(); acc.zzz = F::one();
            return;
        }
        F U2 = F::mul(p.x, acc.zz);
        F S2 = F::mul(py, acc.zzz);
        F Pp = F::sub(U2, acc.x);
        F R = F::sub(S2, acc.y);
        if (Pp.is_zero()) {
            if (R.is_zero()) acc = dbl_affine(p.x, py);
            else acc = identity();
            return;
        }
        F PP = F::sqr(Pp);
        F PPP = F::mul(Pp, PP);
        F Q = F::mul(acc.x, PP);
        F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        F Y3 = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(acc.y, PPP));
        acc.x = X3;
        acc.y = Y3;
        acc.zz = F::mul(acc.zz, PP);
        acc.zzz = F::mul(acc.zzz, PPP);
    }

    // add-2008-s
    B2_HD_NI static xyzz_t add(const xyzz_t& a, const xyzz_t& b) {
        if (a.is_inf()) return b;
        if (b.is_inf()) return a;
        F U1 = F::mul(a.x, b.zz);
        F U2 = F::mul(b.x, a.zz);
        F S1 = F::mul(a.y, b.zzz);
        F S2 = F::mul(b.y, a.zzz);
        F Pp = F::sub(U2, U1);
        F R = F::sub(S2, S1);
        if (Pp.is_zero()) {
            if (R.is_zero()) return dbl(a);
            return identity();
        }
        F PP = F::sqr(Pp);
        F PPP = F::mul(Pp, PP);
        F Q = F::mul(U1, PP);
        xyzz_t r;
        r.x = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        r.y = F::sub(F::mul(R, F::sub(Q, r.x)), F::mul(S1, PPP));
        r.zz = F::mul(F::mul(a.zz, b.zz), PP);
        r.zzz = F::mul(F::mul(a.zzz, b.zzz), PPP);
        return r;
    }

    B2_HD static xyzz_t neg(const xyzz_t& a) { xyzz_t r = a; r.y = F::neg(a.y); return r; }

    // canonical affine (x = X/ZZ, y = Y/ZZZ); infinity -> (0,0)
    B2_HD_NI static affine_t<F> to_affine(const xyzz_t& a) {
        if (a.is_inf()) return affine_t<F>::infinity();
        F izzz = F::inv(a.zzz);                      // 1/ZZZ
        F izz = F::sqr(F::mul(izzz, a.zz));          // ZZ^3 = ZZZ^2  =>  1/ZZ = (ZZ/ZZZ)^2
        affine_t<F> r;
        r.x = F::mul(a.x, izz);
        r.y = F::mul(a.y, izzz);
        return r;
    }

    // k * p for a 256-bit canonical (non-Montgomery) scalar k given as 8 x u32
    B2_HD_NI static xyzz_t mul_scalar(const xyzz_t& p, const uint32_t k[8]) {
        xyzz_t acc = identity();
        for (int bit = 255; bit >= 0; --bit) {
            acc = dbl(acc);
            if ((k[bit >> 5] >> (bit & 31)) & 1) acc = add(acc, p);
        }
        return acc;
    }
};

struct G1Curve {
    typedef Fq F;
    static constexpr int LIMBS64 = 8;     // u64 limbs per affine point
};
struct G2Curve {
    typedef Fq2 F;
    static constexpr int LIMBS64 = 16;
};

}  // namespace b200zk
