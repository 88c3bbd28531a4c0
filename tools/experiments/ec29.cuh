// ec29.cuh -- the bucket-accumulation group law (XYZZ += affine) over the carry-free 9 x 29-bit field of fp29.cuh.
//
// Same formulas as xyzz_t::madd in ec.cuh (madd-2008-s, 8M + 2S; replaces the bucket additions inside arkworks'
// `VariableBaseMSM`, reached from /root/reference/dist-primitives/src/dmsm/mod.rs:82), different number system:
// coordinates live in the 2^261 Montgomery domain as loose 29-bit limbs, additions are limb-wise without carries,
// subtractions add a limb-wise "borrow-free" multiple of p, and a carry pass (norm) runs only before a value becomes
// a product operand.  The comments carry the value bounds (in multiples of p) that make every step safe:
//   product:  out < p + a b / 2^261, and p / 2^261 < 0.0059, so e.g. 11p x 11p -> out < 1.72 p;
//   operands: limbs < 2^30 on both sides (or 2^31 x 2^29) keep the 64-bit column sums from overflowing.
// Points enter as stored affine coordinates (8 x u32, 2^256 domain; a 5-bit shift re-limbs them, fp29.cuh) and the
// finished bucket leaves as an ordinary xyzz_t<Fq> (4 products with the constant 2^256), so nothing outside the
// bucket kernel sees this representation.
#pragma once
#include "../../distributed_groth16_b200/csrc/ec.cuh"
#include "fp29.cuh"

namespace b200zk {

// ---- G1: coordinates in Fq29 -------------------------------------------------------------------------------------
struct xyzz29_g1 {
    typedef Fq29 F;
    F x, y, zz, zzz;        // invariant: x, y normalised and < 8 p; zz, zzz product outputs (< 2 p); zz == 0 limbs <=> identity

    B2_HD static xyzz29_g1 identity() { xyzz29_g1 r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r; }
    B2_HD bool is_inf() const { return zz.limbs_zero(); }

    // 2 * (px, py) for an affine point already in the 2^261 domain (< 32 p): the P + P corner of madd, rare
    B2_HD_NI static xyzz29_g1 dbl_affine(const F& px, const F& py) {
        // px < 32 p, py < 33 p
        F U = F::norm(F::dbl_lazy(py));                         // < 66 p
        F V = F::sqr(U);                                        // < p + 66^2 * .0059 p  < 27 p   (still a fine operand)
        F W = F::mul(U, V);                                     // < p + 66 * 27 * .0059 p < 12 p
        F S = F::mul(px, V);                                    // < p + 32 * 27 * .0059 p < 6.1 p
        F X2 = F::sqr(px);                                      // < p + 1024 * .0059 p < 7.1 p
        F M = F::norm(F::add_lazy(F::dbl_lazy(X2), X2));        // < 21.3 p, limbs < 3 * 2^29 before norm
        F MM = F::sqr(M);                                       // < 3.7 p
        const F one = F::one();
        xyzz29_g1 r;
        r.x = F::norm(F::template sub_lazy<13, 31>(MM, F::dbl_lazy(S)));        // 2 S < 12.2 p  ->  < 16.7 p
        r.x = F::mul(r.x, one);                                 // same residue, back under 1.1 p
        F T = F::template sub<2>(S, r.x);                       // < 8.1 p
        F T1 = F::mul(M, T);                                    // < p + 21.3 * 8.1 * .0059 p < 2.1 p
        F T2 = F::mul(W, py);                                   // < p + 12 * 33 * .0059 p < 3.4 p
        r.y = F::template sub<4>(T1, T2);                       // < 6.1 p
        r.zz = F::mul(V, one);                                  // < 1.2 p
        r.zzz = F::mul(W, one);
        return r;
    }

    // acc += (negate ? -p : p) for a stored affine point p (2^256 domain, canonical); handles p = inf, acc = inf, acc = +-p
    B2_HD static void madd(xyzz29_g1& acc, const affine_t<Fq>& p, bool negate) {
        if (p.is_inf()) return;
        F X2 = F::from_mont256(p.x);                            // < 32 p, limbs < 2^29
        F Y2 = F::from_mont256(p.y);
        if (acc.is_inf()) {
            const F one = F::one();
            acc.x = F::mul(X2, one);                            // same residue, < p + 32 * .0059 p: restores the x, y < 8 p invariant
            F y1 = F::mul(Y2, one);
            acc.y = negate ? F::norm(F::template neg_lazy<2, 29>(y1)) : y1;
            acc.zz = one;
            acc.zzz = one;
            return;
        }
        F U2 = F::mul(X2, acc.zz);                              // < p + 32 * 2 * .0059 p = 1.38 p
        F S2 = F::mul(Y2, acc.zzz);                             // < 1.38 p
        F Pp = F::template sub<9>(U2, acc.x);                   // U2 - X1 + 9 p < 10.4 p
        // R = +-S2 - Y1: the sign of the point goes on S2 (2 p - S2), limb-wise, lazily
        F Ts;
        {
            F n = F::template neg_lazy<2, 29>(S2);
#pragma unroll
            for (int k = 0; k < 9; ++k) Ts.l[k] = negate ? n.l[k] : S2.l[k];
        }
        F R = F::norm(F::template sub_lazy<9, 29>(Ts, acc.y));  // < 2 p + 9 p = 11 p
        F PP = F::sqr(Pp);                                      // < p + 108 * .0059 p = 1.64 p
        F PPP = F::mul(Pp, PP);                                 // < 1.11 p
        F Q = F::mul(acc.x, PP);                                // < 1.08 p
        F RR = F::sqr(R);                                       // < 1.72 p
        F ZZ3 = F::mul(acc.zz, PP);                             // < 1.02 p
        if (ZZ3.is_zero_mod_2p()) {                             // <=> Pp == 0 mod p (zz != 0): same x, i.e. acc = +-p
            bool same = F::is_zero_mod_slow(R, 11);
            if (same) acc = dbl_affine(X2, negate ? F::norm(F::template neg_lazy<33, 29>(Y2)) : Y2);
            else acc = identity();
            return;
        }
        // X3 = RR - PPP - 2 Q + 4 p   (PPP + 2 Q < 3.3 p as a lazy sum with limbs < 3 * 2^29)
        F X3 = F::norm(F::template sub_lazy<4, 31>(RR, F::add_lazy(PPP, F::dbl_lazy(Q))));      // < 5.72 p
        F QX = F::template sub<6>(Q, X3);                       // < 7.1 p
        F T1 = F::mul(R, QX);                                   // < p + 11 * 7.1 * .0059 p = 1.46 p
        F T2 = F::mul(acc.y, PPP);                              // < 1.06 p
        acc.x = X3;
        acc.y = F::template sub<2>(T1, T2);                     // < 3.46 p
        acc.zz = ZZ3;
        acc.zzz = F::mul(acc.zzz, PPP);
    }

    // leave the 29-bit world: canonical 2^256-domain coordinates
    B2_HD static xyzz_t<Fq> to_xyzz(const xyzz29_g1& a) {
        xyzz_t<Fq> r;
        if (a.is_inf()) return xyzz_t<Fq>::identity();
        r.x = F::to_mont256(a.x);
        r.y = F::to_mont256(a.y);
        r.zz = F::to_mont256(a.zz);
        r.zzz = F::to_mont256(a.zzz);
        return r;
    }
};

}  // namespace b200zk
