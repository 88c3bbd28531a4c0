// batch_affine.cuh -- EXPERIMENT (see msm_chains.cuh for the outcome): affine point additions with a shared (batched)
// inversion, the arithmetic of a first bucket accumulation level made of "chains".
//
// Replaces the bucket additions inside arkworks' `VariableBaseMSM` (reached from
// /root/reference/dist-primitives/src/dmsm/mod.rs:82).  An XYZZ mixed addition costs 8M + 2S = 10 field products;
// an affine addition costs 1 inversion + 2M + 1S, and Montgomery's trick turns N inversions into one inversion plus
// 3 (N - 1) products: 6 products per addition once the one inversion is shared by enough additions.  The bucket kernel
// of round 1 ran at 91% of the multiplier pipe, so doing fewer products is the only way to make it faster.
//
// A slot (one running affine sum `acc` and the point `pt` to add to it) goes through two calls around the shared inversion:
//   prepare(): classify the pair and hand back the denominator d of the slope (x2 - x1, or 2 y1 for a doubling);
//   finish():  given 1/d, replace acc by acc + pt.
// The corner cases need no inversion and return NO denominator (the caller multiplies nothing into its running product):
//   pt = inf -> unchanged, acc = inf -> pt, acc = -pt -> inf.  BN254 has no point with y = 0 (no 2-torsion), so 2 y1 != 0.
// Infinity is the all-zero pair (the zkey convention, ark-circom/src/zkey.rs:353-373; not on y^2 = x^3 + b).
// Everything is B2_HD and unit-tested on the host against the XYZZ group law (tests/host/fp_host_test.cpp).
#pragma once
#include "../../distributed_groth16_b200/csrc/ec.cuh"

namespace b200zk {

template <class F>
struct batch_affine {
    enum Case : int { NOP = 0, COPY = 1, CANCEL = 2, DBL = 3, ADD = 4 };

    B2_HD static bool needs_inverse(int cs) { return cs >= DBL; }

    B2_HD static int prepare(const affine_t<F>& acc, const affine_t<F>& pt, F& d) {
        if (pt.is_inf()) return NOP;
        if (acc.is_inf()) return COPY;
        if (acc.x == pt.x) {
            if (acc.y == pt.y) { d = F::dbl(acc.y); return DBL; }
            return CANCEL;                               // same x, y2 = -y1
        }
        d = F::sub(pt.x, acc.x);
        return ADD;
    }

    // acc <- acc + pt;  dinv = 1/d for the DBL / ADD cases (ignored otherwise)
    B2_HD static void finish(int cs, affine_t<F>& acc, const affine_t<F>& pt, const F& dinv) {
        if (cs == NOP) return;
        if (cs == COPY) { acc = pt; return; }
        if (cs == CANCEL) { acc = affine_t<F>::infinity(); return; }
        F num;
        if (cs == DBL) {
            F xx = F::sqr(acc.x);
            num = F::add(F::dbl(xx), xx);                // 3 x1^2  (a = 0)
        } else {
            num = F::sub(pt.y, acc.y);
        }
        F lam = F::mul(num, dinv);
        F x3 = F::sub(F::sub(F::sqr(lam), acc.x), pt.x);
        F y3 = F::sub(F::mul(lam, F::sub(acc.x, x3)), acc.y);
        acc.x = x3;
        acc.y = y3;
    }

    B2_HD static affine_t<F> signed_point(const affine_t<F>& p, bool negate) {
        affine_t<F> r = p;
        if (negate && !p.is_inf()) r.y = F::neg(p.y);
        return r;
    }
};

}  // namespace b200zk
