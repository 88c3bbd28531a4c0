// pairing.cuh -- optimal-ate pairing on BN254 and the Groth16 verification equation.
//
// Closes the loop the reference closes with `Groth16::verify_with_processed_vk`
// (/root/reference/groth16/examples/sha256.rs:229-254, mpc-api/src/main.rs:187-247):
//     e(A, B) = e(alpha, beta) * e(sum_i x_i IC_i, gamma) * e(C, delta).
// arkworks' pairing lives in ark-ec / ark-bn254 (third-party, absent from /root/reference); what is restated here is
// the textbook construction: Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi), xi = 9 + u; D-type twist
// (x, y) -> (x w^2, y w^3); Miller loop over 6u + 2 with affine line functions, the two Frobenius additions, and the
// final exponentiation (p^12 - 1)/r = (p^6 - 1)(p^2 + 1) * (p^4 - p^2 + 1)/r.  Verification is one scalar operation per
// proof -- latency, not throughput -- so the code favours obviousness: affine steps, no sparse multiplications.
// Everything is host+device so that tests/host/pairing_host_test.cpp checks it against the oracle without a GPU.
#pragma once
#include "ec.cuh"

namespace b200zk {

B2_HD Fq2 fq2_mul_xi(const Fq2& a) {          // (a0 + a1 u)(9 + u) = (9 a0 - a1) + (9 a1 + a0) u
    Fq t0 = Fq::dbl(Fq::dbl(Fq::dbl(a.c0))), t1 = Fq::dbl(Fq::dbl(Fq::dbl(a.c1)));
    Fq2 r;
    r.c0 = Fq::sub(Fq::add(t0, a.c0), a.c1);
    r.c1 = Fq::add(Fq::add(t1, a.c1), a.c0);
    return r;
}
B2_HD Fq2 fq2_conj(const Fq2& a) { Fq2 r; r.c0 = a.c0; r.c1 = Fq::neg(a.c1); return r; }
B2_HD Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) { Fq2 r; r.c0 = Fq::mul_ni(a.c0, k); r.c1 = Fq::mul_ni(a.c1, k); return r; }

struct Fq6 {
    Fq2 a, b, c;                               // a + b v + c v^2
    B2_HD static Fq6 zero() { Fq6 r; r.a = Fq2::zero(); r.b = Fq2::zero(); r.c = Fq2::zero(); return r; }
    B2_HD static Fq6 one() { Fq6 r = zero(); r.a = Fq2::one(); return r; }
    B2_HD bool operator==(const Fq6& o) const { return a == o.a && b == o.b && c == o.c; }
    B2_HD static Fq6 add(const Fq6& x, const Fq6& y) { Fq6 r; r.a = Fq2::add(x.a, y.a); r.b = Fq2::add(x.b, y.b); r.c = Fq2::add(x.c, y.c); return r; }
    B2_HD static Fq6 sub(const Fq6& x, const Fq6& y) { Fq6 r; r.a = Fq2::sub(x.a, y.a); r.b = Fq2::sub(x.b, y.b); r.c = Fq2::sub(x.c, y.c); return r; }
    B2_HD static Fq6 neg(const Fq6& x) { Fq6 r; r.a = Fq2::neg(x.a); r.b = Fq2::neg(x.b); r.c = Fq2::neg(x.c); return r; }
    B2_HD static Fq6 mul_v(const Fq6& x) { Fq6 r; r.a = fq2_mul_xi(x.c); r.b = x.a; r.c = x.b; return r; }
    B2_HD_NI static Fq6 mul(const Fq6& x, const Fq6& y) {
        Fq2 t0 = Fq2::mul(x.a, y.a), t1 = Fq2::mul(x.b, y.b), t2 = Fq2::mul(x.c, y.c);
        Fq6 r;
        r.a = Fq2::add(t0, fq2_mul_xi(Fq2::sub(Fq2::sub(Fq2::mul(Fq2::add(x.b, x.c), Fq2::add(y.b, y.c)), t1), t2)));
        r.b = Fq2::add(Fq2::sub(Fq2::sub(Fq2::mul(Fq2::add(x.a, x.b), Fq2::add(y.a, y.b)), t0), t1), fq2_mul_xi(t2));
        r.c = Fq2::add(Fq2::sub(Fq2::sub(Fq2::mul(Fq2::add(x.a, x.c), Fq2::add(y.a, y.c)), t0), t2), t1);
        return r;
    }
    B2_HD_NI static Fq6 inv(const Fq6& x) {
        Fq2 A = Fq2::sub(Fq2::sqr(x.a), fq2_mul_xi(Fq2::mul(x.b, x.c)));
        Fq2 B = Fq2::sub(fq2_mul_xi(Fq2::sqr(x.c)), Fq2::mul(x.a, x.b));
        Fq2 C = Fq2::sub(Fq2::sqr(x.b), Fq2::mul(x.a, x.c));
        Fq2 F = Fq2::add(Fq2::mul(x.a, A), fq2_mul_xi(Fq2::add(Fq2::mul(x.c, B), Fq2::mul(x.b, C))));
        Fq2 Fi = Fq2::inv(F);
        Fq6 r; r.a = Fq2::mul(A, Fi); r.b = Fq2::mul(B, Fi); r.c = Fq2::mul(C, Fi);
        return r;
    }
};

struct Fq12 {
    Fq6 c0, c1;                                // c0 + c1 w;  as a polynomial in w: c0 = (w^0, w^2, w^4), c1 = (w^1, w^3, w^5)
    B2_HD static Fq12 one() { Fq12 r; r.c0 = Fq6::one(); r.c1 = Fq6::zero(); return r; }
    B2_HD bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    B2_HD_NI static Fq12 mul(const Fq12& x, const Fq12& y) {
        Fq6 t0 = Fq6::mul(x.c0, y.c0), t1 = Fq6::mul(x.c1, y.c1);
        Fq12 r;
        r.c1 = Fq6::sub(Fq6::sub(Fq6::mul(Fq6::add(x.c0, x.c1), Fq6::add(y.c0, y.c1)), t0), t1);
        r.c0 = Fq6::add(t0, Fq6::mul_v(t1));
        return r;
    }
    B2_HD static Fq12 conj(const Fq12& x) { Fq12 r; r.c0 = x.c0; r.c1 = Fq6::neg(x.c1); return r; }      // x^(p^6)
    B2_HD_NI static Fq12 inv(const Fq12& x) {
        Fq6 t = Fq6::inv(Fq6::sub(Fq6::mul(x.c0, x.c0), Fq6::mul_v(Fq6::mul(x.c1, x.c1))));
        Fq12 r; r.c0 = Fq6::mul(x.c0, t); r.c1 = Fq6::neg(Fq6::mul(x.c1, t));
        return r;
    }
    // x^(p^2): the Fq2 coefficients are fixed, w^i picks up xi^(i (p^2 - 1)/6) (an element of Fq)
    B2_HD_NI static Fq12 frob2(const Fq12& x) {
        Fq g[6];
        g[0] = Fq::one();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            g[1].l[i] = PairingConst::frob2_1(i); g[2].l[i] = PairingConst::frob2_2(i); g[3].l[i] = PairingConst::frob2_3(i);
            g[4].l[i] = PairingConst::frob2_4(i); g[5].l[i] = PairingConst::frob2_5(i);
        }
        Fq12 r;
        r.c0.a = x.c0.a;                     r.c1.a = fq2_mul_fq(x.c1.a, g[1]);
        r.c0.b = fq2_mul_fq(x.c0.b, g[2]);   r.c1.b = fq2_mul_fq(x.c1.b, g[3]);
        r.c0.c = fq2_mul_fq(x.c0.c, g[4]);   r.c1.c = fq2_mul_fq(x.c1.c, g[5]);
        return r;
    }
};

// line through the untwisted T with slope lambda (on the twist), evaluated at P = (xP, yP) in G1:
//   yP - lambda xP w + (lambda xT - yT) w^3
B2_HD Fq12 pairing_line(const Fq2& lambda, const Fq2& xT, const Fq2& yT, const Fq& xP, const Fq& yP) {
    Fq12 l;
    l.c0 = Fq6::zero(); l.c1 = Fq6::zero();
    l.c0.a.c0 = yP;
    l.c1.a = Fq2::neg(fq2_mul_fq(lambda, xP));
    l.c1.b = Fq2::sub(Fq2::mul(lambda, xT), yT);
    return l;
}

B2_HD affine_t<Fq2> g2_frobenius_twist(const affine_t<Fq2>& q) {
    Fq2 gx, gy;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        gx.c0.l[i] = PairingConst::tw_x_c0(i); gx.c1.l[i] = PairingConst::tw_x_c1(i);
        gy.c0.l[i] = PairingConst::tw_y_c0(i); gy.c1.l[i] = PairingConst::tw_y_c1(i);
    }
    affine_t<Fq2> r;
    r.x = Fq2::mul(fq2_conj(q.x), gx);
    r.y = Fq2::mul(fq2_conj(q.y), gy);
    return r;
}

// f_{6u+2,Q}(P) * l_{T,pi(Q)}(P) * l_{T+pi(Q),-pi^2(Q)}(P); 1 when either point is the identity
B2_HD_NI Fq12 miller_loop(const affine_t<Fq>& P, const affine_t<Fq2>& Q) {
    Fq12 f = Fq12::one();
    if (P.is_inf() || Q.is_inf()) return f;
    Fq2 tx = Q.x, ty = Q.y;
    auto add_step = [&](const Fq2& qx, const Fq2& qy) {
        Fq2 lam = Fq2::mul(Fq2::sub(qy, ty), Fq2::inv(Fq2::sub(qx, tx)));
        Fq12 l = pairing_line(lam, tx, ty, P.x, P.y);
        Fq2 x3 = Fq2::sub(Fq2::sub(Fq2::sqr(lam), tx), qx);
        Fq2 y3 = Fq2::sub(Fq2::mul(lam, Fq2::sub(tx, x3)), ty);
        tx = x3; ty = y3;
        f = Fq12::mul(f, l);
    };
    // 6u + 2 = 2^64 + ATE_LOOP_LO: the leading bit is consumed by T = Q, f = 1
    for (int bit = 63; bit >= 0; --bit) {
        Fq2 x2 = Fq2::sqr(tx);
        Fq2 lam = Fq2::mul(Fq2::add(Fq2::dbl(x2), x2), Fq2::inv(Fq2::dbl(ty)));
        Fq12 l = pairing_line(lam, tx, ty, P.x, P.y);
        Fq2 x3 = Fq2::sub(Fq2::sqr(lam), Fq2::dbl(tx));
        Fq2 y3 = Fq2::sub(Fq2::mul(lam, Fq2::sub(tx, x3)), ty);
        tx = x3; ty = y3;
        f = Fq12::mul(Fq12::mul(f, f), l);
        if ((PairingConst::ATE_LOOP_LO >> bit) & 1) add_step(Q.x, Q.y);
    }
    affine_t<Fq2> q1 = g2_frobenius_twist(Q);
    affine_t<Fq2> q2 = g2_frobenius_twist(q1);
    add_step(q1.x, q1.y);
    add_step(q2.x, Fq2::neg(q2.y));
    return f;
}

B2_HD_NI Fq12 final_exponentiation(const Fq12& f) {
    Fq12 t = Fq12::mul(Fq12::conj(f), Fq12::inv(f));          // f^(p^6 - 1)
    t = Fq12::mul(Fq12::frob2(t), t);                         // ^(p^2 + 1)
    Fq12 res = Fq12::one();                                   // ^((p^4 - p^2 + 1)/r), plain square-and-multiply
    for (int bit = PairingConst::HARD_EXP_BITS - 1; bit >= 0; --bit) {
        res = Fq12::mul(res, res);
        if ((PairingConst::hard_exp(bit >> 5) >> (bit & 31)) & 1) res = Fq12::mul(res, t);
    }
    return res;
}

}  // namespace b200zk
