// fp29.cuh -- EXPERIMENT (not part of the product library): carry-free 9 x 29-bit-limb Montgomery arithmetic.
//
// Outcome on B200 (profiles/r2_microbench.txt): correct (bit-identical to fp.cuh through the domain changes, host +
// device) but SLOWER: 52 G products/s against 66-69 G/s for the 8 x 32-bit-limb product of fp.cuh.  The premise below
// was wrong: round 1's "IMAD.WIDE.U32 without carry = 60 lanes/clk" probe had its multiplies hoisted out of the loop
// by ptxas (the loop measured IADD3).  Every 32x32->64 multiply-add issues at 32 lanes/clk/SM whatever its addend,
// so 171 of them (9 x 9 x 2 + 9) lose to 128 + 16.  Kept for the record and for tools/microbench.cu.
//
// Why a second representation.  tools/microbench.cu on B200 (profiles/r1_microbench_pipes.txt): a wide multiply-add
// that consumes or produces a carry (IMAD.WIDE.U32.X, what a saturated 8 x 32-bit-limb product is made of, fp.cuh)
// issues at 32 lanes/clk/SM, the plain 64-bit-accumulating IMAD.WIDE.U32 at 60.  With 29-bit limbs the 64-bit column
// sums of a 9 x 9 product *and* of its Montgomery reduction never overflow (18 terms < 2^58 each, + carries), so the
// whole product is 2 x 81 carry-less IMAD.WIDE on the FMA pipe; the limb carries are shifts / adds on the otherwise
// idle ALU pipe.  162 full-rate instead of 128 half-rate multiply-adds: ~1.5x fewer pipe cycles per product.
//
// Representation: value = sum l[k] 2^(29 k), k < 9, limbs "loose" (see the bounds on each function), value NOT
// canonical.  Montgomery radix here is 2^261 (= 9 x 29): mul(a, b) = a b 2^-261 mod p.
// Domains: memory (and the C ABI) keeps arkworks' R = 2^256 form, x~ = x 2^256 mod p, canonical, 8 x u32.  Since
// x 2^261 = x~ 2^5, a *shift by 5 while re-limbing* (from_mont256) turns a stored element into a valid (non-reduced,
// < 32 p) residue of the 2^261 domain at the cost of a few ALU shifts -- no multiplication.  Going back
// (to_mont256) is one product with the integer 2^256 (a 2^261-domain value times a 2^256-domain value is a
// 2^256-domain value: mixed products "drop" one 2^261), then one conditional subtraction of p.
//
// Replaces arkworks' `Fp256<MontBackend>` products inside `G::msm` (dist-primitives/src/dmsm/mod.rs:82) and the
// butterflies of dist-primitives/src/dfft/mod.rs:128-131, like fp.cuh; the results that leave a kernel are the same
// canonical 8 x u32 Montgomery words.  Plain C++ on purpose (nvcc turns `t += (u64)a * b` into IMAD.WIDE.U32 with a
// 64-bit addend), so the identical code is unit-tested on the host (tests/host/fp_host_test.cpp).
#pragma once
#include "../../distributed_groth16_b200/csrc/fp.cuh"

namespace b200zk {

template <class P>
struct Fp29 {
    uint32_t l[9];
    static constexpr uint32_t MASK = (1u << 29) - 1;
    static constexpr uint32_t INV29 = P::INV & MASK;      // -p^-1 mod 2^29

    // 29-bit limb k of p
    B2_HD static constexpr uint32_t modl(int k) {
        const int w = (29 * k) >> 5, s = (29 * k) & 31;
        const uint64_t v = (uint64_t)P::mod(w) | (w + 1 < 8 ? (uint64_t)P::mod(w + 1) << 32 : 0ull);
        return (uint32_t)(v >> s) & MASK;
    }

    // acc + a * b: nvcc turns this C++ form into ONE IMAD.WIDE.U32 with a 64-bit addend.  (Spelling it as PTX
    // mad.wide.u32 makes ptxas split it into IMAD.WIDE + a three-input 64-bit add again, which moves the bottleneck
    // to the ALU pipe: 410 instead of 240 instructions per product, see profiles/r2_sass_hist.md.)
    B2_HD static uint64_t madw(uint64_t acc, uint32_t a, uint32_t b) { return acc + (uint64_t)a * b; }
    // a 32-bit value the front end must keep 32 bits wide: without it LLVM widens the Montgomery factor m to 64 bits
    // and m * p_j becomes a 64 x 64-bit multiply (extra high-word terms after every multiply-add)
    B2_HD static uint32_t opaque(uint32_t c) {
#ifdef __CUDA_ARCH__
        uint32_t r;
        asm("mov.u32 %0, %1;" : "=r"(r) : "r"(c));
        return r;
#else
        return c;
#endif
    }

    B2_HD static Fp29 zero() { Fp29 r; for (int i = 0; i < 9; ++i) r.l[i] = 0; return r; }

    // ---- domain changes ---------------------------------------------------------------------------------------
    // stored 2^256-form element (canonical, < p) -> 2^261-domain residue x~ 2^5 < 32 p, limbs < 2^29
    B2_HD static Fp29 from_mont256(const Fp<P>& a) {
        Fp29 r;
        r.l[0] = (a.l[0] << 5) & MASK;
#pragma unroll
        for (int k = 1; k < 9; ++k) {
            const int off = 29 * k - 5, w = off >> 5, s = off & 31;      // bits [off, off + 29) of a
            uint32_t lo = a.l[w] >> s;
            uint32_t hi = (s > 3 && w + 1 < 8) ? a.l[w + 1] << (32 - s) : 0u;
            r.l[k] = (lo | hi) & MASK;
        }
        return r;
    }
    // same integer, plain re-limbing (no shift): a 2^256-domain multiplier for mixed products (twiddles, constants)
    B2_HD static Fp29 relimb(const Fp<P>& a) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int off = 29 * k, w = off >> 5, s = off & 31;
            uint32_t lo = a.l[w] >> s;
            uint32_t hi = (s > 3 && w + 1 < 8) ? a.l[w + 1] << (32 - s) : 0u;
            r.l[k] = (lo | hi) & MASK;
        }
        return r;
    }
    // normalised limbs (< 2^29), value < 2^256 -> 8 x u32 words of the same integer
    B2_HD static void pack(uint32_t t[8], const Fp29& a) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int k = (32 * w) / 29, s = 32 * w - 29 * k;            // word w starts at bit s of limb k
            uint32_t v = a.l[k] >> s;
            v |= a.l[k + 1] << (29 - s);
            if (58 - s < 32 && k + 2 < 9) v |= a.l[k + 2] << (58 - s);
            t[w] = v;
        }
    }
    // 2^261-domain residue (value < 32 p, limbs < 2^30) -> canonical stored 2^256 form
    B2_HD static Fp<P> to_mont256(const Fp29& a) {
        Fp29 c = zero();
        c.l[8] = opaque(1u << 24);                                       // the integer 2^256 (only 9 of the 81 a*b terms survive)
        Fp29 m = mul(a, c);                                              // a 2^-5 mod p, < 2 p
        uint32_t t[8];
        pack(t, m);
        Fp<P> r; Fp<P>::final_sub(r, t); return r;
    }
    // value already in the 2^256 domain (e.g. after a mixed product), < 2 p, normalised
    B2_HD static Fp<P> canon256(const Fp29& a) {
        uint32_t t[8];
        pack(t, a);
        Fp<P> r; Fp<P>::final_sub(r, t); return r;
    }

    // ---- Montgomery product -----------------------------------------------------------------------------------
    // a b 2^-261 mod p.  Requires 9 max(a.l) max(b.l) + 9 * 2^58 + 2^36 < 2^64 (e.g. both limbs < 2^30, or < 2^31
    // against < 2^29) and a b < 2^261 p' for the output bound: result < p + a b / 2^261, limbs < 2^29 (top < 2^24).
    B2_HD static Fp29 mul(const Fp29& a, const Fp29& b) {
        uint64_t t[18];
#pragma unroll
        for (int i = 0; i < 18; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
#pragma unroll
            for (int j = 0; j < 9; ++j) t[i + j] = madw(t[i + j], a.l[j], b.l[i]);
            reduce_row(t, i);
        }
        return collect(t);
    }
    // one Montgomery step on column i: t += m p 2^(29 i) with m = -t_i / p mod 2^29, then push column i's carry up
    B2_HD static void reduce_row(uint64_t* t, int i) {
        uint32_t m = opaque(((uint32_t)t[i] * INV29) & MASK);
#pragma unroll
        for (int j = 0; j < 9; ++j) t[i + j] = madw(t[i + j], m, modl(j));
        t[i + 1] += t[i] >> 29;
    }
    // limbs of the reduced product from columns 9..17
    B2_HD static Fp29 collect(uint64_t* t) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r.l[k] = opaque((uint32_t)t[9 + k] & MASK);      // opaque: keeps the limb 32 bits wide for the next product (see opaque())
            t[10 + k] += t[9 + k] >> 29;
        }
        r.l[8] = opaque((uint32_t)t[17]);
        return r;
    }
    // a^2 2^-261: 45 products instead of 81 (cross terms once, against the doubled operand).  a.l < 2^30.
    B2_HD static Fp29 sqr(const Fp29& a) {
        uint64_t t[18];
        uint32_t d[9];
#pragma unroll
        for (int i = 0; i < 18; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) d[i] = a.l[i] << 1;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            t[2 * i] = madw(t[2 * i], a.l[i], a.l[i]);
#pragma unroll
            for (int j = i + 1; j < 9; ++j) t[i + j] = madw(t[i + j], d[i], a.l[j]);
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) reduce_row(t, i);
        return collect(t);
    }

    // ---- additive operations (lazy) ----------------------------------------------------------------------------
    // limb-wise sum, no carry propagation: limb bounds add up
    B2_HD static Fp29 add_lazy(const Fp29& a, const Fp29& b) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 9; ++k) r.l[k] = a.l[k] + b.l[k];
        return r;
    }
    B2_HD static Fp29 dbl_lazy(const Fp29& a) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 9; ++k) r.l[k] = a.l[k] << 1;
        return r;
    }
    // carry propagation: limbs 0..7 < 2^29 again, the top limb keeps the excess; value unchanged.  Input limbs < 2^32 - 8.
    B2_HD static Fp29 norm(const Fp29& a) {
        Fp29 r;
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t v = a.l[k] + c;
            r.l[k] = v & MASK;
            c = v >> 29;
        }
        r.l[8] = a.l[8] + c;
        return r;
    }
    // limb k of the integer K p (K <= 64), top limb unreduced
    B2_HD static constexpr uint32_t kp_limb(int K, int k) {
        uint64_t carry = 0;
        uint32_t v = 0;
        for (int i = 0; i <= k; ++i) {
            uint64_t x = (uint64_t)modl(i) * (uint64_t)K + carry;
            v = i == 8 ? (uint32_t)x : (uint32_t)(x & MASK);
            carry = x >> 29;
        }
        return v;
    }
    // K p re-written with every limb below the top raised by 2^S (and the borrow repaid one limb up), so that
    // a.l[k] + bias(k) - b.l[k] never goes negative for b.l[k] <= 2^S - 2^(S-29): a borrow-free limb-wise a - b + K p.
    // S = 29 for a normalised b, 31 for a lazy sum of up to four normalised values.
    template <int K, int S>
    B2_HD static constexpr uint32_t bias(int k) {
        return k == 0 ? kp_limb(K, 0) + (1u << S) : (k < 8 ? kp_limb(K, k) + (1u << S) - (1u << (S - 29)) : kp_limb(K, 8) - (1u << (S - 29)));
    }
    // a - b + K p limb-wise.  Needs value(b) + 2^(232 + S - 29) <= K p for the top limb.  Result limbs < max a.l + 2^S +
    // 2^29: norm() it before it becomes a product operand.
    template <int K, int S>
    B2_HD static Fp29 sub_lazy(const Fp29& a, const Fp29& b) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 9; ++k) r.l[k] = a.l[k] + bias<K, S>(k) - b.l[k];
        return r;
    }
    template <int K>
    B2_HD static Fp29 sub(const Fp29& a, const Fp29& b) { return norm(sub_lazy<K, 29>(a, b)); }
    // K p - b limb-wise (lazy)
    template <int K, int S>
    B2_HD static Fp29 neg_lazy(const Fp29& b) {
        Fp29 r;
#pragma unroll
        for (int k = 0; k < 9; ++k) r.l[k] = bias<K, S>(k) - b.l[k];
        return r;
    }

    B2_HD bool limbs_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) o |= l[k];
        return o == 0;
    }
    // a product output (normalised, < 2 p) is 0 mod p iff it is the integer 0 or p; the low limb filters first
    B2_HD bool is_zero_mod_2p() const {
        if (l[0] != 0 && l[0] != modl(0)) return false;
        uint32_t z = 0, e = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { z |= l[k]; e |= l[k] ^ modl(k); }
        return z == 0 || e == 0;
    }
    // general (slow, rare paths): normalised value < (KMAX + 1) p
    B2_HD_NI static bool is_zero_mod_slow(const Fp29& a, int kmax) {
        uint32_t kp[9];
        for (int i = 0; i < 9; ++i) kp[i] = 0;
        for (int k = 0; k <= kmax; ++k) {
            uint32_t diff = 0;
            for (int i = 0; i < 9; ++i) diff |= a.l[i] ^ kp[i];
            if (diff == 0) return true;
            uint32_t c = 0;
            for (int i = 0; i < 9; ++i) {
                uint32_t v = kp[i] + modl_rt(i) + c;
                if (i < 8) { kp[i] = v & MASK; c = v >> 29; } else kp[i] = v;
            }
        }
        return false;
    }
    B2_HD static uint32_t modl_rt(int k) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) if (i == k) v = modl(i);
        return v;
    }
    // the 2^261-domain residue of 1, canonical (< p)
    B2_HD static Fp29 one() { Fp29 r; for (int i = 0; i < 9; ++i) r.l[i] = opaque(P::one261(i)); return r; }
};

typedef Fp29<FqParams> Fq29;
typedef Fp29<FrParams> Fr29;

}  // namespace b200zk
