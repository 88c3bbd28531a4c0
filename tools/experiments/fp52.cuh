// fp52.cuh -- EXPERIMENT (not part of the product library): 5 x 52-bit-limb Montgomery arithmetic on the FP64 pipe.
//
// Why.  tools/microbench.cu on B200: DFMA issues at 64 lanes/clk/SM -- twice the rate of the 32 x 32 -> 64-bit integer
// multiply-add every product of fp.cuh is made of (32 lanes/clk/SM, profiles/r2_microbench_29bit.txt) -- and on a pipe
// the bucket kernels leave idle.  Two DFMAs in round-towards-zero mode give the exact 104-bit product of two 52-bit
// integers held in doubles:
//     x = fma_rz(a, b, 2^104)                = 2^104 + floor(a b / 2^52) 2^52       (ulp of [2^104, 2^105) is 2^52)
//     y = fma_rz(a, b, (2^104 + 2^52) - x)   = 2^52 + (a b mod 2^52)                (the addend is exact; so is y)
// so the mantissa fields of x and y ARE the high and low halves of the product.  Their raw 64-bit patterns are added
// into integer column sums (IADD3 pairs on the ALU pipe); the exponent fields sum to constants known at compile time
// and are subtracted up front.  One 254-bit Montgomery product = 55 limb products = 110 DFMA + 70 DADD on the FP64
// pipe and ~135 integer instructions, against 136 half-rate IMAD.WIDE for the 8 x 32-bit-limb product.
//
// Representation: value = sum l[k] 2^(52 k), k < 5, limbs are non-negative integers < 2^52 stored as doubles, value
// < 8 p (NOT canonical).  Montgomery radix 2^260: mul(a, b) = a b 2^-260 mod p, result < 2 p for operands < 8 p
// ((8p)^2 / 2^260 < p because p < 2^254).
// Domains: memory and the C ABI keep arkworks' R = 2^256 form, x~ = x 2^256 mod p, canonical, 8 x u32.  x 2^260 =
// x~ 2^4: re-limbing with a shift by 4 (from_mont256) enters the 2^260 domain without a product (value < 16 p, see
// from_mont256); to_mont256 is one product with the integer 2^256 followed by the conditional subtraction.
//
// Outcome on B200 (profiles/r2_microbench_fp64.txt, tools/microbench52.cu): bit-identical to fp.cuh after 2000-step chains,
// but NOT faster: 55-61.5 G products/s alone (integer product: 65-66 G/s), and with half of the warps of an SM running
// each kind the total is 66-72 G/s, not the sum -- the FP64 and the integer multiply-adds evidently share a datapath or
// its operand bandwidth (a product costs ~590 SMSP cycles for 180 FP64 instructions, 3.3 cycles each, although a bare
// DFMA chain issues every 2).  Below the 72 G/s gate: recorded, not used.
//
// Plain C++ with two intrinsics, so the identical code runs on the host under fesetround(FE_TOWARDZERO)
// (tests/host/fp_host_test.cpp).  Technique: Emmart, Zheng, Weems, "Faster modular exponentiation using double
// precision floating point arithmetic on the GPU" (ARITH 2018) -- restated here from the idea, not from code.
#pragma once
#include <cmath>

#include "../../distributed_groth16_b200/csrc/fp.cuh"

namespace b200zk {

template <class P>
struct Fp52 {
    double l[5];
    static constexpr uint64_t M52 = (1ull << 52) - 1;
    static constexpr uint64_t K1 = 0x467ull << 52;      // raw bits of 2^104
    static constexpr uint64_t K2 = 0x433ull << 52;      // raw bits of 2^52

    // 52-bit limb k of p
    B2_HD static constexpr uint64_t modl(int k) {
        uint64_t v = 0;
        for (int b = 0; b < 52; ++b) {
            const int bit = 52 * k + b;
            if (bit < 256 && ((P::mod(bit >> 5) >> (bit & 31)) & 1u)) v |= 1ull << b;
        }
        return v;
    }
    // -p^-1 mod 2^52
    B2_HD static constexpr uint64_t np0() {
        const uint64_t p0 = (uint64_t)P::mod(0) | ((uint64_t)P::mod(1) << 32);
        uint64_t y = 1;
        for (int i = 0; i < 6; ++i) y *= 2 - p0 * y;
        return (0 - y) & M52;
    }

    B2_HD static double fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
        return __fma_rz(a, b, c);
#else
        return std::fma(a, b, c);            // the host test runs under fesetround(FE_TOWARDZERO)
#endif
    }
    B2_HD static uint64_t raw(double d) {
#ifdef __CUDA_ARCH__
        return (uint64_t)__double_as_longlong(d);
#else
        uint64_t u; memcpy(&u, &d, 8); return u;
#endif
    }
    B2_HD static double from_raw(uint64_t u) {
#ifdef __CUDA_ARCH__
        return __longlong_as_double((long long)u);
#else
        double d; memcpy(&d, &u, 8); return d;
#endif
    }
    // integer < 2^52 -> double, exactly: OR the exponent of 2^52 in, subtract 2^52
    B2_HD static double to_double(uint64_t v) { return from_raw(v | K2) - 0x1p52; }
    // keeps a constant out of the constant folder / rematerialiser
    B2_HD static double opaque(double c) {
#ifdef __CUDA_ARCH__
        double r;
        asm("mov.f64 %0, %1;" : "=d"(r) : "d"(c));
        return r;
#else
        return c;
#endif
    }

    // number of (i, j), 0 <= i, j < 5, with i + j = k
    B2_HD static constexpr uint64_t cnt(int k) { return (k < 0 || k > 8) ? 0 : (uint64_t)(5 - (k < 4 ? 4 - k : k - 4)); }

    // a b 2^-260 mod p, < 2 p; limbs < 2^52 (the top one smaller).  Operands: limbs < 2^52, value < 8 p.
    B2_HD static Fp52 mul(const Fp52& a, const Fp52& b) {
        const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52;
        uint64_t t[11];
        // every column k receives, over the whole algorithm, 2 cnt(k) low halves (exponent bits K2) and 2 cnt(k - 1) high
        // halves (K1): from a x b and from the five m x p rows
#pragma unroll
        for (int k = 0; k < 11; ++k) t[k] = 0 - (2 * cnt(k) * K2 + 2 * cnt(k - 1) * K1);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const double x = fma_rz(a.l[j], b.l[i], C1);
                const double y = fma_rz(a.l[j], b.l[i], C2 - x);
                t[i + j + 1] += raw(x);
                t[i + j] += raw(y);
            }
        }
        const double n0 = (double)np0();
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            // m = (t[i] mod 2^52) (-p^-1) mod 2^52; the exponent constants are multiples of 2^52 and do not disturb the low bits
            const double lo = to_double(t[i] & M52);
            const double xm = fma_rz(lo, n0, C1);
            const double m = fma_rz(lo, n0, C2 - xm) - 0x1p52;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const double pj = (double)modl(j);
                const double x = fma_rz(m, pj, C1);
                const double y = fma_rz(m, pj, C2 - x);
                t[i + j + 1] += raw(x);
                t[i + j] += raw(y);
            }
            t[i + 1] += t[i] >> 52;             // column i is complete (all its exponent constants cancelled) and = 0 mod 2^52
        }
        Fp52 r;
#pragma unroll
        for (int k = 5; k < 9; ++k) {
            t[k + 1] += t[k] >> 52;
            r.l[k - 5] = to_double(t[k] & M52);
        }
        r.l[4] = to_double(t[9]);
        return r;
    }
    B2_HD static Fp52 sqr(const Fp52& a) { return mul(a, a); }

    // ---- domain changes (integer side) ----------------------------------------------------------------------------------
    // canonical 8 x u32 value v (any 256-bit integer) -> limbs of v 2^s, s < 8 (value must stay < 2^260)
    B2_HD static Fp52 from_words(const uint32_t* w, int s) {
        Fp52 r;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            // bits [52 k - s, 52 k - s + 52) of v
            uint64_t v = 0;
            const int lo = 52 * k - s;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int wi = (lo >> 5) + q;                      // arithmetic shift: lo may be negative for k = 0
                if (wi < 0 || wi >= 8) continue;
                const int off = 32 * wi - lo;                      // position of word wi inside the limb
                if (off >= 0) { if (off < 52) v |= (uint64_t)w[wi] << off; }
                else v |= (uint64_t)w[wi] >> (-off);
            }
            r.l[k] = to_double(v & M52);
        }
        return r;
    }
    // stored element x~ = x 2^256 (canonical) -> x 2^260 = 16 x~ < 16 p: limbs of 16 x~.  16 p < 2^258 fits; products of such
    // values: (16 p)^2 / 2^260 < 4 p, + p -> results < 5 p, still fine for the next product.  Callers that add several such
    // values first should fold them with one product by `one260` instead.
    B2_HD static Fp52 from_mont256(const Fp<P>& a) { return from_words(a.l, 4); }

    // limbs (value < 2 p) -> canonical words of value mod p
    B2_HD static void to_words_canonical(const Fp52& a, uint32_t* w) {
        uint64_t v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = raw(a.l[k] + 0x1p52) & M52;
        // subtract p once if >= p
        uint64_t d[5];
        uint64_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint64_t s = v[k] - modl(k) - borrow;
            d[k] = s & M52;
            borrow = (s >> 63) & 1;
        }
        if (!borrow) {
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] = d[k];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int lo = 32 * i, k = lo / 52, off = lo % 52;
            uint64_t x = v[k] >> off;
            if (off > 20 && k + 1 < 5) x |= v[k + 1] << (52 - off);
            w[i] = (uint32_t)x;
        }
    }
    // x 2^260 (< 8 p) -> canonical x 2^256 : one product with the integer 2^256, then canonicalise
    B2_HD static Fp<P> to_mont256(const Fp52& a) {
        Fp52 c;
#pragma unroll
        for (int k = 0; k < 5; ++k) c.l[k] = 0.0;
        c.l[4] = (double)(1ull << (256 - 208));
        Fp52 r = mul(a, c);
        Fp<P> o;
        to_words_canonical(r, o.l);
        return o;
    }
};

}  // namespace b200zk
