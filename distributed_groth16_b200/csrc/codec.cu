// codec.cu -- batch (de)compression of BN254 points in ark-serialize `Compress::Yes` form.
//
// The reference moves every proving / verifying key and proof through `CanonicalSerialize` with compression
// (/root/reference/common/src/utils/serializer.rs:20-49; `proving_key.bin` written at mpc-api/src/main.rs:161-165,
// proofs at zk-cli/src/main.rs:130-136).  Encoding (pinned on zk-cli/test-circuits/sha256/proof.bin):
//   G1: x as 32 little-endian bytes; top two bits of the last byte are flags: 0x80 = y is the larger of (y, -y),
//       0x40 = point at infinity (all other bits zero).
//   G2: x.c0 then x.c1 (32 bytes each), flags in the last byte of x.c1; Fq2 order compares c1 first, then c0.
// Decompression needs one square root per point (p = 3 mod 4: a^((p+1)/4), 253 squarings + ~127 products), which
// for a 2^20-constraint key is ~5 * 2^20 independent exponentiations: a data-parallel kernel, one thread per point.
#include "common.cuh"

namespace b200zk {

// a^((p+1)/4) in Fq
__device__ Fq fq_pow_p1_4(const Fq& a) {
    // e = (p + 1) / 4, from the modulus limbs
    uint32_t e[8];
    uint32_t carry = 1;
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)FqParams::mod(i) + carry;
        e[i] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
    for (int i = 0; i < 8; ++i) e[i] = (e[i] >> 2) | (i < 7 ? e[i + 1] << 30 : 0);
    Fq res = Fq::one();
    for (int i = 253; i >= 0; --i) {
        res = Fq::mul_ni(res, res);
        if ((e[i >> 5] >> (i & 31)) & 1) res = Fq::mul_ni(res, a);
    }
    return res;
}

__device__ bool fq_sqrt(const Fq& a, Fq* out) {
    Fq s = fq_pow_p1_4(a);
    *out = s;
    return Fq::mul_ni(s, s) == a;
}

// x / 2 (works on Montgomery representatives as on plain ones: the map is linear)
__device__ Fq fq_half(const Fq& a) {
    uint32_t t[9];
    for (int i = 0; i < 8; ++i) t[i] = a.l[i];
    t[8] = 0;
    if (a.l[0] & 1) {
        uint32_t carry = 0;
        for (int i = 0; i < 8; ++i) {
            uint64_t s = (uint64_t)t[i] + FqParams::mod(i) + carry;
            t[i] = (uint32_t)s;
            carry = (uint32_t)(s >> 32);
        }
        t[8] = carry;
    }
    Fq r;
    for (int i = 0; i < 8; ++i) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
    return r;
}

// square root in Fq2 = Fq[u]/(u^2+1), "complex method": sqrt(a0 + a1 u) = x0 + x1 u with
// x0^2 = (a0 +- |a|) / 2, x1 = a1 / (2 x0), |a| = sqrt(a0^2 + a1^2)
__device__ bool fq2_sqrt(const Fq2& a, Fq2* out) {
    if (a.is_zero()) { *out = Fq2::zero(); return true; }
    if (a.c1.is_zero()) {
        Fq s;
        if (fq_sqrt(a.c0, &s)) { out->c0 = s; out->c1 = Fq::zero(); return true; }
        if (!fq_sqrt(Fq::neg(a.c0), &s)) return false;       // cannot happen: -1 is a non-residue
        out->c0 = Fq::zero(); out->c1 = s;
        return true;
    }
    Fq norm = Fq::add(Fq::mul_ni(a.c0, a.c0), Fq::mul_ni(a.c1, a.c1));
    Fq alpha;
    if (!fq_sqrt(norm, &alpha)) return false;
    Fq delta = fq_half(Fq::add(a.c0, alpha));
    Fq x0;
    if (!fq_sqrt(delta, &x0)) {
        delta = fq_half(Fq::sub(a.c0, alpha));
        if (!fq_sqrt(delta, &x0)) return false;
    }
    Fq x1 = Fq::mul_ni(a.c1, Fq::inv(Fq::dbl(x0)));
    out->c0 = x0; out->c1 = x1;
    return Fq2::sqr(*out) == a;
}

__device__ bool fq_is_larger(const Fq& y) {       // arkworks: y > -y as canonical integers
    Fq a = Fq::from_mont(y), b = Fq::from_mont(Fq::neg(y));
    for (int i = 7; i >= 0; --i) {
        if (a.l[i] > b.l[i]) return true;
        if (a.l[i] < b.l[i]) return false;
    }
    return false;
}
__device__ bool fq2_is_larger(const Fq2& y) { return y.c1.is_zero() ? fq_is_larger(y.c0) : fq_is_larger(y.c1); }

// 32 little-endian bytes (flags already masked off) -> canonical limbs; false when >= p
__device__ bool fq_from_bytes(const uint8_t* in, uint8_t top_mask, Fq* out) {
    Fq x;
    for (int i = 0; i < 8; ++i) {
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) {
            uint32_t byte = in[4 * i + b];
            if (4 * i + b == 31) byte &= top_mask;
            w |= byte << (8 * b);
        }
        x.l[i] = w;
    }
    bool lt = false;                                   // x < p ?
    for (int i = 7; i >= 0; --i) {
        if (x.l[i] < FqParams::mod(i)) { lt = true; break; }
        if (x.l[i] > FqParams::mod(i)) break;
    }
    *out = Fq::to_mont(x);
    return lt;
}
__device__ void fq_to_bytes(const Fq& xm, uint8_t* out) {
    Fq x = Fq::from_mont(xm);
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(x.l[i >> 2] >> (8 * (i & 3)));
}

__global__ void __launch_bounds__(128) k_g1_compress(const affine_t<Fq>* in, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t<Fq> p = in[i];
    __align__(16) uint8_t b[32];
    if (p.is_inf()) {
        for (int k = 0; k < 32; ++k) b[k] = 0;
        b[31] = 0x40;
    } else {
        fq_to_bytes(p.x, b);
        if (fq_is_larger(p.y)) b[31] |= 0x80;
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 32);
    o[0] = *reinterpret_cast<uint4*>(b);
    o[1] = *reinterpret_cast<uint4*>(b + 16);
}

__global__ void __launch_bounds__(128) k_g2_compress(const affine_t<Fq2>* in, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t<Fq2> p = in[i];
    __align__(16) uint8_t b[64];
    if (p.is_inf()) {
        for (int k = 0; k < 64; ++k) b[k] = 0;
        b[63] = 0x40;
    } else {
        fq_to_bytes(p.x.c0, b);
        fq_to_bytes(p.x.c1, b + 32);
        if (fq2_is_larger(p.y)) b[63] |= 0x80;
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 64);
    for (int k = 0; k < 4; ++k) o[k] = *reinterpret_cast<uint4*>(b + 16 * k);
}

// bad[0] += 1 for every encoding that is not a point (x >= p, x not on the curve, infinity flag with other bits set,
// or -- G2 with check_subgroup -- a twist point outside the order-r subgroup); its slot is left as infinity.
__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* in, uint32_t n, affine_t<Fq>* out, uint32_t* bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 32;
    affine_t<Fq> p = affine_t<Fq>::infinity();
    uint8_t flags = b[31] & 0xC0;
    bool ok = true;
    if (flags & 0x40) {
        for (int k = 0; k < 32; ++k) if ((k == 31 ? (b[k] & 0x3F) : b[k]) != 0) ok = false;
        if (flags & 0x80) ok = false;
    } else {
        Fq x;
        ok = fq_from_bytes(b, 0x3F, &x);
        if (ok) {
            Fq bb;
            for (int k = 0; k < 8; ++k) bb.l[k] = CurveConst::g1_b(k);
            Fq y2 = Fq::add(Fq::mul_ni(Fq::mul_ni(x, x), x), bb), y;
            ok = fq_sqrt(y2, &y);
            if (ok) {
                if (fq_is_larger(y) != ((flags & 0x80) != 0)) y = Fq::neg(y);
                p.x = x; p.y = y;
            }
        }
    }
    if (!ok) { atomicAdd(bad, 1u); p = affine_t<Fq>::infinity(); }
    out[i] = p;
}

__global__ void __launch_bounds__(128) k_g2_decompress(const uint8_t* in, uint32_t n, int check_subgroup, affine_t<Fq2>* out,
                                                       uint32_t* bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 64;
    affine_t<Fq2> p = affine_t<Fq2>::infinity();
    uint8_t flags = b[63] & 0xC0;
    bool ok = true;
    if (flags & 0x40) {
        for (int k = 0; k < 64; ++k) if ((k == 63 ? (b[k] & 0x3F) : b[k]) != 0) ok = false;
        if (flags & 0x80) ok = false;
    } else {
        Fq2 x;
        ok = fq_from_bytes(b, 0xFF, &x.c0);
        ok = fq_from_bytes(b + 32, 0x3F, &x.c1) && ok;
        if (ok) {
            Fq2 bb;
            for (int k = 0; k < 8; ++k) { bb.c0.l[k] = CurveConst::g2_b_c0(k); bb.c1.l[k] = CurveConst::g2_b_c1(k); }
            Fq2 y2 = Fq2::add(Fq2::mul(Fq2::sqr(x), x), bb), y;
            ok = fq2_sqrt(y2, &y);
            if (ok) {
                if (fq2_is_larger(y) != ((flags & 0x80) != 0)) y = Fq2::neg(y);
                p.x = x; p.y = y;
                if (check_subgroup) {                       // [r] P == O  (the twist has a large cofactor)
                    uint32_t r[8];
                    for (int k = 0; k < 8; ++k) r[k] = FrParams::mod(k);
                    ok = xyzz_t<Fq2>::mul_scalar(xyzz_t<Fq2>::from_affine(p), r).is_inf();
                }
            }
        }
    }
    if (!ok) { atomicAdd(bad, 1u); p = affine_t<Fq2>::infinity(); }
    out[i] = p;
}

int points_compress_dev(b200zk_ctx* ctx, Slot& sl, int g2, const void* d_affine, size_t n, void* d_bytes) {
    if (n == 0) return B200ZK_OK;
    if (n >= (1ull << 32)) return set_error(ctx, B200ZK_ERR_ARG, "too many points");
    {
        LaunchScope ls(ctx, sl.stream, "points_compress");
        unsigned grid = (unsigned)((n + 127) / 128);
        if (g2) k_g2_compress<<<grid, 128, 0, sl.stream>>>(reinterpret_cast<const affine_t<Fq2>*>(d_affine), (uint32_t)n, (uint8_t*)d_bytes);
        else k_g1_compress<<<grid, 128, 0, sl.stream>>>(reinterpret_cast<const affine_t<Fq>*>(d_affine), (uint32_t)n, (uint8_t*)d_bytes);
    }
    return check_launch(ctx, "k_compress");
}

int points_decompress_dev(b200zk_ctx* ctx, Slot& sl, int g2, const void* d_bytes, size_t n, int check_subgroup, void* d_affine,
                          size_t* n_invalid) {
    if (n_invalid) *n_invalid = 0;
    if (n == 0) return B200ZK_OK;
    if (n >= (1ull << 32)) return set_error(ctx, B200ZK_ERR_ARG, "too many points");
    B2_CUDA_OK(ctx, sl.small.reserve(1024));
    uint32_t* bad = reinterpret_cast<uint32_t*>(sl.small.p);
    B2_CUDA_OK(ctx, cudaMemsetAsync(bad, 0, 4, sl.stream));
    {
        LaunchScope ls(ctx, sl.stream, "points_decompress");
        unsigned grid = (unsigned)((n + 127) / 128);
        if (g2) k_g2_decompress<<<grid, 128, 0, sl.stream>>>((const uint8_t*)d_bytes, (uint32_t)n, check_subgroup,
                                                            reinterpret_cast<affine_t<Fq2>*>(d_affine), bad);
        else k_g1_decompress<<<grid, 128, 0, sl.stream>>>((const uint8_t*)d_bytes, (uint32_t)n, reinterpret_cast<affine_t<Fq>*>(d_affine), bad);
    }
    B2_TRY(check_launch(ctx, "k_decompress"));
    uint32_t h_bad = 0;
    B2_CUDA_OK(ctx, cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, sl.stream));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(sl.stream));
    if (n_invalid) *n_invalid = h_bad;
    if (h_bad) {
        char msg[96];
        snprintf(msg, sizeof(msg), "%u of %zu encodings are not valid curve points", h_bad, n);
        return set_error(ctx, B200ZK_ERR_ARG, msg);
    }
    return B200ZK_OK;
}

}  // namespace b200zk
