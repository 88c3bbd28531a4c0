// setup.cu -- kernels for a circuit-specific Groth16 setup on the device: fixed-base scalar
// multiplication of the G1 / G2 generators, power vectors, generic CSR mat-vec and the linear combinations that turn
// QAP evaluations at tau into query scalars.  Mirrors what `Groth16::circuit_specific_setup` does in the reference's
// drivers (/root/reference/groth16/examples/sha256.rs:133-137, mpc-api/src/main.rs:148-152) with the CircomReduction
// h-query of ark-circom/src/circom/qap.rs:94-110; orchestration in distributed_groth16_b200/groth16/setup.py.
#include "common.cuh"

namespace b200zk {

template <class T>
__device__ __forceinline__ T lds(const T* p) {
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
    return r;
}
template <class T>
__device__ __forceinline__ void sts(T* p, const T& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
}

template <class F> __device__ affine_t<F> generator_of();
template <> __device__ affine_t<Fq> generator_of<Fq>() {
    affine_t<Fq> g;
    for (int i = 0; i < 8; ++i) { g.x.l[i] = CurveConst::g1_gen_x(i); g.y.l[i] = CurveConst::g1_gen_y(i); }
    return g;
}
template <> __device__ affine_t<Fq2> generator_of<Fq2>() {
    affine_t<Fq2> g;
    for (int i = 0; i < 8; ++i) {
        g.x.c0.l[i] = CurveConst::g2_gen_x0(i); g.x.c1.l[i] = CurveConst::g2_gen_x1(i);
        g.y.c0.l[i] = CurveConst::g2_gen_y0(i); g.y.c1.l[i] = CurveConst::g2_gen_y1(i);
    }
    return g;
}

// table[w * 15 + (d - 1)] = d * 16^w * G,  w < 64, d in 1..15 (affine).  64 threads: thread w first walks to 16^w G.
template <class F>
__global__ void k_fixed_base_table(affine_t<F>* table) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 64) return;
    xyzz_t<F> base = xyzz_t<F>::from_affine(generator_of<F>());
    for (uint32_t k = 0; k < 4 * w; ++k) base = xyzz_t<F>::dbl(base);
    affine_t<F> b = xyzz_t<F>::to_affine(base);
    xyzz_t<F> acc = xyzz_t<F>::identity();
    for (uint32_t d = 1; d <= 15; ++d) {
        xyzz_t<F>::madd(acc, b, false);
        sts(table + w * 15 + (d - 1), xyzz_t<F>::to_affine(acc));
    }
}

// out[i] = scalars[i] * G  (scalars Montgomery); 64 mixed additions, no doublings
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_mul(const affine_t<F>* table, const Fr* scalars, size_t n, affine_t<F>* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k = Fr::from_mont(lds(scalars + i));
    xyzz_t<F> acc = xyzz_t<F>::identity();
    for (uint32_t w = 0; w < 64; ++w) {
        uint32_t d = (k.l[w >> 3] >> ((w & 7) * 4)) & 15;
        if (d) xyzz_t<F>::madd(acc, lds(table + w * 15 + (d - 1)), false);
    }
    sts(out + i, xyzz_t<F>::to_affine(acc));
}

// out[i] = scale * base^i
__global__ void k_fr_powers(const Fr* consts /* base, scale */, size_t n, Fr* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr b = consts[0], res = consts[1];
    uint64_t e = i;
    while (e) {
        if (e & 1) res = Fr::mul(res, b);
        b = Fr::sqr(b);
        e >>= 1;
    }
    sts(out + i, res);
}

// out[r] = sum_k val[k] * x[idx[k]],  k in [ptr[r], ptr[r+1])
__global__ void k_spmv(const uint32_t* ptr, const uint32_t* idx, const Fr* val, const Fr* x, size_t n_rows, Fr* out) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    Fr acc = Fr::zero();
    for (uint32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) acc = Fr::add(acc, Fr::mul(lds(val + k), lds(x + idx[k])));
    sts(out + r, acc);
}

// out[i] = (a[i] * s[0] + b[i] * s[1] + c[i] * s[2]) * s[3]
__global__ void k_fr_lincomb(const Fr* a, const Fr* b, const Fr* c, const Fr* s, size_t n, Fr* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::add(Fr::add(Fr::mul(lds(a + i), s[0]), Fr::mul(lds(b + i), s[1])), Fr::mul(lds(c + i), s[2]));
    sts(out + i, Fr::mul(v, s[3]));
}

template <class F>
static int fixed_base_impl(b200zk_ctx* ctx, Slot& sl, void** table_slot, const void* d_scalars, size_t n, void* d_out) {
    cudaStream_t st = sl.stream;
    if (!*table_slot) {
        B2_CUDA_OK(ctx, cudaMalloc(table_slot, 64 * 15 * sizeof(affine_t<F>)));
        {
            LaunchScope ls(ctx, st, "fixed_base_table");
            k_fixed_base_table<F><<<2, 32, 0, st>>>(reinterpret_cast<affine_t<F>*>(*table_slot));
        }
        B2_TRY(check_launch(ctx, "k_fixed_base_table"));
        B2_CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    if (n == 0) return B200ZK_OK;
    {
        LaunchScope ls(ctx, st, "fixed_base_mul");
        k_fixed_base_mul<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(reinterpret_cast<const affine_t<F>*>(*table_slot),
                                                                           reinterpret_cast<const Fr*>(d_scalars), n,
                                                                           reinterpret_cast<affine_t<F>*>(d_out));
    }
    return check_launch(ctx, "k_fixed_base_mul");
}

int fixed_base_mul_dev(b200zk_ctx* ctx, Slot& sl, int g2, const void* d_scalars, size_t n, void* d_out) {
    std::lock_guard<std::mutex> g(ctx->plan_mu);
    return g2 ? fixed_base_impl<Fq2>(ctx, sl, &ctx->fb_table_g2, d_scalars, n, d_out)
              : fixed_base_impl<Fq>(ctx, sl, &ctx->fb_table_g1, d_scalars, n, d_out);
}

int fr_powers_dev(b200zk_ctx* ctx, Slot& sl, const uint64_t base[4], const uint64_t scale[4], size_t n, void* d_out) {
    if (n == 0) return B200ZK_OK;
    B2_CUDA_OK(ctx, sl.small.reserve(1024));
    uint64_t h[8];
    memcpy(h, base, 32); memcpy(h + 4, scale, 32);
    B2_CUDA_OK(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(sl.small.p) + 896, h, 64, cudaMemcpyHostToDevice, sl.stream));
    {
        LaunchScope ls(ctx, sl.stream, "fr_powers");
        k_fr_powers<<<(unsigned)((n + 255) / 256), 256, 0, sl.stream>>>(reinterpret_cast<const Fr*>(reinterpret_cast<char*>(sl.small.p) + 896), n,
                                                                         reinterpret_cast<Fr*>(d_out));
    }
    B2_TRY(check_launch(ctx, "k_fr_powers"));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(sl.stream));      // the staging words are reused by the next call
    return B200ZK_OK;
}

int spmv_dev(b200zk_ctx* ctx, Slot& sl, const void* ptr, const void* idx, const void* val, const void* x, size_t n_rows, void* out) {
    if (n_rows == 0) return B200ZK_OK;
    {
        LaunchScope ls(ctx, sl.stream, "spmv");
        k_spmv<<<(unsigned)((n_rows + 127) / 128), 128, 0, sl.stream>>>((const uint32_t*)ptr, (const uint32_t*)idx, (const Fr*)val,
                                                                         (const Fr*)x, n_rows, (Fr*)out);
    }
    return check_launch(ctx, "k_spmv");
}

int fr_lincomb_dev(b200zk_ctx* ctx, Slot& sl, const void* a, const void* b, const void* c, const uint64_t s[16], size_t n, void* out) {
    if (n == 0) return B200ZK_OK;
    B2_CUDA_OK(ctx, sl.small.reserve(1024));
    char* stage = reinterpret_cast<char*>(sl.small.p) + 768;
    B2_CUDA_OK(ctx, cudaMemcpyAsync(stage, s, 128, cudaMemcpyHostToDevice, sl.stream));
    {
        LaunchScope ls(ctx, sl.stream, "fr_lincomb");
        k_fr_lincomb<<<(unsigned)((n + 255) / 256), 256, 0, sl.stream>>>((const Fr*)a, (const Fr*)b, (const Fr*)c,
                                                                          reinterpret_cast<const Fr*>(stage), n, (Fr*)out);
    }
    B2_TRY(check_launch(ctx, "k_fr_lincomb"));
    B2_CUDA_OK(ctx, cudaStreamSynchronize(sl.stream));
    return B200ZK_OK;
}

}  // namespace b200zk
