// tools/microbench.cu -- instruction-throughput probes used to choose the field-multiplier schedule.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "experiments/ec29.cuh"
using namespace b200zk;

#define ITERS 2000

__global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_hi(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[i]) : "r"(a + i), "r"(b));
    }
    uint64_t s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// IMAD.WIDE.U32 with RZ addend, operands fed back from the previous result (nothing loop-invariant to hoist:
// round 1's k_imad_wide above had its products hoisted by ptxas and measured the IADD3 adds instead)
__global__ void k_imad_wide_rz(uint64_t* out, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i + a;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint64_t w;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(b));
            x[i] = (uint32_t)w ^ (uint32_t)(w >> 32);
        }
    }
    uint64_t s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double a, double b) {
    double r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r[i]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 4 fused (lo.cc, hi.cc) pairs in one carry chain = 4 IMAD.WIDE.U32(.X) per "op group"
__global__ void k_imad_wide_cc(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
        asm volatile(
            "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %9, %2;\n\tmadc.hi.cc.u32 %3, %10, %9, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %9, %4;\n\tmadc.hi.cc.u32 %5, %11, %9, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %9, %6;\n\tmadc.hi.u32 %7, %12, %9, %7;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
            : "r"(a), "r"(b), "r"(a + 1), "r"(a + 2), "r"(a + 3));
    }
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
        asm volatile(
            "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %9;\n\t"
            "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %9;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %9;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
            : "r"(a), "r"(b));
    }
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mul_chain(Fq* out, int iters) {
    Fq x = Fq::one(), y = Fq::one();
    x.l[0] += threadIdx.x; y.l[1] += blockIdx.x + 3;
    for (int it = 0; it < iters; ++it) x = Fq::mul(x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_mul_chain2(Fq* out, int iters) {   // two independent chains per thread (ILP)
    Fq x = Fq::one(), y = Fq::one(), z = Fq::one();
    x.l[0] += threadIdx.x; y.l[1] += blockIdx.x + 3; z.l[2] += threadIdx.x * 7;
    for (int it = 0; it < iters; ++it) { x = Fq::mul(x, y); z = Fq::mul(z, y); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = Fq::add(x, z);
}
__global__ void __launch_bounds__(128) k_madd_chain(xyzz_t<Fq>* out, int iters) {
    affine_t<Fq> p;
    for (int i = 0; i < 8; ++i) { p.x.l[i] = CurveConst::g1_gen_x(i); p.y.l[i] = CurveConst::g1_gen_y(i); }
    xyzz_t<Fq> acc = xyzz_t<Fq>::dbl_affine(p.x, p.y);
    acc.x.l[0] ^= 0;   // keep
    for (int it = 0; it < iters; ++it) xyzz_t<Fq>::madd(acc, p, (it & 1) && false);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- 9 x 29-bit carry-free field (fp29.cuh): operands come from memory so nothing is constant-folded ----
__global__ void k_mul29_chain(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq29 a = Fq29::from_mont256(in[t & 1023]), b = Fq29::from_mont256(in[(t + 1) & 1023]);
    for (int it = 0; it < iters; ++it) a = Fq29::mul(a, b);
    out[t] = Fq29::to_mont256(a);
}
__global__ void k_sqr29_chain(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq29 a = Fq29::from_mont256(in[t & 1023]);
    for (int it = 0; it < iters; ++it) a = Fq29::sqr(a);
    out[t] = Fq29::to_mont256(a);
}
__global__ void k_mul32_chain_mem(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[t & 1023], y = in[(t + 1) & 1023];
    for (int it = 0; it < iters; ++it) x = Fq::mul(x, y);
    out[t] = x;
}
__global__ void __launch_bounds__(128, 4) k_madd29_chain(const affine_t<Fq>* pts, xyzz_t<Fq>* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz29_g1 acc = xyzz29_g1::identity();
    for (int it = 0; it < iters; ++it) xyzz29_g1::madd(acc, pts[(t * 7 + it) & 1023], (it & 3) == 1);
    out[t] = xyzz29_g1::to_xyzz(acc);
}
__global__ void __launch_bounds__(128, 4) k_madd32_chain_mem(const affine_t<Fq>* pts, xyzz_t<Fq>* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz_t<Fq> acc = xyzz_t<Fq>::identity();
    for (int it = 0; it < iters; ++it) xyzz_t<Fq>::madd(acc, pts[(t * 7 + it) & 1023], (it & 3) == 1);
    out[t] = acc;
}
// 1024 distinct curve points k G, k = 1..1024 (running sum), and 1024 field elements (their x coordinates)
__global__ void k_make_points(affine_t<Fq>* pts) {
    if (threadIdx.x || blockIdx.x) return;
    affine_t<Fq> g;
    for (int i = 0; i < 8; ++i) { g.x.l[i] = CurveConst::g1_gen_x(i); g.y.l[i] = CurveConst::g1_gen_y(i); }
    xyzz_t<Fq> acc = xyzz_t<Fq>::identity();
    for (int i = 0; i < 1024; ++i) { xyzz_t<Fq>::madd(acc, g, false); pts[i] = xyzz_t<Fq>::to_affine(acc); }
}
__global__ void k_cmp(const uint32_t* a, const uint32_t* b, size_t n, int* bad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(bad, 1);
}

template <class K>
static float timeit(K launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    void* buf; cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 256);
    int blocks = sms * 8, threads = 256;
    double nthreads = (double)blocks * threads;
    float ms;
    ms = timeit([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    printf("IMAD            : %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM @1.9GHz)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_imad_hi<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    printf("IMAD.HI         : %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)buf, 3, 5); });
    printf("IMAD.WIDE (hoisted by ptxas: measures IADD3): %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_imad_wide_rz<<<blocks, threads>>>((uint64_t*)buf, 3, 0x9E3779B9u); });
    printf("IMAD.WIDE (RZ addend, fed back) + LOP3: %.3f ms  %.2f T wide-ops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.0000001, 0.5); });
    printf("DFMA            : %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_imad_wide_cc<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    printf("IMAD.WIDE.X x4  : %.3f ms  %.2f T wide-ops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 4 / ms / 1e9, nthreads * ITERS * 4 / (ms * 1e-3) / sms / 1.9e9);
    ms = timeit([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)buf, 3, 5); });
    printf("IADD3(.X) x8    : %.3f ms  %.2f Tops/s  (%.1f lanes/clk/SM)\n", ms, nthreads * ITERS * 8 / ms / 1e9, nthreads * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    int iters = 2000;
    for (int t : {64, 128, 256, 512}) {
        int bl = sms * (2048 / t > 16 ? 16 : 2048 / t);
        ms = timeit([&] { k_mul_chain<<<bl, t>>>((Fq*)buf, iters); });
        printf("Fq::mul chain   : blocks %d x %d thr: %.3f ms  %.1f Gmul/s\n", bl, t, ms, (double)bl * t * iters / ms / 1e6);
    }
    ms = timeit([&] { k_mul_chain2<<<sms * 4, 256>>>((Fq*)buf, iters); });
    printf("Fq::mul 2chains : %.3f ms  %.1f Gmul/s\n", ms, (double)sms * 4 * 256 * iters * 2 / ms / 1e6);
    ms = timeit([&] { k_mul_chain<<<1, 32>>>((Fq*)buf, iters); });
    printf("Fq::mul latency : single warp: %.1f ns per dependent mul\n", ms * 1e6 / iters);
    ms = timeit([&] { k_madd_chain<<<sms * 4, 128>>>((xyzz_t<Fq>*)buf, 500); });
    printf("madd chain      : %.3f ms  %.2f Gmadd/s\n", ms, (double)sms * 4 * 128 * 500 / ms / 1e6);
    ms = timeit([&] { k_madd_chain<<<sms * 8, 128>>>((xyzz_t<Fq>*)buf, 500); });
    printf("madd chain x2occ: %.3f ms  %.2f Gmadd/s\n", ms, (double)sms * 8 * 128 * 500 / ms / 1e6);
    ms = timeit([&] { k_madd_chain<<<1, 32>>>((xyzz_t<Fq>*)buf, 500); });
    printf("madd latency    : single warp: %.1f ns per dependent madd\n", ms * 1e6 / 500);
    // ---- 29-bit-limb field vs the 32-bit-limb one, same operands from memory ----
    affine_t<Fq>* pts; cudaMalloc(&pts, 1024 * sizeof(affine_t<Fq>));
    k_make_points<<<1, 32>>>(pts); cudaDeviceSynchronize();
    void* buf2; cudaMalloc(&buf2, (size_t)sms * 8 * 1024 * 256);
    int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    const Fq* fin = (const Fq*)pts;
    for (int t : {128, 256}) {
        int bl = sms * (t == 128 ? 8 : 4);
        ms = timeit([&] { k_mul32_chain_mem<<<bl, t>>>(fin, (Fq*)buf, iters); });
        printf("Fq::mul (32-bit limbs, mem operands) blocks %d x %d: %.3f ms  %.1f Gmul/s\n", bl, t, ms, (double)bl * t * iters / ms / 1e6);
        ms = timeit([&] { k_mul29_chain<<<bl, t>>>(fin, (Fq*)buf2, iters); });
        printf("Fq29::mul (29-bit limbs)             blocks %d x %d: %.3f ms  %.1f Gmul/s\n", bl, t, ms, (double)bl * t * iters / ms / 1e6);
        k_cmp<<<(unsigned)(((size_t)bl * t * 8 + 255) / 256), 256>>>((uint32_t*)buf, (uint32_t*)buf2, (size_t)bl * t * 8, bad);
        ms = timeit([&] { k_sqr29_chain<<<bl, t>>>(fin, (Fq*)buf2, iters); });
        printf("Fq29::sqr                            blocks %d x %d: %.3f ms  %.1f Gsqr/s\n", bl, t, ms, (double)bl * t * iters / ms / 1e6);
    }
    for (int occ : {4, 8}) {
        ms = timeit([&] { k_madd32_chain_mem<<<sms * occ, 128>>>(pts, (xyzz_t<Fq>*)buf, 500); });
        printf("madd 32-bit limbs x%d blocks/SM: %.3f ms  %.2f Gmadd/s\n", occ, ms, (double)sms * occ * 128 * 500 / ms / 1e6);
        ms = timeit([&] { k_madd29_chain<<<sms * occ, 128>>>(pts, (xyzz_t<Fq>*)buf2, 500); });
        printf("madd 29-bit limbs x%d blocks/SM: %.3f ms  %.2f Gmadd/s\n", occ, ms, (double)sms * occ * 128 * 500 / ms / 1e6);
    }
    int hbad = 0; cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("mul29 == mul32 on %d-thread chains: %s (%d words differ)\n", sms * 8 * 128, hbad ? "MISMATCH" : "ok", hbad);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cuda status: %s\n", cudaGetErrorString(e));
    return 0;
}
