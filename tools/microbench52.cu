// EXPERIMENT: Montgomery products on the FP64 pipe (tools/experiments/fp52.cuh) against the integer product of csrc/fp.cuh,
// alone and sharing an SM (warp-specialised and interleaved in one thread).  VERDICT r1 item 10.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Iinclude -o tools/microbench52 tools/microbench52.cu
#include <cstdio>
#include <cstdint>
#include <functional>
#include "../distributed_groth16_b200/csrc/fp.cuh"
#include "experiments/fp52.cuh"
using namespace b200zk;
typedef Fp52<FqParams> Fq52;

__global__ void k_fill(Fq* in) {      // 1024 pseudo-random canonical field elements: powers of a fixed element
    if (threadIdx.x || blockIdx.x) return;
    Fq g = Fq::one(); g.l[0] += 12345; g.l[3] ^= 0x5a5a5a5a; g.l[7] &= 0x0fffffff;
    Fq x = g;
    for (int i = 0; i < 1024; ++i) { in[i] = x; x = Fq::mul(x, g); }
}
__global__ void k_mul32(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = in[t & 1023], y = in[(t + 1) & 1023];
    for (int it = 0; it < iters; ++it) x = Fq::mul(x, y);
    out[t] = x;
}
__global__ void k_mul52(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq52 x = Fq52::from_mont256(in[t & 1023]), y = Fq52::from_mont256(in[(t + 1) & 1023]);
    for (int it = 0; it < iters; ++it) x = Fq52::mul(x, y);
    out[t] = Fq52::to_mont256(x);
}
__global__ void k_mul52x2(const Fq* in, Fq* out, int iters) {      // two independent chains per thread
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq52 x = Fq52::from_mont256(in[t & 1023]), y = Fq52::from_mont256(in[(t + 1) & 1023]), z = Fq52::from_mont256(in[(t + 2) & 1023]);
    for (int it = 0; it < iters; ++it) { x = Fq52::mul(x, y); z = Fq52::mul(z, y); }
    out[t] = Fq::add(Fq52::to_mont256(x), Fq52::to_mont256(z));
}
// even warps: integer products, odd warps: FP64 products; i32 / i52 iterations each
__global__ void k_mixed_warps(const Fq* in, Fq* out, int i32, int i52) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ((threadIdx.x >> 5) & 1) {
        Fq52 x = Fq52::from_mont256(in[t & 1023]), y = Fq52::from_mont256(in[(t + 1) & 1023]);
        for (int it = 0; it < i52; ++it) x = Fq52::mul(x, y);
        out[t] = Fq52::to_mont256(x);
    } else {
        Fq x = in[t & 1023], y = in[(t + 1) & 1023];
        for (int it = 0; it < i32; ++it) x = Fq::mul(x, y);
        out[t] = x;
    }
}
// one integer chain and one FP64 chain interleaved in every thread
__global__ void k_mixed_ilp(const Fq* in, Fq* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a = in[t & 1023], y = in[(t + 1) & 1023];
    Fq52 x = Fq52::from_mont256(in[(t + 2) & 1023]), y5 = Fq52::from_mont256(y);
    for (int it = 0; it < iters; ++it) { a = Fq::mul(a, y); x = Fq52::mul(x, y5); }
    out[t] = Fq::add(a, Fq52::to_mont256(x));
}
__global__ void k_cmp(const uint32_t* a, const uint32_t* b, size_t n, int* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(bad, 1);
}
static float timeit(std::function<void()> f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, sms);
    Fq *in, *o1, *o2; cudaMalloc(&in, 1024 * sizeof(Fq)); cudaMalloc(&o1, (size_t)sms * 16 * 256 * sizeof(Fq)); cudaMalloc(&o2, (size_t)sms * 16 * 256 * sizeof(Fq));
    int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    k_fill<<<1, 32>>>(in); cudaDeviceSynchronize();
    const int iters = 2000;
    float ms;
    for (int occ : {2, 4, 8}) {
        int bl = sms * occ, t = 128;
        ms = timeit([&] { k_mul32<<<bl, t>>>(in, o1, iters); });
        printf("integer product  (8 x 32-bit limbs) %2d x 128 thr/SM: %.3f ms  %.1f Gmul/s\n", occ, ms, (double)bl * t * iters / ms / 1e6);
        ms = timeit([&] { k_mul52<<<bl, t>>>(in, o2, iters); });
        printf("FP64 product     (5 x 52-bit limbs) %2d x 128 thr/SM: %.3f ms  %.1f Gmul/s\n", occ, ms, (double)bl * t * iters / ms / 1e6);
        k_cmp<<<(unsigned)(((size_t)bl * t * 8 + 255) / 256), 256>>>((uint32_t*)o1, (uint32_t*)o2, (size_t)bl * t * 8, bad);
        ms = timeit([&] { k_mul52x2<<<bl, t>>>(in, o2, iters); });
        printf("FP64 product, 2 chains / thread     %2d x 128 thr/SM: %.3f ms  %.1f Gmul/s\n", occ, ms, (double)bl * t * iters * 2 / ms / 1e6);
    }
    int hbad = 0; cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
    printf("FP64 product == integer product after 2000-step chains: %s (%d words differ)\n", hbad ? "MISMATCH" : "ok", hbad);
    for (int occ : {4, 8}) {
        int bl = sms * occ, t = 128;
        const int pairs[5][2] = {{2000, 0}, {0, 2000}, {2000, 2000}, {2000, 3000}, {2000, 1500}};
        for (auto& pr2 : pairs) {
            ms = timeit([&] { k_mixed_warps<<<bl, t>>>(in, o2, pr2[0], pr2[1]); });
            printf("warp-specialised %2d x 128 thr/SM, %4d integer + %4d FP64 iterations: %.3f ms  %.1f Gmul/s total\n", occ, pr2[0], pr2[1], ms,
                   (double)bl * t / 2 * (pr2[0] + pr2[1]) / ms / 1e6);
        }
        ms = timeit([&] { k_mixed_ilp<<<bl, t>>>(in, o2, iters); });
        printf("interleaved in one thread %2d x 128 thr/SM: %.3f ms  %.1f Gmul/s total\n", occ, ms, (double)bl * t * iters * 2 / ms / 1e6);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
