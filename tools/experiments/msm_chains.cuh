// msm_chains.cuh -- EXPERIMENT (not compiled into libb200zk.so): first bucket-accumulation level as chains of affine
// additions with batched inversions (batch_affine.cuh), VERDICT r1 item 1.
//
// Outcome on B200, 2^20 G1 pairs (commit 4b1f0b6.. of round 2; profiles/r2_chains_affine.md): BIT-EXACT (all of
// tests/test_gpu_msm.py and tests/test_gpu_prove.py passed with it enabled) but SLOWER than the XYZZ bucket kernel:
//   chains kernel 3.22 ms (K = 32) / 3.45 ms (K = 16) / 4.03 ms (K = 8) + 0.31 ms second level, against 2.57 ms XYZZ.
// ncu --set full (K = 16): multiplier pipe (fmaheavy) 35.8% busy, top stall long_scoreboard 7.5 warps per issue, barrier
// 3.5, wait 3.4; DRAM 6.0 GB read + 1.5 GB written per launch, L2 hit rate 35%.  The 6-products-per-addition arithmetic
// is real, but (i) every addition moves ~470 B of accumulator / prefix-product / point traffic through a working set
// (2.4 M chains x 96 B + the base array) that does not fit the 126 MB L2 -- the XYZZ kernel keeps its accumulator in
// registers and moves 68 B per addition; (ii) the shared inversion is a 45 us single-thread binary GCD per block-step,
// which needs >= 16 slots per thread to amortise, which is exactly what blows the working set up.  A start skew between
// the blocks of an SM (lockstep hypothesis) and L2 prefetches of the gathers changed nothing (3.78 -> 3.88 ms).
// Keeping the state in shared memory caps an SM at ~2000 slots = one block, whose inversion bubble then has nothing to
// hide behind.  Conclusion: on this machine the XYZZ mixed addition stays; see DESIGN.md section 7.
//
// The code below is the kernel as measured (it slots into msm.cu between the task tables and k_msm_accumulate; host side:
// scatter everything, build the task tables with task_len = CHAIN_LEN plus chain_lo / chain_len / chain_pos in sorted
// order, run this kernel, then run the XYZZ pipeline with bases = acc_pos, entries = chain_pos, offsets = chain offsets).
#pragma once
#include "batch_affine.cuh"

namespace b200zk {

// ---------------------------------------------------------------------------------------------
// 4b. first accumulation level: chains of <= CHAIN_LEN entries summed in AFFINE coordinates with batched inversions
//     (batch_affine.cuh): 6 field products per addition instead of the 10 of an XYZZ mixed addition.
//     Every bucket's sorted entry range is cut into chains (the task machinery above with task_len = CHAIN_LEN); the
//     chains are sorted by length and dealt to blocks of CHAIN_THREADS threads x K slots: thread `tid`, slot `k` of
//     block `b` owns the chain at sorted position b * (CHAIN_THREADS * K) + k * CHAIN_THREADS + tid (a warp touches 32
//     consecutive positions: the running sums acc_pos[] and the prefix products pref[] are read and written coalesced).
//     Step s adds entry s of every chain that has one.  All additions of a step in a block share ONE field inversion:
//     each thread multiplies its slots' denominators up (forward pass, prefix products parked in pref[]), the 128 thread
//     products go through a product tree in shared memory, thread 0 inverts the root, the tree is walked back down
//     (inverse of a child = inverse of the parent x the sibling's product) and each thread unwinds its own slots
//     (backward pass).  6 products per addition + ~3 per thread-step for the tree + one binary-GCD inversion per
//     block-step (integer ALU work, off the multiplier pipe; the block waits on it, the SM's other blocks do not).
//     The chain sums stay affine in acc_pos[]; the second level adds the few chains of every bucket with the XYZZ
//     kernel below (bases = acc_pos, entries = chain_pos, offsets = the chain offsets per bucket).
// ---------------------------------------------------------------------------------------------
static const uint32_t CHAIN_LEN = 8;
static const int CHAIN_THREADS = 128;

template <class F, int K>
__global__ void __launch_bounds__(CHAIN_THREADS, sizeof(F) > 32 ? 2 : 4) k_msm_chains_affine(
        const affine_t<F>* bases, const uint32_t* entries, const uint32_t* chain_lo, const uint32_t* chain_len,
        const uint32_t* nchains_ptr, affine_t<F>* acc_pos, F* pref, uint32_t skew_ns, uint32_t sm_count) {
    typedef batch_affine<F> BA;
    __shared__ F tree[2 * CHAIN_THREADS];              // [1] root, [2 i], [2 i + 1] children of [i]; leaves at [128 + tid]
    const uint32_t nch = *nchains_ptr;
    const uint32_t base = blockIdx.x * (uint32_t)(CHAIN_THREADS * K);
    if (base >= nch) return;                           // whole block
    const uint32_t tid = threadIdx.x;
    const uint32_t lmax = chain_len[base];             // chains are sorted by length, longest first
    // The blocks that start together on one SM would otherwise stay in lockstep (equal chain lengths) and sit in their
    // single-thread inversions at the same moments; a start skew keeps other blocks' products flowing meanwhile.
    for (uint32_t q = (blockIdx.x / sm_count) & 3u; skew_ns && q; --q) __nanosleep(skew_ns);
    // step 0: every chain starts as its first entry
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const uint32_t p = base + k * CHAIN_THREADS + tid;
        if (p >= nch) break;
        const uint32_t e = entries[chain_lo[p]];
        st16(acc_pos + p, BA::signed_point(ld16(bases + (e & 0x7FFFFFFFu)), (e >> 31) != 0));
    }
    for (uint32_t s = 1; s < lmax; ++s) {
        // warm the caches for this step: the (chain_lo -> entries -> bases) load chains of all slots in flight at once
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const uint32_t p = base + k * CHAIN_THREADS + tid;
            if (p >= nch || chain_len[p] <= s) break;
            const uint32_t e = entries[chain_lo[p] + s];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(bases + (e & 0x7FFFFFFFu)));
        }
        // forward: denominators and their running product
        F run = F::one();
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const uint32_t p = base + k * CHAIN_THREADS + tid;
            if (p >= nch || chain_len[p] <= s) break;   // sorted: the later slots are no longer either
            const uint32_t e = entries[chain_lo[p] + s];
            const affine_t<F> pt = BA::signed_point(ld16(bases + (e & 0x7FFFFFFFu)), (e >> 31) != 0);
            const affine_t<F> acc = ld16(acc_pos + p);
            F d;
            const int cs = BA::prepare(acc, pt, d);
            st16(pref + p, run);
            if (BA::needs_inverse(cs)) run = F::mul(run, d);
        }
        // one inversion for the block: product tree up, invert the root, inverses down
        tree[CHAIN_THREADS + tid] = run;
        __syncthreads();
#pragma unroll 1
        for (uint32_t n = CHAIN_THREADS / 2; n >= 1; n >>= 1) {
            if (tid < n) tree[n + tid] = F::mul(tree[2 * (n + tid)], tree[2 * (n + tid) + 1]);
            __syncthreads();
        }
        if (tid == 0) tree[1] = F::inv(tree[1]);
        __syncthreads();
#pragma unroll 1
        for (uint32_t n = 1; n <= CHAIN_THREADS / 2; n <<= 1) {
            if (tid < n) {
                const uint32_t i = n + tid;
                const F a = tree[2 * i], b = tree[2 * i + 1], v = tree[i];
                tree[2 * i] = F::mul(v, b);
                tree[2 * i + 1] = F::mul(v, a);
            }
            __syncthreads();
        }
        F inv = tree[CHAIN_THREADS + tid];              // 1 / (this thread's product)
        // backward: unwind the prefix products, finish the additions
        int kmax = 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const uint32_t p = base + k * CHAIN_THREADS + tid;
            if (p >= nch || chain_len[p] <= s) break;
            kmax = k + 1;
        }
#pragma unroll 1
        for (int k = kmax - 1; k >= 0; --k) {
            const uint32_t p = base + k * CHAIN_THREADS + tid;
            const uint32_t e = entries[chain_lo[p] + s];
            const affine_t<F> pt = BA::signed_point(ld16(bases + (e & 0x7FFFFFFFu)), (e >> 31) != 0);
            affine_t<F> acc = ld16(acc_pos + p);
            F d;
            const int cs = BA::prepare(acc, pt, d);
            F dinv = inv;
            if (BA::needs_inverse(cs)) {
                dinv = F::mul(inv, ld16(pref + p));     // 1 / d
                inv = F::mul(inv, d);                   // drop d from the running inverse
            }
            BA::finish(cs, acc, pt, dinv);
            st16(acc_pos + p, acc);
        }
    }
}


}  // namespace b200zk
