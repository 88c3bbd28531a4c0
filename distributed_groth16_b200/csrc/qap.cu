// qap.cu -- R1CS x witness -> QAP evaluation vectors on the device, plus the Montgomery
// conversions the file readers need.
//
// Replaces `qap::qap` (/root/reference/groth16/src/qap.rs:44-91; identical logic in
// ark-circom/src/circom/qap.rs:38-62): a_i = <A_i, z>, b_i = <B_i, z> for i < num_constraints (rayon
// `evaluate_constraint` per row, qap.rs:60-67), a[num_constraints + j] = z[j] for j < num_inputs (:69-73),
// c_i = a_i * b_i (:75-81), everything zero-padded to the domain size m.
// HBM-bound sparse mat-vec: per non-zero 4 B column index + 32 B coefficient + a 32 B gather from z.
#include "common.cuh"

namespace b200zk {

__device__ __forceinline__ Fr ldq(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stq(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__global__ void k_fr_convert(const Fr* in, Fr* out, size_t n, int to_mont, int times) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = ldq(in + i);
    for (int t = 0; t < times; ++t) v = to_mont ? Fr::to_mont(v) : Fr::from_mont(v);
    stq(out + i, v);
}

__device__ __forceinline__ Fr row_dot(const uint32_t* ptr, const uint32_t* col, const Fr* val, const Fr* z, uint32_t i) {
    Fr acc = Fr::zero();
    for (uint32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k) acc = Fr::add(acc, Fr::mul(ldq(val + k), ldq(z + col[k])));
    return acc;
}

__global__ void k_qap(const uint32_t* a_ptr, const uint32_t* a_col, const Fr* a_val, const uint32_t* b_ptr, const uint32_t* b_col,
                      const Fr* b_val, const Fr* z, uint32_t nc, uint32_t n_inputs, uint32_t m, Fr* a, Fr* b, Fr* c) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fr va = Fr::zero(), vb = Fr::zero(), vc = Fr::zero();
    if (i < nc) {
        va = row_dot(a_ptr, a_col, a_val, z, i);
        vb = row_dot(b_ptr, b_col, b_val, z, i);
        vc = Fr::mul(va, vb);
    } else if (i < nc + n_inputs) {
        va = ldq(z + (i - nc));
    }
    stq(a + i, va); stq(b + i, vb); stq(c + i, vc);
}

// One warp per row: the lanes stride over the row's non-zeros and the partial sums meet in a shuffle tree.  Small circuits have
// far fewer rows than the machine has lanes and a few long rows (the reference's sha256 circuit: 30 134 rows, the longest row
// is a serial chain of ~0.5 ms for one thread); field addition is exact, so the order of summation does not change the result.
__device__ __forceinline__ Fr warp_sum(Fr v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.l[i] = __shfl_down_sync(0xFFFFFFFFu, v.l[i], d);
        v = Fr::add(v, o);
    }
    return v;
}
__device__ __forceinline__ Fr row_dot_warp(const uint32_t* ptr, const uint32_t* col, const Fr* val, const Fr* z, uint32_t i, uint32_t lane) {
    Fr acc = Fr::zero();
    for (uint32_t k = ptr[i] + lane, e = ptr[i + 1]; k < e; k += 32) acc = Fr::add(acc, Fr::mul(ldq(val + k), ldq(z + col[k])));
    return warp_sum(acc);
}
__global__ void __launch_bounds__(256) k_qap_warp(const uint32_t* a_ptr, const uint32_t* a_col, const Fr* a_val, const uint32_t* b_ptr,
                                                  const uint32_t* b_col, const Fr* b_val, const Fr* z, uint32_t nc, uint32_t n_inputs,
                                                  uint32_t m, Fr* a, Fr* b, Fr* c) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= m) return;
    Fr va = Fr::zero(), vb = Fr::zero(), vc = Fr::zero();
    if (i < nc) {
        va = row_dot_warp(a_ptr, a_col, a_val, z, i, lane);
        vb = row_dot_warp(b_ptr, b_col, b_val, z, i, lane);
        vc = Fr::mul(va, vb);
    } else if (i < nc + n_inputs) {
        va = ldq(z + (i - nc));
    }
    if (lane == 0) { stq(a + i, va); stq(b + i, vb); stq(c + i, vc); }
}

int fr_convert_dev(b200zk_ctx* ctx, Slot& sl, const void* d_in, void* d_out, size_t n, int to_mont, int times) {
    if (n == 0) return B200ZK_OK;
    {
        LaunchScope ls(ctx, sl.stream, "fr_convert");
        k_fr_convert<<<(unsigned)((n + 255) / 256), 256, 0, sl.stream>>>((const Fr*)d_in, (Fr*)d_out, n, to_mont, times);
    }
    return check_launch(ctx, "k_fr_convert");
}

int qap_dev(b200zk_ctx* ctx, Slot& sl, const void* a_ptr, const void* a_col, const void* a_val, const void* b_ptr,
            const void* b_col, const void* b_val, size_t nc, size_t n_inputs, const void* d_z, unsigned log_m, void* d_a,
            void* d_b, void* d_c) {
    if (log_m > 28) return set_error(ctx, B200ZK_ERR_DOMAIN, "domain too large (PolynomialDegreeTooLarge)");
    size_t m = (size_t)1 << log_m;
    if (nc + n_inputs > m) return set_error(ctx, B200ZK_ERR_DOMAIN, "num_constraints + num_inputs exceeds the domain size");
    {
        LaunchScope ls(ctx, sl.stream, "qap_matvec");
        const uint32_t *ap = (const uint32_t*)a_ptr, *ac = (const uint32_t*)a_col, *bp = (const uint32_t*)b_ptr, *bc = (const uint32_t*)b_col;
        if (m <= ((size_t)1 << 18))          // warp per row while rows x 32 lanes still fit a few waves
            k_qap_warp<<<(unsigned)((m * 32 + 255) / 256), 256, 0, sl.stream>>>(ap, ac, (const Fr*)a_val, bp, bc, (const Fr*)b_val, (const Fr*)d_z,
                                                                                (uint32_t)nc, (uint32_t)n_inputs, (uint32_t)m, (Fr*)d_a, (Fr*)d_b, (Fr*)d_c);
        else
            k_qap<<<(unsigned)((m + 127) / 128), 128, 0, sl.stream>>>(ap, ac, (const Fr*)a_val, bp, bc, (const Fr*)b_val, (const Fr*)d_z,
                                                                       (uint32_t)nc, (uint32_t)n_inputs, (uint32_t)m, (Fr*)d_a, (Fr*)d_b, (Fr*)d_c);
    }
    return check_launch(ctx, "k_qap");
}

}  // namespace b200zk
