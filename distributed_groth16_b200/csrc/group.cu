// group.cu -- the multi-GPU hot path behind the C ABI: one host process, one b200zk_ctx per device, NVLink peer access.
//
// Replaces, for the single-box setting, the reference's king/client star (/root/reference/mpc-net/src/lib.rs:61-139 used
// through dist-primitives/src/channel/mod.rs:8-56) under the very calls the Rust prover makes:
//   d_msm   dist-primitives/src/dmsm/mod.rs:70-98      -> b200zk_group_msm_g1 / _g2   (bases / scalars split by index range,
//                                                          one XYZZ partial per GPU, peer copies to GPU 0, point sum)
//   d_fft / d_ifft  dist-primitives/src/dfft/mod.rs:17-95 -> b200zk_group_ntt_fr        (four-step transform: column NTTs whose last
//                                                          pass stores straight into the owning peer's memory -- the kernel IS
//                                                          the all-to-all -- then row NTTs)
//   ext_wit::h      groth16/src/ext_wit.rs:16-101         -> b200zk_group_h_circom      (3 inverse + 3 forward sharded transforms)
//   prove::{A,B,C}  groth16/src/prove.rs:21-136 + sha256.rs:208-212 -> b200zk_group_groth16_prove (BASELINE config 5)
// It is the C++ port of distributed_groth16_b200/parallel.py (the torch.distributed orchestration used by bench.py under
// torchrun): same layouts, same kernels (ntt.cu / msm.cu / prove.cu entry points), the NCCL collectives replaced by peer
// stores and cross-device events because everything lives in one address space.  The host thread only enqueues: all
// devices run concurrently, ordered by events; the only host synchronisation is the final read of the result.
//
// Distributed vectors use the "column layout" of parallel.py: for a length-N vector seen as rows x cols (row-major),
// GPU g of P owns columns [g cols/P, (g+1) cols/P), each stored contiguously: local[c][r] = x[r cols + g cols/P + c].
// A transform with (rows, cols) maps layout(cols) to layout(rows), so iNTT -> NTT chains need no redistribution.
#include "common.cuh"

#include <algorithm>

using namespace b200zk;

struct b200zk_group {
    int n = 0;
    std::vector<b200zk_ctx*> ctx;
    void* recv[2][8] = {};            // four-step receive buffers (2 alternate) on every device
    size_t recv_bytes = 0;
    int turn = 0;
    std::vector<cudaEvent_t> ev_col, ev_done;   // per device
    std::string last_error;
};

struct b200zk_group_pk {
    int n = 0;
    size_t n_vars = 0, n_inputs = 0, m = 0;
    unsigned log_m = 0;
    // per device: slices of the five queries (a / b_g1 / b_g2 rows [lo_g, hi_g) of n_vars; l rows of n_vars - n_inputs;
    // h in the column layout) and their fixed-base tables
    struct Shard {
        size_t v0 = 0, v1 = 0, l0 = 0, l1 = 0, hn = 0;
        void* q[5] = {};              // a, b_g1, b_g2, l, h
        void* tab[5] = {};
        unsigned tab_c[5] = {};
    };
    std::vector<Shard> shard;
    b200zk_pk* pk0 = nullptr;         // device 0: vk points for the final assembly
    size_t table_bytes = 0;
};

namespace {

// the group calls hop between devices; the caller's current device (which frameworks such as torch cache) is restored on exit
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int gerr(b200zk_group* g, int code, const std::string& msg) {
    if (g) g->last_error = msg;
    return code;
}
int gfrom(b200zk_group* g, int rank, int rc) {
    if (rc != B200ZK_OK && g) g->last_error = "device " + std::to_string(g->ctx[rank]->device) + ": " + g->ctx[rank]->last_error;
    return rc;
}
#define G_CUDA_OK(grp, expr)                                                                                  \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            char _b[512];                                                                                     \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
            return gerr((grp), _e == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA, _b);         \
        }                                                                                                     \
    } while (0)
#define G_TRY(grp, rank, expr)                     \
    do {                                           \
        int _rc = gfrom((grp), (rank), (expr));    \
        if (_rc != B200ZK_OK) return _rc;          \
    } while (0)

// out[c][r] = in[r][c] for an R x C matrix of T (Fr: 32 B, G1 point: 64 B, G2 point: 128 B)
template <class T>
__global__ void k_transpose(const T* in, T* out, uint32_t R, uint32_t C) {
    __shared__ T tile[16][17];
    const uint32_t c = blockIdx.x * 16 + threadIdx.x, r = blockIdx.y * 16 + threadIdx.y;
    if (r < R && c < C) tile[threadIdx.y][threadIdx.x] = in[(size_t)r * C + c];
    __syncthreads();
    const uint32_t oc = blockIdx.x * 16 + threadIdx.y, orow = blockIdx.y * 16 + threadIdx.x;   // out is C x R
    if (oc < C && orow < R) out[(size_t)oc * R + orow] = tile[threadIdx.x][threadIdx.y];
}
struct Blob32 { uint4 v[2]; };
struct Blob64 { uint4 v[4]; };
struct Blob128 { uint4 v[8]; };

template <class T>
int transpose_dev(b200zk_ctx* ctx, cudaStream_t st, const void* in, void* out, size_t R, size_t C) {
    dim3 grid((unsigned)((C + 15) / 16), (unsigned)((R + 15) / 16)), block(16, 16);
    {
        LaunchScope ls(ctx, st, "transpose");
        k_transpose<T><<<grid, block, 0, st>>>(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), (uint32_t)R, (uint32_t)C);
    }
    return check_launch(ctx, "k_transpose");
}

void split_log(unsigned log_n, unsigned* log_rows, unsigned* log_cols) {
    *log_rows = (log_n + 1) / 2;
    *log_cols = log_n / 2;
}
unsigned ilog2(size_t x) { unsigned l = 0; while (((size_t)1 << l) < x) ++l; return l; }

// every device's stream waits until all devices have passed `ev[g]` (recorded on their streams just before)
int cross_barrier(b200zk_group* grp, std::vector<cudaEvent_t>& ev) {
    for (int g = 0; g < grp->n; ++g) {
        G_CUDA_OK(grp, cudaSetDevice(grp->ctx[g]->device));
        G_CUDA_OK(grp, cudaEventRecord(ev[g], grp->ctx[g]->slots[0].stream));
    }
    for (int g = 0; g < grp->n; ++g) {
        G_CUDA_OK(grp, cudaSetDevice(grp->ctx[g]->device));
        for (int o = 0; o < grp->n; ++o)
            if (o != g) G_CUDA_OK(grp, cudaStreamWaitEvent(grp->ctx[g]->slots[0].stream, ev[o], 0));
    }
    return B200ZK_OK;
}

int ensure_recv(b200zk_group* grp, size_t bytes) {
    if (bytes <= grp->recv_bytes) return B200ZK_OK;
    for (int g = 0; g < grp->n; ++g) {
        G_CUDA_OK(grp, cudaSetDevice(grp->ctx[g]->device));
        G_CUDA_OK(grp, cudaDeviceSynchronize());
        for (int b = 0; b < 2; ++b) {
            if (grp->recv[b][g]) cudaFree(grp->recv[b][g]);
            grp->recv[b][g] = nullptr;
            G_CUDA_OK(grp, cudaMalloc(&grp->recv[b][g], bytes));
        }
    }
    grp->recv_bytes = bytes;
    return B200ZK_OK;
}

// One sharded four-step transform (parallel.sharded_ntt_p2p).  in[g]: (cols / P) x rows on device g; out[g]: (rows / P) x cols.
// shift_log_m != 0 (inverse transforms of the h pipeline): output coefficient j is also multiplied by w_2m^j.
int sharded_ntt(b200zk_group* grp, Fr* const* in, Fr* const* out, unsigned log_rows, unsigned log_cols, bool inverse, unsigned shift_log_m) {
    const int P = grp->n;
    const size_t rows = (size_t)1 << log_rows, cols = (size_t)1 << log_cols;
    if (rows % P || cols % P) return gerr(grp, B200ZK_ERR_ARG, "rows and cols of the four-step split must be divisible by the number of GPUs");
    const size_t cg = cols / P, rl = rows / P;
    int rc = ensure_recv(grp, rl * cols * sizeof(Fr));
    if (rc) return rc;
    const int b = grp->turn;
    grp->turn ^= 1;
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        G_TRY(grp, g, fourstep_cols_p2p_dev(ctx, ctx->slots[0], in[g], grp->recv[b], (unsigned)P, log_rows, ilog2(cg), log_rows + log_cols,
                                            (uint64_t)g * cg, inverse));
    }
    rc = cross_barrier(grp, grp->ev_col);          // all column kernels (= all peer stores) are complete
    if (rc) return rc;
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        const Fr* rows_in = reinterpret_cast<const Fr*>(grp->recv[b][g]);
        if (!shift_log_m) G_TRY(grp, g, ntt_dev(ctx, ctx->slots[0], rows_in, out[g], log_cols, inverse, false, (unsigned)rl));
        else G_TRY(grp, g, ntt_batched_post_dev(ctx, ctx->slots[0], rows_in, out[g], log_cols, (unsigned)rl, inverse, shift_log_m + 1, true,
                                                 (uint64_t)g * rl, 0, 1, rows));
    }
    return B200ZK_OK;
}

// host vector (natural order, N = rows x cols row-major) -> column layout on the devices: a strided H2D copy of the
// device's column slab (rows x cg) followed by an on-device transpose to cg x rows
template <class T>
int scatter_columns(b200zk_group* grp, const void* host, size_t rows, size_t cols, void* const* d_tmp, void* const* d_out) {
    const int P = grp->n;
    const size_t cg = cols / P;
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->slots[0].stream;
        G_CUDA_OK(grp, cudaMemcpy2DAsync(d_tmp[g], cg * sizeof(T), reinterpret_cast<const char*>(host) + g * cg * sizeof(T), cols * sizeof(T),
                                         cg * sizeof(T), rows, cudaMemcpyHostToDevice, st));
        G_TRY(grp, g, transpose_dev<T>(ctx, st, d_tmp[g], d_out[g], rows, cg));
    }
    return B200ZK_OK;
}
// column layout (cg x rows per device) -> host natural order
template <class T>
int gather_columns(b200zk_group* grp, void* host, size_t rows, size_t cols, void* const* d_in, void* const* d_tmp) {
    const int P = grp->n;
    const size_t cg = cols / P;
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->slots[0].stream;
        G_TRY(grp, g, transpose_dev<T>(ctx, st, d_in[g], d_tmp[g], cg, rows));
        G_CUDA_OK(grp, cudaMemcpy2DAsync(reinterpret_cast<char*>(host) + g * cg * sizeof(T), cols * sizeof(T), d_tmp[g], cg * sizeof(T),
                                         cg * sizeof(T), rows, cudaMemcpyDeviceToHost, st));
    }
    return B200ZK_OK;
}

int sync_all(b200zk_group* grp) {
    for (int g = 0; g < grp->n; ++g) {
        G_CUDA_OK(grp, cudaSetDevice(grp->ctx[g]->device));
        G_CUDA_OK(grp, cudaStreamSynchronize(grp->ctx[g]->slots[0].stream));
    }
    return B200ZK_OK;
}

struct SlotLocks {           // slot 0 of every device, in rank order (same order everywhere: no deadlock)
    std::vector<std::unique_lock<std::mutex>> l;
    explicit SlotLocks(b200zk_group* grp) { for (int g = 0; g < grp->n; ++g) l.emplace_back(grp->ctx[g]->slots[0].mu); }
};

// balanced index-range split
size_t cut(size_t n, int g, int P) { return (size_t)(((unsigned __int128)n * (unsigned)g) / (unsigned)P); }

template <int G2>
int group_msm(b200zk_group* grp, const uint64_t* bases, size_t n_bases, const uint64_t* scalars, size_t n_scalars, uint64_t* out_affine,
              int* out_is_inf) {
    DeviceGuard dev_guard;
    if (!grp || !out_affine || !out_is_inf) return B200ZK_ERR_ARG;
    if (n_bases != n_scalars) return gerr(grp, B200ZK_ERR_LENGTH, std::to_string(n_bases < n_scalars ? n_bases : n_scalars));
    const size_t n = n_bases;
    if (n && (!bases || !scalars)) return gerr(grp, B200ZK_ERR_ARG, "null input");
    const size_t PB = G2 ? 128 : 64, XB = G2 ? 256 : 128;
    const int P = grp->n;
    SlotLocks locks(grp);
    // every device: stage its slice (scalars on the compute stream, bases on the copy stream behind the sort phases), MSM
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        Slot& sl = ctx->slots[0];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        const size_t lo = cut(n, g, P), cnt = cut(n, g + 1, P) - lo;
        G_CUDA_OK(grp, sl.small.reserve(4096));
        // this device's index range, staged in parts behind its own bucket kernels (api.cu, msm_staged_dev)
        G_TRY(grp, g, msm_staged_dev(ctx, sl, G2, reinterpret_cast<const char*>(bases) + lo * PB, reinterpret_cast<const char*>(scalars) + lo * 32,
                                     cnt, sl.small.p, false));
        G_CUDA_OK(grp, cudaEventRecord(grp->ev_done[g], sl.stream));
    }
    // partials -> device 0 (peer copies), sum, normalise
    b200zk_ctx* c0 = grp->ctx[0];
    Slot& s0 = c0->slots[0];
    G_CUDA_OK(grp, cudaSetDevice(c0->device));
    char* sm = reinterpret_cast<char*>(s0.small.p);
    char* gathered = sm + 1024;                               // P x XB <= 2 KB
    for (int g = 0; g < P; ++g) {
        if (g) G_CUDA_OK(grp, cudaStreamWaitEvent(s0.stream, grp->ev_done[g], 0));
        G_CUDA_OK(grp, cudaMemcpyPeerAsync(gathered + g * XB, c0->device, grp->ctx[g]->slots[0].small.p, grp->ctx[g]->device, XB, s0.stream));
    }
    G_TRY(grp, 0, G2 ? g2_sum_dev(c0, s0, gathered, P, sm + 3584) : g1_sum_dev(c0, s0, gathered, P, sm + 3584));
    uint64_t host[17];
    G_CUDA_OK(grp, cudaMemcpyAsync(host, sm + 3584, PB + 8, cudaMemcpyDeviceToHost, s0.stream));
    G_CUDA_OK(grp, cudaStreamSynchronize(s0.stream));
    int rc = sync_all(grp);
    for (int g = 0; g < P && !rc; ++g) {                      // the host buffers are free again once every copy stream drained
        cudaSetDevice(grp->ctx[g]->device);
        if (cudaStreamSynchronize(grp->ctx[g]->slots[0].copy_stream) != cudaSuccess) rc = gerr(grp, B200ZK_ERR_CUDA, "copy stream failed");
    }
    if (rc) return rc;
    memcpy(out_affine, host, PB);
    *out_is_inf = (int)host[PB / 8];
    return B200ZK_OK;
}

// the h pipeline on vectors already in the column layout: a, b, c, h: (cols / P) x rows per device
int sharded_h(b200zk_group* grp, Fr* const* a, Fr* const* b, Fr* const* c, Fr* const* tmp1, Fr* const* tmp2, unsigned log_m, Fr* const* h) {
    unsigned log_rows, log_cols;
    split_log(log_m, &log_rows, &log_cols);
    // a, b, c are transformed in place (through tmp1): coef = iNTT + shift, ev = NTT
    Fr* const* vecs[3] = {a, b, c};
    for (int k = 0; k < 3; ++k) {
        int rc = sharded_ntt(grp, vecs[k], tmp1, log_rows, log_cols, true, log_m);
        if (rc) return rc;
        rc = sharded_ntt(grp, tmp1, vecs[k], log_cols, log_rows, false, 0);
        if (rc) return rc;
    }
    const size_t local = ((size_t)1 << log_m) / grp->n;
    for (int g = 0; g < grp->n; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        G_TRY(grp, g, mul_sub_dev(ctx, ctx->slots[0], a[g], b[g], c[g], h[g], local));
    }
    (void)tmp2;
    return B200ZK_OK;
}

}  // namespace

extern "C" {

int b200zk_group_create(const int* device_ids, int n_dev, b200zk_group** out) {
    DeviceGuard dev_guard;
    if (!device_ids || !out || n_dev < 1 || n_dev > 8 || (n_dev & (n_dev - 1))) return B200ZK_ERR_ARG;
    b200zk_group* grp = new b200zk_group();
    grp->n = n_dev;
    for (int g = 0; g < n_dev; ++g) {
        b200zk_ctx* c = nullptr;
        int rc = b200zk_ctx_create(device_ids[g], &c);
        if (rc) { b200zk_group_destroy(grp); return rc; }
        grp->ctx.push_back(c);
    }
    for (int g = 0; g < n_dev; ++g) {
        cudaSetDevice(device_ids[g]);
        for (int o = 0; o < n_dev; ++o) {
            if (o == g || device_ids[o] == device_ids[g]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, device_ids[g], device_ids[o]);
            if (!can) { b200zk_group_destroy(grp); return B200ZK_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[o], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { b200zk_group_destroy(grp); return B200ZK_ERR_CUDA; }
            cudaGetLastError();
        }
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        if (cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
            b200zk_group_destroy(grp);
            return B200ZK_ERR_CUDA;
        }
        grp->ev_col.push_back(e1);
        grp->ev_done.push_back(e2);
    }
    *out = grp;
    return B200ZK_OK;
}

void b200zk_group_destroy(b200zk_group* grp) {
    DeviceGuard dev_guard;
    if (!grp) return;
    for (size_t g = 0; g < grp->ctx.size(); ++g) {
        cudaSetDevice(grp->ctx[g]->device);
        cudaDeviceSynchronize();
        for (int b = 0; b < 2; ++b) if (grp->recv[b][g]) cudaFree(grp->recv[b][g]);
        if (g < grp->ev_col.size()) cudaEventDestroy(grp->ev_col[g]);
        if (g < grp->ev_done.size()) cudaEventDestroy(grp->ev_done[g]);
    }
    for (b200zk_ctx* c : grp->ctx) b200zk_ctx_destroy(c);
    delete grp;
}

int b200zk_group_size(const b200zk_group* grp) { return grp ? grp->n : 0; }
b200zk_ctx* b200zk_group_ctx(b200zk_group* grp, int rank) { return grp && rank >= 0 && rank < grp->n ? grp->ctx[rank] : nullptr; }
const char* b200zk_group_last_error(const b200zk_group* grp) { return grp ? grp->last_error.c_str() : "null group"; }

int b200zk_group_msm_g1(b200zk_group* grp, const uint64_t* bases, size_t n_bases, const uint64_t* scalars, size_t n_scalars,
                        uint64_t out_affine[8], int* out_is_inf) {
    return group_msm<0>(grp, bases, n_bases, scalars, n_scalars, out_affine, out_is_inf);
}
int b200zk_group_msm_g2(b200zk_group* grp, const uint64_t* bases, size_t n_bases, const uint64_t* scalars, size_t n_scalars,
                        uint64_t out_affine[16], int* out_is_inf) {
    return group_msm<1>(grp, bases, n_bases, scalars, n_scalars, out_affine, out_is_inf);
}

int b200zk_group_ntt_fr(b200zk_group* grp, uint64_t* data, unsigned log_n, int inverse) {
    DeviceGuard dev_guard;
    if (!grp || !data) return B200ZK_ERR_ARG;
    if (log_n > 28) return gerr(grp, B200ZK_ERR_DOMAIN, "log n > 28");
    const int P = grp->n;
    unsigned log_rows, log_cols;
    split_log(log_n, &log_rows, &log_cols);
    const size_t rows = (size_t)1 << log_rows, cols = (size_t)1 << log_cols, N = (size_t)1 << log_n;
    if (cols < (size_t)P) return gerr(grp, B200ZK_ERR_ARG, "transform too small for this many GPUs");
    SlotLocks locks(grp);
    std::vector<void*> bufA(P), bufB(P);
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        G_CUDA_OK(grp, ctx->slots[0].io_a.reserve(2 * (N / P) * sizeof(Fr)));
        bufA[g] = ctx->slots[0].io_a.p;
        bufB[g] = reinterpret_cast<Fr*>(bufA[g]) + N / P;
    }
    int rc = scatter_columns<Blob32>(grp, data, rows, cols, bufB.data(), bufA.data());         // layout(cols) in A
    if (rc) return rc;
    std::vector<Fr*> in(P), out(P);
    for (int g = 0; g < P; ++g) { in[g] = reinterpret_cast<Fr*>(bufA[g]); out[g] = reinterpret_cast<Fr*>(bufB[g]); }
    rc = sharded_ntt(grp, in.data(), out.data(), log_rows, log_cols, inverse != 0, 0);          // layout(rows) in B: (rows / P) x cols
    if (rc) return rc;
    // X[k1 + rows k2] sits at out[g][k1 local][k2]: as a cols x rows row-major matrix X[k2][k1] the device owns columns k1
    rc = gather_columns<Blob32>(grp, data, cols, rows, bufB.data(), bufA.data());
    if (rc) return rc;
    return sync_all(grp);
}

int b200zk_group_h_circom(b200zk_group* grp, const uint64_t* a, const uint64_t* b, const uint64_t* c, unsigned log_m, uint64_t* h_out) {
    DeviceGuard dev_guard;
    if (!grp || !a || !b || !c || !h_out) return B200ZK_ERR_ARG;
    if (log_m + 1 > 28) return gerr(grp, B200ZK_ERR_DOMAIN, "2m exceeds the 2^28 subgroup (PolynomialDegreeTooLarge)");
    const int P = grp->n;
    unsigned log_rows, log_cols;
    split_log(log_m, &log_rows, &log_cols);
    const size_t rows = (size_t)1 << log_rows, cols = (size_t)1 << log_cols, m = (size_t)1 << log_m, loc = m / P;
    if (cols < (size_t)P) return gerr(grp, B200ZK_ERR_ARG, "domain too small for this many GPUs");
    SlotLocks locks(grp);
    std::vector<void*> va(P), vb(P), vc(P), t1(P), t2(P);
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        G_CUDA_OK(grp, ctx->slots[0].io_a.reserve(5 * loc * sizeof(Fr)));
        Fr* p = reinterpret_cast<Fr*>(ctx->slots[0].io_a.p);
        va[g] = p; vb[g] = p + loc; vc[g] = p + 2 * loc; t1[g] = p + 3 * loc; t2[g] = p + 4 * loc;
    }
    int rc = scatter_columns<Blob32>(grp, a, rows, cols, t1.data(), va.data());
    if (!rc) rc = scatter_columns<Blob32>(grp, b, rows, cols, t1.data(), vb.data());
    if (!rc) rc = scatter_columns<Blob32>(grp, c, rows, cols, t1.data(), vc.data());
    if (rc) return rc;
    rc = sharded_h(grp, reinterpret_cast<Fr* const*>(va.data()), reinterpret_cast<Fr* const*>(vb.data()), reinterpret_cast<Fr* const*>(vc.data()),
                   reinterpret_cast<Fr* const*>(t1.data()), reinterpret_cast<Fr* const*>(t2.data()), log_m, reinterpret_cast<Fr* const*>(t2.data()));
    if (rc) return rc;
    rc = gather_columns<Blob32>(grp, h_out, rows, cols, t2.data(), t1.data());
    if (rc) return rc;
    return sync_all(grp);
}

// ---- sharded proving key + prove (BASELINE config 5) ------------------------------------------------------------------------
void b200zk_group_pk_free(b200zk_group* grp, b200zk_group_pk* pk) {
    DeviceGuard dev_guard;
    if (!pk) return;
    for (size_t g = 0; g < pk->shard.size(); ++g) {
        if (grp && g < grp->ctx.size()) cudaSetDevice(grp->ctx[g]->device);
        for (int k = 0; k < 5; ++k) {
            if (pk->shard[g].q[k]) cudaFree(pk->shard[g].q[k]);
            if (pk->shard[g].tab[k]) cudaFree(pk->shard[g].tab[k]);
        }
    }
    if (pk->pk0) b200zk_pk_free(grp ? grp->ctx[0] : nullptr, pk->pk0);
    delete pk;
}

int b200zk_group_pk_upload(b200zk_group* grp, const uint64_t* a_query, const uint64_t* b_g1_query, const uint64_t* b_g2_query,
                           const uint64_t* l_query, const uint64_t* h_query, size_t n_vars, size_t n_inputs, size_t m,
                           const uint64_t* vk_points, b200zk_group_pk** out) {
    DeviceGuard dev_guard;
    if (!grp || !out || !a_query || !b_g1_query || !b_g2_query || !h_query || !vk_points) return B200ZK_ERR_ARG;
    if (n_vars == 0 || n_inputs == 0 || n_inputs > n_vars) return gerr(grp, B200ZK_ERR_ARG, "need 1 <= n_inputs <= n_vars");
    if (m == 0 || (m & (m - 1))) return gerr(grp, B200ZK_ERR_DOMAIN, "h_query length must be a power of two");
    const int P = grp->n;
    const unsigned log_m = ilog2(m);
    unsigned log_rows, log_cols;
    split_log(log_m, &log_rows, &log_cols);
    const size_t rows = (size_t)1 << log_rows, cols = (size_t)1 << log_cols;
    if (cols < (size_t)P) return gerr(grp, B200ZK_ERR_ARG, "domain too small for this many GPUs");
    SlotLocks locks(grp);
    b200zk_group_pk* pk = new b200zk_group_pk();
    pk->n = P; pk->n_vars = n_vars; pk->n_inputs = n_inputs; pk->m = m; pk->log_m = log_m;
    pk->shard.resize(P);
    const size_t n_aux = n_vars - n_inputs;
    const char* tab_env = getenv("B200ZK_PK_TABLES");
    const bool want_tables = !(tab_env && tab_env[0] == '0');
    auto fail = [&](int rc) { b200zk_group_pk_free(grp, pk); return rc; };
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        Slot& sl = ctx->slots[0];
        if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(gerr(grp, B200ZK_ERR_CUDA, "cudaSetDevice failed"));
        b200zk_group_pk::Shard& S = pk->shard[g];
        S.v0 = cut(n_vars, g, P); S.v1 = cut(n_vars, g + 1, P);
        S.l0 = cut(n_aux, g, P); S.l1 = cut(n_aux, g + 1, P);
        S.hn = m / P;
        const size_t cnt[5] = {S.v1 - S.v0, S.v1 - S.v0, S.v1 - S.v0, S.l1 - S.l0, S.hn};
        const size_t psz[5] = {64, 64, 128, 64, 64};
        const void* src[4] = {reinterpret_cast<const char*>(a_query) + S.v0 * 64, reinterpret_cast<const char*>(b_g1_query) + S.v0 * 64,
                              reinterpret_cast<const char*>(b_g2_query) + S.v0 * 128,
                              l_query ? reinterpret_cast<const char*>(l_query) + S.l0 * 64 : nullptr};
        for (int k = 0; k < 5; ++k) {
            if (cudaMalloc(&S.q[k], cnt[k] * psz[k] + 16) != cudaSuccess) return fail(gerr(grp, B200ZK_ERR_OOM, "proving-key shard allocation failed"));
            if (k < 4 && cnt[k] && cudaMemcpyAsync(S.q[k], src[k], cnt[k] * psz[k], cudaMemcpyHostToDevice, sl.stream) != cudaSuccess)
                return fail(gerr(grp, B200ZK_ERR_CUDA, "proving-key upload failed"));
        }
        {   // h_query in the column layout of the h this device will compute: local[c][r] = h_query[r cols + g cg + c]
            const size_t cg = cols / P;
            if (sl.io_a.reserve(S.hn * 64) != cudaSuccess) return fail(gerr(grp, B200ZK_ERR_OOM, "staging allocation failed"));
            if (cudaMemcpy2DAsync(sl.io_a.p, cg * 64, reinterpret_cast<const char*>(h_query) + g * cg * 64, cols * 64, cg * 64, rows,
                                  cudaMemcpyHostToDevice, sl.stream) != cudaSuccess)
                return fail(gerr(grp, B200ZK_ERR_CUDA, "h_query upload failed"));
            int rc = gfrom(grp, g, transpose_dev<Blob64>(ctx, sl.stream, sl.io_a.p, S.q[4], rows, cg));
            if (rc) return fail(rc);
        }
        if (want_tables) {
            // same policy as pk_precompute_dev (prove.cu): window ~ log2(count), all five tables or none, <= 60% of the free HBM
            double max_gb = 48.0;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) max_gb = 0.6 * (double)free_b / 1073741824.0;
            if (const char* env = getenv("B200ZK_PK_TABLE_MAX_GB")) max_gb = atof(env);
            unsigned cs[5];
            size_t total = 0;
            bool ok = true;
            for (int k = 0; k < 5; ++k) {
                cs[k] = msm_table_auto_window(cnt[k]);
                if ((uint64_t)msm_table_windows(cs[k]) * cnt[k] >= (1ull << 31)) ok = false;
                total += (size_t)msm_table_windows(cs[k]) * cnt[k] * psz[k];
            }
            if (ok && (double)total <= max_gb * 1073741824.0) {
                for (int k = 0; k < 5 && ok; ++k) {
                    if (cnt[k] < 64) continue;
                    const size_t bytes = (size_t)msm_table_windows(cs[k]) * cnt[k] * psz[k];
                    if (cudaMalloc(&S.tab[k], bytes) != cudaSuccess) { cudaGetLastError(); S.tab[k] = nullptr; ok = false; break; }
                    int rc = gfrom(grp, g, msm_table_build_dev(ctx, sl, k == 2, S.q[k], cnt[k], cs[k], S.tab[k]));
                    if (rc) return fail(rc);
                    S.tab_c[k] = cs[k];
                    pk->table_bytes += bytes;
                }
                if (!ok) for (int k = 0; k < 5; ++k) { if (S.tab[k]) { cudaFree(S.tab[k]); S.tab[k] = nullptr; } S.tab_c[k] = 0; }
            }
        }
    }
    // device 0 keeps the vk points for the final assembly (queries of the tiny pk object are never read: include_zero_terms = 0)
    {
        b200zk_ctx* c0 = grp->ctx[0];
        cudaSetDevice(c0->device);
        pk->pk0 = new b200zk_pk();
        pk->pk0->n_vars = n_vars; pk->pk0->n_inputs = n_inputs; pk->pk0->m = m;
        if (cudaMalloc(&pk->pk0->vk, 56 * 8) != cudaSuccess ||
            cudaMemcpyAsync(pk->pk0->vk, vk_points, 56 * 8, cudaMemcpyHostToDevice, c0->slots[0].stream) != cudaSuccess)
            return fail(gerr(grp, B200ZK_ERR_CUDA, "vk upload failed"));
    }
    int rc = sync_all(grp);
    if (rc) return fail(rc);
    *out = pk;
    return B200ZK_OK;
}

size_t b200zk_group_pk_table_bytes(const b200zk_group_pk* pk) { return pk ? pk->table_bytes : 0; }

int b200zk_group_groth16_prove(b200zk_group* grp, const b200zk_group_pk* pk, const uint64_t* z, const uint64_t* a, const uint64_t* b,
                               const uint64_t* c, const uint64_t r[4], const uint64_t s[4], uint8_t proof_out[128]) {
    DeviceGuard dev_guard;
    if (!grp || !pk || !z || !a || !b || !c || !r || !s || !proof_out) return B200ZK_ERR_ARG;
    if (pk->n != grp->n) return gerr(grp, B200ZK_ERR_ARG, "proving key was sharded for a different group");
    const int P = grp->n;
    const unsigned log_m = pk->log_m;
    unsigned log_rows, log_cols;
    split_log(log_m, &log_rows, &log_cols);
    const size_t rows = (size_t)1 << log_rows, cols = (size_t)1 << log_cols, m = pk->m, loc = m / P;
    const bool need_b1 = (r[0] | r[1] | r[2] | r[3]) != 0;
    SlotLocks locks(grp);
    // staging per device: a, b, c (column layout) + 2 temporaries (t2 ends up holding h), then this device's rows of z
    std::vector<void*> va(P), vb(P), vc(P), t1(P), t2(P), vz(P);
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        const b200zk_group_pk::Shard& S = pk->shard[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        const size_t zrows = std::max(S.v1 - S.v0, (S.l1 - S.l0));
        G_CUDA_OK(grp, ctx->slots[0].io_a.reserve((5 * loc + (S.v1 - S.v0) + (S.l1 - S.l0) + zrows) * sizeof(Fr)));
        G_CUDA_OK(grp, ctx->slots[0].small.reserve(4096));
        Fr* p = reinterpret_cast<Fr*>(ctx->slots[0].io_a.p);
        va[g] = p; vb[g] = p + loc; vc[g] = p + 2 * loc; t1[g] = p + 3 * loc; t2[g] = p + 4 * loc; vz[g] = p + 5 * loc;
        cudaStream_t st = ctx->slots[0].stream;
        Fr* dz = reinterpret_cast<Fr*>(vz[g]);
        Fr* dzaux = dz + (S.v1 - S.v0);
        G_CUDA_OK(grp, cudaMemcpyAsync(dz, reinterpret_cast<const char*>(z) + S.v0 * 32, (S.v1 - S.v0) * 32, cudaMemcpyHostToDevice, st));
        G_CUDA_OK(grp, cudaMemcpyAsync(dzaux, reinterpret_cast<const char*>(z) + (pk->n_inputs + S.l0) * 32, (S.l1 - S.l0) * 32, cudaMemcpyHostToDevice, st));
    }
    int rc = scatter_columns<Blob32>(grp, a, rows, cols, t1.data(), va.data());
    if (!rc) rc = scatter_columns<Blob32>(grp, b, rows, cols, t1.data(), vb.data());
    if (!rc) rc = scatter_columns<Blob32>(grp, c, rows, cols, t1.data(), vc.data());
    if (rc) return rc;
    rc = sharded_h(grp, reinterpret_cast<Fr* const*>(va.data()), reinterpret_cast<Fr* const*>(vb.data()), reinterpret_cast<Fr* const*>(vc.data()),
                   reinterpret_cast<Fr* const*>(t1.data()), reinterpret_cast<Fr* const*>(t2.data()), log_m, reinterpret_cast<Fr* const*>(t2.data()));
    if (rc) return rc;
    // partial MSMs: small[0..16) A, [16..32) L, [32..48) H, [48..64) B1 (u64 words), [64..96) B2
    for (int g = 0; g < P; ++g) {
        b200zk_ctx* ctx = grp->ctx[g];
        Slot& sl = ctx->slots[0];
        const b200zk_group_pk::Shard& S = pk->shard[g];
        G_CUDA_OK(grp, cudaSetDevice(ctx->device));
        char* sm = reinterpret_cast<char*>(sl.small.p);
        const Fr* dz = reinterpret_cast<const Fr*>(vz[g]);
        const Fr* dzaux = dz + (S.v1 - S.v0);
        auto msm = [&](int k, int g2, const Fr* scalars, size_t n, void* out) -> int {
            if (S.tab_c[k]) return msm_table_dev(ctx, sl, g2, S.tab[k], scalars, n, S.tab_c[k], out);
            return g2 ? msm_g2_dev(ctx, sl, S.q[k], scalars, n, out) : msm_g1_dev(ctx, sl, S.q[k], scalars, n, out);
        };
        G_TRY(grp, g, msm(2, 1, dz, S.v1 - S.v0, sm + 512));                       // the G2 MSM first: the longest
        G_TRY(grp, g, msm(0, 0, dz, S.v1 - S.v0, sm + 0));
        G_TRY(grp, g, msm(3, 0, dzaux, S.l1 - S.l0, sm + 128));
        G_TRY(grp, g, msm(4, 0, reinterpret_cast<const Fr*>(t2[g]), S.hn, sm + 256));
        if (need_b1) G_TRY(grp, g, msm(1, 0, dz, S.v1 - S.v0, sm + 384));
        else G_CUDA_OK(grp, cudaMemsetAsync(sm + 384, 0, 128, sl.stream));
        G_CUDA_OK(grp, cudaEventRecord(grp->ev_done[g], sl.stream));
    }
    // gather on device 0, add up, assemble (the sharded MSMs ran over index 0 too: z[0] = 1 makes a_query[0] z[0] the term the
    // reference's driver adds, sha256.rs:208-212 -> include_zero_terms = 0)
    b200zk_ctx* c0 = grp->ctx[0];
    Slot& s0 = c0->slots[0];
    G_CUDA_OK(grp, cudaSetDevice(c0->device));
    G_CUDA_OK(grp, s0.io_b.reserve((size_t)(P + 1) * 768));
    char* gathered = reinterpret_cast<char*>(s0.io_b.p);
    char* tot = gathered + (size_t)P * 768;
    for (int g = 0; g < P; ++g) {
        if (g) G_CUDA_OK(grp, cudaStreamWaitEvent(s0.stream, grp->ev_done[g], 0));
        G_CUDA_OK(grp, cudaMemcpyPeerAsync(gathered + (size_t)g * 768, c0->device, grp->ctx[g]->slots[0].small.p, grp->ctx[g]->device, 768, s0.stream));
    }
    for (int k = 0; k < 4; ++k) G_TRY(grp, 0, xyzz_sum_dev(c0, s0, 0, gathered + k * 128, P, 6, tot + k * 128));      // stride: 768 B = 6 G1 XYZZ
    G_TRY(grp, 0, xyzz_sum_dev(c0, s0, 1, gathered + 512, P, 3, tot + 512));                                          // 3 G2 XYZZ
    rc = gfrom(grp, 0, assemble_dev(c0, s0, pk->pk0, tot, tot + 512, tot + 128, tot + 256, need_b1 ? tot + 384 : nullptr, r, s, 0, proof_out));
    if (rc) return rc;
    return sync_all(grp);
}

}  // extern "C"
