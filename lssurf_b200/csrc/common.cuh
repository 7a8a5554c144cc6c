// Shared helpers for the lssurf_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/lssurf_b200.h"

namespace lss {

extern thread_local char g_err[512];
int fail(int code, const char *fmt, ...);

#define LSS_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return lss::fail(LSS_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define LSS_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != LSS_OK) return rc__; \
    } while (0)

extern long long g_launches;
}
#define LSS_LAUNCH_CHECK()            \
    do {                              \
        ++lss::g_launches;            \
        LSS_CUDA(cudaGetLastError()); \
    } while (0)
namespace lss {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids of grid-stride kernels are multiples of this

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride kernel: enough CTAs to cover n once, capped at `waves` CTAs per SM.
static inline int stride_grid(int64_t n, int block, int per_sm = 8) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)kSMs * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Stream-ordered scratch allocation (cudaMallocAsync pool); freed in stream order by the destructor.
struct Scratch {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        if (bytes == 0) bytes = 16;
        configure_pool();
        return cudaMallocAsync(&p, bytes, stream);
    }
    // Keep freed scratch in the stream-ordered pool across synchronisations (the default threshold of 0
    // returns it to the driver at every sync, which makes each later allocation a real cudaMalloc).
    static void configure_pool() {
        static int done_for = -1;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            // bounded: scratch beyond this is handed back to the driver at the next synchronisation, so the library does
            // not sit on memory PyTorch's caching allocator could use (cfg 4 keeps ~5 GB of staging alive between fits)
            unsigned long long keep = 24ull << 30;
            if (const char *e = getenv("LSS_POOL_KEEP_GB")) keep = (unsigned long long)atoll(e) << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        done_for = dev;
    }
    template <typename T>
    T *as() { return reinterpret_cast<T *>(p); }
    ~Scratch() {
        if (p) cudaFreeAsync(p, s);
    }
};

// ---- device-side helpers ------------------------------------------------------------------

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Streaming (read-once) loads that do not allocate in L1: keeps L1 for the gathered vector.
__device__ __forceinline__ double ldg_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_stream(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// L2 eviction policies: the matrix stream of a product is read once per CG iteration and is larger than L2
// (evict_first), while the vectors and the preconditioner blocks are re-read every iteration (evict_last).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ldg_stream(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ldg_stream(const int *p, uint64_t pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_keep(const float *p, uint64_t pol) {
    float v;
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

// Exclusive prefix sum of uint32 (device, any n).  `total` (device, optional) receives the sum.
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, cudaStream_t stream);

// Stable LSD radix sort of (key, payload) pairs on key bits [bit_lo, bit_lo + n_bits).
// keys_a/vals_a hold the input and are clobbered; the result lands in *keys_out/*vals_out, which
// point to either the a or the b buffers.  With iota_first the input payload is not read: the first
// pass writes the element's input position instead.
int radix_sort_pairs(uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b, int64_t n, int bit_lo,
                     int n_bits, int iota_first, uint64_t **keys_out, uint32_t **vals_out, cudaStream_t stream);

// y = A x family (spmv.cu), rows [0, n_rows).  A plan caches what depends only on the sparsity structure
// (the first row of every non-zero tile) and the per-tile carry buffers; without one they are rebuilt per call.
struct SpmvPlan {
    int n_tiles;
    int tile;
    cudaStream_t stream;
    int *bounds;
    double *carry_val;
    int *carry_row;
};
int spmv_plan_create(SpmvPlan *plan, int64_t n_rows, const int32_t *indptr, int64_t nnz, cudaStream_t stream);
void spmv_plan_destroy(SpmvPlan *plan);
int spmv_launch(int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data,
                const double *x, double *y, int mode, const double *s, int64_t nnz, const SpmvPlan *plan,
                const double *stop, cudaStream_t stream, int64_t *launches);

int spmv_launch_atomic(int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y,
                       int64_t nnz, const SpmvPlan *plan, const double *stop, cudaStream_t stream, int64_t *launches,
                       double *dot_slots = nullptr,    // y pre-zeroed; dot_slots[256] += partial sums of x . (A x)
                       const int32_t *row_ids = nullptr,    // compacted row list: row r of the CSR is row row_ids[r] of y and x
                       bool accumulate = false);            // every row is added with an atomic (y is shared with another product)

// Products specialised for the explicit normal matrix (long overlapping rows); variant 0 is spmv_launch_atomic.
int normal_product_variant();
void set_normal_product_variant(int v);
int spmv_launch_normal(int variant, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x,
                       double *y, int64_t nnz, const SpmvPlan *plan, const double *stop, cudaStream_t stream, int64_t *launches,
                       double *dot_slots = nullptr);

// variant 10: row-aligned tiles streamed through shared memory by TMA bulk copies (plan per row count)
struct SpmvTmaPlan {
    int64_t n_rows;
    int n_tiles, cap;          // cap: entries a ring stage must hold
    int *tile_row, *tile_k;    // [n_tiles + 1]
    // dictionary form (variant 11): per tile the sorted distinct columns and 16-bit positions in them per entry
    unsigned short *lcol;
    int *dict, *dict_n;
    int dstride, dcap, cap8;
    cudaStream_t stream;
};
int spmv_tma_plan_add_dict(SpmvTmaPlan *plan, const int32_t *indices, int64_t nnz);
int spmv_launch_tmad(const SpmvTmaPlan *plan, const int32_t *indptr, const double *data, const double *x, double *y, const double *stop,
                     cudaStream_t stream, int64_t *launches, double *dot_slots = nullptr);
int spmv_tma_plan_create(SpmvTmaPlan *plan, int64_t n_rows, const int32_t *indptr, int64_t nnz, cudaStream_t stream);
void spmv_tma_plan_destroy(SpmvTmaPlan *plan);
bool spmv_tma_fits(const SpmvTmaPlan *plan);
int spmv_launch_tma(const SpmvTmaPlan *plan, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y,
                    const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots = nullptr);

// variants 7..9: product from the chunk-of-8 padded copy (U = 4, 6, 8 loads per lane)
int spmv_launch_chunk8(int variant, int64_t n_chunks, const int32_t *chunk_row, const int32_t *cols, const double *vals, const double *x,
                       double *y, const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots = nullptr);

// Column-major fixed-row-length ("ELL") copy of the leading rows of a CSR matrix and its product (spmv.cu).
int ell_build(int64_t n_rows, int K, const int32_t *indptr, const int32_t *indices, const double *data, int32_t *ell_col,
              double *ell_val, int64_t n_pad, cudaStream_t stream);
int ell_spmv_launch(int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val, const double *x, double *y,
                    int mode, const double *s, const double *stop, cudaStream_t stream);

// q += G^T (w .* (G x)) (mode 0), q += G^T (w .* b) (mode 1), q += diag(G^T W G) (mode 2) from the ELL copy of G.
int ell_normal_launch(int mode, int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val, const double *x,
                      const double *w, const double *b, double *q, double *pq_slots, const double *stop, cudaStream_t stream);

// Explicit normal matrix N = A^T W A built from the ELL rows of A (normal.cu).  The ELL arrays and the fixed-column
// flags are borrowed; indptr / indices / data / w_cur are owned.
struct NormalMatrix {
    int64_t n_rows, n_pad, n_cols, nnz, n_pairs;
    int K;
    const int32_t *ell_col;
    const double *ell_val;
    const unsigned char *fixed;
    int32_t *indptr, *indices;
    double *data, *w_cur;
    unsigned long long *missing;
    // chunk-of-8 padded copy for the CG product (built at the first fast product, values refreshed after every update)
    int32_t *c8_ptr, *c8_row, *c8_cols;     // chunk start per row [n_cols + 1], row per chunk, padded columns [8 * n_chunks]
    double *c8_vals;
    int64_t n_chunks, c8_part_rows, c8_part_chunks, c8_part_nnz;
    bool c8_dirty;
    SpmvTmaPlan tma_plan;      // built at the first fast product for the row count it is asked for
    int64_t tma_nnz;           // entries of the rows the plan covers
    SpmvPlan plan;
    const unsigned char *owned;   // row-sharded N: flags of this rank's own columns (borrowed; null: all rows)
    int64_t own_e0, own_e1;       // eliminated block whose rows keep their entries at own columns
    SpmvPlan plan_part;        // plan of the leading rows [0, part_rows) (an eliminated trailing block is left out of the product)
    int64_t part_rows, part_nnz;
    // row-sharded N: most rows are empty (columns of other ranks); the product runs over the compacted list of non-empty rows
    int32_t *cp_indptr, *cp_rows;
    int64_t cp_n, cp_nnz, cp_for_rows;
    SpmvPlan plan_cp;
    cudaStream_t stream;
};
extern int g_opt_step_kernel;       // -1: unset (environment LSS_STEP_KERNEL decides, default 2), 0: warp per block, 1: thread per block row, 2: CTA tile
extern int g_opt_explicit_normal;   // -1: unset (environment LSS_NO_EXPLICIT_N decides), 0 / 1: lss_set_option
int normal_create(NormalMatrix **out, int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val,
                  int64_t n_cols, const unsigned char *fixed, cudaStream_t stream, const unsigned char *owned = nullptr,
                  int64_t own_e0 = 0, int64_t own_e1 = 0);
int normal_destroy(NormalMatrix *m);
int normal_update(NormalMatrix *m, const double *w);                 // N += A^T diag(w - w_cur) A ; w_cur <- w
int normal_diag(NormalMatrix *m, double *diag);
int normal_apply(NormalMatrix *m, const double *x, double *y, const double *stop, int64_t *launches);
// same product in one launch (atomic carries) over the leading n_rows rows; y must hold zeros (y_is_zero) or is cleared first
int normal_apply_fast(NormalMatrix *m, int64_t n_rows, const double *x, double *y, bool y_is_zero, const double *stop, int64_t *launches,
                      double *dot_slots = nullptr);

}  // namespace lss

// NCCL communicator wrapper (comm.cu)
constexpr int LSS_MAX_RANKS = 8;
struct lss_comm {
    void *nccl;  // ncclComm_t
    int rank, n_ranks;
    // peer-to-peer mailbox (comm.cu: lss_comm_box_*): one cudaMalloc'ed block per rank, mapped into every other rank of the
    // node through CUDA IPC; kernels exchange data with plain stores over NVLink and sequence-numbered flags
    void *box_local = nullptr;
    void *box_peer[LSS_MAX_RANKS] = {};      // box of rank r as seen from this process (box_peer[rank] == box_local)
    size_t box_bytes = 0;
    unsigned long long seq = 16;             // next free sequence number (monotonic over the life of the communicator)
    bool box_open = false;
};
namespace lss {
int comm_allreduce(lss_comm *c, void *buf, int64_t n, int dtype /*0 f64, 1 i64, 2 u64*/, int op /*0 sum,1 min,2 max*/,
                   cudaStream_t stream);
}
