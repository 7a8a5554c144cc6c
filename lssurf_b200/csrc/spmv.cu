// CSR products y = A x: the explicit normal matrix N of the CG iteration, G for residuals / parse_model, A and A^T of the generic form.
//
// Replaces scipy's csr_matvec behind `.dot` (LSsurf/smooth_fit.py:142, 146, 326-345).  HBM-bound:
// the matrix (12 B per non-zero) is streamed exactly once with coalesced, L1-bypassing loads;
// x is gathered through L1/L2 (it is a few MB: resident in the 126 MB L2).
//
// Work decomposition is by NON-ZEROS, not rows: CTA t owns non-zeros [t*TILE, (t+1)*TILE), so the
// same kernel is balanced for G (12-13 entries per row), for the constraint stencils (2-8 per
// row) and for G^T, whose rows range from empty (most grid columns between tracks see no data) to
// >10^4 entries (bias columns).  Products are staged in shared memory, rows are then reduced
// from shared memory by one thread (short rows) or one warp (long rows) each.  Rows cut by a tile
// boundary are completed by a tiny fix-up kernel from per-tile carries, so the result is
// deterministic (no atomics).  The first row of every tile depends only on the sparsity structure
// and is computed once per matrix (SpmvPlan), not searched for by every CTA of every product.
#include <stdlib.h>

#include "common.cuh"

namespace lss {

// (threads per CTA, non-zeros per thread); selectable at run time for tuning (lss_set_option("spmv_variant")).
struct SpmvShape { int block, ipt; };
static const SpmvShape kShapes[] = {{256, 8}, {128, 8}, {128, 16}, {256, 4}, {512, 4}, {256, 16}, {128, 4},
                                    {128, 8}, {128, 8}, {256, 8}, {128, 4}};   // 7..10: persistent kernels
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);
static int g_variant = -1;

static int current_variant() {
    if (g_variant < 0) {
        const char *e = getenv("LSS_SPMV_VARIANT");
        g_variant = e ? atoi(e) : 7;   // persistent 128 x 8 (tools/spmv_bench.py, profiles/spmv_variants_r01.md)
        if (g_variant < 0 || g_variant >= kNumShapes) g_variant = 0;
    }
    return g_variant;
}
int spmv_tile_size() {
    const SpmvShape &s = kShapes[current_variant()];
    return s.block * s.ipt;
}

__device__ __forceinline__ double epilogue(double v, int mode, const double *__restrict__ s, int64_t r) {
    if (mode == LSS_SPMV_SCALE) return s[r] * v;
    if (mode == LSS_SPMV_RESIDUAL) return s[r] - v;
    return v;
}

// first r in [0, n] with a[r] >= key  (a is non-decreasing and readable up to a[n])
__device__ __forceinline__ int64_t lower_bound_i32(const int *__restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n + 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Row range touched by tile t (non-zeros [t*tile, min((t+1)*tile, nnz))): bounds[2t] = the row holding non-zero t*tile
// (first row whose end lies beyond it), bounds[2t+1] = one past the row holding the tile's last non-zero.  Empty rows
// before / after that range belong to no tile (they are pre-filled), so a tile never walks a long block of empty rows
// (A^T of a row shard has >10^5 consecutive empty rows).
__global__ void __launch_bounds__(256)
spmv_tile_bounds_kernel(int64_t n_rows, int64_t nnz, const int *__restrict__ indptr, int n_tiles, int tile, int *__restrict__ bounds) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t k0 = (int64_t)t * tile;
    const int64_t k1 = (k0 + tile < nnz) ? k0 + tile : nnz;
    // first r in [0, n_rows) with indptr[r + 1] > k0   /   first r with indptr[r + 1] >= k1
    bounds[2 * t] = (int)lower_bound_i32(indptr + 1, n_rows - 1, k0 + 1);
    bounds[2 * t + 1] = (int)lower_bound_i32(indptr + 1, n_rows - 1, k1) + 1;
}

// ATOMIC (plain products into a pre-zeroed y only): partial sums of rows cut by a tile boundary are added with RED
// atomics instead of going through the carry buffers and the fix-up kernel; complete rows are stored.
template <int BLOCK, bool ATOMIC = false>
__device__ __forceinline__ void reduce_rows(const double *__restrict__ prod, int tile, int64_t k0, int64_t k1, int64_t r_lo, int64_t r_hi,
                                            const int *__restrict__ indptr, double *__restrict__ y, int mode,
                                            const double *__restrict__ s, double *__restrict__ carry_val, int *__restrict__ carry_row,
                                            const double *__restrict__ xdot = nullptr, double *__restrict__ dot_slots = nullptr,
                                            const int *__restrict__ row_ids = nullptr) {
    // row_ids (ATOMIC only): the CSR holds a compacted list of rows, row r of it is row row_ids[r] of y / x
    double dot = 0.0;      // ATOMIC only: this thread's share of x . (A x), summed per warp into dot_slots
    // carry-in: the row that started in an earlier tile and still has entries here
    const bool has_carry = (int64_t)indptr[r_lo] < k0;                // r_lo itself started in an earlier tile
    const int64_t first = r_lo + (has_carry ? 1 : 0);                 // first row that starts in this tile
    const int64_t n_virtual = (r_hi - first) + 1;                     // virtual row 0 is the carry-in
    // g lanes per row, g the largest power of two with n_virtual * g <= BLOCK: one thread per row for the 13-entry rows
    // of G (~80 rows per tile), 8 lanes per row for the ~90-entry rows of the explicit normal matrix, a warp per row
    // for the long rows of A^T.  Lanes of a group stride the row's staged products, then fold with g-wide shuffles.
    int g = 1;
    while (g < 32 && n_virtual * (2 * g) <= BLOCK) g <<= 1;
    const int gl = threadIdx.x & (g - 1), rows_per_pass = BLOCK / g;
    for (int64_t base = 0; base < n_virtual; base += rows_per_pass) {
        const int64_t v = base + threadIdx.x / g;
        const bool valid = v < n_virtual && !(v == 0 && !has_carry);
        int64_t b = 0, e = 0;
        const int64_t r = first + v - 1, rr = v == 0 ? r_lo : r;
        if (valid) {
            if (v == 0) {
                b = k0;
                e = (int64_t)indptr[r_lo + 1];
            } else {
                b = indptr[r];
                e = indptr[r + 1];
            }
        }
        const bool open = e > k1;   // continues in the next tile
        if (open) e = k1;
        const int64_t ro = (ATOMIC && row_ids && valid && gl == 0) ? (int64_t)__ldg(row_ids + rr) : rr;
        const double xr = (ATOMIC && xdot && valid && gl == 0) ? __ldg(xdot + ro) : 0.0;   // in flight during the sum
        double acc = 0.0;
        for (int64_t k = b - k0 + gl; k < e - k0; k += g) acc += prod[k];
        for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && gl == 0) {
            if (ATOMIC) {
                if (v == 0 || open || mode == LSS_SPMV_ACCUM) atomicAdd(y + ro, acc);
                else y[ro] = acc;
                dot += acc * xr;
            } else if (v == 0) carry_val[tile] = acc;
            else y[r] = open ? acc : epilogue(acc, mode, s, r);
        }
    }
    if (ATOMIC && xdot) {          // x . (A x) as a by-product (partial sums of cut rows add up by linearity)
        dot = warp_sum(dot);
        if (lane_id() == 0 && dot != 0.0) atomicAdd(dot_slots + (((unsigned)tile * (BLOCK / 32) + (threadIdx.x >> 5)) & 255u), dot);
    }
    if (!ATOMIC && threadIdx.x == 0) carry_row[tile] = has_carry ? (int)r_lo : -1;
}

template <int BLOCK, int IPT>
__global__ void __launch_bounds__(BLOCK)
spmv_tile_kernel(int64_t n_rows, int64_t nnz, const int *__restrict__ indptr, const int *__restrict__ indices,
                 const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y, int mode,
                 const double *__restrict__ s, double *__restrict__ carry_val, int *__restrict__ carry_row,
                 const int *__restrict__ tile_bounds, const double *__restrict__ stop) {
    constexpr int TILE = BLOCK * IPT;
    __shared__ double prod[TILE];
    const int64_t k0 = (int64_t)blockIdx.x * TILE;
    const int64_t k1 = (k0 + TILE < nnz) ? k0 + TILE : nnz;
    // issued first, consumed after the staging loads are in flight (no serialised round trip per CTA)
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    const int64_t r_lo = tile_bounds[2 * blockIdx.x], r_hi = tile_bounds[2 * blockIdx.x + 1];
    // stage the products: coalesced streaming loads of the matrix, gathered loads of x
    {
        int col[IPT];
        double val[IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int64_t k = k0 + j * BLOCK + threadIdx.x;
            const bool in = k < k1;
            col[j] = in ? ldg_stream(indices + k) : 0;
            val[j] = in ? ldg_stream(data + k) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < IPT; ++j) prod[j * BLOCK + threadIdx.x] = val[j] * __ldg(x + col[j]);
    }
    if (stop_v != 0.0) return;   // solver converged: the queued tail of the batch is a no-op (uniform per launch)
    __syncthreads();
    reduce_rows<BLOCK>(prod, blockIdx.x, k0, k1, r_lo, r_hi, indptr, y, mode, s, carry_val, carry_row);
}

// Persistent variant: CTAs loop over tiles (static round robin) and prefetch the NEXT tile's matrix stream
// (column indices + values, the HBM traffic) into registers before reducing the current tile's rows, so the
// DRAM pipe stays busy during the shared-memory phase.  DBUF double-buffers the staged products (one barrier
// per tile instead of two).
template <int BLOCK, int IPT, bool DBUF, bool ATOMIC = false>
__global__ void __launch_bounds__(BLOCK)
spmv_persistent_kernel(int64_t n_rows, int64_t nnz, int n_tiles, const int *__restrict__ indptr, const int *__restrict__ indices,
                       const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y, int mode,
                       const double *__restrict__ s, double *__restrict__ carry_val, int *__restrict__ carry_row,
                       const int *__restrict__ tile_bounds, const double *__restrict__ stop, double *__restrict__ dot_slots = nullptr,
                       const int *__restrict__ row_ids = nullptr) {
    constexpr int TILE = BLOCK * IPT;
    __shared__ double prod_s[DBUF ? 2 : 1][TILE];
    if (stop && *stop != 0.0) return;
    int col[IPT];
    double val[IPT];
    int t = blockIdx.x;
    if (t >= n_tiles) return;
    const uint64_t once = l2_policy_evict_first();     // the matrix does not fit in L2: leave L2 to x, y and the preconditioner
    auto fetch = [&](int tile) {
        const int64_t k0 = (int64_t)tile * TILE;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int64_t k = k0 + j * BLOCK + threadIdx.x;
            const bool in = k < nnz;
            col[j] = in ? ldg_stream(indices + k, once) : 0;
            val[j] = in ? ldg_stream(data + k, once) : 0.0;
        }
    };
    fetch(t);
    int buf = 0;
    for (; t < n_tiles; t += gridDim.x) {
        const int64_t k0 = (int64_t)t * TILE;
        const int64_t k1 = (k0 + TILE < nnz) ? k0 + TILE : nnz;
        const int64_t r_lo = tile_bounds[2 * t], r_hi = tile_bounds[2 * t + 1];
        double *prod = prod_s[buf];
        double xg[IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) xg[j] = __ldg(x + col[j]);
#pragma unroll
        for (int j = 0; j < IPT; ++j) prod[j * BLOCK + threadIdx.x] = val[j] * xg[j];
        const int nt = t + gridDim.x;
        if (nt < n_tiles) fetch(nt);            // in flight during the barrier and the row reduction below
        __syncthreads();
        reduce_rows<BLOCK, ATOMIC>(prod, t, k0, k1, r_lo, r_hi, indptr, y, mode, s, carry_val, carry_row, dot_slots ? x : nullptr, dot_slots, row_ids);
        if (DBUF) buf ^= 1; else __syncthreads();
    }
}

// Completes rows cut by tile boundaries: the first carry tile of each such row adds the carries of
// the whole chain in tile order and applies the epilogue.
__global__ void __launch_bounds__(256)
spmv_fixup_kernel(int n_tiles, const double *__restrict__ carry_val, const int *__restrict__ carry_row, double *__restrict__ y,
                  int mode, const double *__restrict__ s, const double *__restrict__ stop) {
    if (stop && *stop != 0.0) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int r = carry_row[t];
    if (r < 0) return;
    if (t > 0 && carry_row[t - 1] == r) return;   // not the head of the chain
    double acc = y[r];
    for (int u = t; u < n_tiles && carry_row[u] == r; ++u) acc += carry_val[u];
    y[r] = epilogue(acc, mode, s, r);
}

__global__ void __launch_bounds__(256)
spmv_empty_kernel(int64_t n_rows, double *__restrict__ y, int mode, const double *__restrict__ s, const double *__restrict__ stop) {
    if (stop && *stop != 0.0) return;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        y[r] = epilogue(0.0, mode, s, r);
}

int spmv_plan_create(SpmvPlan *plan, int64_t n_rows, const int32_t *indptr, int64_t nnz, cudaStream_t stream) {
    plan->tile = spmv_tile_size();
    plan->n_tiles = nnz > 0 ? ceil_div(nnz, plan->tile) : 0;
    plan->bounds = nullptr;
    plan->carry_val = nullptr;
    plan->carry_row = nullptr;
    if (plan->n_tiles == 0) return LSS_OK;
    plan->stream = stream;
    Scratch::configure_pool();
    LSS_CUDA(cudaMallocAsync((void **)&plan->bounds, sizeof(int) * 2 * plan->n_tiles, stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->carry_val, sizeof(double) * plan->n_tiles, stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->carry_row, sizeof(int) * plan->n_tiles, stream));
    spmv_tile_bounds_kernel<<<ceil_div(plan->n_tiles, 256), 256, 0, stream>>>(n_rows, nnz, indptr, plan->n_tiles, plan->tile, plan->bounds);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

void spmv_plan_destroy(SpmvPlan *plan) {
    if (plan->bounds) cudaFreeAsync(plan->bounds, plan->stream);
    if (plan->carry_val) cudaFreeAsync(plan->carry_val, plan->stream);
    if (plan->carry_row) cudaFreeAsync(plan->carry_row, plan->stream);
    plan->bounds = nullptr;
    plan->carry_val = nullptr;
    plan->carry_row = nullptr;
}

template <int BLOCK, int IPT>
static void launch_tile(int n_tiles, cudaStream_t stream, int64_t n_rows, int64_t nnz, const int *indptr, const int *indices,
                        const double *data, const double *x, double *y, int mode, const double *s, double *cv, int *cr,
                        const int *bounds, const double *stop) {
    spmv_tile_kernel<BLOCK, IPT><<<n_tiles, BLOCK, 0, stream>>>(n_rows, nnz, indptr, indices, data, x, y, mode, s, cv, cr, bounds, stop);
}

template <int BLOCK, int IPT, bool DBUF>
static void launch_persistent(int n_tiles, cudaStream_t stream, int64_t n_rows, int64_t nnz, const int *indptr, const int *indices,
                              const double *data, const double *x, double *y, int mode, const double *s, double *cv, int *cr,
                              const int *bounds, const double *stop) {
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_persistent_kernel<BLOCK, IPT, DBUF>, BLOCK, 0);
        if (per_sm < 1) per_sm = 1;
    }
    const int grid = n_tiles < kSMs * per_sm ? n_tiles : kSMs * per_sm;   // one resident wave: a multiple of the SM count
    spmv_persistent_kernel<BLOCK, IPT, DBUF><<<grid, BLOCK, 0, stream>>>(n_rows, nnz, n_tiles, indptr, indices, data, x, y, mode, s,
                                                                         cv, cr, bounds, stop);
}

int spmv_launch(int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x,
                double *y, int mode, const double *s, int64_t nnz, const SpmvPlan *plan, const double *stop,
                cudaStream_t stream, int64_t *launches) {
    if (n_rows <= 0) return LSS_OK;
    if (nnz == 0) {
        spmv_empty_kernel<<<stride_grid(n_rows, 256), 256, 0, stream>>>(n_rows, y, mode, s, nullptr);
        LSS_LAUNCH_CHECK();
        if (launches) *launches += 1;
        return LSS_OK;
    }
    const int variant = current_variant();
    const int tile = kShapes[variant].block * kShapes[variant].ipt;
    const int n_tiles = ceil_div(nnz, tile);
    Scratch cv, cr, tb;
    double *carry_val;
    int *carry_row, *bounds;
    if (plan && plan->bounds && plan->tile == tile) {
        carry_val = plan->carry_val; carry_row = plan->carry_row; bounds = plan->bounds;
    } else {
        LSS_CUDA(cv.alloc(sizeof(double) * n_tiles, stream));
        LSS_CUDA(cr.alloc(sizeof(int) * n_tiles, stream));
        LSS_CUDA(tb.alloc(sizeof(int) * 2 * n_tiles, stream));
        carry_val = cv.as<double>(); carry_row = cr.as<int>(); bounds = tb.as<int>();
        spmv_tile_bounds_kernel<<<ceil_div(n_tiles, 256), 256, 0, stream>>>(n_rows, nnz, indptr, n_tiles, tile, bounds);
        LSS_LAUNCH_CHECK();
        if (launches) *launches += 1;
    }
    // rows without entries (and nothing else) keep this value: y = epilogue(0)
    spmv_empty_kernel<<<stride_grid(n_rows, 256), 256, 0, stream>>>(n_rows, y, mode, s, stop);
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
#define LSS_SPMV_CASE(V, B, I)                                                                                     \
    case V: launch_tile<B, I>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
    switch (variant) {
        LSS_SPMV_CASE(0, 256, 8)
        LSS_SPMV_CASE(1, 128, 8)
        LSS_SPMV_CASE(2, 128, 16)
        LSS_SPMV_CASE(3, 256, 4)
        LSS_SPMV_CASE(4, 512, 4)
        LSS_SPMV_CASE(5, 256, 16)
        case 6: launch_tile<128, 4>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
        case 7: launch_persistent<128, 8, false>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
        case 8: launch_persistent<128, 8, true>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
        case 9: launch_persistent<256, 8, false>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
        default: launch_persistent<128, 4, true>(n_tiles, stream, n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop); break;
    }
#undef LSS_SPMV_CASE
    LSS_LAUNCH_CHECK();
    spmv_fixup_kernel<<<ceil_div(n_tiles, 256), 256, 0, stream>>>(n_tiles, carry_val, carry_row, y, mode, s, stop);
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 2;
    return LSS_OK;
}


// dot_slots (optional, 256 doubles, accumulated into): x . (A x) as a by-product, one RED per warp and tile.
// y += A x in ONE launch for a y that holds zeros (at least in the rows cut by tile boundaries and the empty rows):
// the persistent kernel with atomic carries.  Used for the explicit normal matrix inside CG, where the vector step
// that consumes q clears it for the next product.
int spmv_launch_atomic(int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y,
                       int64_t nnz, const SpmvPlan *plan, const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots,
                       const int32_t *row_ids, bool accumulate) {
    if (n_rows <= 0 || nnz == 0) return LSS_OK;
    constexpr int BLOCK = 128, IPT = 8;
    if (!plan || !plan->bounds || plan->tile != BLOCK * IPT) return fail(LSS_EINVAL, "spmv_launch_atomic: needs a plan with %d-entry tiles", BLOCK * IPT);
    const int n_tiles = ceil_div(nnz, BLOCK * IPT);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_persistent_kernel<BLOCK, IPT, false, true>, BLOCK, 0);
        if (per_sm < 1) per_sm = 1;
    }
    const int grid = n_tiles < kSMs * per_sm ? n_tiles : kSMs * per_sm;
    spmv_persistent_kernel<BLOCK, IPT, false, true><<<grid, BLOCK, 0, stream>>>(n_rows, nnz, n_tiles, indptr, indices, data, x, y,
                                                                                accumulate ? LSS_SPMV_ACCUM : LSS_SPMV_PLAIN,
                                                                                nullptr, nullptr, nullptr, plan->bounds, stop, dot_slots, row_ids);
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
    return LSS_OK;
}

// ============================================================================ products for the explicit normal matrix
// The explicit normal matrix has long, uneven rows (mean 80, 0..277 entries for the 10^7-point workload) whose column
// sets overlap heavily between neighbouring rows (neighbouring grid nodes).  The nnz-tiled kernel above (variant 0) is
// still the one the solver uses: 0.092 ms per product, 0.61 of the measured HBM peak.  The kernels below are the
// alternatives tried in round 1, all parity-tested (tests/test_gpu_parity.py) and selectable with
// lss_set_option("normal_product", v) / LSS_N_PRODUCT for profiling; profiles/README.md has their timings and what ncu
// says limits each of them (none is HBM-bound: L1 wavefronts of the gathers, issue slots, or per-tile latency chains).
//
//   1, 5, 6  rows    G lanes per row, everything in registers: each lane loads up to U (column, value) pairs of its row
//                    with streaming loads, gathers x, folds with G-wide shuffles and stores y[row].
//   7, 8, 9  chunk8  padded copy of N in chunks of 8 entries with one row id per chunk: no load address depends on a
//                    row pointer; 8-lane sums merged across equal rows, one RED per (warp load, row).
//   10       tma     row-aligned tiles streamed into a shared-memory ring by TMA bulk copies, rows reduced from the ring.
//   11       tmad    tma with a per-tile column dictionary: x is gathered once per distinct column of a tile, the row
//                    sums run from shared memory only, 10.3 instead of 12 bytes per entry.
template <int G, int U>
__global__ void __launch_bounds__(256)
spmv_rows_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                 const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ stop, double *__restrict__ dot_slots) {
    if (stop && *stop != 0.0) return;
    const uint64_t once = l2_policy_evict_first();
    const int gl = threadIdx.x & (G - 1);
    const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
    int b = 0, e = 0;
    if (row < n_rows) {
        b = __ldg(indptr + row);
        e = __ldg(indptr + row + 1);
    }
    double acc = 0.0;
    for (int k0 = b; k0 < e; k0 += G * U) {
        int col[U];
        double val[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int k = k0 + j * G + gl;
            const bool in = k < e;
            col[j] = in ? ldg_stream(indices + k, once) : -1;
            val[j] = in ? ldg_stream(data + k, once) : 0.0;
        }
        double xg[U];
#pragma unroll
        for (int j = 0; j < U; ++j) xg[j] = col[j] >= 0 ? __ldg(x + col[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < U; ++j) acc += val[j] * xg[j];
    }
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    double dot = 0.0;
    if (row < n_rows && gl == 0) {
        y[row] = acc;
        if (dot_slots) dot = acc * __ldg(x + row);
    }
    if (dot_slots) {
        dot = warp_sum(dot);
        if (lane_id() == 0 && dot != 0.0) atomicAdd(dot_slots + ((blockIdx.x * 8u + (threadIdx.x >> 5)) & 255u), dot);
    }
}

//   chunk8  The row kernels above (and the nnz-tiled one) start every row or tile with a chain of dependent loads
//           (tile bounds -> row pointers -> matrix stream -> gather).  Here the row structure is taken out of the
//           address computation altogether: a padded copy of N keeps every row in chunks of 8 entries (zero padding,
//           +4 %; one row id per chunk), so the address of every matrix load depends on the thread index only.  A warp
//           owns 4U consecutive chunks: U perfectly coalesced loads of 32 (column, value) pairs per lane are issued
//           back to back, then the gathers, then 8-lane sums; chunks of the same row inside one load are merged with
//           two more shuffles and the sum is added to y[row] with one RED (y must hold zeros).
template <int U>
__global__ void __launch_bounds__(256)
spmv_chunk8_kernel(int64_t n_chunks, const int *__restrict__ chunk_row, const int *__restrict__ cols, const double *__restrict__ vals,
                   const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ stop, double *__restrict__ dot_slots) {
    static_assert(4 * U <= 32, "one lane per chunk row id");
    const unsigned lane = lane_id();
    const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t c0 = warp * (4 * U);
    if (c0 >= n_chunks) return;
    const uint64_t once = l2_policy_evict_first();
    int col[U];
    double val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t c = c0 + 4 * u + (lane >> 3);
        const int64_t s = (c0 + 4 * u) * 8 + lane;
        const bool in = c < n_chunks;
        col[u] = in ? ldg_stream(cols + s, once) : -1;
        val[u] = in ? ldg_stream(vals + s, once) : 0.0;
    }
    const int my_row = ((int)lane < 4 * U && c0 + lane < n_chunks) ? __ldg(chunk_row + c0 + lane) : -1;
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    const double my_xr = (dot_slots && my_row >= 0) ? __ldg(x + my_row) : 0.0;       // x[row] for the x . (A x) by-product
    // ptxas is free to sink the (read-only) value loads below the gathers, where they would sit behind the wait for the
    // column loads: make the gather base depend on every value (a NaN pattern no matrix entry holds), so all 2U matrix
    // loads are issued before the first wait
    int dep = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) dep |= (__double2hiint(val[u]) == (int)0xffffffffu) ? 1 : 0;
    const double *xd = x + dep;
    double xg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xg[u] = col[u] >= 0 ? __ldg(xd + col[u]) : 0.0;
    if (stop_v != 0.0) return;
    double dot = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        double p = val[u] * xg[u];
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 4);
        const int r = __shfl_sync(0xffffffffu, my_row, 4 * u + (int)(lane >> 3));
        const double xr = __shfl_sync(0xffffffffu, my_xr, 4 * u + (int)(lane >> 3));
        // chunk rows are non-decreasing: merge the sums of equal rows (groups 0|1, 2|3, then 0|2)
        double o = __shfl_xor_sync(0xffffffffu, p, 8);
        int ro = __shfl_xor_sync(0xffffffffu, r, 8);
        int dead = 0;
        if (ro == r) {
            if (lane & 8u) dead = 1; else p += o;
        }
        o = __shfl_xor_sync(0xffffffffu, p, 16);
        ro = __shfl_xor_sync(0xffffffffu, r, 16);
        const int odead = __shfl_xor_sync(0xffffffffu, dead, 16);
        if (ro == r && !dead && !odead) {
            if (lane & 16u) dead = 1; else p += o;
        }
        if (!dead && (lane & 7u) == 0u && r >= 0) {
            atomicAdd(y + r, p);
            dot += p * xr;
        }
    }
    if (dot_slots) {
        dot = warp_sum(dot);
        if (lane == 0 && dot != 0.0) atomicAdd(dot_slots + ((unsigned)warp & 255u), dot);
    }
}

int spmv_launch_chunk8(int variant, int64_t n_chunks, const int32_t *chunk_row, const int32_t *cols, const double *vals, const double *x,
                       double *y, const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots) {
    if (n_chunks <= 0) return LSS_OK;
    if (variant == 7) {
        spmv_chunk8_kernel<4><<<ceil_div(ceil_div(n_chunks, 16), 8), 256, 0, stream>>>(n_chunks, chunk_row, cols, vals, x, y, stop, dot_slots);
    } else if (variant == 9) {
        spmv_chunk8_kernel<8><<<ceil_div(ceil_div(n_chunks, 32), 8), 256, 0, stream>>>(n_chunks, chunk_row, cols, vals, x, y, stop, dot_slots);
    } else {
        spmv_chunk8_kernel<6><<<ceil_div(ceil_div(n_chunks, 24), 8), 256, 0, stream>>>(n_chunks, chunk_row, cols, vals, x, y, stop, dot_slots);
    }
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
    return LSS_OK;
}

//   tma     row-aligned tiles streamed by the TMA engine.  Tile t holds the rows whose first entry lies in
//           [1024 t, 1024 (t + 1)), so no row is cut (plain stores: y needs no clearing, no atomics, deterministic) and
//           a tile's values, column indices and row pointers are three contiguous byte ranges.  One thread per CTA
//           issues them as 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a STAGES-deep shared-memory
//           ring, several tiles ahead: the stream costs no registers and no LSU wavefronts, and the descriptors of
//           the tile after next are prefetched so that no load address waits on another load.  Consumers wait on the
//           stage's mbarrier and reduce rows straight from shared memory with g lanes per row (neighbouring rows side
//           by side in a warp: their gathers share lines), one shuffle fold per row.
constexpr int TMA_TILE = 1024, TMA_RP_CAP = 132, TMA_HEAD = 512;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// tile_row[t] = first row whose entries start at or after 1024 t (tile_row[n_tiles] = n_rows), tile_k[t] = indptr[tile_row[t]];
// *max_span = largest number of entries a stage must hold (from the 4-aligned start of the tile)
__global__ void __launch_bounds__(256)
spmv_tma_plan_kernel(int64_t n_rows, const int *__restrict__ indptr, int n_tiles, int *__restrict__ tile_row, int *__restrict__ tile_k) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const int r = t == n_tiles ? (int)n_rows : (int)lower_bound_i32(indptr, n_rows, (int64_t)t * TMA_TILE);
    tile_row[t] = r;
    tile_k[t] = indptr[r];
}
__global__ void __launch_bounds__(256)
spmv_tma_span_kernel(int n_tiles, const int *__restrict__ tile_k, int *__restrict__ max_span) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    atomicMax(max_span, tile_k[t + 1] - (tile_k[t] & ~3));
}

template <int BLOCK, int STAGES>
__global__ void __launch_bounds__(BLOCK)
spmv_tma_kernel(int64_t n_rows, int n_tiles, int cap, const int *__restrict__ indptr, const int *__restrict__ indices,
                const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ tile_row,
                const int *__restrict__ tile_k, const double *__restrict__ stop, double *__restrict__ dot_slots) {
    extern __shared__ __align__(128) unsigned char tma_ring[];
    unsigned char *ring = tma_ring;
    // [STAGES mbarriers | STAGES x 8 descriptor ints | STAGES x { cap doubles | cap ints | TMA_RP_CAP ints }]
    static_assert(STAGES <= 8, "header layout");
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ring);          // [0, 64): mbarriers
    int *desc = reinterpret_cast<int *>(ring + 64);                                    // [64, 320): 8 ints per stage
    unsigned char *stages = ring + TMA_HEAD;
    const size_t stage_bytes = (size_t)cap * 12 + TMA_RP_CAP * 4;
    if (stop && *stop != 0.0) return;
    const uint64_t once = l2_policy_evict_first();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0 only: descriptor of the next tile to issue, loaded one iteration before it is needed
    int d_row0 = 0, d_row1 = 0, d_k0 = 0, d_k1 = 0;
    auto load_desc = [&](int tile) {
        if (tile < n_tiles) {
            d_row0 = __ldg(tile_row + tile);
            d_row1 = __ldg(tile_row + tile + 1);
            d_k0 = __ldg(tile_k + tile);
            d_k1 = __ldg(tile_k + tile + 1);
        }
    };
    auto issue = [&](int tile, int stage) {          // thread 0; uses the descriptor registers
        if (tile >= n_tiles) return;
        unsigned char *base = stages + (size_t)stage * stage_bytes;
        const int k0a = d_k0 & ~3, nk = d_k1 - k0a;
        const int r0a = d_row0 & ~3;
        int nrp = d_row1 - r0a + 1;
        if (nrp > TMA_RP_CAP) nrp = TMA_RP_CAP;
        const unsigned b_val = (unsigned)((nk + 1) & ~1) * 8u, b_col = (unsigned)((nk + 3) & ~3) * 4u, b_rp = (unsigned)((nrp + 3) & ~3) * 4u;
        int *dd = desc + stage * 8;
        dd[0] = d_row0; dd[1] = d_row1; dd[2] = k0a; dd[3] = r0a;
        mbar_arrive_expect_tx(full + stage, b_val + b_col + b_rp);
        if (b_val) bulk_g2s(base, data + k0a, b_val, full + stage, once);
        if (b_col) bulk_g2s(base + (size_t)cap * 8, indices + k0a, b_col, full + stage, once);
        bulk_g2s(base + (size_t)cap * 12, indptr + r0a, b_rp, full + stage, once);
    };
    const int t0 = blockIdx.x, dt = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            load_desc(t0 + s * dt);
            issue(t0 + s * dt, s);
        }
        load_desc(t0 + (STAGES - 1) * dt);
    }
    double dot = 0.0;
    int i = 0;
    for (int t = t0; t < n_tiles; t += dt, ++i) {
        const int stage = i % STAGES;
        if (threadIdx.x == 0) {
            issue(t + (STAGES - 1) * dt, (i + STAGES - 1) % STAGES);       // the stage tile i - 1 was reduced from
            load_desc(t + STAGES * dt);
        }
        const unsigned parity = (unsigned)(i / STAGES) & 1u;
        while (!mbar_try_wait(full + stage, parity)) {}
        const unsigned char *base = stages + (size_t)stage * stage_bytes;
        const double *sval = reinterpret_cast<const double *>(base);
        const int *scol = reinterpret_cast<const int *>(base + (size_t)cap * 8);
        const int *srp = reinterpret_cast<const int *>(base + (size_t)cap * 12);
        const int *dd = desc + stage * 8;
        const int row0 = dd[0], row1 = dd[1], k0a = dd[2], rp_off = row0 - dd[3];
        const int nr = row1 - row0;
        int g = 1;
        while (g < 32 && nr * (2 * g) <= BLOCK) g <<= 1;
        const int gl = threadIdx.x & (g - 1), n_groups = BLOCK / g;
        for (int rr = threadIdx.x / g; rr < nr; rr += n_groups) {       // rr is uniform inside a group; g divides 32
            const int row = row0 + rr;
            const int j = rr + rp_off;
            const int b = (j < TMA_RP_CAP ? srp[j] : __ldg(indptr + row)) - k0a;
            const int e = (j + 1 < TMA_RP_CAP ? srp[j + 1] : __ldg(indptr + row + 1)) - k0a;
            const double xr = (dot_slots && gl == 0) ? __ldg(x + row) : 0.0;
            // all gathers of a row are issued in ONE round (rows of up to 12 g entries): a second round would wait for
            // the first one's products, and the rounds of a tile are what its CTA's time consists of
            constexpr int UG = 12;
            double acc = 0.0;
            for (int k = b + gl; k < e; k += UG * g) {
                // branch-free: slots past the row's end re-read the lane's first entry with a zero value
                int c[UG];
                double v[UG];
#pragma unroll
                for (int j = 0; j < UG; ++j) {
                    const bool in = k + j * g < e;
                    const int idx = in ? k + j * g : k;
                    c[j] = scol[idx];
                    v[j] = sval[idx];
                    if (!in) v[j] = 0.0;
                }
                double xv[UG];
#pragma unroll
                for (int j = 0; j < UG; ++j) xv[j] = __ldg(x + c[j]);
                // ptxas would otherwise put product j (which waits for gather j) in front of gather j + 1: the running
                // sum is made to depend on every gathered word (a NaN pattern x never holds), so all gathers are issued
                // before the first product
                int dep = 0;
#pragma unroll
                for (int j = 0; j < UG; ++j) dep |= (__double2hiint(xv[j]) == (int)0xffffffffu) ? 1 : 0;
                acc = __hiloint2double(__double2hiint(acc) ^ dep, __double2loint(acc));
#pragma unroll
                for (int j = 0; j < UG; ++j) acc += v[j] * xv[j];
            }
            // fold inside the group: groups are aligned power-of-two lane ranges; rows of one warp may have
            // different trip counts, so name the participating lanes explicitly
            const unsigned lane = lane_id();
            const unsigned gmask = g == 32 ? 0xffffffffu : (((1u << g) - 1u) << (lane & ~(unsigned)(g - 1)));
            for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
            if (gl == 0) {
                y[row] = acc;
                dot += acc * xr;
            }
        }
        __syncthreads();                                                   // the stage is free for the copy issued next
    }
    if (dot_slots) {
        dot = warp_sum(dot);
        if (lane_id() == 0 && dot != 0.0) atomicAdd(dot_slots + ((blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5)) & 255u), dot);
    }
}

int spmv_tma_plan_create(SpmvTmaPlan *plan, int64_t n_rows, const int32_t *indptr, int64_t nnz, cudaStream_t stream) {
    plan->n_rows = n_rows;
    plan->n_tiles = (int)(nnz / TMA_TILE) + 1;
    plan->stream = stream;
    plan->cap = 0;
    plan->lcol = nullptr;
    plan->dict = plan->dict_n = nullptr;
    Scratch mx;
    LSS_CUDA(mx.alloc(sizeof(int), stream));
    LSS_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->tile_row, sizeof(int) * (plan->n_tiles + 1), stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->tile_k, sizeof(int) * (plan->n_tiles + 1), stream));
    spmv_tma_plan_kernel<<<ceil_div(plan->n_tiles + 1, 256), 256, 0, stream>>>(n_rows, indptr, plan->n_tiles, plan->tile_row, plan->tile_k);
    LSS_LAUNCH_CHECK();
    spmv_tma_span_kernel<<<ceil_div(plan->n_tiles, 256), 256, 0, stream>>>(plan->n_tiles, plan->tile_k, mx.as<int>());
    LSS_LAUNCH_CHECK();
    int span = 0;
    LSS_CUDA(cudaMemcpyAsync(&span, mx.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LSS_CUDA(cudaStreamSynchronize(stream));
    plan->cap = (span + 15) / 16 * 16 + 16;
    return LSS_OK;
}

void spmv_tma_plan_destroy(SpmvTmaPlan *plan) {
    if (plan->tile_row) cudaFreeAsync(plan->tile_row, plan->stream);
    if (plan->tile_k) cudaFreeAsync(plan->tile_k, plan->stream);
    if (plan->lcol) cudaFreeAsync(plan->lcol, plan->stream);
    if (plan->dict) cudaFreeAsync(plan->dict, plan->stream);
    if (plan->dict_n) cudaFreeAsync(plan->dict_n, plan->stream);
    plan->tile_row = plan->tile_k = plan->dict = plan->dict_n = nullptr;
    plan->lcol = nullptr;
}

constexpr int TMA_BLOCK = 128, TMA_MAX_CAP = 2304;
static int tma_stages() {
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("LSS_TMA_STAGES");
        v = e ? atoi(e) : 2;
        if (v < 2 || v > 5) v = 2;
    }
    return v;
}
bool spmv_tma_fits(const SpmvTmaPlan *plan) { return plan->tile_row && plan->cap > 0 && plan->cap <= TMA_MAX_CAP; }

// y = A x over the rows of the plan (every row is written; rows of the arrays must be readable 16 bytes past their end)
int spmv_launch_tma(const SpmvTmaPlan *plan, const int32_t *indptr, const int32_t *indices, const double *data, const double *x, double *y,
                    const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots) {
    if (plan->n_rows <= 0) return LSS_OK;
    int S = tma_stages();
    if (S > 4) S = 4;
    const size_t smem = TMA_HEAD + (size_t)S * ((size_t)plan->cap * 12 + TMA_RP_CAP * 4);
    const void *fn = S == 2 ? (const void *)spmv_tma_kernel<TMA_BLOCK, 2> : S == 3 ? (const void *)spmv_tma_kernel<TMA_BLOCK, 3>
                                                                                   : (const void *)spmv_tma_kernel<TMA_BLOCK, 4>;
    static size_t attr_for = 0;
    static int per_sm = 0;
    if (attr_for != smem) {
        LSS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TMA_BLOCK, smem));
        if (per_sm < 1) return fail(LSS_EINVAL, "spmv_tma_kernel: %zu bytes of shared memory do not fit", smem);
        attr_for = smem;
    }
    const int grid = plan->n_tiles < kSMs * per_sm ? plan->n_tiles : kSMs * per_sm;
    int64_t n_rows = plan->n_rows;
    int n_tiles = plan->n_tiles, cap = plan->cap;
    const int *tr = plan->tile_row, *tk = plan->tile_k;
    void *args[] = {&n_rows, &n_tiles, &cap, &indptr, &indices, &data, &x, &y, &tr, &tk, &stop, &dot_slots};
    LSS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(TMA_BLOCK), args, smem, stream));
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
    return LSS_OK;
}

//   tmad    "tma" with a per-tile column dictionary.  The gathers are what every kernel above pays for (8.7 L1 lines per
//           32 entries of a row, every x entry fetched ~4 times per tile by neighbouring rows).  Once per pattern each
//           tile's distinct columns are sorted into a dictionary (~300 for 13 neighbouring rows) and the column indices
//           are rewritten as 16-bit positions in it.  Per tile the TMA engine then brings values (8 B), local indices
//           (2 B) and the dictionary; the CTA gathers x ONCE per distinct column into shared memory and the row sums
//           run entirely from shared memory.  10.3 instead of 12 bytes per entry from HBM, ~2.5x fewer L1 wavefronts.
__global__ void __launch_bounds__(256)
spmv_tmad_build_kernel(int n_tiles, const int *__restrict__ tile_k, const int *__restrict__ indices, int dstride, int *__restrict__ dict,
                       int *__restrict__ dict_n, unsigned short *__restrict__ lcol, int *__restrict__ dict_max) {
    constexpr int PMAX = 4096;
    __shared__ int keys[PMAX];
    __shared__ int uniq[PMAX / 2 + 512];
    __shared__ int wsum[8];
    __shared__ int total_s;
    const int t = blockIdx.x;
    const int k0 = tile_k[t], n = tile_k[t + 1] - k0;
    int P = 256;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += 256) keys[i] = i < n ? indices[k0 + i] : 0x7fffffff;
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < P; i += 256) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if (up ? (a > b) : (a < b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    __syncthreads();
    // compaction of the run heads: thread j owns the contiguous slice [j * chunk, (j + 1) * chunk)
    const int chunk = P / 256;
    const int lo = threadIdx.x * chunk;
    int cnt = 0;
    for (int i = lo; i < lo + chunk; ++i) cnt += (keys[i] != 0x7fffffff && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += v;
    }
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += wsum[w];
    if (threadIdx.x == 255) total_s = base + incl;
    int pos = base + incl - cnt;
    for (int i = lo; i < lo + chunk; ++i)
        if (keys[i] != 0x7fffffff && (i == 0 || keys[i] != keys[i - 1])) uniq[pos++] = keys[i];
    __syncthreads();
    const int nd = total_s;
    if (threadIdx.x == 0) {
        dict_n[t] = nd;
        atomicMax(dict_max, nd);
    }
    int *out = dict + (size_t)t * dstride;
    const int nd4 = (nd + 3) & ~3;
    for (int i = threadIdx.x; i < nd4 && i < dstride; i += 256) out[i] = uniq[i < nd ? i : nd - 1];
    for (int i = threadIdx.x; i < n; i += 256) {
        const int c = indices[k0 + i];
        int a = 0, b = nd;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (uniq[mid] < c) a = mid + 1; else b = mid;
        }
        lcol[k0 + i] = (unsigned short)a;
    }
}

template <int BLOCK, int STAGES, bool PIPE>
__global__ void __launch_bounds__(BLOCK)
spmv_tmad_kernel(int64_t n_rows, int n_tiles, int cap, int dcap, int dstride, const int *__restrict__ indptr,
                 const unsigned short *__restrict__ lcol, const double *__restrict__ data, const int *__restrict__ dict,
                 const int *__restrict__ dict_n, const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ tile_row,
                 const int *__restrict__ tile_k, const double *__restrict__ stop, double *__restrict__ dot_slots) {
    extern __shared__ __align__(128) unsigned char tmad_ring[];
    static_assert(STAGES <= 8, "header layout");
    static_assert(!PIPE || STAGES >= 3, "PIPE: tile i is reduced while the dictionary of tile i + 1 is gathered and tile i + 2 is in flight");
    // [0, 64) mbarriers | [64, 320) 8 descriptor ints per stage | xs: 1 (2 with PIPE) x dcap doubles |
    // STAGES x { cap doubles | cap u16 | dcap ints | RP ints }
    unsigned long long *full = reinterpret_cast<unsigned long long *>(tmad_ring);
    int *desc = reinterpret_cast<int *>(tmad_ring + 64);
    double *xs = reinterpret_cast<double *>(tmad_ring + TMA_HEAD);
    unsigned char *stages = tmad_ring + TMA_HEAD + (size_t)dcap * (PIPE ? 16 : 8);
    const size_t stage_bytes = (size_t)cap * 10 + (size_t)dcap * 4 + TMA_RP_CAP * 4;
    if (stop && *stop != 0.0) return;
    const uint64_t once = l2_policy_evict_first();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int d_row0 = 0, d_row1 = 0, d_k0 = 0, d_k1 = 0, d_nd = 0;
    auto load_desc = [&](int tile) {
        if (tile < n_tiles) {
            d_row0 = __ldg(tile_row + tile);
            d_row1 = __ldg(tile_row + tile + 1);
            d_k0 = __ldg(tile_k + tile);
            d_k1 = __ldg(tile_k + tile + 1);
            d_nd = __ldg(dict_n + tile);
        }
    };
    auto issue = [&](int tile, int stage) {          // thread 0; uses the descriptor registers
        if (tile >= n_tiles) return;
        unsigned char *base = stages + (size_t)stage * stage_bytes;
        const int k0a = d_k0 & ~7, nk = d_k1 - k0a;
        const int r0a = d_row0 & ~3;
        int nrp = d_row1 - r0a + 1;
        if (nrp > TMA_RP_CAP) nrp = TMA_RP_CAP;
        const unsigned b_val = (unsigned)((nk + 1) & ~1) * 8u, b_col = (unsigned)((nk + 7) & ~7) * 2u, b_rp = (unsigned)((nrp + 3) & ~3) * 4u;
        const unsigned b_dict = (unsigned)((d_nd + 3) & ~3) * 4u;
        int *dd = desc + stage * 8;
        dd[0] = d_row0; dd[1] = d_row1; dd[2] = k0a; dd[3] = r0a; dd[4] = d_nd; dd[5] = d_k0 - k0a; dd[6] = nk;
        mbar_arrive_expect_tx(full + stage, b_val + b_col + b_rp + b_dict);
        if (b_val) bulk_g2s(base, data + k0a, b_val, full + stage, once);
        if (b_col) bulk_g2s(base + (size_t)cap * 8, lcol + k0a, b_col, full + stage, once);
        if (b_dict) bulk_g2s(base + (size_t)cap * 10, dict + (size_t)tile * dstride, b_dict, full + stage, once);
        bulk_g2s(base + (size_t)cap * 10 + (size_t)dcap * 4, indptr + r0a, b_rp, full + stage, once);
    };
    constexpr int DREG = 8;                               // PIPE: dictionary entries gathered per thread (dcap <= DREG * BLOCK)
    const int t0 = blockIdx.x, dt = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            load_desc(t0 + s * dt);
            issue(t0 + s * dt, s);
        }
        load_desc(t0 + (STAGES - 1) * dt);
    }
    if (PIPE) {                                           // x at the distinct columns of tile 0
        if (t0 < n_tiles) {
            while (!mbar_try_wait(full, 0u)) {}
            const int *sdict = reinterpret_cast<const int *>(stages + (size_t)cap * 10);
            const int nd = desc[4];
            for (int j = threadIdx.x; j < nd; j += BLOCK) xs[j] = __ldg(x + sdict[j]);
        }
        __syncthreads();
    }
    double dot = 0.0;
    int i = 0;
    for (int t = t0; t < n_tiles; t += dt, ++i) {
        const int stage = i % STAGES;
        if (threadIdx.x == 0) {
            issue(t + (STAGES - 1) * dt, (i + STAGES - 1) % STAGES);        // the stage tile i - 1 was reduced from
            load_desc(t + STAGES * dt);
        }
        unsigned char *base = stages + (size_t)stage * stage_bytes;
        double *sval = reinterpret_cast<double *>(base);
        const unsigned short *slc = reinterpret_cast<const unsigned short *>(base + (size_t)cap * 8);
        const int *srp = reinterpret_cast<const int *>(base + (size_t)cap * 10 + (size_t)dcap * 4);
        const int *dd = desc + stage * 8;
        double xn[DREG];
        int nd_next = 0;
        const double *xsc = xs;
        if (PIPE) {
            // gathers for the NEXT tile's dictionary are issued now and land while this tile is reduced
            if (t + dt < n_tiles) {
                const int sn = (i + 1) % STAGES;
                while (!mbar_try_wait(full + sn, (unsigned)((i + 1) / STAGES) & 1u)) {}
                const int *sdict = reinterpret_cast<const int *>(stages + (size_t)sn * stage_bytes + (size_t)cap * 10);
                nd_next = desc[sn * 8 + 4];
#pragma unroll
                for (int u = 0; u < DREG; ++u) {
                    const int j = threadIdx.x + u * BLOCK;
                    xn[u] = j < nd_next ? __ldg(x + sdict[j]) : 0.0;
                }
            }
            xsc = xs + (size_t)(i & 1) * dcap;
        } else {
            while (!mbar_try_wait(full + stage, (unsigned)(i / STAGES) & 1u)) {}
            const int *sdict = reinterpret_cast<const int *>(base + (size_t)cap * 10);
            const int nd = dd[4];
            for (int j = threadIdx.x; j < nd; j += BLOCK) xs[j] = __ldg(x + sdict[j]);     // x once per distinct column of the tile
            __syncthreads();
        }
        const int row0 = dd[0], row1 = dd[1], k0a = dd[2], rp_off = row0 - dd[3], off0 = dd[5], span = dd[6];
        // phase A, one thread per entry: products in place of the values (conflict-free strided shared-memory accesses)
        {
            int k = off0 + threadIdx.x;
            for (; k + 3 * BLOCK < span; k += 4 * BLOCK) {                   // four independent load chains in flight
                const int c0 = slc[k], c1 = slc[k + BLOCK], c2 = slc[k + 2 * BLOCK], c3 = slc[k + 3 * BLOCK];
                const double v0 = sval[k], v1 = sval[k + BLOCK], v2 = sval[k + 2 * BLOCK], v3 = sval[k + 3 * BLOCK];
                const double x0 = xsc[c0], x1 = xsc[c1], x2 = xsc[c2], x3 = xsc[c3];
                sval[k] = v0 * x0;
                sval[k + BLOCK] = v1 * x1;
                sval[k + 2 * BLOCK] = v2 * x2;
                sval[k + 3 * BLOCK] = v3 * x3;
            }
            for (; k < span; k += BLOCK) sval[k] *= xsc[slc[k]];
        }
        __syncthreads();
        // phase B, g lanes per row: sums of the staged products
        const int nr = row1 - row0;
        int g = 1;
        while (g < 32 && nr * (2 * g) <= BLOCK) g <<= 1;
        const int gl = threadIdx.x & (g - 1), n_groups = BLOCK / g;
        for (int rr = threadIdx.x / g; rr < nr; rr += n_groups) {
            const int row = row0 + rr;
            const int j = rr + rp_off;
            const int b = (j < TMA_RP_CAP ? srp[j] : __ldg(indptr + row)) - k0a;
            const int e = (j + 1 < TMA_RP_CAP ? srp[j + 1] : __ldg(indptr + row + 1)) - k0a;
            const double xr = (dot_slots && gl == 0) ? __ldg(x + row) : 0.0;
            double a0 = 0.0, a1 = 0.0;
            int k = b + gl;
            for (; k + g < e; k += 2 * g) {
                a0 += sval[k];
                a1 += sval[k + g];
            }
            if (k < e) a0 += sval[k];
            double acc = a0 + a1;
            const unsigned lane = lane_id();
            const unsigned gmask = g == 32 ? 0xffffffffu : (((1u << g) - 1u) << (lane & ~(unsigned)(g - 1)));
            for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
            if (gl == 0) {
                y[row] = acc;
                dot += acc * xr;
            }
        }
        if (PIPE && nd_next > 0) {
            double *xnext = xs + (size_t)((i + 1) & 1) * dcap;
#pragma unroll
            for (int u = 0; u < DREG; ++u) {
                const int j = threadIdx.x + u * BLOCK;
                if (j < nd_next) xnext[j] = xn[u];
            }
        }
        // the products were written through the generic proxy into a buffer the TMA engine (async proxy) refills next
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                                   // stage i (and its x copy) are free
    }
    if (dot_slots) {
        dot = warp_sum(dot);
        if (lane_id() == 0 && dot != 0.0) atomicAdd(dot_slots + ((blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5)) & 255u), dot);
    }
}

// adds the dictionary form to a plan (indices: the CSR column indices of the planned rows)
int spmv_tma_plan_add_dict(SpmvTmaPlan *plan, const int32_t *indices, int64_t nnz) {
    cudaStream_t stream = plan->stream;
    if (plan->lcol || !spmv_tma_fits(plan)) return LSS_OK;
    plan->dstride = (plan->cap + 3) & ~3;
    Scratch mx;
    LSS_CUDA(mx.alloc(sizeof(int), stream));
    LSS_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->lcol, sizeof(unsigned short) * ((size_t)nnz + 16), stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->dict, sizeof(int) * ((size_t)plan->n_tiles * plan->dstride + 16), stream));
    LSS_CUDA(cudaMallocAsync((void **)&plan->dict_n, sizeof(int) * ((size_t)plan->n_tiles + 1), stream));
    spmv_tmad_build_kernel<<<plan->n_tiles, 256, 0, stream>>>(plan->n_tiles, plan->tile_k, indices, plan->dstride, plan->dict, plan->dict_n,
                                                              plan->lcol, mx.as<int>());
    LSS_LAUNCH_CHECK();
    int dmax = 0;
    LSS_CUDA(cudaMemcpyAsync(&dmax, mx.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LSS_CUDA(cudaStreamSynchronize(stream));
    plan->dcap = ((dmax + 3) & ~3) + 4;
    // entries counted from the 8-aligned start of a tile: the stage capacity may need 4 more than the 4-aligned span
    plan->cap8 = plan->cap + 16;
    return LSS_OK;
}

int spmv_launch_tmad(const SpmvTmaPlan *plan, const int32_t *indptr, const double *data, const double *x, double *y, const double *stop,
                     cudaStream_t stream, int64_t *launches, double *dot_slots) {
    if (plan->n_rows <= 0) return LSS_OK;
    const int S = tma_stages();                  // 2: dictionary gather inside the tile's own iteration; 3..5: pipelined one tile ahead
    const bool pipe = S >= 3;
    if (plan->dcap > 8 * TMA_BLOCK) return fail(LSS_EINVAL, "spmv_tmad_kernel: %d distinct columns in a tile", plan->dcap);
    const size_t smem = TMA_HEAD + (size_t)plan->dcap * (pipe ? 16 : 8) +
                        (size_t)S * ((size_t)plan->cap8 * 10 + (size_t)plan->dcap * 4 + TMA_RP_CAP * 4);
    const void *fn = S == 2 ? (const void *)spmv_tmad_kernel<TMA_BLOCK, 2, false>
                   : S == 3 ? (const void *)spmv_tmad_kernel<TMA_BLOCK, 3, true>
                   : S == 4 ? (const void *)spmv_tmad_kernel<TMA_BLOCK, 4, true> : (const void *)spmv_tmad_kernel<TMA_BLOCK, 5, true>;
    static size_t attr_for = 0;
    static int per_sm = 0;
    if (attr_for != smem) {
        LSS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TMA_BLOCK, smem));
        if (per_sm < 1) return fail(LSS_EINVAL, "spmv_tmad_kernel: %zu bytes of shared memory do not fit", smem);
        attr_for = smem;
    }
    const int grid = plan->n_tiles < kSMs * per_sm ? plan->n_tiles : kSMs * per_sm;
    int64_t n_rows = plan->n_rows;
    int n_tiles = plan->n_tiles, cap = plan->cap8, dcap = plan->dcap, dstride = plan->dstride;
    const int *tr = plan->tile_row, *tk = plan->tile_k, *dict = plan->dict, *dn = plan->dict_n;
    const unsigned short *lc = plan->lcol;
    void *args[] = {&n_rows, &n_tiles, &cap, &dcap, &dstride, &indptr, &lc, &data, &dict, &dn, &x, &y, &tr, &tk, &stop, &dot_slots};
    LSS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(TMA_BLOCK), args, smem, stream));
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
    return LSS_OK;
}

static int g_normal_product = -1;
int normal_product_variant() {
    if (g_normal_product < 0) {
        const char *e = getenv("LSS_N_PRODUCT");
        g_normal_product = e ? atoi(e) : 0;     // measured on the 10^7-point workload (profiles/README.md): none of 1..10 beats the tiled kernel yet
        if (g_normal_product < 0 || g_normal_product > 11) g_normal_product = 0;
    }
    return g_normal_product;
}
void set_normal_product_variant(int v) { g_normal_product = (v < 0 || v > 11) ? 0 : v; }

// y (+)= A x for the explicit normal matrix, one launch; y must hold zeros (see above).  variant: normal_product_variant()
int spmv_launch_normal(int variant, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x,
                       double *y, int64_t nnz, const SpmvPlan *plan, const double *stop, cudaStream_t stream, int64_t *launches,
                       double *dot_slots) {
    if (n_rows <= 0 || nnz == 0) return LSS_OK;
    if (variant == 1 || variant == 5 || variant == 6) {
        if (variant == 1) {
            spmv_rows_kernel<16, 8><<<ceil_div(n_rows * 16, 256), 256, 0, stream>>>(n_rows, indptr, indices, data, x, y, stop, dot_slots);
        } else if (variant == 5) {
            spmv_rows_kernel<8, 12><<<ceil_div(n_rows * 8, 256), 256, 0, stream>>>(n_rows, indptr, indices, data, x, y, stop, dot_slots);
        } else {
            spmv_rows_kernel<32, 4><<<ceil_div(n_rows * 32, 256), 256, 0, stream>>>(n_rows, indptr, indices, data, x, y, stop, dot_slots);
        }
        LSS_LAUNCH_CHECK();
        if (launches) *launches += 1;
        return LSS_OK;
    }
    return spmv_launch_atomic(n_rows, indptr, indices, data, x, y, nnz, plan, stop, stream, launches, dot_slots);
}

// ============================================================================ ELL (fixed row length) product
// The data block G has the same number of entries in every row (2^2 z0 corners + 2^3 dz corners + bias columns),
// so it is also kept column-major ("ELL"): entry k of row r at [k * n_pad + r].  One thread per row then reads
// perfectly coalesced index / value streams (1 + 2 L1 wavefronts per warp instruction) and gathers x for 32
// CONSECUTIVE rows at a time -- points that are neighbours on the ground and share grid nodes -- instead of
// 32 consecutive non-zeros of 2-3 rows.  The nnz-tiled CSR kernel is L1TEX-wavefront bound on those gathers
// (ncu: l1tex 75 %, DRAM 59-64 %); this layout needs ~2.5x fewer wavefronts per byte streamed.
__global__ void __launch_bounds__(256)
ell_build_kernel(int64_t n_rows, int64_t n_pad, int K, const int *__restrict__ indptr, const int *__restrict__ indices,
                 const double *__restrict__ data, int *__restrict__ ell_col, double *__restrict__ ell_val) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int b = indptr[r], e = indptr[r + 1];
    for (int k = 0; k < K; ++k) {
        const bool in = b + k < e;
        ell_col[(int64_t)k * n_pad + r] = in ? indices[b + k] : 0;
        ell_val[(int64_t)k * n_pad + r] = in ? data[b + k] : 0.0;
    }
}

template <int K>
__global__ void __launch_bounds__(256)
ell_spmv_kernel(int64_t n_rows, int64_t n_pad, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                const double *__restrict__ x, double *__restrict__ y, int mode, const double *__restrict__ s,
                const double *__restrict__ stop) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    int c[K];
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        c[k] = ldg_stream(ell_col + (int64_t)k * n_pad + r);
        v[k] = ldg_stream(ell_val + (int64_t)k * n_pad + r);
    }
    if (stop_v != 0.0) return;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) acc += v[k] * __ldg(x + c[k]);     // same left-to-right order as the CSR row sum
    y[r] = epilogue(acc, mode, s, r);
}

// q += A^T u straight from the CSR rows of A (one thread per row): for the handful of transpose products per fit that do
// not justify building and storing A^T (parse_model's per-node sums).  The rows are sorted by grid cell, so entry k of
// neighbouring rows usually names the same column: the warp folds runs of equal columns with a segmented shuffle scan and
// only the tail of a run issues a RED.ADD.F64 (a bias column shared by a whole track segment, or the four nodes of a
// 10 km bias grid under 10^5 consecutive points, would otherwise serialise thousands of atomics on one address).
__global__ void __launch_bounds__(256)
csr_rmatvec_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                   const double *__restrict__ u, double *__restrict__ q) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int k0 = 0, len = 0;
    double ur = 0.0;
    if (r < n_rows) {
        ur = u[r];
        if (ur != 0.0) {
            k0 = indptr[r];
            len = indptr[r + 1] - k0;
        }
    }
    int max_len = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    for (int k = 0; k < max_len; ++k) {
        const bool in = k < len;
        const int key = in ? indices[k0 + k] : -1;
        double val = in ? data[k0 + k] * ur : 0.0;
        const int kp = __shfl_up_sync(0xffffffffu, key, 1);
        int head = (lane == 0u || kp != key) ? 1 : 0;          // equal columns need not be adjacent: scan inside runs only
        const int is_head = head;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v2 = __shfl_up_sync(0xffffffffu, val, o);
            const int h2 = __shfl_up_sync(0xffffffffu, head, o);
            if (lane >= (unsigned)o) {
                if (!head) val += v2;
                head |= h2;
            }
        }
        const int hn = __shfl_down_sync(0xffffffffu, is_head, 1);
        if (key >= 0 && (lane == 31u || hn)) atomicAdd(q + key, val);
    }
}

// Fused normal-equation product of the data block:  q += G^T (w .* (G x))  in ONE pass over the ELL copy of G.
// Thread r forms t = sum_k v[k] x[c[k]], u = w[r] t and scatters v[k] u into q[c[k]].  Neighbouring rows (points that
// are neighbours on the ground) hit the same grid nodes, so contributions are first summed over runs of equal
// columns inside 8-lane groups with a 3-step segmented shuffle scan and only run tails issue a RED.ADD.F64; rows
// with zero weight (edited points) issue none.  This reads G once (12 B per entry) instead of streaming G and an
// explicit G^T: 2.2x less HBM traffic per CG iteration than the two-product form.
//   MODE 0: u = w (G x)      MODE 1: u = w .* b  (right-hand side G^T W z)      MODE 2: scatter w v^2 (Jacobi diagonal)
template <int K, int MODE>
__global__ void __launch_bounds__(256)
ell_normal_kernel(int64_t n_rows, int64_t n_pad, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                  const double *__restrict__ x, const double *__restrict__ w, const double *__restrict__ b,
                  double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ stop) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = r < n_rows;
    const int64_t rr = active ? r : n_rows - 1;
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    int c[K];
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        c[k] = ldg_stream(ell_col + (int64_t)k * n_pad + rr);
        v[k] = ldg_stream(ell_val + (int64_t)k * n_pad + rr);
    }
    if (stop_v != 0.0) return;
    double u = active ? w[rr] : 0.0;
    if (MODE == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) t += v[k] * __ldg(x + c[k]);
        u *= t;
        if (pq_slots) {
            // x . (A^T W A x) = sum_r w_r t_r^2 comes for free here: warp sum, one RED per warp spread over 256 slots
            // (no block barrier: warps must not wait for each other's gathers)
            const double part = warp_sum(u * t);
            if (lane_id() == 0 && part != 0.0)
                atomicAdd(pq_slots + (((unsigned)blockIdx.x * 8u + (threadIdx.x >> 5)) & 255u), part);
        }
    } else if (MODE == 1) {
        u *= b[rr];
    }
    const unsigned lane = lane_id(), sub = lane & 7u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int col = active ? c[k] : -1 - (int)lane;
        double sum = (MODE == 2) ? v[k] * v[k] * u : v[k] * u;
        // runs of equal columns inside each 8-lane group: heads from one shuffle + one ballot, the segmented scan
        // then needs only the value shuffles (distance to the run head comes from the ballot bits)
        const int prev = __shfl_up_sync(0xffffffffu, col, 1);
        const bool head = sub == 0 || prev != col;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned grp = (heads >> (lane & 24u)) & 0xffu;
        const int dist = (int)sub - (31 - __clz((int)(grp & ((2u << sub) - 1u))));
        const bool tail = sub == 7 || ((grp >> (sub + 1)) & 1u);
        if (heads != 0xffffffffu) {          // warp-uniform: some neighbouring rows share this column
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const double s2 = __shfl_up_sync(0xffffffffu, sum, o);
                if (dist >= o) sum += s2;
            }
        }
        if (tail && col >= 0 && sum != 0.0) atomicAdd(q + col, sum);
    }
}

int ell_build(int64_t n_rows, int K, const int32_t *indptr, const int32_t *indices, const double *data, int32_t *ell_col,
              double *ell_val, int64_t n_pad, cudaStream_t stream) {
    if (n_rows <= 0) return LSS_OK;
    ell_build_kernel<<<ceil_div(n_rows, 256), 256, 0, stream>>>(n_rows, n_pad, K, indptr, indices, data, ell_col, ell_val);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int ell_spmv_launch(int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val, const double *x, double *y,
                    int mode, const double *s, const double *stop, cudaStream_t stream) {
    if (n_rows <= 0) return LSS_OK;
    const int grid = ceil_div(n_rows, 256);
#define LSS_ELL_CASE(KK) case KK: ell_spmv_kernel<KK><<<grid, 256, 0, stream>>>(n_rows, n_pad, ell_col, ell_val, x, y, mode, s, stop); break;
    switch (K) {
        LSS_ELL_CASE(1) LSS_ELL_CASE(2) LSS_ELL_CASE(3) LSS_ELL_CASE(4) LSS_ELL_CASE(5) LSS_ELL_CASE(6) LSS_ELL_CASE(7) LSS_ELL_CASE(8)
        LSS_ELL_CASE(9) LSS_ELL_CASE(10) LSS_ELL_CASE(11) LSS_ELL_CASE(12) LSS_ELL_CASE(13) LSS_ELL_CASE(14) LSS_ELL_CASE(15) LSS_ELL_CASE(16)
        LSS_ELL_CASE(17) LSS_ELL_CASE(18) LSS_ELL_CASE(19) LSS_ELL_CASE(20) LSS_ELL_CASE(21) LSS_ELL_CASE(22) LSS_ELL_CASE(23) LSS_ELL_CASE(24)
        default: return fail(LSS_EINVAL, "ELL product: row length %d not in 1..24", K);
    }
#undef LSS_ELL_CASE
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int ell_normal_launch(int mode, int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val, const double *x,
                      const double *w, const double *b, double *q, double *pq_slots, const double *stop, cudaStream_t stream) {
    if (n_rows <= 0) return LSS_OK;
    const int grid = ceil_div(n_rows, 256);
#define LSS_ELLN_CASE(KK)                                                                                                         \
    case KK:                                                                                                                      \
        if (mode == 0) ell_normal_kernel<KK, 0><<<grid, 256, 0, stream>>>(n_rows, n_pad, ell_col, ell_val, x, w, b, q, pq_slots, stop); \
        else if (mode == 1) ell_normal_kernel<KK, 1><<<grid, 256, 0, stream>>>(n_rows, n_pad, ell_col, ell_val, x, w, b, q, pq_slots, stop); \
        else ell_normal_kernel<KK, 2><<<grid, 256, 0, stream>>>(n_rows, n_pad, ell_col, ell_val, x, w, b, q, pq_slots, stop); \
        break;
    switch (K) {
        LSS_ELLN_CASE(1) LSS_ELLN_CASE(2) LSS_ELLN_CASE(3) LSS_ELLN_CASE(4) LSS_ELLN_CASE(5) LSS_ELLN_CASE(6) LSS_ELLN_CASE(7) LSS_ELLN_CASE(8)
        LSS_ELLN_CASE(9) LSS_ELLN_CASE(10) LSS_ELLN_CASE(11) LSS_ELLN_CASE(12) LSS_ELLN_CASE(13) LSS_ELLN_CASE(14) LSS_ELLN_CASE(15) LSS_ELLN_CASE(16)
        LSS_ELLN_CASE(17) LSS_ELLN_CASE(18) LSS_ELLN_CASE(19) LSS_ELLN_CASE(20) LSS_ELLN_CASE(21) LSS_ELLN_CASE(22) LSS_ELLN_CASE(23) LSS_ELLN_CASE(24)
        default: return fail(LSS_EINVAL, "ELL product: row length %d not in 1..24", K);
    }
#undef LSS_ELLN_CASE
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

}  // namespace lss

using namespace lss;

namespace lss { int g_opt_explicit_normal = -1; int g_opt_step_kernel = -1; }

extern "C" int lss_set_option(const char *name, int32_t value) {
    if (name && strcmp(name, "explicit_normal") == 0) {
        lss::g_opt_explicit_normal = value != 0;
        return LSS_OK;
    }
    if (name && strcmp(name, "spmv_variant") == 0) {
        if (value < 0 || value >= kNumShapes) return fail(LSS_EINVAL, "spmv_variant must be in [0, %d)", kNumShapes);
        g_variant = value;
        return LSS_OK;
    }
    if (name && strcmp(name, "normal_product") == 0) {
        if (value < 0 || value > 11) return fail(LSS_EINVAL, "normal_product must be in [0, 11]");
        lss::set_normal_product_variant(value);
        return LSS_OK;
    }
    if (name && strcmp(name, "step_kernel") == 0) {
        if (value < 0 || value > 2) return fail(LSS_EINVAL, "step_kernel must be 0, 1 or 2");
        lss::g_opt_step_kernel = value;
        return LSS_OK;
    }
    return fail(LSS_EINVAL, "unknown option %s", name ? name : "(null)");
}

struct lss_spmv_plan {
    lss::SpmvPlan p;
    int64_t nnz;
};

extern "C" int lss_spmv_plan_create(lss_spmv_plan **plan, int64_t n_rows, const int32_t *indptr, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!plan || n_rows < 0) return fail(LSS_EINVAL, "spmv_plan_create: bad arguments");
    int nnz = 0;
    if (n_rows > 0) {
        LSS_CUDA(cudaMemcpyAsync(&nnz, indptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
    }
    lss_spmv_plan *pl = new lss_spmv_plan();
    pl->nnz = nnz;
    int rc = spmv_plan_create(&pl->p, n_rows, indptr, nnz, st);
    if (rc != LSS_OK) { delete pl; return rc; }
    *plan = pl;
    return LSS_OK;
}

extern "C" int lss_spmv_plan_destroy(lss_spmv_plan *plan) {
    if (!plan) return LSS_OK;
    spmv_plan_destroy(&plan->p);
    delete plan;
    return LSS_OK;
}

extern "C" int lss_spmv_planned(lss_spmv_plan *plan, int64_t n_rows, const int32_t *indptr, const int32_t *indices,
                                const double *data, const double *x, double *y, int32_t mode, const double *s, void *stream) {
    if (!plan) return fail(LSS_EINVAL, "spmv_planned: null plan");
    if (mode < 0 || mode > 2) return fail(LSS_EINVAL, "spmv: unknown mode %d", mode);
    if (mode != LSS_SPMV_PLAIN && !s) return fail(LSS_EINVAL, "spmv: mode %d needs the vector s", mode);
    return spmv_launch(n_rows, indptr, indices, data, x, y, mode, s, plan->nnz, &plan->p, nullptr, (cudaStream_t)stream, nullptr);
}

extern "C" int lss_csr_rmatvec(int64_t n_rows, int64_t n_cols, const int32_t *indptr, const int32_t *indices, const double *data,
                               const double *u, double *q, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cols > 0) LSS_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * n_cols, st));
    if (n_rows <= 0) return LSS_OK;
    csr_rmatvec_kernel<<<ceil_div(n_rows, 256), 256, 0, st>>>(n_rows, indptr, indices, data, u, q);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

extern "C" int lss_spmv(int64_t n_rows, int64_t n_cols, const int32_t *indptr, const int32_t *indices, const double *data,
                        const double *x, double *y, int32_t mode, const double *s, void *stream) {
    (void)n_cols;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode < 0 || mode > 2) return fail(LSS_EINVAL, "spmv: unknown mode %d", mode);
    if (mode != LSS_SPMV_PLAIN && !s) return fail(LSS_EINVAL, "spmv: mode %d needs the vector s", mode);
    if (n_rows <= 0) return LSS_OK;
    int nnz = 0;
    LSS_CUDA(cudaMemcpyAsync(&nnz, indptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    return spmv_launch(n_rows, indptr, indices, data, x, y, mode, s, nnz, nullptr, nullptr, st, nullptr);
}
