// CSR products y = A x for the matrix-free normal-equation operator.
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
                                            const double *__restrict__ xdot = nullptr, double *__restrict__ dot_slots = nullptr) {
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
        const double xr = (ATOMIC && xdot && valid && gl == 0) ? __ldg(xdot + rr) : 0.0;   // in flight during the sum
        double acc = 0.0;
        for (int64_t k = b - k0 + gl; k < e - k0; k += g) acc += prod[k];
        for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && gl == 0) {
            if (ATOMIC) {
                if (v == 0 || open) atomicAdd(y + rr, acc);
                else y[rr] = acc;
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
                       const int *__restrict__ tile_bounds, const double *__restrict__ stop, double *__restrict__ dot_slots = nullptr) {
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
        reduce_rows<BLOCK, ATOMIC>(prod, t, k0, k1, r_lo, r_hi, indptr, y, mode, s, carry_val, carry_row, dot_slots ? x : nullptr, dot_slots);
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
                       int64_t nnz, const SpmvPlan *plan, const double *stop, cudaStream_t stream, int64_t *launches, double *dot_slots) {
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
    spmv_persistent_kernel<BLOCK, IPT, false, true><<<grid, BLOCK, 0, stream>>>(n_rows, nnz, n_tiles, indptr, indices, data, x, y, LSS_SPMV_PLAIN,
                                                                                nullptr, nullptr, nullptr, plan->bounds, stop, dot_slots);
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 1;
    return LSS_OK;
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

// q += A^T u straight from the CSR rows of A (one thread per row, one RED.ADD.F64 per entry): for the handful of
// transpose products per fit that do not justify building and storing A^T (parse_model's per-node sums).
__global__ void __launch_bounds__(256)
csr_rmatvec_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                   const double *__restrict__ u, double *__restrict__ q) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const double ur = u[r];
    if (ur == 0.0) return;
    for (int k = indptr[r]; k < indptr[r + 1]; ++k) atomicAdd(q + indices[k], data[k] * ur);
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
        default: return fail(LSS_EINVAL, "ELL product: row length %d not in 1..16", K);
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
        default: return fail(LSS_EINVAL, "ELL product: row length %d not in 1..16", K);
    }
#undef LSS_ELLN_CASE
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

}  // namespace lss

using namespace lss;

namespace lss { int g_opt_explicit_normal = -1; }

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
