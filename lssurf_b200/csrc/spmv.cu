// CSR products y = A x for the matrix-free normal-equation operator.
//
// Replaces scipy's csr_matvec behind `.dot` (LSsurf/smooth_fit.py:142, 146, 326-345).  HBM-bound:
// the matrix (12 B per non-zero) is streamed exactly once with coalesced, L1-bypassing loads;
// x is gathered through L1/L2 (it is a few MB: resident in the 126 MB L2).
//
// Work decomposition is by NON-ZEROS, not rows: CTA t owns non-zeros [t*TILE, (t+1)*TILE), so the
// same kernel is balanced for G (12-13 entries per row), for the constraint stencils (2-8 per
// row) and for G^T, whose rows range from empty (63 % of the grid columns see no data) to
// >10^4 entries (bias columns).  Products are staged in shared memory, rows are then reduced
// from shared memory by one thread (short rows) or one warp (long rows) each.  Rows cut by a tile
// boundary are completed by a tiny fix-up kernel from per-tile carries, so the result is
// deterministic (no atomics).
#include "common.cuh"

namespace lss {

constexpr int SPMV_BLOCK = 256;
constexpr int SPMV_IPT = 8;
constexpr int SPMV_TILE = SPMV_BLOCK * SPMV_IPT;  // 2048 non-zeros, 16 KB of staged products

__device__ __forceinline__ double epilogue(double v, int mode, const double *__restrict__ s, int64_t r) {
    if (mode == LSS_SPMV_SCALE) return s[r] * v;
    if (mode == LSS_SPMV_RESIDUAL) return s[r] - v;
    return v;
}

// first r in [0, n] with a[r] >= key  (a is non-decreasing, length n + 1 is assumed readable up to a[n])
__device__ __forceinline__ int64_t lower_bound_i32(const int *__restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n + 1;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// bounds[t] = first row whose entries start at or after non-zero t*TILE (bounds[n_tiles] = n_rows): the rows a
// tile owns are [bounds[t], bounds[t+1]).  Depends only on the sparsity structure: computed once per matrix.
__global__ void __launch_bounds__(256)
spmv_tile_bounds_kernel(int64_t n_rows, int64_t nnz, const int *__restrict__ indptr, int n_tiles, int *__restrict__ bounds) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    bounds[t] = (t == n_tiles) ? (int)n_rows : (int)lower_bound_i32(indptr, n_rows, (int64_t)t * SPMV_TILE);
}

__global__ void __launch_bounds__(SPMV_BLOCK)
spmv_tile_kernel(int64_t n_rows, int64_t nnz, const int *__restrict__ indptr, const int *__restrict__ indices,
                 const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y, int mode,
                 const double *__restrict__ s, double *__restrict__ carry_val, int *__restrict__ carry_row,
                 const int *__restrict__ tile_bounds, const double *__restrict__ stop) {
    __shared__ double prod[SPMV_TILE];
    const int64_t k0 = (int64_t)blockIdx.x * SPMV_TILE;
    const int64_t k1 = (k0 + SPMV_TILE < nnz) ? k0 + SPMV_TILE : nnz;
    // issued first, consumed after the staging loads are in flight (no serialised round trip per CTA)
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    const int64_t r_lo = tile_bounds[blockIdx.x], r_hi = tile_bounds[blockIdx.x + 1];
    // stage the products: coalesced streaming loads of the matrix, gathered loads of x
#pragma unroll
    for (int j = 0; j < SPMV_IPT; ++j) {
        const int64_t k = k0 + j * SPMV_BLOCK + threadIdx.x;
        double p = 0.0;
        if (k < k1) p = ldg_stream(data + k) * __ldg(x + ldg_stream(indices + k));
        prod[j * SPMV_BLOCK + threadIdx.x] = p;
    }
    if (stop_v != 0.0) return;   // solver converged: the queued tail of the batch is a no-op (uniform per launch)
    __syncthreads();
    // carry-in: the row that started in an earlier tile and still has entries here
    const bool has_carry = r_lo > 0 && (int64_t)indptr[r_lo] > k0;   // r_lo <= n_rows always
    const int64_t n_virtual = (r_hi - r_lo) + 1;                      // virtual row 0 is the carry-in
    if (n_virtual <= 2 * (SPMV_BLOCK / 32)) {
        // few, long rows: one warp per row
        const int w = threadIdx.x >> 5;
        for (int64_t v = w; v < n_virtual; v += SPMV_BLOCK / 32) {
            int64_t b, e, r = r_lo + v - 1;
            if (v == 0) {
                if (!has_carry) continue;
                b = k0;
                e = (int64_t)indptr[r_lo];
            } else {
                b = indptr[r];
                e = indptr[r + 1];
            }
            const bool open = e > k1;   // continues in the next tile
            if (open) e = k1;
            double acc = 0.0;
            for (int64_t k = b - k0 + lane_id(); k < e - k0; k += 32) acc += prod[k];
            acc = warp_sum(acc);
            if (lane_id() == 0) {
                if (v == 0) carry_val[blockIdx.x] = acc;
                else y[r] = open ? acc : epilogue(acc, mode, s, r);
            }
        }
    } else {
        for (int64_t v = threadIdx.x; v < n_virtual; v += SPMV_BLOCK) {
            int64_t b, e, r = r_lo + v - 1;
            if (v == 0) {
                if (!has_carry) continue;
                b = k0;
                e = (int64_t)indptr[r_lo];
            } else {
                b = indptr[r];
                e = indptr[r + 1];
            }
            const bool open = e > k1;
            if (open) e = k1;
            double acc = 0.0;
            for (int64_t k = b - k0; k < e - k0; ++k) acc += prod[k];
            if (v == 0) carry_val[blockIdx.x] = acc;
            else y[r] = open ? acc : epilogue(acc, mode, s, r);
        }
    }
    if (threadIdx.x == 0) carry_row[blockIdx.x] = has_carry ? (int)(r_lo - 1) : -1;
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
spmv_empty_kernel(int64_t n_rows, double *__restrict__ y, int mode, const double *__restrict__ s) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        y[r] = epilogue(0.0, mode, s, r);
}

int spmv_plan_create(SpmvPlan *plan, int64_t n_rows, const int32_t *indptr, int64_t nnz, cudaStream_t stream) {
    plan->n_tiles = nnz > 0 ? ceil_div(nnz, SPMV_TILE) : 0;
    plan->bounds = nullptr;
    plan->carry_val = nullptr;
    plan->carry_row = nullptr;
    if (plan->n_tiles == 0) return LSS_OK;
    LSS_CUDA(cudaMalloc((void **)&plan->bounds, sizeof(int) * (plan->n_tiles + 1)));
    LSS_CUDA(cudaMalloc((void **)&plan->carry_val, sizeof(double) * plan->n_tiles));
    LSS_CUDA(cudaMalloc((void **)&plan->carry_row, sizeof(int) * plan->n_tiles));
    spmv_tile_bounds_kernel<<<ceil_div(plan->n_tiles + 1, 256), 256, 0, stream>>>(n_rows, nnz, indptr, plan->n_tiles, plan->bounds);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

void spmv_plan_destroy(SpmvPlan *plan) {
    if (plan->bounds) cudaFree(plan->bounds);
    if (plan->carry_val) cudaFree(plan->carry_val);
    if (plan->carry_row) cudaFree(plan->carry_row);
    plan->bounds = nullptr;
    plan->carry_val = nullptr;
    plan->carry_row = nullptr;
}

int spmv_launch(int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, const double *x,
                double *y, int mode, const double *s, int64_t nnz, const SpmvPlan *plan, const double *stop,
                cudaStream_t stream, int64_t *launches) {
    if (n_rows <= 0) return LSS_OK;
    if (nnz == 0) {
        spmv_empty_kernel<<<stride_grid(n_rows, 256), 256, 0, stream>>>(n_rows, y, mode, s);
        LSS_LAUNCH_CHECK();
        if (launches) *launches += 1;
        return LSS_OK;
    }
    const int n_tiles = ceil_div(nnz, SPMV_TILE);
    Scratch cv, cr, tb;
    double *carry_val;
    int *carry_row, *bounds;
    if (plan && plan->bounds) {
        carry_val = plan->carry_val; carry_row = plan->carry_row; bounds = plan->bounds;
    } else {
        LSS_CUDA(cv.alloc(sizeof(double) * n_tiles, stream));
        LSS_CUDA(cr.alloc(sizeof(int) * n_tiles, stream));
        LSS_CUDA(tb.alloc(sizeof(int) * (n_tiles + 1), stream));
        carry_val = cv.as<double>(); carry_row = cr.as<int>(); bounds = tb.as<int>();
        spmv_tile_bounds_kernel<<<ceil_div(n_tiles + 1, 256), 256, 0, stream>>>(n_rows, nnz, indptr, n_tiles, bounds);
        LSS_LAUNCH_CHECK();
        if (launches) *launches += 1;
    }
    spmv_tile_kernel<<<n_tiles, SPMV_BLOCK, 0, stream>>>(n_rows, nnz, indptr, indices, data, x, y, mode, s, carry_val, carry_row, bounds, stop);
    LSS_LAUNCH_CHECK();
    spmv_fixup_kernel<<<ceil_div(n_tiles, 256), 256, 0, stream>>>(n_tiles, carry_val, carry_row, y, mode, s, stop);
    LSS_LAUNCH_CHECK();
    if (launches) *launches += 2;
    return LSS_OK;
}

}  // namespace lss

using namespace lss;

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
