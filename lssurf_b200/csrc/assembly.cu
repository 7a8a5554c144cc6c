// Design-matrix assembly kernels: interpolation rows and finite-difference stencil rows as COO.
//
// Replaces lin_op.interp_mtx (LSsurf/lin_op.py:163-247) and lin_op.diff_op (lin_op.py:80-132).
// Both are HBM-bound streaming kernels: one thread per output row, vectorised stores of the
// row's triplets (each thread writes whole 32-byte sectors).
#include <stdarg.h>

#include "common.cuh"

namespace lss {

thread_local char g_err[512] = "";
long long g_launches = 0;

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

struct GridDev {
    int ndim;
    long long shape[3], stride[3], col_0;
    double lo[3], hi[3], delta[3];
};

static GridDev to_dev(const lss_grid_t *g) {
    GridDev d;
    d.ndim = g->ndim;
    d.col_0 = g->col_0;
    for (int i = 0; i < 3; ++i) {
        d.shape[i] = i < g->ndim ? g->shape[i] : 1;
        d.stride[i] = i < g->ndim ? g->stride[i] : 0;
        d.lo[i] = g->lo[i];
        d.hi[i] = g->hi[i];
        d.delta[i] = g->delta[i];
    }
    return d;
}

struct PtsDev {
    const double *p[3];
};

// One thread per point.  NDIM is a template parameter so the 2^NDIM corner loop fully unrolls and
// the triplets stay in registers until the vector stores.
template <int NDIM>
__global__ void __launch_bounds__(256)
interp_coo_kernel(GridDev g, PtsDev pts, long long n, long long row_0, int *__restrict__ rows,
                  int *__restrict__ cols, double *__restrict__ vals, unsigned char *__restrict__ valid,
                  unsigned long long *__restrict__ n_invalid) {
    constexpr int NC = 1 << NDIM;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n) {
        double c[NDIM];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) c[d] = pts.p[d][i];
        // fd_grid.validate_pts (fd_grid.py:90-96): finite first coordinate, closed bounds.
        // NaN in any coordinate fails the comparisons, as in NumPy.
        bool ok = isfinite(c[0]);
#pragma unroll
        for (int d = 0; d < NDIM; ++d) ok = ok && (c[d] >= g.lo[d]) && (c[d] <= g.hi[d]);
        double frac[NDIM];
        long long base = g.col_0;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) {
            // float_sub / cell_sub_for_pts (fd_grid.py:117, 127), each op rounded separately
            double s = __ddiv_rn(__dsub_rn(c[d], g.lo[d]), g.delta[d]);
            double fl = floor(s);
            frac[d] = __dsub_rn(s, fl);
            base += (long long)fl * g.stride[d];
        }
        int rr[NC], cc[NC];
        double vv[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            long long off = 0;
            double w = 1.0;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                const int bit = (k >> (NDIM - 1 - d)) & 1;  // np.mgrid order: dim 0 slowest (lin_op.py:216-222)
                off += bit ? g.stride[d] : 0;
                w = __dmul_rn(w, bit ? frac[d] : __dsub_rn(1.0, frac[d]));  // lin_op.py:232-236
            }
            rr[k] = (int)(row_0 + i);
            cc[k] = ok ? (int)(base + off) : (int)g.col_0;
            vv[k] = ok ? w : 0.0;
        }
        if (NC % 4 == 0) {
            int4 *r4 = reinterpret_cast<int4 *>(rows + i * NC);
            int4 *c4 = reinterpret_cast<int4 *>(cols + i * NC);
#pragma unroll
            for (int k = 0; k < NC / 4; ++k) {
                r4[k] = make_int4(rr[4 * k], rr[4 * k + 1], rr[4 * k + 2], rr[4 * k + 3]);
                c4[k] = make_int4(cc[4 * k], cc[4 * k + 1], cc[4 * k + 2], cc[4 * k + 3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                rows[i * NC + k] = rr[k];
                cols[i * NC + k] = cc[k];
            }
        }
        double2 *v2 = reinterpret_cast<double2 *>(vals + i * NC);
#pragma unroll
        for (int k = 0; k < NC / 2; ++k) v2[k] = make_double2(vv[2 * k], vv[2 * k + 1]);
        valid[i] = ok ? 1 : 0;
        bad = !ok;
    }
    unsigned m = __ballot_sync(0xffffffffu, bad);
    if (m && lane_id() == 0) atomicAdd(n_invalid, (unsigned long long)__popc(m));
}

// ---- stencils ------------------------------------------------------------------------------

struct StencilDev {
    long long first[3], count[3];  // first valid centre subscript and number of centres per dim
};

// One thread per equation; offsets/coefficients live in global memory (tiny, L1-resident).
__global__ void __launch_bounds__(256)
stencil_coo_kernel(GridDev g, StencilDev st, int n_off, const int *__restrict__ offs,
                   const double *__restrict__ coef, long long row_0, int valid_only, long long n_eq,
                   int *__restrict__ rows, int *__restrict__ cols, double *__restrict__ vals,
                   long long *__restrict__ ind0) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eq) return;
    // unravel i over the valid-centre box in C order (np.meshgrid(indexing='ij').ravel, lin_op.py:102-103)
    long long sub[3] = {0, 0, 0};
    long long rem = i;
    for (int d = g.ndim - 1; d >= 0; --d) {
        sub[d] = st.first[d] + rem % st.count[d];
        rem /= st.count[d];
    }
    long long ctr = g.col_0;
    for (int d = 0; d < g.ndim; ++d) ctr += sub[d] * g.stride[d];
    ind0[i] = ctr;
    for (int k = 0; k < n_off; ++k) {
        long long col = g.col_0;
        bool in = true;
        for (int d = 0; d < g.ndim; ++d) {
            long long s = sub[d] + offs[d * n_off + k];
            in = in && s >= 0 && s < g.shape[d];
            col += s * g.stride[d];
        }
        double v = coef[k];
        if (!valid_only) {  // lin_op.py:118-121: out-of-grid offsets -> column 0, value * 0
            if (!in) col = 0;
            v = __dmul_rn(v, in ? 1.0 : 0.0);
        }
        rows[i * n_off + k] = (int)(row_0 + i);
        cols[i * n_off + k] = (int)col;
        vals[i * n_off + k] = v;
    }
}

// Spatial sort key of a point: (coarse cell, fine cell, time cell), row-major, so that points sharing grid
// nodes become neighbours in memory (gather locality of the CSR / ELL products).  Internal ordering only.
__global__ void __launch_bounds__(256)
cell_key_kernel(GridDev fine, GridDev coarse, PtsDev pts, long long n, unsigned long long *__restrict__ keys) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const GridDev &g = pass == 0 ? coarse : fine;
        const int nd = pass == 0 ? 2 : fine.ndim < 2 ? fine.ndim : 2;
        for (int d = 0; d < nd; ++d) {
            long long c = (long long)floor((pts.p[d][i] - g.lo[d]) / g.delta[d]);
            c = c < 0 ? 0 : (c >= g.shape[d] ? g.shape[d] - 1 : c);
            key = key * (unsigned long long)g.shape[d] + (unsigned long long)c;
        }
    }
    if (coarse.ndim > 2) {
        long long c = (long long)floor((pts.p[2][i] - coarse.lo[2]) / coarse.delta[2]);
        c = c < 0 ? 0 : (c >= coarse.shape[2] ? coarse.shape[2] - 1 : c);
        key = key * (unsigned long long)coarse.shape[2] + (unsigned long long)c;
    }
    keys[i] = key;
}

__global__ void __launch_bounds__(256)
coo_extent_kernel(long long nnz, const int *__restrict__ rows, const double *__restrict__ vals,
                  long long *__restrict__ out /* [0]=max_row, [1]=count */) {
    long long mx = -1, cnt = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        if (vals[i] != 0.0) {  // NaN != 0 is true, as in NumPy (lin_op.py:750)
            long long r = rows[i];
            mx = r > mx ? r : mx;
            ++cnt;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane_id() == 0) {
        atomicMax(out, mx);
        atomicAdd((unsigned long long *)(out + 1), (unsigned long long)cnt);
    }
}

}  // namespace lss

using namespace lss;

extern "C" int lss_version(void) { return 100; }
extern "C" const char *lss_last_error(void) { return lss::g_err; }
extern "C" int64_t lss_launch_count(void) { return lss::g_launches; }
extern "C" int lss_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int check_grid(const lss_grid_t *g) {
    if (!g || g->ndim < 1 || g->ndim > LSS_MAX_DIMS) return fail(LSS_EINVAL, "grid: ndim must be 1..3");
    long long nodes = 1;
    for (int d = 0; d < g->ndim; ++d) {
        if (g->shape[d] < 1) return fail(LSS_EINVAL, "grid: empty dimension %d", d);
        nodes *= g->shape[d];
    }
    if (g->col_0 + 2 * nodes >= (1ll << 31)) return fail(LSS_ERANGE, "grid: column index exceeds int32");
    return LSS_OK;
}

extern "C" int lss_interp_coo(const lss_grid_t *grid, const double *const *pts_h, int64_t n, int64_t row_0,
                              int32_t *rows, int32_t *cols, double *vals, uint8_t *valid, int64_t *n_invalid_h,
                              void *stream) {
    LSS_TRY(check_grid(grid));
    if (n < 0 || row_0 < 0 || row_0 + n >= (1ll << 31)) return fail(LSS_ERANGE, "interp_coo: row index exceeds int32");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_invalid_h) *n_invalid_h = 0;
    if (n == 0) return LSS_OK;
    GridDev g = to_dev(grid);
    PtsDev p{};
    for (int d = 0; d < grid->ndim; ++d) p.p[d] = pts_h[d];
    Scratch cnt;
    LSS_CUDA(cnt.alloc(sizeof(unsigned long long), st));
    LSS_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    int grid_dim = ceil_div(n, 256);
    auto *c = cnt.as<unsigned long long>();
    switch (grid->ndim) {
        case 1: interp_coo_kernel<1><<<grid_dim, 256, 0, st>>>(g, p, n, row_0, rows, cols, vals, valid, c); break;
        case 2: interp_coo_kernel<2><<<grid_dim, 256, 0, st>>>(g, p, n, row_0, rows, cols, vals, valid, c); break;
        default: interp_coo_kernel<3><<<grid_dim, 256, 0, st>>>(g, p, n, row_0, rows, cols, vals, valid, c); break;
    }
    LSS_LAUNCH_CHECK();
    unsigned long long bad = 0;
    LSS_CUDA(cudaMemcpyAsync(&bad, c, sizeof(bad), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if (n_invalid_h) *n_invalid_h = (int64_t)bad;
    return LSS_OK;
}

extern "C" int lss_stencil_count(const lss_grid_t *grid, int32_t n_off, const int32_t *offsets_h, int32_t valid_only,
                                 int64_t *n_eq_h, int64_t *first_h, int64_t *count_h) {
    LSS_TRY(check_grid(grid));
    if (n_off < 1) return fail(LSS_EINVAL, "stencil: n_off must be >= 1");
    int64_t n_eq = 1;
    for (int d = 0; d < grid->ndim; ++d) {
        int mn = 0, mx = 0;
        if (valid_only) {
            mn = mx = offsets_h[d * n_off];
            for (int k = 1; k < n_off; ++k) {
                int o = offsets_h[d * n_off + k];
                mn = o < mn ? o : mn;
                mx = o > mx ? o : mx;
            }
        }
        int64_t first = -mn > 0 ? -mn : 0;                                          // np.maximum(0, -min_delta)
        int64_t last = grid->shape[d] < grid->shape[d] - mx ? grid->shape[d] : grid->shape[d] - mx;  // np.minimum(Ni, Ni-max)
        int64_t cnt = last > first ? last - first : 0;
        if (first_h) first_h[d] = first;
        if (count_h) count_h[d] = cnt;
        n_eq *= cnt;
    }
    *n_eq_h = n_eq;
    return LSS_OK;
}

extern "C" int lss_stencil_coo(const lss_grid_t *grid, int32_t n_off, const int32_t *offsets_h, const double *coeffs_h,
                               int64_t row_0, int32_t valid_only, int64_t n_eq, int32_t *rows, int32_t *cols,
                               double *vals, int64_t *ind0, void *stream) {
    int64_t first[3] = {0, 0, 0}, count[3] = {1, 1, 1}, expect = 0;
    LSS_TRY(lss_stencil_count(grid, n_off, offsets_h, valid_only, &expect, first, count));
    if (expect != n_eq) return fail(LSS_EINVAL, "stencil_coo: n_eq=%lld but the template yields %lld", (long long)n_eq, (long long)expect);
    if (row_0 + n_eq >= (1ll << 31) || n_eq * n_off >= (1ll << 31)) return fail(LSS_ERANGE, "stencil_coo: index exceeds int32");
    if (n_eq == 0) return LSS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Scratch offs, coef;
    size_t ob = sizeof(int) * grid->ndim * n_off, cb = sizeof(double) * n_off;
    LSS_CUDA(offs.alloc(ob, st));
    LSS_CUDA(coef.alloc(cb, st));
    LSS_CUDA(cudaMemcpyAsync(offs.p, offsets_h, ob, cudaMemcpyHostToDevice, st));
    LSS_CUDA(cudaMemcpyAsync(coef.p, coeffs_h, cb, cudaMemcpyHostToDevice, st));
    GridDev g = to_dev(grid);
    StencilDev sd;
    for (int d = 0; d < 3; ++d) {
        sd.first[d] = first[d];
        sd.count[d] = count[d];
    }
    stencil_coo_kernel<<<ceil_div(n_eq, 256), 256, 0, st>>>(g, sd, n_off, offs.as<int>(), coef.as<double>(), row_0,
                                                            valid_only, n_eq, rows, cols, vals, (long long *)ind0);
    LSS_LAUNCH_CHECK();
    // the host staging buffers must outlive the async copies
    LSS_CUDA(cudaStreamSynchronize(st));
    return LSS_OK;
}

extern "C" int lss_coo_extent(int64_t nnz, const int32_t *rows, const double *vals, int64_t *max_row_h,
                              int64_t *n_nonzero_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long h[2] = {-1, 0};
    if (nnz > 0) {
        Scratch out;
        LSS_CUDA(out.alloc(2 * sizeof(long long), st));
        LSS_CUDA(cudaMemcpyAsync(out.p, h, sizeof(h), cudaMemcpyHostToDevice, st));
        coo_extent_kernel<<<stride_grid(nnz, 256), 256, 0, st>>>(nnz, rows, vals, out.as<long long>());
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
    }
    if (max_row_h) *max_row_h = h[0];
    if (n_nonzero_h) *n_nonzero_h = h[1];
    return LSS_OK;
}

extern "C" int lss_cell_keys(const lss_grid_t *fine, const lss_grid_t *coarse, const double *const *pts_h, int64_t n,
                             uint64_t *keys, int32_t *key_bits_h, void *stream) {
    LSS_TRY(check_grid(fine));
    LSS_TRY(check_grid(coarse));
    if (fine->ndim < 2 || coarse->ndim < 2) return fail(LSS_EINVAL, "cell_keys: grids need at least 2 dimensions");
    double span = 1.0;
    for (int d = 0; d < 2; ++d) span *= (double)fine->shape[d] * (double)coarse->shape[d];
    if (coarse->ndim > 2) span *= (double)coarse->shape[2];
    int bits = 1;
    while (bits < 63 && span > (double)(1ull << bits)) ++bits;
    if (bits >= 63) return fail(LSS_ERANGE, "cell_keys: key does not fit in 63 bits");
    if (key_bits_h) *key_bits_h = bits;
    if (n <= 0) return LSS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PtsDev p{};
    for (int d = 0; d < coarse->ndim && d < 3; ++d) p.p[d] = pts_h[d];
    cell_key_kernel<<<ceil_div(n, 256), 256, 0, st>>>(to_dev(fine), to_dev(coarse), p, n, (unsigned long long *)keys);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}
