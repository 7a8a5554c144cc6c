// Explicit normal matrix  N = A^T W A  of the stacked system A = [G; C]  (LSsurf/smooth_fit.py:123-142).
//
// Every row of A is short (4 z0 + 8 dz interpolation corners + bias columns, stencils of <= 9 entries) and the
// points of one grid cell all hit the same nodes, so N has FEWER non-zeros than A (0.45 nnz(A) for the 10^7-point
// ICESat-2 workload: ~155 per unknown) and, unlike A^T (w .* (A p)), the product N p is a pure gather: no
// scatter atomics, no row-weight vector, 2.4x fewer bytes per CG iteration than the single-pass fused kernel.
//
//   symbolic (once per solver)   every row emits the keys (c_i << bits | c_j) of its K x K column pairs, except the
//       pairs the previous row emits too (points are stored in grid-cell order, so most rows repeat their
//       predecessor's columns); LSD radix sort; run heads become the CSR pattern of N.  Columns held at zero
//       (reference epoch) are left out of N altogether.
//   numeric  (per outer iteration)   N += sum_r (w_new[r] - w_cur[r]) a_r a_r^T over the rows whose weight changed:
//       a full pass the first time, afterwards only the rows the 3-sigma edit switched on or off.  One warp works
//       on a 32-row chunk: the K*K pair products of consecutive rows with identical columns (points of one track in
//       one grid cell) are summed in registers, then 32 lanes find the slots of (c_i, c_j) by binary search in row
//       c_i of the pattern and issue one RED.ADD.F64 per pair and run.
//   product   the nnz-tiled CSR kernel of spmv.cu (deterministic, no atomics).
#include <vector>

#include "common.cuh"

namespace lss {

constexpr int NK_MAX = 24;          // longest ELL row (lss_solver_create_blocks checks): 4 z0 + 8 dz + bias + slope + sensor-grid columns
constexpr int NU_WARPS = 4;         // warps per CTA of the numeric kernel

__device__ __forceinline__ bool entry_valid(int c, double v, const unsigned char *__restrict__ fixed) {
    return v != 0.0 && !(fixed && fixed[c]);
}

// Row-sharded N (one rank of several): only the rows of this rank's own columns are kept, plus, in the rows of the
// eliminated block E = [e0, e1), the entries at own columns (N_E,own: what the rank needs for its share of N_Eg v).
struct OwnFilter {
    const unsigned char *owned;      // null: keep everything
    int64_t e0, e1;
};
__device__ __forceinline__ bool pair_kept(const OwnFilter &f, int ci, int cj) {
    if (!f.owned) return true;
    return f.owned[ci] || (ci >= f.e0 && ci < f.e1 && f.owned[cj]);
}

// counts[r] = number of column pairs row r contributes that row r - 1 does not
__global__ void __launch_bounds__(256)
normal_pair_count_kernel(int64_t n_rows, int64_t n_pad, int K, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                         const unsigned char *__restrict__ fixed, uint32_t *__restrict__ counts, OwnFilter flt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int nv = 0, ns = 0;
    int c[NK_MAX];
    unsigned valid = 0, same = 0;
    for (int k = 0; k < K; ++k) {
        c[k] = ell_col[(int64_t)k * n_pad + r];
        const double v = ell_val[(int64_t)k * n_pad + r];
        if (!entry_valid(c[k], v, fixed)) continue;
        ++nv;
        valid |= 1u << k;
        if (r > 0 && ell_col[(int64_t)k * n_pad + r - 1] == c[k] && ell_val[(int64_t)k * n_pad + r - 1] != 0.0) { ++ns; same |= 1u << k; }
    }
    if (!flt.owned) { counts[r] = (uint32_t)(nv * nv - ns * ns); return; }
    int n = 0;
    for (int i = 0; i < K; ++i) {
        if (!((valid >> i) & 1u)) continue;
        for (int j = 0; j < K; ++j) {
            if (!((valid >> j) & 1u)) continue;
            if ((same >> i) & (same >> j) & 1u) continue;
            if (pair_kept(flt, c[i], c[j])) ++n;
        }
    }
    counts[r] = (uint32_t)n;
}

__global__ void __launch_bounds__(256)
normal_pair_emit_kernel(int64_t n_rows, int64_t n_pad, int K, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                        const unsigned char *__restrict__ fixed, const uint32_t *__restrict__ offsets, int cbits,
                        uint64_t *__restrict__ keys, OwnFilter flt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int c[NK_MAX];
    unsigned valid = 0, same = 0;
    for (int k = 0; k < K; ++k) {
        c[k] = ell_col[(int64_t)k * n_pad + r];
        const double v = ell_val[(int64_t)k * n_pad + r];
        if (!entry_valid(c[k], v, fixed)) continue;
        valid |= 1u << k;
        if (r > 0 && ell_col[(int64_t)k * n_pad + r - 1] == c[k] && ell_val[(int64_t)k * n_pad + r - 1] != 0.0) same |= 1u << k;
    }
    uint64_t *out = keys + offsets[r];
    for (int i = 0; i < K; ++i) {
        if (!((valid >> i) & 1u)) continue;
        for (int j = 0; j < K; ++j) {
            if (!((valid >> j) & 1u)) continue;
            if ((same >> i) & (same >> j) & 1u) continue;
            if (!pair_kept(flt, c[i], c[j])) continue;
            *out++ = ((uint64_t)(uint32_t)c[i] << cbits) | (uint64_t)(uint32_t)c[j];
        }
    }
}

__global__ void __launch_bounds__(256)
normal_heads_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *__restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// heads of the sorted key runs -> CSR pattern (pos = exclusive scan of the head flags)
__global__ void __launch_bounds__(256)
normal_pattern_kernel(const uint64_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ pos, int cbits, int64_t n_cols,
                      int *__restrict__ indptr, int *__restrict__ indices, const uint32_t *__restrict__ total) {
    const uint64_t cmask = (1ull << cbits) - 1ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        const int64_t row = (int64_t)(k >> cbits);
        if (i == 0 || keys[i - 1] != k) {
            const uint32_t p = pos[i];
            indices[p] = (int)(k & cmask);
            const int64_t prev_row = (i == 0) ? -1 : (int64_t)(keys[i - 1] >> cbits);
            for (int64_t r = prev_row + 1; r <= row; ++r) indptr[r] = (int)p;
        }
        if (i == n - 1) {
            const int t = (int)(*total);
            for (int64_t r = row + 1; r <= n_cols; ++r) indptr[r] = t;
        }
    }
}

__global__ void __launch_bounds__(256) normal_fill_i32_kernel(int *p, int64_t n, int v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// N += sum_r (w_new[r] - w_cur[r]) a_r a_r^T ; w_cur <- w_new.  Warp per 32-row chunk, rows staged in shared memory,
// pairs spread over the lanes (pair e = i * K + j on lane e mod 32).  KM: compile-time bound of K (16 for plain fits: fewer
// pair registers per lane; 24 for rows with slope / sensor-grid columns).
template <int KM>
__global__ void __launch_bounds__(NU_WARPS * 32)
normal_update_kernel(int64_t n_rows, int64_t n_pad, int K, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                     const unsigned char *__restrict__ fixed, const double *__restrict__ w_new, double *__restrict__ w_cur,
                     const int *__restrict__ indptr, const int *__restrict__ indices, double *__restrict__ data,
                     unsigned long long *__restrict__ n_missing, OwnFilter flt) {
    __shared__ int sc[NU_WARPS][32][KM];
    __shared__ double sv[NU_WARPS][32][KM];
    const int w = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int64_t n_chunks = (n_rows + 31) >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * NU_WARPS + w, n_warps = (int64_t)gridDim.x * NU_WARPS;
    const int KK = K * K;
    for (int64_t chunk = warp0; chunk < n_chunks; chunk += n_warps) {
        const int64_t r = chunk * 32 + lane;
        const bool in = r < n_rows;
        const double wn = in ? w_new[r] : 0.0;
        const double wo = in ? w_cur[r] : 0.0;
        const double dw = wn - wo;
        unsigned act = __ballot_sync(0xffffffffu, dw != 0.0);
        if (act == 0u) continue;
        if (dw != 0.0) {
            w_cur[r] = wn;
            for (int k = 0; k < K; ++k) {
                const int c = ldg_stream(ell_col + (int64_t)k * n_pad + r);
                const double v = ldg_stream(ell_val + (int64_t)k * n_pad + r);
                sc[w][lane][k] = entry_valid(c, v, fixed) ? c : -1;
                sv[w][lane][k] = v;
            }
        }
        __syncwarp();
        // Rows are stored in grid-cell order, so consecutive rows mostly hold the SAME columns (points of one track in one
        // cell): their K x K products are summed in registers and the pattern is searched once per run, not per row.
        constexpr int T = (KM * KM + 31) / 32;
        double accp[T];
#pragma unroll
        for (int t = 0; t < T; ++t) accp[t] = 0.0;
        int cur = -1;
        auto flush = [&](int rep) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int e = (int)lane + 32 * t;
                const double val = accp[t];
                accp[t] = 0.0;
                if (e >= KK || val == 0.0) continue;
                const int i = e / K, j = e - i * K;
                const int ci = sc[w][rep][i], cj = sc[w][rep][j];
                if (ci < 0 || cj < 0) continue;
                if (!pair_kept(flt, ci, cj)) continue;
                int lo = __ldg(indptr + ci), hi = __ldg(indptr + ci + 1);
                const int end = hi;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(indices + mid) < cj) lo = mid + 1; else hi = mid;
                }
                if (lo < end && __ldg(indices + lo) == cj) atomicAdd(data + lo, val);
                else atomicAdd(n_missing, 1ull);     // cannot happen with a pattern built from the same rows
            }
        };
        while (act) {
            const int src = __ffs(act) - 1;
            act &= act - 1u;
            bool same = false;
            if (cur >= 0) {
                const int eq = ((int)lane < K) ? (sc[w][src][lane] == sc[w][cur][lane]) : 1;
                same = __all_sync(0xffffffffu, eq) != 0;
            }
            if (!same) {
                if (cur >= 0) flush(cur);
                cur = src;
            }
            const double dws = __shfl_sync(0xffffffffu, dw, src);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int e = (int)lane + 32 * t;
                if (e < KK) {
                    const int i = e / K, j = e - i * K;
                    accp[t] += dws * sv[w][src][i] * sv[w][src][j];
                }
            }
        }
        if (cur >= 0) flush(cur);
        __syncwarp();
    }
}

// diag[i] = N[i, i]
__global__ void __launch_bounds__(256)
normal_diag_kernel(int64_t n_cols, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                   double *__restrict__ diag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += (int64_t)gridDim.x * blockDim.x) {
        int lo = indptr[i], hi = indptr[i + 1];
        const int end = hi;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (indices[mid] < (int)i) lo = mid + 1; else hi = mid;
        }
        diag[i] = (lo < end && indices[lo] == (int)i) ? data[lo] : 0.0;
    }
}

__global__ void sum_u32_kernel(const uint32_t *__restrict__ in, int64_t n, unsigned long long *__restrict__ out);
static int sum_u32_to_u64(const uint32_t *in, int64_t n, unsigned long long *out, cudaStream_t st);

// ---- chunk-of-8 padded copy of N for the CG product (spmv_chunk8_kernel) ----
__global__ void __launch_bounds__(256)
c8_count_kernel(int64_t n_cols, const int *__restrict__ indptr, uint32_t *__restrict__ n_chunks) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_cols; r += (int64_t)gridDim.x * blockDim.x)
        n_chunks[r] = (uint32_t)((indptr[r + 1] - indptr[r] + 7) >> 3);
}

// one warp per row: row id of each of its chunks, padded columns (padding repeats the row's last column, value 0)
__global__ void __launch_bounds__(256)
c8_structure_kernel(int64_t n_cols, const int *__restrict__ indptr, const int *__restrict__ indices, const int *__restrict__ c8_ptr,
                    int *__restrict__ c8_row, int *__restrict__ c8_cols) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_cols; r += n_warps) {
        const int b = indptr[r], len = indptr[r + 1] - b;
        const int cb = c8_ptr[r], nc = c8_ptr[r + 1] - cb;
        for (int c = (int)lane_id(); c < nc; c += 32) c8_row[cb + c] = (int)r;
        for (int k = (int)lane_id(); k < nc * 8; k += 32) c8_cols[(int64_t)cb * 8 + k] = indices[b + (k < len ? k : len - 1)];
    }
}

// values of the padded copy: one thread per padded slot
__global__ void __launch_bounds__(256)
c8_values_kernel(int64_t n_slots, const int *__restrict__ indptr, const double *__restrict__ data, const int *__restrict__ c8_ptr,
                 const int *__restrict__ c8_row, double *__restrict__ c8_vals) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int r = c8_row[s >> 3];
        const int k = (int)(s - (int64_t)c8_ptr[r] * 8);
        const int b = indptr[r], len = indptr[r + 1] - b;
        c8_vals[s] = k < len ? data[b + k] : 0.0;
    }
}

static int c8_build(NormalMatrix *m) {
    cudaStream_t st = m->stream;
    const int64_t n = m->n_cols;
    Scratch cnt, tot;
    LSS_CUDA(cnt.alloc(sizeof(uint32_t) * (n + 1), st));
    LSS_CUDA(tot.alloc(sizeof(uint32_t), st));
    c8_count_kernel<<<stride_grid(n, 256), 256, 0, st>>>(n, m->indptr, cnt.as<uint32_t>());
    LSS_LAUNCH_CHECK();
    LSS_CUDA(cudaMallocAsync((void **)&m->c8_ptr, sizeof(int) * (n + 1), st));
    LSS_TRY(exclusive_scan_u32(cnt.as<uint32_t>(), reinterpret_cast<uint32_t *>(m->c8_ptr), n, tot.as<uint32_t>(), st));
    uint32_t total = 0;
    LSS_CUDA(cudaMemcpyAsync(&total, tot.p, sizeof(total), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if ((uint64_t)total * 8 >= (1ull << 31)) return fail(LSS_ERANGE, "normal matrix: padded copy exceeds 2^31 entries");
    LSS_CUDA(cudaMemcpyAsync(m->c8_ptr + n, &total, sizeof(int), cudaMemcpyHostToDevice, st));
    m->n_chunks = total;
    const size_t nc = total > 0 ? total : 1;
    LSS_CUDA(cudaMallocAsync((void **)&m->c8_row, sizeof(int) * nc, st));
    LSS_CUDA(cudaMallocAsync((void **)&m->c8_cols, sizeof(int) * nc * 8, st));
    LSS_CUDA(cudaMallocAsync((void **)&m->c8_vals, sizeof(double) * nc * 8, st));
    if (total > 0) {
        c8_structure_kernel<<<stride_grid(n * 32, 256, 16), 256, 0, st>>>(n, m->indptr, m->indices, m->c8_ptr, m->c8_row, m->c8_cols);
        LSS_LAUNCH_CHECK();
    }
    LSS_CUDA(cudaStreamSynchronize(st));        // `total` (host stack) was the source of an async copy
    m->c8_dirty = true;
    m->c8_part_rows = -1;
    return LSS_OK;
}

static int bits_for_cols(int64_t n_cols) {
    int b = 1;
    while (b < 31 && ((n_cols - 1) >> b)) ++b;
    return b;
}

int normal_destroy(NormalMatrix *m) {
    if (!m) return LSS_OK;
    cudaStream_t st = m->stream;
    if (m->indptr) cudaFreeAsync(m->indptr, st);
    if (m->indices) cudaFreeAsync(m->indices, st);
    if (m->data) cudaFreeAsync(m->data, st);
    if (m->w_cur) cudaFreeAsync(m->w_cur, st);
    if (m->missing) cudaFreeAsync(m->missing, st);
    if (m->c8_ptr) cudaFreeAsync(m->c8_ptr, st);
    if (m->c8_row) cudaFreeAsync(m->c8_row, st);
    if (m->c8_cols) cudaFreeAsync(m->c8_cols, st);
    if (m->c8_vals) cudaFreeAsync(m->c8_vals, st);
    spmv_tma_plan_destroy(&m->tma_plan);
    spmv_plan_destroy(&m->plan);
    spmv_plan_destroy(&m->plan_part);
    if (m->cp_indptr) cudaFreeAsync(m->cp_indptr, st);
    if (m->cp_rows) cudaFreeAsync(m->cp_rows, st);
    spmv_plan_destroy(&m->plan_cp);
    delete m;
    return LSS_OK;
}

int normal_create(NormalMatrix **out, int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val,
                  int64_t n_cols, const unsigned char *fixed, cudaStream_t st, const unsigned char *owned, int64_t own_e0, int64_t own_e1) {
    if (K < 1 || K > NK_MAX) return fail(LSS_EINVAL, "normal matrix: row length %d not in 1..%d", K, NK_MAX);
    if (n_cols >= (1ll << 31)) return fail(LSS_ERANGE, "normal matrix: too many columns");
    NormalMatrix *m = new NormalMatrix();
    memset(m, 0, sizeof(*m));
    m->n_rows = n_rows; m->n_pad = n_pad; m->K = K; m->n_cols = n_cols;
    m->ell_col = ell_col; m->ell_val = ell_val; m->fixed = fixed; m->stream = st;
    m->part_rows = -1;
    m->c8_part_rows = -1;
    m->cp_for_rows = -1;
    m->c8_dirty = true;
    m->owned = owned; m->own_e0 = own_e0; m->own_e1 = own_e1;
    const OwnFilter flt{owned, own_e0, own_e1};
    Scratch::configure_pool();
    int rc = LSS_OK;
    auto bail = [&](int code) { normal_destroy(m); return code; };
    if (cudaMallocAsync((void **)&m->indptr, sizeof(int) * (n_cols + 1 + 8), st) != cudaSuccess ||
        cudaMallocAsync((void **)&m->w_cur, sizeof(double) * (n_rows > 0 ? n_rows : 1), st) != cudaSuccess ||
        cudaMallocAsync((void **)&m->missing, sizeof(unsigned long long), st) != cudaSuccess)
        return bail(fail(LSS_ECUDA, "normal matrix: cudaMallocAsync failed"));
    cudaMemsetAsync(m->w_cur, 0, sizeof(double) * (n_rows > 0 ? n_rows : 1), st);
    cudaMemsetAsync(m->missing, 0, sizeof(unsigned long long), st);
    uint32_t n_pairs32 = 0, n_unique = 0;
    if (n_rows > 0) {
        Scratch counts, tot;
        if (counts.alloc(sizeof(uint32_t) * n_rows, st) != cudaSuccess || tot.alloc(2 * sizeof(uint32_t), st) != cudaSuccess)
            return bail(fail(LSS_ECUDA, "normal matrix: scratch allocation failed"));
        normal_pair_count_kernel<<<ceil_div(n_rows, 256), 256, 0, st>>>(n_rows, n_pad, K, ell_col, ell_val, fixed, counts.as<uint32_t>(), flt);
        ++g_launches;
        if (cudaGetLastError() != cudaSuccess) return bail(fail(LSS_ECUDA, "normal matrix: launch failed"));
        // 64-bit total first: the 32-bit scan below would wrap silently
        unsigned long long total64 = 0;
        {
            Scratch t64;
            if (t64.alloc(sizeof(unsigned long long), st) != cudaSuccess) return bail(fail(LSS_ECUDA, "normal matrix: scratch allocation failed"));
            cudaMemsetAsync(t64.p, 0, sizeof(unsigned long long), st);
            rc = sum_u32_to_u64(counts.as<uint32_t>(), n_rows, t64.as<unsigned long long>(), st);
            if (rc != LSS_OK) return bail(rc);
            cudaMemcpyAsync(&total64, t64.p, sizeof(total64), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) return bail(fail(LSS_ECUDA, "normal matrix: sync failed"));
        }
        if (total64 >= (1ull << 32)) return bail(fail(LSS_ERANGE, "normal matrix: %llu column pairs exceed the 32-bit sort range", total64));
        n_pairs32 = (uint32_t)total64;
        m->n_pairs = (int64_t)total64;
        rc = exclusive_scan_u32(counts.as<uint32_t>(), counts.as<uint32_t>(), n_rows, tot.as<uint32_t>(), st);
        if (rc != LSS_OK) return bail(rc);
        if (n_pairs32 > 0) {
            const int64_t np = n_pairs32;
            const int cbits = bits_for_cols(n_cols);
            Scratch ka, kb, va, vb, flags;
            if (ka.alloc(sizeof(uint64_t) * np, st) != cudaSuccess || kb.alloc(sizeof(uint64_t) * np, st) != cudaSuccess ||
                va.alloc(sizeof(uint32_t) * np, st) != cudaSuccess || vb.alloc(sizeof(uint32_t) * np, st) != cudaSuccess)
                return bail(fail(LSS_ECUDA, "normal matrix: %.1f GB of sort scratch not available", 24e-9 * (double)np));
            normal_pair_emit_kernel<<<ceil_div(n_rows, 256), 256, 0, st>>>(n_rows, n_pad, K, ell_col, ell_val, fixed, counts.as<uint32_t>(), cbits,
                                                                           ka.as<uint64_t>(), flt);
            ++g_launches;
            uint64_t *ks;
            uint32_t *vs;
            rc = radix_sort_pairs(ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), np, 0, 2 * cbits, 1, &ks, &vs, st);
            if (rc != LSS_OK) return bail(rc);
            // the payload buffer that does not hold the sorted payload is free: reuse it for the head flags
            uint32_t *fl = (vs == va.as<uint32_t>()) ? vb.as<uint32_t>() : va.as<uint32_t>();
            normal_heads_kernel<<<stride_grid(np, 256), 256, 0, st>>>(ks, np, fl);
            ++g_launches;
            rc = exclusive_scan_u32(fl, fl, np, tot.as<uint32_t>() + 1, st);
            if (rc != LSS_OK) return bail(rc);
            cudaMemcpyAsync(&n_unique, tot.as<uint32_t>() + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) return bail(fail(LSS_ECUDA, "normal matrix: symbolic phase failed: %s", cudaGetErrorString(cudaGetLastError())));
            if (n_unique >= (1u << 31)) return bail(fail(LSS_ERANGE, "normal matrix: more than 2^31 non-zeros"));
            m->nnz = n_unique;
            // + 8: the TMA product reads whole 16-byte pieces, up to 3 entries past the end of a tile
            if (cudaMallocAsync((void **)&m->indices, sizeof(int) * ((size_t)n_unique + 8), st) != cudaSuccess ||
                cudaMallocAsync((void **)&m->data, sizeof(double) * ((size_t)n_unique + 8), st) != cudaSuccess)
                return bail(fail(LSS_ECUDA, "normal matrix: cudaMallocAsync failed"));
            cudaMemsetAsync(m->data, 0, sizeof(double) * (size_t)n_unique, st);
            normal_pattern_kernel<<<stride_grid(np, 256, 16), 256, 0, st>>>(ks, np, fl, cbits, n_cols, m->indptr, m->indices, tot.as<uint32_t>() + 1);
            ++g_launches;
            if (cudaGetLastError() != cudaSuccess) return bail(fail(LSS_ECUDA, "normal matrix: launch failed"));
        }
    }
    if (m->nnz == 0) {
        normal_fill_i32_kernel<<<stride_grid(n_cols + 1, 256), 256, 0, st>>>(m->indptr, n_cols + 1, 0);
        ++g_launches;
    }
    rc = spmv_plan_create(&m->plan, n_cols, m->indptr, m->nnz, st);
    if (rc != LSS_OK) return bail(rc);
    *out = m;
    return LSS_OK;
}

int normal_update(NormalMatrix *m, const double *w) {
    if (m->n_rows <= 0 || m->nnz == 0) return LSS_OK;
    m->c8_dirty = true;
    const int64_t n_chunks = (m->n_rows + 31) / 32;
    int64_t need = (n_chunks + NU_WARPS - 1) / NU_WARPS;
    const int64_t cap = (int64_t)kSMs * 16;
    const int grid = (int)(need < cap ? need : cap);
    (m->K <= 16 ? normal_update_kernel<16> : normal_update_kernel<24>)<<<grid, NU_WARPS * 32, 0, m->stream>>>(m->n_rows, m->n_pad, m->K, m->ell_col, m->ell_val, m->fixed, w, m->w_cur,
                                                                m->indptr, m->indices, m->data, m->missing, OwnFilter{m->owned, m->own_e0, m->own_e1});
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int normal_diag(NormalMatrix *m, double *diag) {
    normal_diag_kernel<<<stride_grid(m->n_cols, 256), 256, 0, m->stream>>>(m->n_cols, m->indptr, m->indices, m->data, diag);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int normal_apply(NormalMatrix *m, const double *x, double *y, const double *stop, int64_t *launches) {
    return spmv_launch(m->n_cols, m->indptr, m->indices, m->data, x, y, LSS_SPMV_PLAIN, nullptr, m->nnz, &m->plan, stop, m->stream, launches);
}

int normal_apply_fast(NormalMatrix *m, int64_t n_rows, const double *x, double *y, bool y_is_zero, const double *stop, int64_t *launches,
                      double *dot_slots) {
    if (m->plan.n_tiles > 0 && m->plan.tile != 1024)
        return fail(LSS_EINVAL, "normal matrix product: the single-launch kernel needs the default spmv_variant (1024-entry tiles)");
    if (!y_is_zero) LSS_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * m->n_cols, m->stream));
    if (m->owned) {
        // row-sharded N: a tile of the nnz-tiled kernel would walk the long runs of empty rows between the owned column ranges
        // one by one, so the product runs over the compacted list of non-empty rows (built once: the pattern is fixed)
        const int64_t rows = n_rows < m->n_cols ? n_rows : m->n_cols;
        if (m->cp_for_rows != rows) {
            std::vector<int32_t> h(rows + 1), hp, hr;
            LSS_CUDA(cudaMemcpyAsync(h.data(), m->indptr, sizeof(int32_t) * (rows + 1), cudaMemcpyDeviceToHost, m->stream));
            LSS_CUDA(cudaStreamSynchronize(m->stream));
            hp.reserve(rows / 2 + 2);
            hr.reserve(rows / 2 + 2);
            for (int64_t r = 0; r < rows; ++r)
                if (h[r + 1] > h[r]) { hp.push_back(h[r]); hr.push_back((int32_t)r); }
            hp.push_back(h[rows]);
            if (m->cp_indptr) cudaFreeAsync(m->cp_indptr, m->stream);
            if (m->cp_rows) cudaFreeAsync(m->cp_rows, m->stream);
            spmv_plan_destroy(&m->plan_cp);
            m->cp_n = (int64_t)hr.size();
            m->cp_nnz = h[rows];
            LSS_CUDA(cudaMallocAsync((void **)&m->cp_indptr, sizeof(int32_t) * hp.size(), m->stream));
            LSS_CUDA(cudaMallocAsync((void **)&m->cp_rows, sizeof(int32_t) * (hr.size() + 1), m->stream));
            LSS_CUDA(cudaMemcpyAsync(m->cp_indptr, hp.data(), sizeof(int32_t) * hp.size(), cudaMemcpyHostToDevice, m->stream));
            if (!hr.empty()) LSS_CUDA(cudaMemcpyAsync(m->cp_rows, hr.data(), sizeof(int32_t) * hr.size(), cudaMemcpyHostToDevice, m->stream));
            LSS_CUDA(cudaStreamSynchronize(m->stream));           // the host vectors go out of scope
            if (m->cp_n > 0) LSS_TRY(spmv_plan_create(&m->plan_cp, m->cp_n, m->cp_indptr, m->cp_nnz, m->stream));
            m->cp_for_rows = rows;
        }
        return spmv_launch_atomic(m->cp_n, m->cp_indptr, m->indices, m->data, x, y, m->cp_nnz, &m->plan_cp, stop, m->stream, launches, dot_slots,
                                  m->cp_rows);
    }
    int variant = normal_product_variant();
    if (variant >= 10 && variant <= 11 && m->nnz > 0) {
        const int64_t rows = n_rows < m->n_cols ? n_rows : m->n_cols;
        if (!m->tma_plan.tile_row || m->tma_plan.n_rows != rows) {
            spmv_tma_plan_destroy(&m->tma_plan);
            int nz = 0;
            LSS_CUDA(cudaMemcpyAsync(&nz, m->indptr + rows, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
            LSS_CUDA(cudaStreamSynchronize(m->stream));
            m->tma_nnz = nz;
            LSS_TRY(spmv_tma_plan_create(&m->tma_plan, rows, m->indptr, nz, m->stream));
        }
        if (variant == 11 && spmv_tma_fits(&m->tma_plan)) {
            LSS_TRY(spmv_tma_plan_add_dict(&m->tma_plan, m->indices, m->tma_nnz));
            return spmv_launch_tmad(&m->tma_plan, m->indptr, m->data, x, y, stop, m->stream, launches, dot_slots);
        }
        if (spmv_tma_fits(&m->tma_plan))
            return spmv_launch_tma(&m->tma_plan, m->indptr, m->indices, m->data, x, y, stop, m->stream, launches, dot_slots);
        variant = 0;                     // rows too long for a ring stage (e.g. bias rows that were not eliminated)
    }
    if (variant >= 7 && m->nnz > 0) {
        if (!m->c8_ptr) LSS_TRY(c8_build(m));
        if (m->c8_dirty) {
            const int64_t n_slots = m->n_chunks * 8;
            if (n_slots > 0) {
                c8_values_kernel<<<stride_grid(n_slots, 256, 16), 256, 0, m->stream>>>(n_slots, m->indptr, m->data, m->c8_ptr, m->c8_row, m->c8_vals);
                LSS_LAUNCH_CHECK();
            }
            m->c8_dirty = false;
        }
        int64_t chunks = m->n_chunks;
        if (n_rows < m->n_cols) {
            if (m->c8_part_rows != n_rows) {
                int pc = 0, nz = 0;
                LSS_CUDA(cudaMemcpyAsync(&pc, m->c8_ptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
                LSS_CUDA(cudaMemcpyAsync(&nz, m->indptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
                LSS_CUDA(cudaStreamSynchronize(m->stream));
                m->c8_part_chunks = pc;
                m->c8_part_nnz = nz;
                m->c8_part_rows = n_rows;
            }
            chunks = m->c8_part_chunks;
        }
        return spmv_launch_chunk8(variant, chunks, m->c8_row, m->c8_cols, m->c8_vals, x, y, stop, m->stream, launches, dot_slots);
    }
    if (n_rows >= m->n_cols)
        return spmv_launch_normal(variant, m->n_cols, m->indptr, m->indices, m->data, x, y, m->nnz, &m->plan, stop, m->stream, launches, dot_slots);
    if (m->part_rows != n_rows) {          // first use with this row count: plan of the leading block
        spmv_plan_destroy(&m->plan_part);
        int nz = 0;
        LSS_CUDA(cudaMemcpyAsync(&nz, m->indptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        LSS_CUDA(cudaStreamSynchronize(m->stream));
        m->part_nnz = nz;
        LSS_TRY(spmv_plan_create(&m->plan_part, n_rows, m->indptr, nz, m->stream));
        m->part_rows = n_rows;
    }
    return spmv_launch_normal(variant, n_rows, m->indptr, m->indices, m->data, x, y, m->part_nnz, &m->plan_part, stop, m->stream, launches,
                              dot_slots);
}

__global__ void __launch_bounds__(256)
sum_u32_kernel(const uint32_t *__restrict__ in, int64_t n, unsigned long long *__restrict__ out) {
    unsigned long long v = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v += in[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0 && v) atomicAdd(out, v);
}

static int sum_u32_to_u64(const uint32_t *in, int64_t n, unsigned long long *out, cudaStream_t st) {
    sum_u32_kernel<<<stride_grid(n, 256), 256, 0, st>>>(in, n, out);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

}  // namespace lss

using namespace lss;

// ------------------------------------------------------------------------------------------ C ABI (tests, tools)
struct lss_normal {
    NormalMatrix *m;
    int32_t *ell_col;
    double *ell_val;
    unsigned char *fixed;
    cudaStream_t stream;
};

extern "C" int lss_normal_destroy(lss_normal_t *h) {
    if (!h) return LSS_OK;
    cudaStreamSynchronize(h->stream);
    normal_destroy(h->m);
    if (h->ell_col) cudaFreeAsync(h->ell_col, h->stream);
    if (h->ell_val) cudaFreeAsync(h->ell_val, h->stream);
    if (h->fixed) cudaFreeAsync(h->fixed, h->stream);
    delete h;
    return LSS_OK;
}

namespace lss {
__global__ void __launch_bounds__(256) normal_mark_fixed_kernel(int64_t n_fixed, const int *__restrict__ cols, unsigned char *__restrict__ fixed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fixed; i += (int64_t)gridDim.x * blockDim.x) fixed[cols[i]] = 1;
}
}

extern "C" int lss_normal_create(lss_normal_t **out, int64_t n_rows, int64_t n_cols, int32_t row_len, const int32_t *indptr,
                                 const int32_t *indices, const double *data, const int32_t *fixed_cols, int64_t n_fixed,
                                 int64_t *nnz_h, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out || n_rows < 0 || n_cols <= 0) return fail(LSS_EINVAL, "normal_create: bad sizes");
    if (row_len < 1 || row_len > NK_MAX) return fail(LSS_EINVAL, "normal_create: rows must hold 1..%d entries", NK_MAX);
    lss_normal *h = new lss_normal();
    memset(h, 0, sizeof(*h));
    h->stream = st;
    Scratch::configure_pool();
    const int64_t n_pad = (n_rows + 31) / 32 * 32;
    const size_t ne = (size_t)(n_pad > 0 ? n_pad : 32) * row_len;
    if (cudaMallocAsync((void **)&h->ell_col, sizeof(int32_t) * ne, st) != cudaSuccess ||
        cudaMallocAsync((void **)&h->ell_val, sizeof(double) * ne, st) != cudaSuccess ||
        cudaMallocAsync((void **)&h->fixed, (size_t)n_cols, st) != cudaSuccess) {
        lss_normal_destroy(h);
        return fail(LSS_ECUDA, "normal_create: cudaMallocAsync failed");
    }
    cudaMemsetAsync(h->fixed, 0, (size_t)n_cols, st);
    if (n_fixed > 0) {
        normal_mark_fixed_kernel<<<stride_grid(n_fixed, 256), 256, 0, st>>>(n_fixed, fixed_cols, h->fixed);
        ++g_launches;
    }
    int rc = ell_build(n_rows, row_len, indptr, indices, data, h->ell_col, h->ell_val, n_pad, st);
    if (rc == LSS_OK) rc = normal_create(&h->m, n_rows, n_pad, row_len, h->ell_col, h->ell_val, n_cols, h->fixed, st);
    if (rc != LSS_OK) { lss_normal_destroy(h); return rc; }
    if (nnz_h) *nnz_h = h->m->nnz;
    *out = h;
    return LSS_OK;
}

extern "C" int lss_normal_update(lss_normal_t *h, const double *w) {
    if (!h || !w) return fail(LSS_EINVAL, "normal_update: null argument");
    return normal_update(h->m, w);
}

extern "C" int lss_normal_export(lss_normal_t *h, int32_t *indptr, int32_t *indices, double *data) {
    if (!h) return fail(LSS_EINVAL, "normal_export: null handle");
    cudaStream_t st = h->stream;
    NormalMatrix *m = h->m;
    LSS_CUDA(cudaMemcpyAsync(indptr, m->indptr, sizeof(int) * (m->n_cols + 1), cudaMemcpyDeviceToDevice, st));
    if (m->nnz > 0) {
        LSS_CUDA(cudaMemcpyAsync(indices, m->indices, sizeof(int) * m->nnz, cudaMemcpyDeviceToDevice, st));
        LSS_CUDA(cudaMemcpyAsync(data, m->data, sizeof(double) * m->nnz, cudaMemcpyDeviceToDevice, st));
    }
    unsigned long long miss = 0;
    LSS_CUDA(cudaMemcpyAsync(&miss, m->missing, sizeof(miss), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    if (miss) return fail(LSS_EINVAL, "normal matrix: %llu updates fell outside the pattern", miss);
    return LSS_OK;
}

extern "C" int lss_normal_apply(lss_normal_t *h, const double *x, double *y) {
    if (!h || !x || !y) return fail(LSS_EINVAL, "normal_apply: null argument");
    return normal_apply(h->m, x, y, nullptr, nullptr);
}
