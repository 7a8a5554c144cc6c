// Coarse level of the two-level preconditioner: Galerkin set-up, dense inverse, stand-alone apply (see coarse.cuh).
#include <stdlib.h>

#include <cooperative_groups.h>

#include "coarse.cuh"

namespace lss {

namespace cgx = cooperative_groups;

// A += P^T N_gg P and B += P^T N_gE: one warp per fine row i, lanes over its non-zeros; every non-zero (i, j, v) adds
// w_ia w_jb v to A[parent_a(i)][parent_b(j)] (<= 16 RED.ADD.F64) or w_ia v to B[parent_a(i)][j - e0] (<= 4).
__global__ void __launch_bounds__(256)
galerkin_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                const int4 *__restrict__ cp_idx, const float4 *__restrict__ cp_w, int n_c, int64_t e0, int64_t e1, int n_e,
                double *__restrict__ A, double *__restrict__ B, int only_b) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = lane_id();
    for (int64_t i = warp; i < n_rows; i += n_warps) {
        if (i >= e0 && i < e1) continue;
        const int4 pi = __ldg(cp_idx + i);
        if (pi.x < 0 && pi.y < 0 && pi.z < 0 && pi.w < 0) continue;
        const float4 wf = __ldg(cp_w + i);
        const int pa[4] = {pi.x, pi.y, pi.z, pi.w};
        const double wa[4] = {(double)wf.x, (double)wf.y, (double)wf.z, (double)wf.w};
        const int kb = indptr[i], ke = indptr[i + 1];
        for (int k = kb + (int)lane; k < ke; k += 32) {
            const int j = ldg_stream(indices + k);
            const double v = ldg_stream(data + k);
            if (j >= e0 && j < e1) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (pa[a] >= 0) atomicAdd(B + (size_t)pa[a] * n_e + (j - e0), wa[a] * v);
                continue;
            }
            if (only_b) continue;
            const int4 pj = __ldg(cp_idx + j);
            const float4 wj = __ldg(cp_w + j);
            const int pb[4] = {pj.x, pj.y, pj.z, pj.w};
            const double wb[4] = {(double)wj.x, (double)wj.y, (double)wj.z, (double)wj.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (pa[a] < 0) continue;
                const double wv = wa[a] * v;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (pb[b] >= 0) atomicAdd(A + (size_t)pa[a] * n_c + pb[b], wv * wb[b]);
            }
        }
    }
}

// A[a][b] -= sum_e B[a][e] einv[e] B[b][e]   (32 x 32 output tile per CTA, 32-wide slabs of e through shared memory)
__global__ void __launch_bounds__(256)
schur_sub_kernel(int n_c, int n_e, const double *__restrict__ B, const double *__restrict__ einv, double *__restrict__ A) {
    __shared__ double sa[32][33], sb[32][33];
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8 threads, 4 outputs each
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int e0 = 0; e0 < n_e; e0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int e = e0 + tx;
            const bool in = e < n_e;
            sa[r][tx] = (in && a0 + r < n_c) ? B[(size_t)(a0 + r) * n_e + e] * einv[e] : 0.0;
            sb[r][tx] = (in && b0 + r < n_c) ? B[(size_t)(b0 + r) * n_e + e] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
            const double bv = sb[tx][e];
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] += sa[ty + 8 * t][e] * bv;
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int a = a0 + ty + 8 * t, b = b0 + tx;
        if (a < n_c && b < n_c) A[(size_t)a * n_c + b] -= acc[t];
    }
}

// Coarse unknowns no fine unknown interpolates from (or whose rows vanished) get a unit diagonal so that S_c stays invertible.
__global__ void __launch_bounds__(256) coarse_fix_diag_kernel(int n_c, double *__restrict__ A) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_c && !(A[(size_t)m * n_c + m] > 0.0)) A[(size_t)m * n_c + m] = 1.0;
}

// Plain Gauss-Jordan inverse without pivoting (S_c is symmetric positive definite), ping-pong between two buffers so that no
// CTA reads a row another CTA is rewriting: one grid barrier per pivot.  Superseded by the blocked sweep below; kept behind
// LSS_GJ_PLAIN=1 as its cross-check.  After n steps the inverse is in (n odd ? Y : X).
__global__ void __launch_bounds__(256)
gj_inverse_kernel(int n, double *X, double *Y, int *bad) {
    cgx::grid_group grid = cgx::this_grid();
    double *src = X, *dst = Y;
    for (int k = 0; k < n; ++k) {
        const double piv = src[(size_t)k * n + k];
        if (!(piv > 0.0) || !(piv < 1e300)) {           // every thread of the grid reads the same value: uniform exit
            if (blockIdx.x == 0 && threadIdx.x == 0) *bad = 1;
            return;
        }
        const double ip = 1.0 / piv;
        const double *rk = src + (size_t)k * n;
        for (int i = blockIdx.x; i < n; i += gridDim.x) {
            const double *ri = src + (size_t)i * n;
            double *ro = dst + (size_t)i * n;
            if (i == k) {
                for (int j = threadIdx.x; j < n; j += blockDim.x) ro[j] = (j == k) ? ip : ri[j] * ip;
            } else {
                const double f = ri[k];
                for (int j = threadIdx.x; j < n; j += blockDim.x) ro[j] = (j == k) ? -f * ip : ri[j] - f * (rk[j] * ip);
            }
        }
        grid.sync();
        double *t = src; src = dst; dst = t;
    }
}

// Blocked Gauss-Jordan inverse (32 x 32 blocks, no pivoting: S_c is symmetric positive definite), ping-pong like the plain
// sweep above but ONE grid barrier per block of 32 pivots instead of one per pivot (721 unknowns: 23 barriers instead of 721,
// 3.7 ms -> a few hundred microseconds).  Step k, with D = src[k][k]:
//     dst[k][k] = D^-1,   dst[k][j] = D^-1 src[k][j],   dst[i][k] = -src[i][k] D^-1,   dst[i][j] = src[i][j] - src[i][k] D^-1 src[k][j].
// Every CTA inverts the pivot block itself in shared memory (the same sweep on 32 x 32, padded with the identity in the last
// block), so no barrier is needed between the pivot inverse and the tile updates; the tiles (i, j) are shared out over the grid.
// After ceil(n / 32) steps the inverse is in (steps odd ? Y : X).
constexpr int GJB = 32;
__global__ void __launch_bounds__(256)
gj_block_inverse_kernel(int n, double *X, double *Y, int *bad) {
    cgx::grid_group grid = cgx::this_grid();
    __shared__ double D[2][GJB][GJB + 1];
    __shared__ double Ta[GJB][GJB + 1], Tb[GJB][GJB + 1];
    __shared__ int s_bad;
    const int nb = (n + GJB - 1) / GJB;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8 threads, rows ty + 8 u
    double *src = X, *dst = Y;
    for (int kb = 0; kb < nb; ++kb) {
        const int k0 = kb * GJB, kn = min(GJB, n - k0);
        for (int r = ty; r < GJB; r += 8)
            D[0][r][tx] = (r < kn && tx < kn) ? src[(size_t)(k0 + r) * n + k0 + tx] : (r == tx ? 1.0 : 0.0);
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
        int cur = 0;
        for (int p = 0; p < GJB; ++p) {
            const double piv = D[cur][p][p];
            if (!(piv > 0.0) || !(piv < 1e300)) {
                if (threadIdx.x == 0) s_bad = 1;
            } else {
                const double ip = 1.0 / piv;
                const double pk = D[cur][p][tx] * ip;
                for (int r = ty; r < GJB; r += 8) {
                    double v;
                    if (r == p) {
                        v = (tx == p) ? ip : pk;
                    } else {
                        const double f = D[cur][r][p];
                        v = (tx == p) ? -f * ip : D[cur][r][tx] - f * pk;
                    }
                    D[cur ^ 1][r][tx] = v;
                }
            }
            __syncthreads();
            if (s_bad) break;               // every CTA holds the same block: the same decision everywhere
            cur ^= 1;
        }
        if (s_bad) {
            if (blockIdx.x == 0 && threadIdx.x == 0) *bad = 1;
            return;
        }
        const double (*Di)[GJB + 1] = D[cur];
        for (int t = blockIdx.x; t < nb * nb; t += gridDim.x) {
            const int ib = t / nb, jb = t - ib * nb;
            const int i0 = ib * GJB, j0 = jb * GJB, in = min(GJB, n - i0), jn = min(GJB, n - j0);
            if (ib == kb && jb == kb) {
                for (int r = ty; r < kn; r += 8)
                    if (tx < kn) dst[(size_t)(k0 + r) * n + k0 + tx] = Di[r][tx];
                continue;                    // uniform per CTA: no barrier skipped by part of the block
            }
            // stage src[i][k] (in x kn) and src[k][j] (kn x jn), zero padded
            for (int r = ty; r < GJB; r += 8) {
                Ta[r][tx] = (ib != kb && r < in && tx < kn) ? src[(size_t)(i0 + r) * n + k0 + tx] : 0.0;
                Tb[r][tx] = (jb != kb && r < kn && tx < jn) ? src[(size_t)(k0 + r) * n + j0 + tx] : 0.0;
            }
            __syncthreads();
            if (ib == kb) {                  // dst[k][j] = D^-1 src[k][j]
                for (int r = ty; r < kn; r += 8) {
                    double acc = 0.0;
#pragma unroll 8
                    for (int m = 0; m < GJB; ++m) acc += Di[r][m] * Tb[m][tx];
                    if (tx < jn) dst[(size_t)(k0 + r) * n + j0 + tx] = acc;
                }
            } else {
                double tv[4];                // T = src[i][k] D^-1
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = ty + 8 * u;
                    double acc = 0.0;
#pragma unroll 8
                    for (int m = 0; m < GJB; ++m) acc += Ta[r][m] * Di[m][tx];
                    tv[u] = acc;
                }
                if (jb == kb) {              // dst[i][k] = -T
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = ty + 8 * u;
                        if (r < in && tx < kn) dst[(size_t)(i0 + r) * n + k0 + tx] = -tv[u];
                    }
                } else {                     // dst[i][j] = src[i][j] - T src[k][j]
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < 4; ++u) Ta[ty + 8 * u][tx] = tv[u];
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = ty + 8 * u;
                        double acc = 0.0;
#pragma unroll 8
                        for (int m = 0; m < GJB; ++m) acc += Ta[r][m] * Tb[m][tx];
                        if (r < in && tx < jn) {
                            const size_t o = (size_t)(i0 + r) * n + j0 + tx;
                            dst[o] = src[o] - acc;
                        }
                    }
                }
            }
            __syncthreads();                 // Ta / Tb are restaged by the next tile
        }
        grid.sync();
        double *tmp = src; src = dst; dst = tmp;
    }
}

// Bt[e][m] = B[m][e]
__global__ void __launch_bounds__(256) transpose_kernel(int n_c, int n_e, const double *__restrict__ B, double *__restrict__ Bt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_c * n_e) return;
    const int e = (int)(t / n_c), m = (int)(t - (int64_t)e * n_c);
    Bt[t] = B[(size_t)m * n_e + e];
}

// out = (in + in^T) / 2
__global__ void __launch_bounds__(256) symmetrise_kernel(int n, const double *__restrict__ in, double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    const int r = (int)(t / n), c = (int)(t - (int64_t)r * n);
    out[t] = 0.5 * (in[t] + in[(size_t)c * n + r]);
}

// rc[m] = sum_k pt_w[k] r[pt_col[k]] over coarse row m: one warp per row
__global__ void __launch_bounds__(256)
coarse_restrict_kernel(int n_c, const int *__restrict__ pt_ptr, const int *__restrict__ pt_col, const float *__restrict__ pt_w,
                       const double *__restrict__ r, double *__restrict__ rc) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= n_c) return;
    double acc = 0.0;
    for (int k = pt_ptr[m] + (int)lane_id(); k < pt_ptr[m + 1]; k += 32) acc += (double)pt_w[k] * r[pt_col[k]];
    acc = warp_sum(acc);
    if (lane_id() == 0) rc[m] = acc;
}

// zc = inv rc: one warp per row
__global__ void __launch_bounds__(256)
coarse_matvec_kernel(int n_c, const double *__restrict__ inv, const double *__restrict__ rc, double *__restrict__ zc) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= n_c) return;
    double acc = 0.0;
    for (int j = (int)lane_id(); j < n_c; j += 32) acc += inv[(size_t)m * n_c + j] * rc[j];
    acc = warp_sum(acc);
    if (lane_id() == 0) zc[m] = acc;
}

// z[i] += sum_a w_ia zc[parent_a(i)]
__global__ void __launch_bounds__(256)
coarse_prolong_add_kernel(int64_t n_cols, const int4 *__restrict__ cp_idx, const float4 *__restrict__ cp_w, const double *__restrict__ zc,
                          double *__restrict__ z) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 p = __ldg(cp_idx + i);
        if (p.x < 0 && p.y < 0 && p.z < 0 && p.w < 0) continue;
        const float4 w = __ldg(cp_w + i);
        double a = 0.0;
        if (p.x >= 0) a += (double)w.x * zc[p.x];
        if (p.y >= 0) a += (double)w.y * zc[p.y];
        if (p.z >= 0) a += (double)w.z * zc[p.z];
        if (p.w >= 0) a += (double)w.w * zc[p.w];
        z[i] += a;
    }
}

int coarse_destroy(CoarseSpace *cs) {
    if (!cs) return LSS_OK;
    cudaStream_t st = cs->stream;
    void *ptrs[] = {cs->cp_idx, cs->cp_w, cs->pt_ptr, cs->pt_col, cs->pt_w, cs->A, cs->A2, cs->B, cs->Bt, cs->rc_part, cs->zc_part, cs->rc, cs->zc, cs->rczc, cs->bad};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, st);
    delete cs;
    return LSS_OK;
}

int coarse_create(CoarseSpace **out, int64_t n_cols, int n_c, const int32_t *cp_idx, const float *cp_w, const int32_t *pt_ptr,
                  const int32_t *pt_col, const float *pt_w, int64_t nnz_p, int n_e, cudaStream_t st) {
    if (n_c < 1 || n_c > CS_MAX) return fail(LSS_EINVAL, "coarse space: %d unknowns not in 1..%d", n_c, CS_MAX);
    if (!cp_idx || !cp_w || !pt_ptr || (nnz_p > 0 && (!pt_col || !pt_w))) return fail(LSS_EINVAL, "coarse space: null argument");
    CoarseSpace *cs = new CoarseSpace();
    memset(cs, 0, sizeof(*cs));
    cs->n_c = n_c; cs->n_cols = n_cols; cs->nnz_p = nnz_p; cs->n_e = n_e; cs->stream = st;
    Scratch::configure_pool();
    const size_t nn = (size_t)n_c * n_c, np = (size_t)(nnz_p > 0 ? nnz_p : 1);
    bool ok = cudaMallocAsync((void **)&cs->cp_idx, sizeof(int) * 4 * n_cols, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->cp_w, sizeof(float) * 4 * n_cols, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->pt_ptr, sizeof(int) * (n_c + 1), st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->pt_col, sizeof(int) * np, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->pt_w, sizeof(float) * np, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->A, sizeof(double) * nn, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->A2, sizeof(double) * nn, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->B, sizeof(double) * (size_t)n_c * (n_e > 0 ? n_e : 1), st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->Bt, sizeof(double) * (size_t)n_c * (n_e > 0 ? n_e : 1), st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->rc_part, sizeof(double) * CS_CHUNKS_MAX * n_c, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->zc_part, sizeof(double) * CS_CHUNKS_MAX * n_c, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->rc, sizeof(double) * n_c, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->zc, sizeof(double) * n_c, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->rczc, sizeof(double) * n_c, st) == cudaSuccess &&
              cudaMallocAsync((void **)&cs->bad, sizeof(int), st) == cudaSuccess;
    if (!ok) { coarse_destroy(cs); return fail(LSS_ECUDA, "coarse space: cudaMallocAsync failed"); }
    cudaMemcpyAsync(cs->cp_idx, cp_idx, sizeof(int) * 4 * n_cols, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(cs->cp_w, cp_w, sizeof(float) * 4 * n_cols, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(cs->pt_ptr, pt_ptr, sizeof(int) * (n_c + 1), cudaMemcpyDeviceToDevice, st);
    if (nnz_p > 0) {
        cudaMemcpyAsync(cs->pt_col, pt_col, sizeof(int) * np, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(cs->pt_w, pt_w, sizeof(float) * np, cudaMemcpyDeviceToDevice, st);
    }
    if (cudaGetLastError() != cudaSuccess) { coarse_destroy(cs); return fail(LSS_ECUDA, "coarse space: copy failed"); }
    cs->inv = cs->A;
    *out = cs;
    return LSS_OK;
}

int coarse_factor(CoarseSpace *cs, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, int64_t e0,
                  int64_t e1, const double *einv, lss_comm *comm) {
    cudaStream_t st = cs->stream;
    const int n = cs->n_c, n_e = (int)(e1 - e0);
    cs->ready = false;
    if (n_e != cs->n_e && n_e > 0) return fail(LSS_EINVAL, "coarse space: eliminated block changed size");
    const size_t nn = (size_t)n * n;
    LSS_CUDA(cudaMemsetAsync(cs->A, 0, sizeof(double) * nn, st));
    if (n_e > 0) LSS_CUDA(cudaMemsetAsync(cs->B, 0, sizeof(double) * (size_t)n * n_e, st));
    LSS_CUDA(cudaMemsetAsync(cs->bad, 0, sizeof(int), st));
    galerkin_kernel<<<stride_grid(n_rows * 32, 256, 16), 256, 0, st>>>(n_rows, indptr, indices, data, reinterpret_cast<const int4 *>(cs->cp_idx),
                                                                       reinterpret_cast<const float4 *>(cs->cp_w), n, e0, e1, n_e > 0 ? n_e : 1,
                                                                       cs->A, cs->B, 0);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(comm, cs->A, (int64_t)nn, 0, 0, st));
    if (n_e > 0) {
        LSS_TRY(comm_allreduce(comm, cs->B, (int64_t)n * n_e, 0, 0, st));
        dim3 g((n + 31) / 32, (n + 31) / 32);
        schur_sub_kernel<<<g, 256, 0, st>>>(n, n_e, cs->B, einv, cs->A);
        LSS_LAUNCH_CHECK();
    }
    coarse_fix_diag_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, cs->A);
    LSS_LAUNCH_CHECK();
    const bool plain = getenv("LSS_GJ_PLAIN") != nullptr;       // the one-barrier-per-pivot sweep, kept for cross-checks
    int steps = n;
    {
        int nn_i = n;
        double *X = cs->A, *Y = cs->A2;
        int *bad = cs->bad;
        void *args[] = {&nn_i, &X, &Y, &bad};
        if (plain) {
            int grid = 2 * kSMs;
            if (grid > n) grid = n;
            LSS_CUDA(cudaLaunchCooperativeKernel((void *)gj_inverse_kernel, dim3(grid), dim3(256), args, 0, st));
        } else {
            steps = (n + GJB - 1) / GJB;
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gj_block_inverse_kernel, 256, 0);
            int grid = kSMs * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
            if (grid > steps * steps) grid = steps * steps;
            LSS_CUDA(cudaLaunchCooperativeKernel((void *)gj_block_inverse_kernel, dim3(grid), dim3(256), args, 0, st));
        }
        ++g_launches;
    }
    double *res = (steps & 1) ? cs->A2 : cs->A, *other = (steps & 1) ? cs->A : cs->A2;
    symmetrise_kernel<<<ceil_div((int64_t)nn, 256), 256, 0, st>>>(n, res, other);
    LSS_LAUNCH_CHECK();
    cs->inv = other;
    int bad_h = 0;
    LSS_CUDA(cudaMemcpyAsync(&bad_h, cs->bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    cs->ready = bad_h == 0;
    return LSS_OK;
}

// B = P^T N_gE (and its transpose) for the CURRENT values of N, summed over ranks: the row-sharded CG needs it exactly
// for T^T (P z_c) even when S_c itself is kept from the first weights.
int coarse_update_b(CoarseSpace *cs, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, int64_t e0,
                    int64_t e1, lss_comm *comm) {
    cudaStream_t st = cs->stream;
    const int n = cs->n_c, n_e = (int)(e1 - e0);
    if (n_e <= 0) return LSS_OK;
    LSS_CUDA(cudaMemsetAsync(cs->B, 0, sizeof(double) * (size_t)n * n_e, st));
    galerkin_kernel<<<stride_grid(n_rows * 32, 256, 16), 256, 0, st>>>(n_rows, indptr, indices, data, reinterpret_cast<const int4 *>(cs->cp_idx),
                                                                       reinterpret_cast<const float4 *>(cs->cp_w), n, e0, e1, n_e, cs->A, cs->B, 1);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(comm, cs->B, (int64_t)n * n_e, 0, 0, st));
    transpose_kernel<<<ceil_div((int64_t)n * n_e, 256), 256, 0, st>>>(n, n_e, cs->B, cs->Bt);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int coarse_apply_add(CoarseSpace *cs, const double *r, double *z, lss_comm *comm) {
    cudaStream_t st = cs->stream;
    const int n = cs->n_c;
    coarse_restrict_kernel<<<ceil_div((int64_t)n * 32, 256), 256, 0, st>>>(n, cs->pt_ptr, cs->pt_col, cs->pt_w, r, cs->rc);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(comm, cs->rc, n, 0, 0, st));       // row-sharded ranks hold P^T at their own columns only
    coarse_matvec_kernel<<<ceil_div((int64_t)n * 32, 256), 256, 0, st>>>(n, cs->inv, cs->rc, cs->zc);
    LSS_LAUNCH_CHECK();
    coarse_prolong_add_kernel<<<stride_grid(cs->n_cols, 256), 256, 0, st>>>(cs->n_cols, reinterpret_cast<const int4 *>(cs->cp_idx),
                                                                            reinterpret_cast<const float4 *>(cs->cp_w), cs->zc, z);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

}  // namespace lss
