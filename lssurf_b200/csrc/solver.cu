// fp64 preconditioned conjugate gradients on the normal equations (A^T W A) x = A^T W b.
//
// Replaces `sparseqr.solve(Ip_r.dot(TCinv.dot(Gcoo)), Ip_r.dot(TCinv.dot(rhs)))`
// (LSsurf/smooth_fit.py:123-142).  The row selection Ip_r and the 1/sigma scaling TCinv are never
// materialised: they are the per-row weight vector w (0 for edited rows).  The column selection
// Ip_c (reference epoch, constraint_functions.py:128-151) is a zero in the preconditioner, which
// keeps those unknowns at exactly 0.
//
// Operator forms (chosen at creation, lss_solver_info reports which):
//   explicit normal matrix (default)  N = A^T W A built on the device (normal.cu), q = N p by ONE gather product per
//                                     iteration (spmv.cu, or the grid-structured layout of structured.cu), then ONE
//                                     cooperative vector-step kernel: alpha, x, r, z = (block-Jacobi + coarse level) r,
//                                     beta, p, eliminated bias columns slaved
//   fused ELL pass                    q += A^T (w .* (A p)) in one pass over an ELL copy of A
//   generic CSR pair                  u = w .* (A p), q = A^T u: two deterministic CSR products (any row length)
// Several GPUs: rows of N sharded by preconditioner blocks, every exchange of the iteration (p.q, z, coarse restriction,
// bias coupling) done by the step kernel itself with stores into peer memory over NVLink (cg_step_shard_kernel below,
// mailbox in comm.cu) -- no collective per iteration; the older NCCL all-reduce per iteration remains for the
// generic forms.
// All scalars (alpha, beta, r.z) stay on the device; the host only polls a convergence flag once
// per batch of iterations, one batch ahead of the GPU, so the queue never drains.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"
#include "precond.cuh"
#include "coarse.cuh"
#include "structured.cuh"

namespace lss {

namespace cgx = cooperative_groups;
constexpr int PQ_SLOTS = 256;   // p.q partial sums written by the fused operator kernel

constexpr int VEC_BLOCK = 256;
constexpr int VEC_GRID = 2 * kSMs;  // fixed: the per-block partial sums have a fixed length

// slots of the device scalar block
// S_RZ0/1 and S_DONE0/1 are double-buffered by iteration parity: kernels of iteration k read slot k&1 and only
// block 0 of the last kernel of the iteration writes slot (k+1)&1, so no block can observe a value written
// by a sibling block of the same launch.
enum { S_RZ0 = 0, S_RZ1 = 1, S_BNORM = 2, S_TOL2 = 3, S_DONE0 = 4, S_DONE1 = 5, S_ITERS = 6, S_MAXIT = 7, S_RZ_LAST = 8, S_COUNT = 10 };   // slot 9: S_ERR (row-sharded step)

__device__ __forceinline__ double block_sum(double v, double *sm /*[VEC_BLOCK/32]*/) {
    v = warp_sum(v);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < VEC_BLOCK / 32 ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;  // valid in warp 0
}

// Sum of the P per-block partials in a fixed order, identical in every block (deterministic).
__device__ __forceinline__ double sum_partials(const double *__restrict__ part, int n_part, double *sm) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n_part; i += VEC_BLOCK) v += part[i];
    v = block_sum(v, sm);
    if (threadIdx.x == 0) sm[0] = v;
    __syncthreads();
    v = sm[0];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(VEC_BLOCK)
dot_partials_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ part,
                    const double *__restrict__ stop) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (stop && *stop != 0.0) return;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) v += a[i] * b[i];
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// minv = free ? 1/diag : 0   and partial sums of g^T minv g
__global__ void __launch_bounds__(VEC_BLOCK)
jacobi_kernel(int64_t n, const double *__restrict__ diag, const unsigned char *__restrict__ fixed, const double *__restrict__ g,
              double *__restrict__ minv, double *__restrict__ part) {
    __shared__ double sm[VEC_BLOCK / 32];
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        const double d = diag[i];
        const double m = (!fixed[i] && d > 0.0) ? 1.0 / d : 0.0;
        minv[i] = m;
        v += g[i] * m * g[i];
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// r = g - q (or r = g when q == nullptr); z = minv r; p = z; partials of r.z; x zeroed at fixed columns
__global__ void __launch_bounds__(VEC_BLOCK)
cg_init_kernel(int64_t n, const double *__restrict__ g, const double *__restrict__ q, const double *__restrict__ minv,
               const unsigned char *__restrict__ fixed, double *__restrict__ x, double *__restrict__ r, double *__restrict__ p,
               double *__restrict__ part) {
    __shared__ double sm[VEC_BLOCK / 32];
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        const double ri = q ? g[i] - q[i] : g[i];
        const double zi = minv[i] * ri;
        r[i] = ri;
        p[i] = zi;
        if (fixed[i]) x[i] = 0.0;
        v += ri * zi;
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// Folds the partial sums of a setup reduction into the scalar block (single block).
__global__ void __launch_bounds__(VEC_BLOCK)
cg_scalar_kernel(const double *__restrict__ part, int n_part, double *__restrict__ scal, int slot, double rtol, int max_iter) {
    __shared__ double sm[VEC_BLOCK / 32];
    const double v = sum_partials(part, n_part, sm);
    if (threadIdx.x == 0) {
        scal[slot] = v;
        if (slot == S_BNORM) scal[S_TOL2] = 0.0;
        if (slot == S_RZ0) {  // start of a solve
            scal[S_TOL2] = rtol * rtol * scal[S_BNORM];
            scal[S_ITERS] = 0.0;
            scal[S_MAXIT] = (double)max_iter;
            scal[S_RZ_LAST] = v;
            scal[S_DONE0] = scal[S_DONE1] = (v <= scal[S_TOL2] || max_iter <= 0 || !(v == v)) ? 1.0 : 0.0;
        }
    }
}

// alpha = rz / p.q;  x += alpha p;  r -= alpha q;  partials of r . (minv r)
__global__ void __launch_bounds__(VEC_BLOCK)
cg_update_kernel(int64_t n, const double *__restrict__ q, const double *__restrict__ pq_part, int n_part, const double *__restrict__ minv,
                 const double *__restrict__ p, double *__restrict__ x, double *__restrict__ r, double *__restrict__ rz_part,
                 const double *__restrict__ scal, int parity) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (scal[S_DONE0 + parity] != 0.0) return;
    const double pq = sum_partials(pq_part, n_part, sm);
    const double rz = scal[S_RZ0 + parity];
    const double alpha = pq > 0.0 ? rz / pq : 0.0;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        const double pi = p[i];
        x[i] += alpha * pi;
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        v += ri * (minv[i] * ri);
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) rz_part[blockIdx.x] = v;
}

// beta = rz_new / rz_old;  p = minv r + beta p;  block 0 publishes rz_new, the iteration count and the stop flag
__global__ void __launch_bounds__(VEC_BLOCK)
cg_direction_kernel(int64_t n, const double *__restrict__ r, const double *__restrict__ minv, double *__restrict__ p,
                    const double *__restrict__ rz_part, int n_part, double *__restrict__ scal, int parity) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;   // stay stopped
        return;
    }
    const double rz_new = sum_partials(rz_part, n_part, sm);
    const double rz_old = scal[S_RZ0 + parity];
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK)
        p[i] = minv[i] * r[i] + beta * p[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[S_RZ0 + (parity ^ 1)] = rz_new;   // read by the NEXT iteration; this one reads slot `parity`
        scal[S_RZ_LAST] = rz_new;
        const double it = scal[S_ITERS] + 1.0;
        scal[S_ITERS] = it;
        scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
    }
}

// update + direction in ONE cooperative launch (grid-wide barrier in between) for the block form of the solver;
// also clears q and the p.q slots for the next operator application, which accumulates into them with atomics.
__global__ void __launch_bounds__(VEC_BLOCK)
cg_step_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
               double *__restrict__ x, double *__restrict__ r, double *__restrict__ rz_part, double *__restrict__ scal, int parity) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (scal[S_DONE0 + parity] != 0.0) {          // uniform over the grid: nobody reaches the barrier
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    const double pq = sum_partials(pq_slots, PQ_SLOTS, sm);
    const double rz_old = scal[S_RZ0 + parity];
    const double alpha = pq > 0.0 ? rz_old / pq : 0.0;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        q[i] = 0.0;
        v += ri * (minv[i] * ri);
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) rz_part[blockIdx.x] = v;
    cgx::this_grid().sync();
    const double rz_new = sum_partials(rz_part, (int)gridDim.x, sm);
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK)
        p[i] = minv[i] * r[i] + beta * p[i];
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;     // every block read them before the barrier
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double it = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = it;
            scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

// Whole vector part of one CG iteration for the explicit-normal-matrix form, where q = N p comes from a gather
// product (no p.q by-product): p.q, grid barrier, alpha / x / r update and r.M^-1 r, grid barrier, beta / p update.
// part holds 2 * gridDim.x partial sums.
__global__ void __launch_bounds__(VEC_BLOCK)
cg_step_n_kernel(int64_t n, double *__restrict__ q, const double *__restrict__ minv, double *__restrict__ p,
                 double *__restrict__ x, double *__restrict__ r, double *__restrict__ part, double *__restrict__ scal, int parity) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (scal[S_DONE0 + parity] != 0.0) {          // uniform over the grid: nobody reaches the barriers
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    cgx::grid_group grid = cgx::this_grid();
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) v += p[i] * q[i];
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    grid.sync();
    const double pq = sum_partials(part, (int)gridDim.x, sm);
    const double rz_old = scal[S_RZ0 + parity];
    const double alpha = pq > 0.0 ? rz_old / pq : 0.0;
    v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        q[i] = 0.0;                       // the next product accumulates its cut rows into a clean q
        v += ri * (minv[i] * ri);
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[gridDim.x + blockIdx.x] = v;
    grid.sync();
    const double rz_new = sum_partials(part + gridDim.x, (int)gridDim.x, sm);
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK)
        p[i] = minv[i] * r[i] + beta * p[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[S_RZ0 + (parity ^ 1)] = rz_new;
        scal[S_RZ_LAST] = rz_new;
        const double it = scal[S_ITERS] + 1.0;
        scal[S_ITERS] = it;
        scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
    }
}

// ---- structured form of the explicit-normal solver: block-Jacobi preconditioner and / or an exactly eliminated
// diagonal block E = [e0, e1) (the per-(rgt, cycle) bias columns: every data row holds one of them, so N_EE is
// diagonal).  CG then runs on the Schur complement in full-length vectors: the E part of every direction is slaved to
// its grid part, p_E = -N_EE^-1 N_Eg p_g, which makes (N p)_E = 0 and (N p)_g = S p_g.

// Number of entries of rows [e0, e1) of N that sit inside [e0, e1) off the diagonal (must be zero).
__global__ void __launch_bounds__(256)
elim_check_kernel(int64_t e0, int64_t e1, const int *__restrict__ indptr, const int *__restrict__ indices, unsigned long long *__restrict__ bad) {
    const int64_t row = e0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= e1) return;
    unsigned long long c = 0;
    for (int k = indptr[row] + (int)lane_id(); k < indptr[row + 1]; k += 32) {
        const int j = indices[k];
        if (j >= e0 && j < e1 && j != row) ++c;
    }
    if (c) atomicAdd(bad, c);
}

// einv = 1 / N_ii over E, and no preconditioner action there (z_E = 0)
__global__ void __launch_bounds__(256)
elim_setup_kernel(int64_t e0, int64_t e1, const double *__restrict__ diag, double *__restrict__ einv, double *__restrict__ minv) {
    const int64_t i = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e1) return;
    const double d = diag[i];
    einv[i - e0] = d > 0.0 ? 1.0 / d : 0.0;
    minv[i] = 0.0;
}

// out_i = einv_i * (c_i - sum_{j not in E} N_ij v_j)  for i in E; one CTA per row (bias rows hold thousands of entries),
// every thread's loads issued before the first use.  out may alias v (only entries outside E are read).
constexpr int ELIM_BLOCK = 512, ELIM_IPT = 4;
__global__ void __launch_bounds__(ELIM_BLOCK)
elim_solve_kernel(int64_t e0, int64_t e1, const int *__restrict__ indptr, const int *__restrict__ indices, const double *__restrict__ data,
                  const double *v, const double *__restrict__ c, const double *__restrict__ einv, double *out,
                  const double *__restrict__ stop, double *__restrict__ raw) {
    __shared__ double sm[ELIM_BLOCK / 32];
    const double stop_v = stop ? *reinterpret_cast<const volatile double *>(stop) : 0.0;
    const int64_t row = e0 + blockIdx.x;
    const int kb = indptr[row], ke = indptr[row + 1];
    double acc = 0.0;
    for (int k0 = kb; k0 < ke; k0 += ELIM_BLOCK * ELIM_IPT) {
        int j[ELIM_IPT];
        double a[ELIM_IPT];
#pragma unroll
        for (int t = 0; t < ELIM_IPT; ++t) {
            const int k = k0 + t * ELIM_BLOCK + (int)threadIdx.x;
            j[t] = k < ke ? __ldg(indices + k) : -1;
            a[t] = k < ke ? __ldg(data + k) : 0.0;
        }
        double x[ELIM_IPT];
#pragma unroll
        for (int t = 0; t < ELIM_IPT; ++t) x[t] = (j[t] >= 0 && (j[t] < e0 || j[t] >= e1)) ? v[j[t]] : 0.0;
#pragma unroll
        for (int t = 0; t < ELIM_IPT; ++t) acc += a[t] * x[t];
    }
    if (stop_v != 0.0) return;
    acc = warp_sum(acc);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < ELIM_BLOCK / 32 ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            if (raw) raw[row - e0] = t;        // row-sharded N: the sums are completed over ranks first (elim_finish_kernel)
            else out[row] = einv[row - e0] * ((c ? c[row] : 0.0) - t);
        }
    }
}

__global__ void __launch_bounds__(256)
elim_finish_kernel(int64_t e0, int64_t e1, const double *__restrict__ t, const double *__restrict__ c, const double *__restrict__ einv,
                   double *__restrict__ out, const double *__restrict__ stop) {
    if (stop && *stop != 0.0) return;
    const int64_t i = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e1) out[i] = einv[i - e0] * ((c ? c[i] : 0.0) - t[i - e0]);
}

// r = g - q (q may be null), r_E = 0, x = 0 at fixed columns
__global__ void __launch_bounds__(VEC_BLOCK)
resid_kernel(int64_t n, const double *__restrict__ g, const double *__restrict__ q, const unsigned char *__restrict__ fixed,
             int64_t e0, int64_t e1, double *__restrict__ x, double *__restrict__ r) {
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        r[i] = (i >= e0 && i < e1) ? 0.0 : (q ? g[i] - q[i] : g[i]);
        if (x && fixed[i]) x[i] = 0.0;
    }
}

// p = z and the partial sums of r.z
__global__ void __launch_bounds__(VEC_BLOCK)
init_p_kernel(int64_t n, const double *__restrict__ r, const double *__restrict__ z, double *__restrict__ p, double *__restrict__ part) {
    __shared__ double sm[VEC_BLOCK / 32];
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        if (p) p[i] = z[i];
        v += r[i] * z[i];
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

// One CG iteration's vector work with the block preconditioner (cooperative launch, 2 CTAs per SM):
// alpha from the p.q slots the product left behind | x, r, z = M^-1 r (warp per block), r.z | grid barrier | beta, p.
// Clears q and the slots for the next product.
// The block inverses (bs x bs fp32, contiguous) are staged through shared memory with cp.async, double buffered per
// warp: while a warp multiplies block b its copy of block b + stride is already in flight, so the phase runs at memory
// bandwidth instead of one latency per block.
constexpr int STEP_PER_SM = 2;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(VEC_BLOCK, STEP_PER_SM)
cg_step_p_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
                 double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, double *__restrict__ part,
                 double *__restrict__ scal, int parity, int n_blocks, int bs, const int *__restrict__ bcols,
                 const int *__restrict__ col_block, const float *__restrict__ binv, int n_stage) {
    __shared__ double sm[VEC_BLOCK / 32];
    extern __shared__ __align__(16) float stage_buf[];      // [warp][n_stage][bs * bs]
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    const int64_t tid = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x, nth = (int64_t)gridDim.x * VEC_BLOCK;
    const unsigned lane = lane_id();
    const int bs2 = bs * bs, n16 = bs2 / 4;                 // bs is a multiple of 4: whole 16-byte pieces
    const int stride = (int)(nth >> 5);
    float *mine = stage_buf + (size_t)(threadIdx.x >> 5) * n_stage * bs2;
    auto fetch = [&](int blk, int st) {
        const float4 *src = reinterpret_cast<const float4 *>(binv + (size_t)blk * bs2);
        float4 *dst = reinterpret_cast<float4 *>(mine + (size_t)st * bs2);
        for (int i = (int)lane; i < n16; i += 32) cp_async16(dst + i, src + i);
    };
    int b = (int)(tid >> 5);
    if (b < n_blocks) fetch(b, 0);
    cp_async_commit();
    int c0 = -1, c1 = -1;
    if (b < n_blocks) {
        c0 = (int)lane < bs ? __ldg(bcols + (size_t)b * bs + lane) : -1;
        c1 = (int)lane + 32 < bs ? __ldg(bcols + (size_t)b * bs + 32 + lane) : -1;
    }
    const double pq = sum_partials(pq_slots, PQ_SLOTS, sm);
    const double rz_old = scal[S_RZ0 + parity];
    const double alpha = pq > 0.0 ? rz_old / pq : 0.0;
    double v = 0.0;
    int st = 0;
    for (; b < n_blocks; b += stride) {
        const int nb = b + stride;
        int d0 = -1, d1 = -1;
        if (n_stage > 1) {
            if (nb < n_blocks) {
                fetch(nb, st ^ 1);
                d0 = (int)lane < bs ? __ldg(bcols + (size_t)nb * bs + lane) : -1;
                d1 = (int)lane + 32 < bs ? __ldg(bcols + (size_t)nb * bs + 32 + lane) : -1;
            }
            cp_async_commit();
        }
        double r0 = 0.0, r1 = 0.0;
        if (c0 >= 0) { x[c0] += alpha * p[c0]; r0 = r[c0] - alpha * q[c0]; r[c0] = r0; q[c0] = 0.0; }
        if (c1 >= 0) { x[c1] += alpha * p[c1]; r1 = r[c1] - alpha * q[c1]; r[c1] = r1; q[c1] = 0.0; }
        if (n_stage > 1) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const float *S = mine + (size_t)st * bs2;
        double z0 = 0.0, z1 = 0.0;
        const bool h0 = (int)lane < bs, h1 = (int)lane + 32 < bs;
#pragma unroll 4
        for (int j = 0; j < bs; ++j) {                      // the stored inverse is symmetric: column l == row l
            const double rj = __shfl_sync(0xffffffffu, (j < 32) ? r0 : r1, j & 31);
            if (h0) z0 += (double)S[j * bs + lane] * rj;
            if (h1) z1 += (double)S[j * bs + 32 + lane] * rj;
        }
        if (c0 >= 0) z[c0] = z0;
        if (c1 >= 0) z[c1] = z1;
        v += r0 * z0 + r1 * z1;
        __syncwarp();                                        // all lanes are done with this stage before it is refilled
        if (n_stage > 1) {
            st ^= 1;
            c0 = d0;
            c1 = d1;
        } else if (nb < n_blocks) {
            fetch(nb, 0);
            cp_async_commit();
            c0 = (int)lane < bs ? __ldg(bcols + (size_t)nb * bs + lane) : -1;
            c1 = (int)lane + 32 < bs ? __ldg(bcols + (size_t)nb * bs + 32 + lane) : -1;
        }
    }
    cp_async_wait<0>();
    for (int64_t i = tid; i < n; i += nth) {
        if (col_block && col_block[i] >= 0) continue;
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        q[i] = 0.0;
        const double zi = minv[i] * ri;
        z[i] = zi;
        v += ri * zi;
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    cgx::this_grid().sync();
    const double rz_new = sum_partials(part, (int)gridDim.x, sm);
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = tid; i < n; i += nth) p[i] = z[i] + beta * p[i];
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;     // every block read them before the barrier
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double it = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = it;
            scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

// The same CG vector step with ONE THREAD PER BLOCK ROW instead of a warp per block (the default; LSS_STEP_KERNEL=0 or
// lss_set_option("step_kernel", 0) selects the kernel above).  Thread (b, l) forms z_l = sum_j inv_b[j][l] r'_j with
// r'_j = r_j - alpha q_j recomputed from the old residual: the 36 loads of a column of the (symmetric) inverse are
// independent and coalesced across l, the member / residual loads are warp-uniform broadcasts, and nothing is exchanged
// between threads -- no shuffles, no shared-memory staging, every thread of every resident warp has bs loads in flight.
// The updated residual goes to a second buffer (r_in / r_out swap with the parity) so that no thread reads an entry a
// neighbour has already overwritten; q is cleared after the grid barrier for the same reason.
constexpr int STEP2_PER_SM = 4;
__global__ void __launch_bounds__(VEC_BLOCK, STEP2_PER_SM)
cg_step_rows_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
                    double *__restrict__ x, const double *__restrict__ r_in, double *__restrict__ r_out, double *__restrict__ z,
                    double *__restrict__ part, double *__restrict__ scal, int parity, int n_blocks, int bs, const int *__restrict__ bcols,
                    const int *__restrict__ col_block, const float *__restrict__ binv) {
    __shared__ double sm[VEC_BLOCK / 32];
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    const int64_t tid = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x, nth = (int64_t)gridDim.x * VEC_BLOCK;
    static_assert(PQ_SLOTS == VEC_BLOCK, "one p.q slot per thread");
    const double slot_v = pq_slots[threadIdx.x];              // in flight while the first block row is accumulated
    const double rz_old = scal[S_RZ0 + parity];
    const int total = n_blocks * bs;
    // z_l = sum_j S_jl (r_j - alpha q_j) = zr - alpha zq: nothing below depends on alpha until the row is finished, so the
    // loads of the first row (the bulk of this kernel's traffic) are issued before the p.q reduction has been waited for
    auto accumulate = [&](int t, double &zr, double &zq) -> int {
        const int b = t / bs, l = t - b * bs;
        const int *bc = bcols + (size_t)b * bs;
        const int c = __ldg(bc + l);
        zr = 0.0;
        zq = 0.0;
        if (c < 0) return c;                                  // padding row of a short block
        const float *S = binv + (size_t)b * bs * bs + l;      // column l == row l of the symmetric inverse
#pragma unroll 6
        for (int j = 0; j < bs; ++j) {
            const int cj = __ldg(bc + j);
            const double sj = (double)__ldg(S + j * bs);
            if (cj >= 0) {
                zr += sj * r_in[cj];
                zq += sj * q[cj];
            }
        }
        return c;
    };
    int t = (int)tid;
    double zr0 = 0.0, zq0 = 0.0;
    const int c_first = t < total ? accumulate(t, zr0, zq0) : -1;
    double pq = block_sum(slot_v, sm);
    if (threadIdx.x == 0) sm[0] = pq;
    __syncthreads();
    pq = sm[0];
    __syncthreads();
    const double alpha = pq > 0.0 ? rz_old / pq : 0.0;
    double v = 0.0;
    auto finish = [&](int c, double zr, double zq) {
        const double zl = zr - alpha * zq;
        const double rl = r_in[c] - alpha * q[c];
        x[c] += alpha * p[c];
        r_out[c] = rl;
        z[c] = zl;
        v += rl * zl;
    };
    if (c_first >= 0) finish(c_first, zr0, zq0);
    for (t += (int)nth; t < total; t += (int)nth) {
        double zr, zq;
        const int c = accumulate(t, zr, zq);
        if (c >= 0) finish(c, zr, zq);
    }
    for (int64_t i = tid; i < n; i += nth) {
        if (col_block && col_block[i] >= 0) continue;
        x[i] += alpha * p[i];
        const double ri = r_in[i] - alpha * q[i];
        r_out[i] = ri;
        const double zi = minv[i] * ri;
        z[i] = zi;
        v += ri * zi;
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    cgx::this_grid().sync();
    const double rz_new = sum_partials(part, (int)gridDim.x, sm);
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        p[i] = z[i] + beta * p[i];
        q[i] = 0.0;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;     // every block read them before the barrier
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double it = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = it;
            scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

// Third form of the same step (the default when the block stride is one of the instantiated sizes; step_kernel = 2):
// a CTA takes 256 / BS whole blocks at a time.  Every thread first issues the BS loads of its column of the inverse
// (addresses known from the thread index alone, coalesced, all in flight at once), then gathers r, q of its own
// member into shared memory; after one barrier the products run from registers and shared-memory broadcasts.  The
// thread-per-row kernel above needs the member list before it can gather (two dependent levels, six times per row).
constexpr int STEP3_PER_SM = 3;
template <int BS>
__global__ void __launch_bounds__(VEC_BLOCK, STEP3_PER_SM)
cg_step_tile_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
                    double *__restrict__ x, const double *__restrict__ r_in, double *__restrict__ r_out, double *__restrict__ z,
                    double *__restrict__ part, double *__restrict__ scal, int parity, int n_blocks, const int *__restrict__ bcols,
                    const int *__restrict__ col_block, const float *__restrict__ binv) {
    constexpr int NBLK = VEC_BLOCK / BS, ROWS = NBLK * BS;
    __shared__ double sm[VEC_BLOCK / 32];
    __shared__ double rs[ROWS], qs[ROWS];
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    const int64_t tid = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x, nth = (int64_t)gridDim.x * VEC_BLOCK;
    static_assert(PQ_SLOTS == VEC_BLOCK, "one p.q slot per thread");
    const double slot_v = pq_slots[threadIdx.x];
    const double rz_old = scal[S_RZ0 + parity];
    const int n_groups = (n_blocks + NBLK - 1) / NBLK;
    const int bl = threadIdx.x / BS, l = threadIdx.x - bl * BS;       // block inside the group, member inside the block
    double alpha = 0.0, v = 0.0;
    bool have_alpha = false;
    auto get_alpha = [&]() {                                            // all threads of the CTA call this together
        double pq = block_sum(slot_v, sm);
        if (threadIdx.x == 0) sm[0] = pq;
        __syncthreads();
        pq = sm[0];
        __syncthreads();
        alpha = pq > 0.0 ? rz_old / pq : 0.0;
        have_alpha = true;
    };
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int b = g * NBLK + bl;
        const bool active = threadIdx.x < ROWS && b < n_blocks;
        float sv[BS];
        int c = -1;
        if (active) {
            const float *S = binv + (size_t)b * BS * BS + l;              // column l == row l of the symmetric inverse
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = __ldg(S + j * BS);
            c = __ldg(bcols + (size_t)b * BS + l);
        } else {
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = 0.f;
        }
        double r_own = 0.0, q_own = 0.0, p_own = 0.0, x_own = 0.0;
        if (c >= 0) {
            r_own = r_in[c];
            q_own = q[c];
            p_own = p[c];
            x_own = x[c];
        }
        if (threadIdx.x < ROWS) {
            rs[threadIdx.x] = r_own;
            qs[threadIdx.x] = q_own;
        }
        __syncthreads();
        double zr = 0.0, zq = 0.0;
        const double *rb = rs + bl * BS, *qb = qs + bl * BS;
        if (threadIdx.x < ROWS) {
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                const double sj = (double)sv[j];
                zr += sj * rb[j];
                zq += sj * qb[j];
            }
        }
        if (!have_alpha) get_alpha();                                       // contains barriers: rs / qs are free afterwards
        else __syncthreads();
        if (c >= 0) {
            const double zl = zr - alpha * zq;
            const double rl = r_own - alpha * q_own;
            x[c] = x_own + alpha * p_own;
            r_out[c] = rl;
            z[c] = zl;
            v += rl * zl;
        }
    }
    if (!have_alpha) get_alpha();
    for (int64_t i = tid; i < n; i += nth) {
        if (col_block && col_block[i] >= 0) continue;
        x[i] += alpha * p[i];
        const double ri = r_in[i] - alpha * q[i];
        r_out[i] = ri;
        const double zi = minv[i] * ri;
        z[i] = zi;
        v += ri * zi;
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    cgx::this_grid().sync();
    const double rz_new = sum_partials(part, (int)gridDim.x, sm);
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        p[i] = z[i] + beta * p[i];
        q[i] = 0.0;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;     // every block read them before the barrier
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double it = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = it;
            scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

static const void *step_tile_kernel_for(int bs) {
    switch (bs) {
        case 12: return (const void *)cg_step_tile_kernel<12>;
        case 24: return (const void *)cg_step_tile_kernel<24>;
        case 36: return (const void *)cg_step_tile_kernel<36>;
        default: return nullptr;
    }
}

// Two-level form of the tile step (step_kernel = 3, the default when a coarse space is attached, lss_solver_set_coarse):
// z = M_BJ^-1 r + P S_c^-1 P^T r.  Phases, separated by grid barriers:
//   B   as cg_step_tile_kernel: alpha, x, r, z_BJ (to z), partials of r . z_BJ
//   C1  r_c = P^T r as a gather over P^T stored by coarse row (deterministic; every coarse row is split over `chunks` warps)
//   C2  z_c = S_c^-1 r_c: one warp per coarse row of the dense inverse (a few MB, L2 resident), r_c . z_c partials
//   D   r.z = r.z_BJ + r_c.z_c, beta, p = z_BJ + P z_c + beta p; q and the p.q slots cleared
struct CoarseArgs {
    int n_c, chunks;
    const int *pt_ptr, *pt_col;
    const float *pt_w;
    const double *inv;
    const int4 *cp_idx;
    const float4 *cp_w;
    double *rc_part, *zc_part;
    int dbg;
};

template <int BS>
__global__ void __launch_bounds__(VEC_BLOCK, STEP3_PER_SM)
cg_step_ml_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
                  double *__restrict__ x, const double *r_in, double *r_out, double *z,
                  double *part, double *__restrict__ scal, int parity, int n_blocks, const int *__restrict__ bcols,
                  const int *__restrict__ col_block, const float *__restrict__ binv, CoarseArgs ca) {
    constexpr int NBLK = VEC_BLOCK / BS, ROWS = NBLK * BS;
    __shared__ double sm[VEC_BLOCK / 32];
    __shared__ double rs[ROWS], qs[ROWS];
    __shared__ double cs[CS_MAX];
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    cgx::grid_group grid = cgx::this_grid();
    const int64_t tid = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x, nth = (int64_t)gridDim.x * VEC_BLOCK;
    static_assert(PQ_SLOTS == VEC_BLOCK, "one p.q slot per thread");
    const double slot_v = pq_slots[threadIdx.x];
    const double rz_old = scal[S_RZ0 + parity];
    const int n_groups = (n_blocks + NBLK - 1) / NBLK;
    const int bl = threadIdx.x / BS, l = threadIdx.x - bl * BS;
    double alpha = 0.0, v = 0.0;
    bool have_alpha = false;
    auto get_alpha = [&]() {
        double pq = block_sum(slot_v, sm);
        if (threadIdx.x == 0) sm[0] = pq;
        __syncthreads();
        pq = sm[0];
        __syncthreads();
        alpha = pq > 0.0 ? rz_old / pq : 0.0;
        have_alpha = true;
    };
    // ---- phase B
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int b = g * NBLK + bl;
        const bool active = threadIdx.x < ROWS && b < n_blocks;
        float sv[BS];
        int c = -1;
        if (active) {
            const float *S = binv + (size_t)b * BS * BS + l;
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = __ldg(S + j * BS);
            c = __ldg(bcols + (size_t)b * BS + l);
        } else {
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = 0.f;
        }
        double r_own = 0.0, q_own = 0.0, p_own = 0.0, x_own = 0.0;
        if (c >= 0) {
            r_own = r_in[c];
            q_own = q[c];
            p_own = p[c];
            x_own = x[c];
        }
        if (threadIdx.x < ROWS) {
            rs[threadIdx.x] = r_own;
            qs[threadIdx.x] = q_own;
        }
        __syncthreads();
        double zr = 0.0, zq = 0.0;
        const double *rb = rs + bl * BS, *qb = qs + bl * BS;
        if (threadIdx.x < ROWS) {
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                const double sj = (double)sv[j];
                zr += sj * rb[j];
                zq += sj * qb[j];
            }
        }
        if (!have_alpha) get_alpha();
        else __syncthreads();
        if (c >= 0) {
            const double zl = zr - alpha * zq;
            const double rl = r_own - alpha * q_own;
            x[c] = x_own + alpha * p_own;
            r_out[c] = rl;
            z[c] = zl;
            v += rl * zl;
        }
    }
    if (!have_alpha) get_alpha();
    for (int64_t i = tid; i < n; i += nth) {
        if (col_block && col_block[i] >= 0) continue;
        x[i] += alpha * p[i];
        const double ri = r_in[i] - alpha * q[i];
        r_out[i] = ri;
        const double zi = minv[i] * ri;
        z[i] = zi;
        v += ri * zi;
    }
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    grid.sync();
    // ---- phase C1: partial restrictions
    const int n_c = ca.n_c, chunks = ca.chunks;
    {
        const int warp = (int)(tid >> 5), n_warps = (int)(nth >> 5);
        for (int w = warp; w < n_c * chunks && !(ca.dbg & 1); w += n_warps) {
            const int m = w % n_c, ch = w / n_c;
            const int kb = ca.pt_ptr[m], len = ca.pt_ptr[m + 1] - kb;
            const int per = (len + chunks - 1) / chunks;
            const int k0 = kb + ch * per, k1 = min(kb + len, k0 + per);
            double acc = 0.0;
            int k = k0 + (int)lane_id();
            for (; k + 96 < k1; k += 128) {            // four independent gathers in flight per lane
                const int c0 = __ldg(ca.pt_col + k), c1 = __ldg(ca.pt_col + k + 32), c2 = __ldg(ca.pt_col + k + 64), c3 = __ldg(ca.pt_col + k + 96);
                const float w0 = __ldg(ca.pt_w + k), w1 = __ldg(ca.pt_w + k + 32), w2 = __ldg(ca.pt_w + k + 64), w3 = __ldg(ca.pt_w + k + 96);
                const double r0 = r_out[c0], r1 = r_out[c1], r2 = r_out[c2], r3 = r_out[c3];
                acc += ((double)w0 * r0 + (double)w1 * r1) + ((double)w2 * r2 + (double)w3 * r3);
            }
            for (; k < k1; k += 32) acc += (double)__ldg(ca.pt_w + k) * r_out[__ldg(ca.pt_col + k)];
            acc = warp_sum(acc);
            if (lane_id() == 0) ca.rc_part[(size_t)ch * n_c + m] = acc;
        }
    }
    if (!(ca.dbg & 8)) grid.sync();
    // ---- phase C2: z_c = S_c^-1 r_c, every row of the inverse split over `chunks` warps
    for (int m = threadIdx.x; m < n_c; m += VEC_BLOCK) {
        double a = 0.0;
        for (int ch = 0; ch < chunks; ++ch) a += ca.rc_part[(size_t)ch * n_c + m];
        cs[m] = a;
    }
    __syncthreads();
    {
        const int warp = (int)(tid >> 5), n_warps = (int)(nth >> 5);
        const int per = ((n_c + chunks - 1) / chunks + 31) & ~31;
        for (int w = warp; w < n_c * chunks && !(ca.dbg & 2); w += n_warps) {
            const int m = w % n_c, ch = w / n_c;
            const double *row = ca.inv + (size_t)m * n_c;
            const int j0 = ch * per, j1 = min(n_c, j0 + per);
            double acc = 0.0;
            int j = j0 + (int)lane_id();
            for (; j + 96 < j1; j += 128) {            // four independent loads in flight per lane
                const double a0 = __ldg(row + j), a1 = __ldg(row + j + 32), a2 = __ldg(row + j + 64), a3 = __ldg(row + j + 96);
                acc += (a0 * cs[j] + a1 * cs[j + 32]) + (a2 * cs[j + 64] + a3 * cs[j + 96]);
            }
            for (; j < j1; j += 32) acc += __ldg(row + j) * cs[j];
            acc = warp_sum(acc);
            if (lane_id() == 0) ca.zc_part[(size_t)ch * n_c + m] = acc;
        }
    }
    if (!(ca.dbg & 8)) grid.sync();
    // ---- phase D
    double rz_new = sum_partials(part, (int)gridDim.x, sm);
    double rczc = 0.0;
    double zreg[(CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK];
#pragma unroll
    for (int u = 0; u < (CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK; ++u) {
        const int m = threadIdx.x + u * VEC_BLOCK;
        double a = 0.0;
        if (m < n_c) {
            for (int ch = 0; ch < chunks; ++ch) a += ca.zc_part[(size_t)ch * n_c + m];
            rczc += cs[m] * a;
        }
        zreg[u] = a;
    }
    rczc = block_sum(rczc, sm);                 // identical in every CTA: same values, same order
    if (threadIdx.x == 0) sm[0] = rczc;
    __syncthreads();
    rz_new += sm[0];
    __syncthreads();
#pragma unroll
    for (int u = 0; u < (CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK; ++u) {
        const int m = threadIdx.x + u * VEC_BLOCK;
        if (m < n_c) cs[m] = zreg[u];
    }
    __syncthreads();
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        double zi = z[i];
        const int4 pi = __ldg(ca.cp_idx + i);
        if (!(ca.dbg & 4) && (pi.x >= 0 || pi.y >= 0 || pi.z >= 0 || pi.w >= 0)) {
            const float4 w = __ldg(ca.cp_w + i);
            if (pi.x >= 0) zi += (double)w.x * cs[pi.x];
            if (pi.y >= 0) zi += (double)w.y * cs[pi.y];
            if (pi.z >= 0) zi += (double)w.z * cs[pi.z];
            if (pi.w >= 0) zi += (double)w.w * cs[pi.w];
        }
        p[i] = zi + beta * p[i];
        q[i] = 0.0;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double it = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = it;
            scal[S_DONE0 + (parity ^ 1)] = (rz_new <= scal[S_TOL2] || it >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

static const void *step_ml_kernel_for(int bs) {
    switch (bs) {
        case 12: return (const void *)cg_step_ml_kernel<12>;
        case 24: return (const void *)cg_step_ml_kernel<24>;
        case 36: return (const void *)cg_step_ml_kernel<36>;
        default: return nullptr;
    }
}


// ---- row-sharded N over the GPUs of one node: the CG step with peer-to-peer exchanges -------------------------------
// Every rank owns a contiguous range of preconditioner blocks: the rows of N at their columns, the block inverses, the
// matching entries of x and r.  p is replicated.  One iteration = the local product q_own = N_own p (same kernel as on
// one GPU) + this kernel, which talks to the peers through the mailbox (comm.cu) with plain stores over NVLink:
//   A   p.q: every rank stores its partial into every box, flags it, and sums the P partials in rank order
//   B   own blocks: x, r, z_BJ; z_BJ goes straight into every peer's packed z buffer (coalesced 8-byte stores)
//   C1  own part of the restriction r_c,k = P_k^T r and of the bias coupling t_k = N_E,own z_BJ
//   X   packet [r.z_BJ partial | r_c,k | t_k] -> every box, flag; wait for the P packets; sum in rank order
//   C2  z_c = S_c^-1 r_c (replicated, spread over the CTAs)
//   D   beta, p = z_BJ + P z_c + beta p for ALL columns from the packed z of all ranks;
//       u = T^T p carried along: u <- sum_k t_k + B^T z_c + beta u, p_E = -N_EE^-1 u
// All ranks form every scalar from the same numbers in the same order: alpha, beta and the stop flag are bit-identical
// everywhere, so no rank can leave the loop alone.  A wait that exceeds ~4 s sets S_ERR and every later wait is skipped.
constexpr int S_ERR = 9;
constexpr int SHARD_NS = 8;                // slices the exchange packet is sent in (one CTA and one flag per peer and slice)
constexpr int SHARD_ECH_MAX = 32;           // pieces per row of N_E in phase C1
struct ShardArgs {
    int rank, n_ranks, blk_lo, blk_hi;
    unsigned long long seq;                 // sequence base of this solve; iteration k uses seq + 2k + 1 (A) and seq + 2k + 2 (X)
    unsigned long long *box[LSS_MAX_RANKS]; // the mailboxes (box[rank] is local)
    long long off_lla, off_zf, off_llx, off_z;          // word offsets: p.q cells / z-complete flags / packet cells / packed z
    int LB;                                 // words per packet: 1 + n_c + n_e (padded)
    long long ZP;                           // words of one packed z buffer: n_blocks * BS
    int n_e, ech;                           // eliminated columns; pieces every row of N_E is cut into (phase C1)
    long long e0;
    const int *n_indptr, *n_indices;        // rows of E of the local N: entries at own columns
    const double *n_data;
    const double *einv, *Bt;
    double *u_e, *tk;
    // phase C1 as flat entry lists (row key per entry): restriction entries [0, n_fr), rows of N_E [kE0, kE0 + n_fe)
    const int *fr_key;                      // coarse row of every entry of P^T (own columns)
    const int *fe_key;                      // n_c + bias row of every entry of the rows of N_E
    int n_fr, n_fe;
    long long kE0;
    double *tgt;                            // [n_c + n_e] sums of phase C1 (cleared in phase D for the next iteration)
    const unsigned char *blk_peers;         // per own block: bit t set when rank t references one of its columns (null: all ranks)
    const int *dlist;                       // phase D: packed-z slots of the columns this rank needs p at (null: all)
    int n_dlist;
    unsigned long long *trace;              // LSS_SHARD_TRACE: %globaltimer of CTA 0 at the phase boundaries, [iteration % 1024][16]
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SHARD_STAMP(k)                                                                                \
    do {                                                                                           \
        if (sa.trace && blockIdx.x == 0 && threadIdx.x == 0) sa.trace[(it & 1023ull) * 16 + (k)] = global_ns(); \
    } while (0)
__device__ __forceinline__ double ld_box(const double *p) { return __ldcg(p); }     // written by a peer: L2 is the coherence point

// "LL" cells (the protocol NCCL uses for small messages): a double travels as two 8-byte words, each carrying 32 data bits and
// the 32-bit sequence number of the exchange, written with ONE 16-byte store.  The receiver polls the cell until both words
// show the expected sequence: no separate flag, no fence between data and flag, one NVLink crossing per exchange.
__device__ __forceinline__ void ll_store(unsigned long long *cell, double v, unsigned int seq32) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), hi = (unsigned long long)seq32 << 32;
    const unsigned long long w0 = hi | (b & 0xffffffffull), w1 = hi | (b >> 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(cell), "l"(w0), "l"(w1) : "memory");
}
// returns the value once it has arrived; gives up after ~4 s (a peer died): raises S_ERR and returns 0
__device__ __forceinline__ double ll_wait(const unsigned long long *cell, unsigned int seq32, double *scal) {
    unsigned long long w0, w1;
    unsigned int spins = 0;
    long long t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(cell) : "memory");
        if ((unsigned int)(w0 >> 32) == seq32 && (unsigned int)(w1 >> 32) == seq32) break;
        if ((++spins & 255u) == 0) {
            if (*reinterpret_cast<const volatile double *>(scal + 9) != 0.0) return 0.0;       // S_ERR
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000ll) { scal[9] = 1.0; __threadfence(); return 0.0; }
        }
    }
    return __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
}

// sum over the ranks (in rank order) of the cells cell0 + t * stride: all loads of a round are issued together, only the
// cells that have not arrived are polled again
__device__ __forceinline__ double ll_wait_sum(const unsigned long long *cell0, size_t stride, int P, unsigned int seq32, double *scal) {
    double v[LSS_MAX_RANKS];
    unsigned int pending = (1u << P) - 1u, spins = 0;
    long long t0 = 0;
#pragma unroll
    for (int t = 0; t < LSS_MAX_RANKS; ++t) v[t] = 0.0;
    while (pending) {
        unsigned long long w0[LSS_MAX_RANKS], w1[LSS_MAX_RANKS];
#pragma unroll
        for (int t = 0; t < LSS_MAX_RANKS; ++t)
            if (pending & (1u << t))
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[t]), "=l"(w1[t]) : "l"(cell0 + t * stride) : "memory");
#pragma unroll
        for (int t = 0; t < LSS_MAX_RANKS; ++t)
            if ((pending & (1u << t)) && (unsigned int)(w0[t] >> 32) == seq32 && (unsigned int)(w1[t] >> 32) == seq32) {
                v[t] = __longlong_as_double((long long)((w1[t] << 32) | (w0[t] & 0xffffffffull)));
                pending &= ~(1u << t);
            }
        if (pending && (++spins & 255u) == 0) {
            if (*reinterpret_cast<const volatile double *>(scal + 9) != 0.0) break;       // S_ERR
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000ll) { scal[9] = 1.0; __threadfence(); break; }
        }
    }
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < LSS_MAX_RANKS; ++t) a += v[t];
    return a;
}

// thread t < n waits until flags[t] >= want; gives up after ~4 s (another rank died) and raises S_ERR
__device__ __forceinline__ void wait_flags(const unsigned long long *flags, int n, unsigned long long want, double *scal) {
    if ((int)threadIdx.x < n) {
        const volatile double *err = scal + S_ERR;
        if (*err == 0.0) {
            const long long t0 = clock64();
            unsigned int spins = 0;
            while (ld_acquire_sys(flags + threadIdx.x) < want) {
                if ((++spins & 1023u) == 0) {
                    if (*err != 0.0) break;
                    if (clock64() - t0 > 8000000000ll) { scal[S_ERR] = 1.0; __threadfence(); break; }
                }
            }
        }
    }
    __syncthreads();
}

template <int BS>
__global__ void __launch_bounds__(VEC_BLOCK, STEP3_PER_SM)
cg_step_shard_kernel(int64_t n, double *__restrict__ q, double *__restrict__ pq_slots, const double *__restrict__ minv, double *__restrict__ p,
                     double *__restrict__ x, const double *r_in, double *r_out, double *z,
                     double *part, double *__restrict__ scal, int parity, int n_blocks, const int *__restrict__ bcols,
                     const float *__restrict__ binv, CoarseArgs ca, ShardArgs sa) {
    constexpr int NBLK = VEC_BLOCK / BS, ROWS = NBLK * BS;
    __shared__ double sm[VEC_BLOCK / 32];
    __shared__ double rs[ROWS], qs[ROWS];
    __shared__ double cs[CS_MAX];
    __shared__ double llv[LSS_MAX_RANKS];
    if (scal[S_DONE0 + parity] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) scal[S_DONE0 + (parity ^ 1)] = 1.0;
        return;
    }
    cgx::grid_group grid = cgx::this_grid();
    const int64_t tid = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x, nth = (int64_t)gridDim.x * VEC_BLOCK;
    const int P = sa.n_ranks, me = sa.rank;
    const unsigned long long it = (unsigned long long)scal[S_ITERS];
    const int par = (int)(it & 1ull);
    const unsigned long long seq_a = sa.seq + 2ull * it + 1ull, seq_x = seq_a + 1ull;
    const unsigned int seq32_a = (unsigned int)seq_a, seq32_x = (unsigned int)seq_x;
    unsigned long long *mine = sa.box[me];
    static_assert(PQ_SLOTS == VEC_BLOCK, "one p.q slot per thread");
    SHARD_STAMP(0);
    const double slot_v = pq_slots[threadIdx.x];
    const double rz_old = scal[S_RZ0 + parity];
    const int n_own = sa.blk_hi - sa.blk_lo;
    const int n_groups = (n_own + NBLK - 1) / NBLK;
    const int bl = threadIdx.x / BS, l = threadIdx.x - bl * BS;
    double alpha = 0.0, v = 0.0;
    bool have_alpha = false;
    // ---- phase A (folded behind the first block's loads): p.q over all ranks
    auto get_alpha = [&]() {
        double pq = block_sum(slot_v, sm);
        if (threadIdx.x == 0) sm[0] = pq;
        __syncthreads();
        pq = sm[0];
        __syncthreads();
        if (blockIdx.x == 0 && (int)threadIdx.x < P)            // one thread per peer: value and sequence number in one store
            ll_store(sa.box[threadIdx.x] + sa.off_lla + 2 * (par * P + me), pq, seq32_a);
        if ((int)threadIdx.x < P) llv[threadIdx.x] = ll_wait(mine + sa.off_lla + 2 * (par * P + threadIdx.x), seq32_a, scal);
        __syncthreads();
        double tot = 0.0;
        for (int t = 0; t < P; ++t) tot += llv[t];        // rank order: the same sum on every rank
        alpha = tot > 0.0 ? rz_old / tot : 0.0;
        have_alpha = true;
        SHARD_STAMP(1);
    };
    // ---- phase B: own blocks
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int b = sa.blk_lo + g * NBLK + bl;
        const bool active = threadIdx.x < ROWS && b < sa.blk_hi;
        float sv[BS];
        int c = -1;
        if (active) {
            const float *S = binv + (size_t)b * BS * BS + l;
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = __ldg(S + j * BS);
            c = __ldg(bcols + (size_t)b * BS + l);
        } else {
#pragma unroll
            for (int j = 0; j < BS; ++j) sv[j] = 0.f;
        }
        double r_own = 0.0, q_own = 0.0, p_own = 0.0, x_own = 0.0;
        if (c >= 0) {
            r_own = r_in[c];
            q_own = q[c];
            p_own = p[c];
            x_own = x[c];
        }
        if (threadIdx.x < ROWS) {
            rs[threadIdx.x] = r_own;
            qs[threadIdx.x] = q_own;
        }
        __syncthreads();
        double zr = 0.0, zq = 0.0;
        const double *rb = rs + bl * BS, *qb = qs + bl * BS;
        if (threadIdx.x < ROWS) {
#pragma unroll
            for (int j = 0; j < BS; ++j) {
                const double sj = (double)sv[j];
                zr += sj * rb[j];
                zq += sj * qb[j];
            }
        }
        if (!have_alpha) get_alpha();
        else __syncthreads();
        if (active) {
            double zl = 0.0;
            if (c >= 0) {
                zl = zr - alpha * zq;
                const double rl = r_own - alpha * q_own;
                x[c] = x_own + alpha * p_own;
                r_out[c] = rl;
                z[c] = zl;
                v += rl * zl;
            }
            // z of the own blocks goes to the ranks that reference them, as LL cells (16-byte stores, coalesced over the CTA's
            // blocks): no fence and no flag are needed before the peers may read them
            const long long zi = sa.off_z + 2 * ((long long)par * sa.ZP + (long long)b * BS + l);
            const unsigned int peers = sa.blk_peers ? ((unsigned int)__ldg(sa.blk_peers + (b - sa.blk_lo)) | (1u << me)) : 0xffffffffu;
            for (int t = 0; t < P; ++t)
                if (peers & (1u << t)) ll_store(sa.box[t] + zi, zl, seq32_x);
        }
    }
    if (!have_alpha) get_alpha();
    v = block_sum(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    SHARD_STAMP(2);
    SHARD_STAMP(3);
    grid.sync();
    SHARD_STAMP(4);
    // ---- phase C1: own part of the restriction r_c,k = P_k^T r and of t_k = N_E,own z.  Both are sums over flat entry lists
    // sorted by row: every warp takes 128 consecutive entries at a time (4 independent loads per lane, then 4 gathers), folds
    // equal rows with a segmented warp scan and adds one value per (warp, row) to tgt with a RED.
    const int n_c = ca.n_c, chunks = ca.chunks;
    const unsigned long long tc0 = sa.trace ? global_ns() : 0ull;
    {
        const int warp = (int)(tid >> 5), n_warps = (int)(nth >> 5), lane = (int)lane_id();
        auto seg_add = [&](int key, double val) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v2 = __shfl_up_sync(0xffffffffu, val, o);
                const int k2 = __shfl_up_sync(0xffffffffu, key, o);
                if (lane >= o && k2 == key) val += v2;
            }
            const int kn = __shfl_down_sync(0xffffffffu, key, 1);
            if (key >= 0 && (lane == 31 || kn != key)) atomicAdd(sa.tgt + key, val);
        };
        for (int base = warp * 128; base < sa.n_fr; base += n_warps * 128) {
            int key[4], col[4];
            float wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = base + 32 * j + lane;
                const bool in = k < sa.n_fr;
                key[j] = in ? __ldg(sa.fr_key + k) : -1;
                col[j] = in ? __ldg(ca.pt_col + k) : 0;
                wv[j] = in ? __ldg(ca.pt_w + k) : 0.f;
            }
            double val[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = key[j] >= 0 ? (double)wv[j] * r_out[col[j]] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) seg_add(key[j], val[j]);
        }
        const long long e_lo = sa.e0, e_hi = sa.e0 + sa.n_e;
        for (int base = warp * 128; base < sa.n_fe; base += n_warps * 128) {
            int key[4], col[4];
            double av[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = base + 32 * j + lane;
                const bool in = k < sa.n_fe;
                key[j] = in ? __ldg(sa.fe_key + k) : -1;
                col[j] = in ? __ldg(sa.n_indices + sa.kE0 + k) : 0;
                av[j] = in ? __ldg(sa.n_data + sa.kE0 + k) : 0.0;
            }
            double val[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = (key[j] >= 0 && (col[j] < e_lo || col[j] >= e_hi)) ? av[j] * z[col[j]] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) seg_add(key[j], val[j]);
        }
    }
    const unsigned long long tc1 = sa.trace ? global_ns() : 0ull;
    if (sa.trace && threadIdx.x == 0) {         // slowest CTA of the phase: work, then fence
        atomicMax(sa.trace + (it & 1023ull) * 16 + 11, tc1 - tc0);
        atomicMax(sa.trace + (it & 1023ull) * 16 + 12, global_ns() - tc1);
    }
    grid.sync();
    SHARD_STAMP(5);
    // ---- exchange X: packet [r.z_BJ | r_c,k | t_k] to every rank as LL cells, in SHARD_NS slices (CTA b serves peer b % P,
    // slice b / P)
    if ((int)blockIdx.x < P * SHARD_NS) {
        const int peer = (int)blockIdx.x % P, sl = (int)blockIdx.x / P;
        unsigned long long *dst = sa.box[peer] + sa.off_llx + 2 * (size_t)(par * P + me) * sa.LB;
        const int used = 1 + n_c + sa.n_e, per_s = (used + SHARD_NS - 1) / SHARD_NS;
        const int i0 = sl * per_s, i1 = min(used, i0 + per_s);
        if (sl == 0) {
            const double rzk = sum_partials(part, (int)gridDim.x, sm);
            if (threadIdx.x == 0) ll_store(dst, rzk, seq32_x);
        }
        for (int i = max(i0, 1) + (int)threadIdx.x; i < i1; i += VEC_BLOCK)
            ll_store(dst + 2 * (size_t)i, __ldcg(sa.tgt + (i - 1)), seq32_x);     // tgt: written with REDs (L2) before the grid barrier
    }
    SHARD_STAMP(6);
    const unsigned long long *db = mine + sa.off_llx + 2 * (size_t)par * P * sa.LB;
    if ((int)threadIdx.x < P) llv[threadIdx.x] = ll_wait(db + 2 * (size_t)threadIdx.x * sa.LB, seq32_x, scal);
    for (int m = threadIdx.x; m < n_c; m += VEC_BLOCK) {
        cs[m] = ll_wait_sum(db + 2 * (size_t)(1 + m), 2 * (size_t)sa.LB, P, seq32_x, scal);
    }
    __syncthreads();
    SHARD_STAMP(7);
    double rz_new = 0.0;
    for (int t = 0; t < P; ++t) rz_new += llv[t];
    // ---- phase C2: z_c = S_c^-1 r_c
    {
        const int warp = (int)(tid >> 5), n_warps = (int)(nth >> 5);
        const int per = ((n_c + chunks - 1) / chunks + 31) & ~31;
        for (int w = warp; w < n_c * chunks; w += n_warps) {
            const int m = w % n_c, ch = w / n_c;
            const double *row = ca.inv + (size_t)m * n_c;
            const int j0 = ch * per, j1 = min(n_c, j0 + per);
            double acc = 0.0;
            int j = j0 + (int)lane_id();
            for (; j + 96 < j1; j += 128) {
                const double a0 = __ldg(row + j), a1 = __ldg(row + j + 32), a2 = __ldg(row + j + 64), a3 = __ldg(row + j + 96);
                acc += (a0 * cs[j] + a1 * cs[j + 32]) + (a2 * cs[j + 64] + a3 * cs[j + 96]);
            }
            for (; j < j1; j += 32) acc += __ldg(row + j) * cs[j];
            acc = warp_sum(acc);
            if (lane_id() == 0) ca.zc_part[(size_t)ch * n_c + m] = acc;
        }
    }
    SHARD_STAMP(8);
    grid.sync();
    SHARD_STAMP(9);
    // ---- phase D
    double rczc = 0.0;
    double zreg[(CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK];
#pragma unroll
    for (int u = 0; u < (CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK; ++u) {
        const int m = threadIdx.x + u * VEC_BLOCK;
        double a = 0.0;
        if (m < n_c) {
            for (int ch = 0; ch < chunks; ++ch) a += ca.zc_part[(size_t)ch * n_c + m];
            rczc += cs[m] * a;
        }
        zreg[u] = a;
    }
    rczc = block_sum(rczc, sm);
    if (threadIdx.x == 0) sm[0] = rczc;
    __syncthreads();
    rz_new += sm[0];
    __syncthreads();
#pragma unroll
    for (int u = 0; u < (CS_MAX + VEC_BLOCK - 1) / VEC_BLOCK; ++u) {
        const int m = threadIdx.x + u * VEC_BLOCK;
        if (m < n_c) cs[m] = zreg[u];
    }
    __syncthreads();
    const double beta = rz_old > 0.0 ? rz_new / rz_old : 0.0;
    // p from the packed z of the owners, at the columns this rank's rows of N reference (the rest of p is never read here)
    {
        const unsigned long long *zp = mine + sa.off_z + 2 * (size_t)par * sa.ZP;      // LL cells, sent in phase B of this iteration
        const int64_t total = sa.dlist ? (int64_t)sa.n_dlist : (int64_t)n_blocks * BS;
        for (int64_t i = tid; i < total; i += nth) {
            const int64_t k = sa.dlist ? (int64_t)__ldg(sa.dlist + i) : i;
            const int c = __ldg(bcols + k);
            if (c < 0) continue;
            double zi = ll_wait(zp + 2 * k, seq32_x, scal);
            if (n_c > 0) {
                const int4 pi = __ldg(ca.cp_idx + c);
                const float4 w = __ldg(ca.cp_w + c);
                if (pi.x >= 0) zi += (double)w.x * cs[pi.x];
                if (pi.y >= 0) zi += (double)w.y * cs[pi.y];
                if (pi.z >= 0) zi += (double)w.z * cs[pi.z];
                if (pi.w >= 0) zi += (double)w.w * cs[pi.w];
            }
            p[c] = zi + beta * p[c];
        }
    }
    for (int64_t i = tid; i < n; i += nth) q[i] = 0.0;
    for (int64_t i = tid; i < n_c + sa.n_e; i += nth) sa.tgt[i] = 0.0;      // for phase C1 of the next iteration
    // the eliminated block: u = T^T p by linearity, p_E = -N_EE^-1 u (one warp per bias column)
    {
        const int warp = (int)(tid >> 5), n_warps = (int)(nth >> 5);
        for (int e = warp; e < sa.n_e; e += n_warps) {
            double acc = 0.0;
            if (n_c > 0) {
                const double *row = sa.Bt + (size_t)e * n_c;
                for (int m = (int)lane_id(); m < n_c; m += 32) acc += __ldg(row + m) * cs[m];
                acc = warp_sum(acc);
            }
            if (lane_id() == 0) {
                const double t = ll_wait_sum(db + 2 * (size_t)(1 + n_c + e), 2 * (size_t)sa.LB, P, seq32_x, scal);
                const double un = t + acc + beta * sa.u_e[e];
                sa.u_e[e] = un;
                p[sa.e0 + e] = -sa.einv[e] * un;
            }
        }
    }
    SHARD_STAMP(10);
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < PQ_SLOTS; i += VEC_BLOCK) pq_slots[i] = 0.0;
        if (threadIdx.x == 0) {
            scal[S_RZ0 + (parity ^ 1)] = rz_new;
            scal[S_RZ_LAST] = rz_new;
            const double itn = scal[S_ITERS] + 1.0;
            scal[S_ITERS] = itn;
            const bool err = *reinterpret_cast<volatile double *>(scal + S_ERR) != 0.0;
            scal[S_DONE0 + (parity ^ 1)] = (err || rz_new <= scal[S_TOL2] || itn >= scal[S_MAXIT] || !(rz_new == rz_new)) ? 1.0 : 0.0;
        }
    }
}

static const void *step_shard_kernel_for(int bs) {
    switch (bs) {
        case 12: return (const void *)cg_step_shard_kernel<12>;
        case 24: return (const void *)cg_step_shard_kernel<24>;
        case 36: return (const void *)cg_step_shard_kernel<36>;
        default: return nullptr;
    }
}

// row-sharded step, set-up: need[c] = 1 for every column the local rows of N reference; bit (owner of c) set in the mask of the
// preconditioner block of every own row that references c (block b belongs to the rank r with floor(r n_b / P) <= b < floor((r + 1) n_b / P))
__global__ void __launch_bounds__(256)
shard_need_kernel(int64_t n_rows, const int *__restrict__ indptr, const int *__restrict__ indices, const int *__restrict__ col_block, int n_b, int P,
                  int blk_lo, int blk_hi, unsigned char *__restrict__ need, unsigned int *__restrict__ blk_mask) {
    const int64_t r = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const int b_row = col_block[r];
    const bool own_row = b_row >= blk_lo && b_row < blk_hi;
    unsigned int mask = 0;
    for (int k = indptr[r] + (int)lane_id(); k < indptr[r + 1]; k += 32) {
        const int c = indices[k];
        need[c] = 1;
        const int b = col_block[c];
        if (own_row && b >= 0) {
            int t = (int)(((long long)b * P) / n_b);
            while (t + 1 < P && (int)(((long long)(t + 1) * n_b) / P) <= b) ++t;
            while (t > 0 && (int)(((long long)t * n_b) / P) > b) --t;
            mask |= 1u << t;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
    if (own_row && lane_id() == 0 && mask) atomicOr(blk_mask + (b_row - blk_lo), mask);
}

// slots of the packed z buffer whose column is needed, appended in (roughly) ascending order: one append per warp
__global__ void __launch_bounds__(256)
shard_dlist_kernel(int64_t n_slots, const int *__restrict__ cols_pad, const unsigned char *__restrict__ need, int *__restrict__ dlist,
                   int *__restrict__ counter) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int c = k < n_slots ? cols_pad[k] : -1;
    const bool keep = c >= 0 && need[c];
    const unsigned int m = __ballot_sync(0xffffffffu, keep);
    if (!m) return;
    int base = 0;
    if (lane_id() == 0) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) dlist[base + __popc(m & ((1u << lane_id()) - 1u))] = (int)k;
}

__global__ void __launch_bounds__(256)
narrow_mask_kernel(int n, const unsigned int *__restrict__ in, unsigned char *__restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = (unsigned char)(in[i] & 0xffu);
}

// out[k - base] = key0 + r for every entry k of row r of a CSR pointer array (one warp per row)
__global__ void __launch_bounds__(256)
expand_rows_kernel(int n_rows, const int *__restrict__ ptr, int base, int key0, int *__restrict__ out) {
    const int r = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    if (r >= n_rows) return;
    for (int k = ptr[r] + (int)lane_id(); k < ptr[r + 1]; k += 32) out[k - base] = key0 + r;
}

// v[i] = 0 unless column i is this rank's own (or, with keep_e, inside the eliminated block)
__global__ void __launch_bounds__(VEC_BLOCK)
mask_own_kernel(int64_t n, const unsigned char *__restrict__ owned, int64_t e0, int64_t e1, int keep_e, double *__restrict__ v) {
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) {
        const bool in_e = i >= e0 && i < e1;
        if (!(owned[i] || (in_e && keep_e))) v[i] = 0.0;
    }
}

// Diagonal and right-hand side of the eliminated block from this rank's OWN rows (halo copies of other ranks' points
// are left out; the constraint rows are counted by the caller on one rank only):
// esum[e] += w a^2, esum[n_e + e] += w a b for every entry a of row r at column e0 + e.
__global__ void __launch_bounds__(256)
own_bias_sums_kernel(int64_t row_lo, int64_t row_hi, int64_t n_pad, int K, const int *__restrict__ ell_col, const double *__restrict__ ell_val,
                     const double *__restrict__ w, const double *__restrict__ b, int64_t e0, int n_e, double *__restrict__ esum) {
    for (int64_t r0 = row_lo + (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); r0 < row_hi; r0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = r0 + lane_id();
        const bool in = r < row_hi;
        const double wr = in ? w[r] : 0.0, br = in ? b[r] : 0.0;
        for (int k = 0; k < K; ++k) {
            int e = -1;
            double a = 0.0;
            if (in) {
                const int c = ell_col[(int64_t)k * n_pad + r];
                a = ell_val[(int64_t)k * n_pad + r];
                if (a != 0.0 && c >= e0 && c < e0 + n_e) e = (int)(c - e0);
            }
            if (__ballot_sync(0xffffffffu, e >= 0) == 0u) continue;
            // neighbouring points mostly share their bias column: one RED per distinct column and warp
            const unsigned peers = __match_any_sync(0xffffffffu, e);
            double d = wr * a * a, g = wr * a * br;
            double ds = 0.0, gs = 0.0;
            for (unsigned m = peers; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                ds += __shfl_sync(peers, d, src);
                gs += __shfl_sync(peers, g, src);
            }
            if (e >= 0 && lane_id() == (unsigned)(__ffs(peers) - 1)) {
                atomicAdd(esum + e, ds);
                atomicAdd(esum + n_e + e, gs);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
set_block_kernel(int64_t e0, int n_e, const double *__restrict__ esum, double *__restrict__ diag, double *__restrict__ g) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_e) { diag[e0 + e] = esum[e]; g[e0 + e] = esum[n_e + e]; }
}

// out_E = scale * einv .* c_E (scale -1: p_E = -N_EE^-1 u; +1 with c = g: the E part of N^-1-eliminated right-hand side)
__global__ void __launch_bounds__(256)
elim_scale_kernel(int64_t e0, int n_e, const double *__restrict__ einv, const double *__restrict__ c, double scale, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_e) out[e0 + e] = scale * einv[e] * c[e];
}

__global__ void __launch_bounds__(VEC_BLOCK)
mul_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK) out[i] = a[i] * b[i];
}

__global__ void __launch_bounds__(VEC_BLOCK)
mark_fixed_kernel(int64_t n_fixed, const int *__restrict__ cols, unsigned char *__restrict__ fixed) {
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n_fixed; i += (int64_t)gridDim.x * VEC_BLOCK) fixed[cols[i]] = 1;
}

__global__ void __launch_bounds__(VEC_BLOCK)
zero_fixed_kernel(int64_t n, const unsigned char *__restrict__ fixed, double *__restrict__ x) {
    for (int64_t i = (int64_t)blockIdx.x * VEC_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * VEC_BLOCK)
        if (fixed[i]) x[i] = 0.0;
}

// diag_j = sum_k At.data[k]^2 * w[At.indices[k]] : one warp per row of A^T (a one-off per outer iteration)
__global__ void __launch_bounds__(256)
diag_kernel(int64_t n_cols, const int *__restrict__ t_indptr, const int *__restrict__ t_indices, const double *__restrict__ t_data,
            const double *__restrict__ w, double *__restrict__ diag) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n_cols; c += n_warps) {
        const int b = t_indptr[c], e = t_indptr[c + 1];
        double acc = 0.0;
        for (int k = b + (int)lane_id(); k < e; k += 32) {
            const double a = t_data[k];
            acc += a * a * w[t_indices[k]];
        }
        acc = warp_sum(acc);
        if (lane_id() == 0) diag[c] = acc;
    }
}

}  // namespace lss

using namespace lss;

struct lss_solver {
    int64_t n_rows, n_cols, nnz;
    const int32_t *indptr, *indices, *t_indptr, *t_indices;
    const double *data, *t_data, *rhs;
    lss_comm_t *comm;
    cudaStream_t stream;
    // owned device memory
    double *w, *u, *buf /* q[n_cols] | partials[VEC_GRID] */, *r, *p, *minv, *g, *diag, *part2, *scal;
    unsigned char *fixed;
    double *host_scal;  // pinned
    cudaEvent_t ev0, ev1, ev_batch[2];
    bool have_weights;
    bool profile;
    // stats of the last solve
    double ms_total, ms_a, ms_at, ms_step;
    int64_t n_apply;
    int64_t launches;
    std::vector<cudaEvent_t> *prof_events;   // fixed pool: 3 events per operator application, 2 batches deep
    int64_t n_prof;
    lss::SpmvPlan plan_a, plan_at, plan_c, plan_ct;
    // block form (lss_solver_create_blocks): data rows as an ELL copy applied by the fused G^T W G kernel,
    // constraint rows as a CSR pair (C, C^T); the A / A^T arrays above are then unused
    bool blocks;
    bool q_clean;          // q and the p.q slots were zeroed by the last cg_step_kernel
    int64_t n_data, n_con, ell_pad;
    int ell_k;
    int32_t *ell_col;
    double *ell_val;
    const int32_t *c_indptr, *c_indices, *ct_indptr, *ct_indices;
    const double *c_data, *ct_data;
    int64_t c_nnz;
    // explicit form: N = A^T W A (normal.cu), built from the ELL rows; the ELL copy stays for the right-hand side
    lss::NormalMatrix *nm;
    // structured form (lss_solver_set_structure): block-Jacobi preconditioner and / or eliminated diagonal block
    lss::BlockPrecond *bp;
    int64_t e0, e1;          // eliminated columns [e0, e1); e0 == e1: none
    double *einv, *z, *eraw;   // eraw: rank-local N_Eg v sums (multi-rank only)
    double *r2;              // second residual buffer of the thread-per-row step kernel (swapped with r by parity)
    int step_rows;           // 2: cg_step_tile_kernel, 1: cg_step_rows_kernel, 0: cg_step_p_kernel
    int step_grid;           // CTAs of the cooperative step kernel
    int step_stages;         // shared-memory stages per warp for the block inverses
    size_t step_smem;
    bool l2_window;          // an L2 persisting window on the preconditioner blocks is installed on the stream
    lss::CoarseSpace *cs;    // coarse level of the two-level preconditioner (lss_solver_set_coarse); null: block-Jacobi alone
    int cs_refresh;          // 1: S_c rebuilt with every weight change, 0: built with the first weights of the solver only
    bool cs_built;
    int step_ml_grid;
    // row-sharded N over the ranks of one node (lss_solver_create_sharded): p exchanged through the P2P mailbox
    bool shard;
    lss_comm *pc;            // communicator with an open mailbox; NCCL only in set-up / start / end of a solve (`comm` stays null)
    unsigned char *owned;    // [n_cols] flags of this rank's own columns
    int blk_lo, blk_hi;      // its preconditioner blocks
    int64_t own_row_lo, own_row_hi, con_row_lo;   // own data rows and first constraint row inside the ELL block
    double *u_e, *tk, *esum; // T^T p carried through the iterations, this rank's N_E,own z, [2 n_e] set-up sums
    int32_t *fr_key = nullptr, *fe_key = nullptr;   // row keys of the flat entry lists of phase C1 (row-sharded step)
    int n_fr = 0, n_fe = 0;
    int64_t kE0 = 0;
    double *tgt = nullptr;
    lss::StructuredN *sn = nullptr;        // grid-structured copy of N for the CG product (lss_solver_set_structured)
    unsigned char *blk_peers = nullptr;   // per own block: ranks that reference it
    int32_t *dlist = nullptr; // slots of the packed z buffer whose columns this rank's rows of N reference (own + halo): the only entries of p it keeps current
    int n_dlist = 0;
    int step_sh_grid;
};


// Stream-ordered pool allocations: a solver is created and destroyed once per smooth_fit call, and cudaFree of
// GB-sized blocks costs up to a second; freed pool memory is simply reused by the next solver.
static int dev_alloc(double **p, int64_t n, cudaStream_t st) {
    lss::Scratch::configure_pool();
    LSS_CUDA(cudaMallocAsync((void **)p, sizeof(double) * (size_t)(n > 0 ? n : 1), st));
    return LSS_OK;
}

extern "C" int lss_solver_destroy(lss_solver_t *s) {
    if (!s) return LSS_OK;
    cudaStreamSynchronize(s->stream);
    double **ptrs[] = {&s->w, &s->u, &s->buf, &s->r, &s->p, &s->minv, &s->g, &s->diag, &s->part2, &s->scal};
    for (auto pp : ptrs)
        if (*pp) cudaFreeAsync(*pp, s->stream);
    lss::spmv_plan_destroy(&s->plan_a);
    lss::spmv_plan_destroy(&s->plan_at);
    lss::spmv_plan_destroy(&s->plan_c);
    lss::spmv_plan_destroy(&s->plan_ct);
    if (s->l2_window) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &attr);     // num_bytes = 0: no window
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    if (s->bp) lss::block_precond_destroy(s->bp);
    if (s->cs) lss::coarse_destroy(s->cs);
    if (s->owned) cudaFreeAsync(s->owned, s->stream);
    if (s->u_e) cudaFreeAsync(s->u_e, s->stream);
    if (s->tk) cudaFreeAsync(s->tk, s->stream);
    if (s->sn) lss::structured_destroy(s->sn);
    if (s->dlist) cudaFreeAsync(s->dlist, s->stream);
    if (s->blk_peers) cudaFreeAsync(s->blk_peers, s->stream);
    if (s->fr_key) cudaFreeAsync(s->fr_key, s->stream);
    if (s->fe_key) cudaFreeAsync(s->fe_key, s->stream);
    if (s->tgt) cudaFreeAsync(s->tgt, s->stream);
    if (s->esum) cudaFreeAsync(s->esum, s->stream);
    if (s->einv) cudaFreeAsync(s->einv, s->stream);
    if (s->eraw) cudaFreeAsync(s->eraw, s->stream);
    if (s->z) cudaFreeAsync(s->z, s->stream);
    if (s->r2) cudaFreeAsync(s->r2, s->stream);
    if (s->nm) lss::normal_destroy(s->nm);
    if (s->ell_col) cudaFreeAsync(s->ell_col, s->stream);
    if (s->ell_val) cudaFreeAsync(s->ell_val, s->stream);
    if (s->fixed) cudaFreeAsync(s->fixed, s->stream);
    if (s->host_scal) cudaFreeHost(s->host_scal);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (int i = 0; i < 2; ++i)
        if (s->ev_batch[i]) cudaEventDestroy(s->ev_batch[i]);
    if (s->prof_events) {
        for (auto e : *s->prof_events) cudaEventDestroy(e);
        delete s->prof_events;
    }
    delete s;
    return LSS_OK;
}

static int solver_init_common(lss_solver *s, int64_t n_rows, int64_t n_cols, const double *rhs, const int32_t *fixed_cols,
                              int64_t n_fixed, lss_comm_t *comm, void *stream) {
    memset(s, 0, sizeof(*s));
    s->n_rows = n_rows; s->n_cols = n_cols;
    s->rhs = rhs; s->comm = comm; s->stream = (cudaStream_t)stream;
    s->prof_events = new std::vector<cudaEvent_t>();
    int rc = LSS_OK;
    auto chk = [&](int r) { if (rc == LSS_OK) rc = r; };
    chk(dev_alloc(&s->w, n_rows, s->stream)); chk(dev_alloc(&s->u, n_rows, s->stream));
    chk(dev_alloc(&s->buf, n_cols + VEC_GRID, s->stream)); chk(dev_alloc(&s->r, n_cols, s->stream)); chk(dev_alloc(&s->p, n_cols, s->stream));
    chk(dev_alloc(&s->minv, n_cols, s->stream)); chk(dev_alloc(&s->g, n_cols, s->stream)); chk(dev_alloc(&s->diag, n_cols + VEC_GRID, s->stream));
    chk(dev_alloc(&s->part2, 2 * kSMs * 8, s->stream)); chk(dev_alloc(&s->scal, S_COUNT, s->stream));
    if (rc == LSS_OK && cudaMallocAsync((void **)&s->fixed, (size_t)n_cols, s->stream) != cudaSuccess) rc = fail(LSS_ECUDA, "solver_create: cudaMalloc failed");
    if (rc == LSS_OK && cudaMallocHost((void **)&s->host_scal, sizeof(double) * S_COUNT * 2) != cudaSuccess) rc = fail(LSS_ECUDA, "solver_create: cudaMallocHost failed");
    if (rc == LSS_OK) {
        cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);
        cudaEventCreateWithFlags(&s->ev_batch[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s->ev_batch[1], cudaEventDisableTiming);
        cudaMemsetAsync(s->fixed, 0, (size_t)n_cols, s->stream);
        cudaMemsetAsync(s->scal, 0, sizeof(double) * S_COUNT, s->stream);
        if (n_fixed > 0) mark_fixed_kernel<<<stride_grid(n_fixed, VEC_BLOCK), VEC_BLOCK, 0, s->stream>>>(n_fixed, fixed_cols, s->fixed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(LSS_ECUDA, "solver_create: launch failed");
    }
    s->profile = getenv("LSS_PROFILE_SPMV") != nullptr;
    return rc;
}

static int read_nnz(const int32_t *indptr, int64_t n_rows, cudaStream_t st, int64_t *nnz) {
    int v = 0;
    if (n_rows > 0) {
        LSS_CUDA(cudaMemcpyAsync(&v, indptr + n_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
    }
    *nnz = v;
    return LSS_OK;
}

extern "C" int lss_solver_create(lss_solver_t **solver, int64_t n_rows, int64_t n_cols, const int32_t *indptr,
                                 const int32_t *indices, const double *data, const int32_t *t_indptr,
                                 const int32_t *t_indices, const double *t_data, const double *rhs,
                                 const int32_t *fixed_cols, int64_t n_fixed, lss_comm_t *comm, void *stream) {
    if (!solver || n_rows < 0 || n_cols <= 0) return fail(LSS_EINVAL, "solver_create: bad sizes");
    lss_solver *s = new lss_solver();
    int rc = solver_init_common(s, n_rows, n_cols, rhs, fixed_cols, n_fixed, comm, stream);
    s->indptr = indptr; s->indices = indices; s->data = data;
    s->t_indptr = t_indptr; s->t_indices = t_indices; s->t_data = t_data;
    if (rc == LSS_OK) rc = read_nnz(indptr, n_rows, s->stream, &s->nnz);
    if (rc == LSS_OK) rc = spmv_plan_create(&s->plan_a, n_rows, indptr, s->nnz, s->stream);
    if (rc == LSS_OK) rc = spmv_plan_create(&s->plan_at, n_cols, t_indptr, s->nnz, s->stream);
    if (rc != LSS_OK) { lss_solver_destroy(s); return rc; }
    *solver = s;
    return LSS_OK;
}

struct ShardSpec {
    const uint8_t *owned_cols;     // device [n_cols]
    int32_t block_lo, block_hi;
    int64_t own_row_lo, own_row_hi, con_row_lo, elim_begin, elim_end;
    lss_comm *pc;
};

static int create_blocks_impl(lss_solver_t **solver, int64_t n_data, int64_t n_con, int64_t n_cols, int32_t row_len,
                              const int32_t *g_indptr, const int32_t *g_indices, const double *g_data,
                              const int32_t *c_indptr, const int32_t *c_indices, const double *c_data,
                              const int32_t *ct_indptr, const int32_t *ct_indices, const double *ct_data,
                              const double *rhs, const int32_t *fixed_cols, int64_t n_fixed, lss_comm_t *comm,
                              void *stream, const ShardSpec *sh) {
    if (!solver || n_data < 0 || n_con < 0 || n_cols <= 0) return fail(LSS_EINVAL, "solver_create_blocks: bad sizes");
    if (n_data > 0 && (row_len < 1 || row_len > 24)) return fail(LSS_EINVAL, "solver_create_blocks: data rows must hold 1..24 entries");
    lss_solver *s = new lss_solver();
    int rc = solver_init_common(s, n_data + n_con, n_cols, rhs, fixed_cols, n_fixed, comm, stream);
    cudaStream_t st = s->stream;
    if (sh && rc == LSS_OK) {
        s->shard = true;
        s->pc = sh->pc;
        s->blk_lo = sh->block_lo; s->blk_hi = sh->block_hi;
        s->own_row_lo = sh->own_row_lo; s->own_row_hi = sh->own_row_hi; s->con_row_lo = sh->con_row_lo;
        if (cudaMallocAsync((void **)&s->owned, (size_t)n_cols, st) != cudaSuccess) rc = fail(LSS_ECUDA, "solver_create_sharded: cudaMallocAsync failed");
        else cudaMemcpyAsync(s->owned, sh->owned_cols, (size_t)n_cols, cudaMemcpyDeviceToDevice, st);
    }
    s->blocks = true;
    s->n_data = n_data; s->n_con = n_con;
    s->c_indptr = c_indptr; s->c_indices = c_indices; s->c_data = c_data;
    s->ct_indptr = ct_indptr; s->ct_indices = ct_indices; s->ct_data = ct_data;
    int64_t g_nnz = 0;
    if (rc == LSS_OK) rc = read_nnz(g_indptr, n_data, st, &g_nnz);
    if (rc == LSS_OK && n_con > 0) rc = read_nnz(c_indptr, n_con, st, &s->c_nnz);
    s->nnz = g_nnz + s->c_nnz;
    if (rc == LSS_OK && n_data > 0) {
        s->ell_k = row_len;
        s->ell_pad = (n_data + 31) / 32 * 32;          // every ELL column stays 256-byte aligned
        if (cudaMallocAsync((void **)&s->ell_col, sizeof(int32_t) * s->ell_pad * row_len, st) != cudaSuccess ||
            cudaMallocAsync((void **)&s->ell_val, sizeof(double) * s->ell_pad * row_len, st) != cudaSuccess)
            rc = fail(LSS_ECUDA, "solver_create_blocks: cudaMallocAsync failed");
        if (rc == LSS_OK) rc = ell_build(n_data, row_len, g_indptr, g_indices, g_data, s->ell_col, s->ell_val, s->ell_pad, st);
    }
    if (rc == LSS_OK && n_data > 0 && n_con == 0) {
        int want = lss::g_opt_explicit_normal;
        if (want < 0) want = getenv("LSS_NO_EXPLICIT_N") ? 0 : 1;
        if (want) {
            rc = normal_create(&s->nm, n_data, s->ell_pad, s->ell_k, s->ell_col, s->ell_val, n_cols, s->fixed, st, s->owned,
                               sh ? sh->elim_begin : 0, sh ? sh->elim_end : 0);
            if (rc == LSS_ERANGE && !sh) { s->nm = nullptr; rc = LSS_OK; }   // too many column pairs for the 32-bit sort: fused ELL form
        }
    }
    if (rc == LSS_OK && sh && !s->nm) rc = fail(LSS_EINVAL, "solver_create_sharded: needs the explicit-normal form (all rows <= 24 entries in one block)");
    if (rc == LSS_OK && n_con > 0) rc = spmv_plan_create(&s->plan_c, n_con, c_indptr, s->c_nnz, st);
    if (rc == LSS_OK && n_con > 0) rc = spmv_plan_create(&s->plan_ct, n_cols, ct_indptr, s->c_nnz, st);
    if (rc != LSS_OK) { lss_solver_destroy(s); return rc; }
    *solver = s;
    return LSS_OK;
}

extern "C" int lss_solver_create_blocks(lss_solver_t **solver, int64_t n_data, int64_t n_con, int64_t n_cols, int32_t row_len,
                                        const int32_t *g_indptr, const int32_t *g_indices, const double *g_data,
                                        const int32_t *c_indptr, const int32_t *c_indices, const double *c_data,
                                        const int32_t *ct_indptr, const int32_t *ct_indices, const double *ct_data,
                                        const double *rhs, const int32_t *fixed_cols, int64_t n_fixed, lss_comm_t *comm,
                                        void *stream) {
    return create_blocks_impl(solver, n_data, n_con, n_cols, row_len, g_indptr, g_indices, g_data, c_indptr, c_indices, c_data, ct_indptr,
                              ct_indices, ct_data, rhs, fixed_cols, n_fixed, comm, stream, nullptr);
}

extern "C" int lss_solver_create_sharded(lss_solver_t **solver, int64_t n_rows, int64_t n_cols, int32_t row_len, const int32_t *indptr,
                                         const int32_t *indices, const double *data, const double *rhs, const int32_t *fixed_cols,
                                         int64_t n_fixed, const uint8_t *owned_cols, int32_t block_lo, int32_t block_hi,
                                         int64_t own_row_lo, int64_t own_row_hi, int64_t con_row_lo, int64_t elim_begin, int64_t elim_end,
                                         lss_comm_t *comm, void *stream) {
    if (!comm || !comm->box_open) return fail(LSS_EINVAL, "solver_create_sharded: the communicator needs an open mailbox (lss_comm_box_open)");
    if (!owned_cols || block_hi < block_lo || own_row_lo < 0 || own_row_hi < own_row_lo || con_row_lo < own_row_hi || con_row_lo > n_rows)
        return fail(LSS_EINVAL, "solver_create_sharded: bad ownership arguments");
    ShardSpec sh{owned_cols, block_lo, block_hi, own_row_lo, own_row_hi, con_row_lo, elim_begin, elim_end, comm};
    return create_blocks_impl(solver, n_rows, 0, n_cols, row_len, indptr, indices, data, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                              rhs, fixed_cols, n_fixed, nullptr, stream, &sh);
}

extern "C" int lss_solver_set_structure(lss_solver_t *s, int32_t n_blocks, const int32_t *block_ptr, const int32_t *block_cols,
                                        int32_t max_block, int64_t elim_begin, int64_t elim_end) {
    if (!s) return fail(LSS_EINVAL, "solver_set_structure: null solver");
    if (!s->nm) return LSS_OK;          // only the explicit-normal form has the matrix entries these need
    cudaStream_t st = s->stream;
    const bool multi = s->comm && s->comm->n_ranks > 1;
    const bool want_elim = elim_end > elim_begin;
    if (want_elim) {
        if (elim_begin < 0 || elim_end > s->n_cols) return fail(LSS_EINVAL, "solver_set_structure: eliminated range outside the columns");
        Scratch bad;
        LSS_CUDA(bad.alloc(sizeof(unsigned long long), st));
        LSS_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(unsigned long long), st));
        elim_check_kernel<<<ceil_div((elim_end - elim_begin) * 32, 256), 256, 0, st>>>(elim_begin, elim_end, s->nm->indptr, s->nm->indices,
                                                                                     bad.as<unsigned long long>());
        LSS_LAUNCH_CHECK();
        LSS_TRY(comm_allreduce(s->comm, bad.p, 1, 2, 0, st));      // the block must be diagonal on every rank
        unsigned long long nb = 0;
        LSS_CUDA(cudaMemcpyAsync(&nb, bad.p, sizeof(nb), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
        if (nb) return fail(LSS_EINVAL, "solver_set_structure: columns [%lld, %lld) are coupled to each other (%llu entries): not a diagonal block",
                            (long long)elim_begin, (long long)elim_end, nb);
        s->e0 = elim_begin; s->e1 = elim_end;
        LSS_TRY(dev_alloc(&s->einv, elim_end - elim_begin, st));
        if (multi || s->shard) LSS_TRY(dev_alloc(&s->eraw, elim_end - elim_begin, st));
        if (s->shard) {
            LSS_TRY(dev_alloc(&s->u_e, elim_end - elim_begin, st));
            LSS_TRY(dev_alloc(&s->tk, (elim_end - elim_begin) * SHARD_ECH_MAX, st));
            LSS_TRY(dev_alloc(&s->esum, 2 * (elim_end - elim_begin), st));
        }
    }
    if (n_blocks > 0) LSS_TRY(block_precond_create(&s->bp, s->n_cols, n_blocks, block_ptr, block_cols, max_block, st));
    if (s->bp && !getenv("LSS_NO_L2_WINDOW")) {
        // keep the block inverses (re-read every CG iteration, ~50 MB at 10^4 blocks) in the persisting part of L2
        // while the normal matrix streams through the rest
        // device limits are queried once per process and the carve-out is only set when it changes: cudaGetDeviceProperties
        // and cudaDeviceSetLimit cost up to 10^2 ms on some hosts, and a solver is created per smooth_fit call
        static int prop_dev = -1, max_persist = 0, max_window = 0;
        static size_t carve_set = 0;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev != prop_dev) {
            max_persist = max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            prop_dev = dev;
            carve_set = 0;
        }
        if (max_persist > 0 && max_window > 0) {
            const size_t bytes = sizeof(float) * (size_t)s->bp->n_blocks * s->bp->bs * s->bp->bs;
            size_t carve = bytes + (8u << 20);
            if (carve > (size_t)max_persist) carve = (size_t)max_persist;
            size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
            bool carve_ok = carve_set == carve;
            if (!carve_ok && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                carve_set = carve;
                carve_ok = true;
            }
            if (carve_ok) {
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = s->bp->inv;
                attr.accessPolicyWindow.num_bytes = win;
                attr.accessPolicyWindow.hitRatio = carve >= win ? 1.0f : (float)((double)carve / (double)win);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) s->l2_window = true;
            }
            cudaGetLastError();     // the window is an optimisation: never fail the solve over it
        }
    }
    if (s->bp || s->e1 > s->e0) {
        LSS_TRY(dev_alloc(&s->z, s->n_cols, st));
        // shared-memory staging of the block inverses: two stages per warp when they fit next to a second CTA
        const size_t one = s->bp ? sizeof(float) * (size_t)s->bp->bs * s->bp->bs * (VEC_BLOCK / 32) : 0;
        s->step_stages = (2 * one <= 100 * 1024) ? 2 : 1;
        s->step_smem = one * s->step_stages;
        if (s->step_smem > 48 * 1024)
            LSS_CUDA(cudaFuncSetAttribute(cg_step_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->step_smem));
        int per_sm = 0;
        int want_rows = lss::g_opt_step_kernel;
        if (want_rows < 0) {
            const char *e = getenv("LSS_STEP_KERNEL");
            want_rows = e ? atoi(e) : 2;
        }
        if (want_rows < 0 || want_rows > 2) want_rows = 2;
        if (want_rows == 2 && (!s->bp || !step_tile_kernel_for(s->bp->bs))) want_rows = 1;      // block stride without an instance
        s->step_rows = want_rows;
        if (s->step_rows == 2) {
            if (!s->r2) LSS_TRY(dev_alloc(&s->r2, s->n_cols, st));
            LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_tile_kernel_for(s->bp->bs), VEC_BLOCK, 0));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > STEP3_PER_SM) per_sm = STEP3_PER_SM;
        } else if (s->step_rows) {
            if (!s->r2) LSS_TRY(dev_alloc(&s->r2, s->n_cols, st));
            LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_step_rows_kernel, VEC_BLOCK, 0));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > STEP2_PER_SM) per_sm = STEP2_PER_SM;
        } else {
            LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_step_p_kernel, VEC_BLOCK, s->step_smem));
            if (per_sm < 1) return fail(LSS_EINVAL, "solver_set_structure: %zu bytes of staging memory per CTA do not fit", s->step_smem);
            if (per_sm > STEP_PER_SM) per_sm = STEP_PER_SM;
        }
        s->step_grid = kSMs * per_sm;
    }
    if (s->shard) {
        if (!s->bp || s->step_rows != 2 || !step_shard_kernel_for(s->bp->bs))
            return fail(LSS_EINVAL, "solver_set_structure: the row-sharded solver needs preconditioner blocks of 12, 24 or 36 unknowns");
        if (s->blk_hi > s->bp->n_blocks) return fail(LSS_EINVAL, "solver_set_structure: own block range outside the blocks");
        int per_sm = 0;
        LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_shard_kernel_for(s->bp->bs), VEC_BLOCK, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > STEP3_PER_SM) per_sm = STEP3_PER_SM;
        s->step_sh_grid = kSMs * per_sm;
    }
    s->have_weights = false;
    return LSS_OK;
}

// q(buf[0:n]) = A^T (w .* (A v)), partials of v.q appended at buf[n:], summed over ranks.
constexpr int kBatch = 16;   // CG iterations enqueued between two looks at the convergence flag
constexpr int kProfEvery = 4; // with lss_solver_set_profile: timing events around every 4th iteration (events cost ~10 us of stream time each)

static int apply_normal(lss_solver *s, const double *v, bool with_dot, const double *stop, int prof_slot = -1) {
    cudaStream_t st = s->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    const bool timed = s->profile && prof_slot >= 0;
    if (timed) {
        while ((int)s->prof_events->size() < 4 * 2 * kBatch) {   // created once, reused for every batch
            cudaEvent_t e;
            LSS_CUDA(cudaEventCreate(&e));
            s->prof_events->push_back(e);
        }
        e0 = (*s->prof_events)[4 * prof_slot];
        e1 = (*s->prof_events)[4 * prof_slot + 1];
        e2 = (*s->prof_events)[4 * prof_slot + 2];
        LSS_CUDA(cudaEventRecord(e0, st));
    }
    if (s->nm) {
        if (timed) LSS_CUDA(cudaEventRecord(e1, st));
        // q = N v : one gather product, one launch (rows of an eliminated trailing block are left out: (N p)_E = 0)
        const int64_t rows = (s->e1 > s->e0 && s->e1 == s->n_cols) ? s->e0 : s->n_cols;
        double *slots = (s->z && with_dot) ? s->buf + s->n_cols : nullptr;      // p.q by-product for the structured step kernel
        if (slots && !s->q_clean) LSS_CUDA(cudaMemsetAsync(slots, 0, sizeof(double) * PQ_SLOTS, st));
        if (s->sn && s->sn->d.rows == rows) {
            if (!s->q_clean) LSS_CUDA(cudaMemsetAsync(s->buf, 0, sizeof(double) * s->n_cols, st));
            LSS_TRY(lss::structured_apply(s->sn, v, s->buf, stop, &s->launches, slots));
        } else
        LSS_TRY(normal_apply_fast(s->nm, rows, v, s->buf, s->q_clean, stop, &s->launches, slots));
        s->q_clean = false;
    } else if (s->blocks) {
        // constraint rows: q = C^T (w_c .* (C v)) with the tiled CSR kernel (writes every entry of q) ...
        if (s->n_con > 0) {
            LSS_TRY(spmv_launch(s->n_con, s->c_indptr, s->c_indices, s->c_data, v, s->u + s->n_data, LSS_SPMV_SCALE,
                                s->w + s->n_data, s->c_nnz, &s->plan_c, stop, st, &s->launches));
            LSS_TRY(spmv_launch(s->n_cols, s->ct_indptr, s->ct_indices, s->ct_data, s->u + s->n_data, s->buf, LSS_SPMV_PLAIN, nullptr,
                                s->c_nnz, &s->plan_ct, stop, st, &s->launches));
            LSS_CUDA(cudaMemsetAsync(s->buf + s->n_cols, 0, sizeof(double) * PQ_SLOTS, st));
        } else if (!s->q_clean) {
            LSS_CUDA(cudaMemsetAsync(s->buf, 0, sizeof(double) * (s->n_cols + PQ_SLOTS), st));
        }
        if (timed) LSS_CUDA(cudaEventRecord(e1, st));
        // ... data rows: q += G^T (w .* (G v)) and the p.q partials, one fused pass over the ELL copy of G
        LSS_TRY(ell_normal_launch(0, s->n_data, s->ell_pad, s->ell_k, s->ell_col, s->ell_val, v, s->w, nullptr, s->buf,
                                  s->buf + s->n_cols, stop, st));
        if (s->n_data > 0) s->launches += 1;
        s->q_clean = false;
    } else {
        LSS_TRY(spmv_launch(s->n_rows, s->indptr, s->indices, s->data, v, s->u, LSS_SPMV_SCALE, s->w, s->nnz, &s->plan_a, stop, st, &s->launches));
        if (timed) LSS_CUDA(cudaEventRecord(e1, st));
        LSS_TRY(spmv_launch(s->n_cols, s->t_indptr, s->t_indices, s->t_data, s->u, s->buf, LSS_SPMV_PLAIN, nullptr, s->nnz, &s->plan_at, stop, st, &s->launches));
    }
    if (timed) LSS_CUDA(cudaEventRecord(e2, st));
    int64_t n_red = s->n_cols;
    if (s->nm) {
        if (s->z && with_dot) n_red += PQ_SLOTS;   // structured form: the product left its p.q partials behind q
        // plain form: p.q is formed from the reduced q by cg_step_n_kernel on every rank
    } else if (s->blocks) {
        n_red += PQ_SLOTS;              // the fused kernel already left the p.q partials behind q
    } else if (with_dot) {
        dot_partials_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(s->n_cols, v, s->buf, s->buf + s->n_cols, stop);
        LSS_LAUNCH_CHECK();
        s->launches += 1;
        n_red += VEC_GRID;
    }
    LSS_TRY(comm_allreduce(s->comm, s->buf, n_red, 0, 0, st));
    s->n_apply += 1;
    return LSS_OK;
}

// v_E = N_EE^-1 (c_E - N_Eg v_g): one kernel on one rank; local sums, all-reduce, finish when the rows of N are sharded
static int elim_slave(lss_solver *s, double *v, const double *c, const double *stop) {
    cudaStream_t st = s->stream;
    const int n_e = (int)(s->e1 - s->e0);
    elim_solve_kernel<<<n_e, ELIM_BLOCK, 0, st>>>(s->e0, s->e1, s->nm->indptr, s->nm->indices, s->nm->data, v, c, s->einv, v, stop, s->eraw);
    LSS_LAUNCH_CHECK();
    s->launches += 1;
    if (s->eraw) {
        LSS_TRY(comm_allreduce(s->comm, s->eraw, n_e, 0, 0, st));
        elim_finish_kernel<<<ceil_div(n_e, 256), 256, 0, st>>>(s->e0, s->e1, s->eraw, c, s->einv, v, stop);
        LSS_LAUNCH_CHECK();
        s->launches += 1;
    }
    return LSS_OK;
}

// z = M^-1 r for the structured form
static int apply_precond(lss_solver *s, const double *r, double *z) {
    if (s->bp) {
        LSS_TRY(block_precond_apply(s->bp, s->n_cols, r, s->minv, z, s->stream));
        if (s->cs && s->cs->ready) LSS_TRY(coarse_apply_add(s->cs, r, z));
        return LSS_OK;
    }
    mul_kernel<<<VEC_GRID, VEC_BLOCK, 0, s->stream>>>(s->n_cols, s->minv, r, z);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

extern "C" int lss_solver_set_coarse(lss_solver_t *s, int32_t n_coarse, const int32_t *cp_idx, const float *cp_w, const int32_t *pt_ptr,
                                     const int32_t *pt_col, const float *pt_w, int64_t nnz_p, int32_t refresh) {
    if (!s) return fail(LSS_EINVAL, "solver_set_coarse: null solver");
    if (!s->nm || !s->bp || s->step_rows != 2) return LSS_OK;     // needs the explicit-normal form with the tile step kernel
    if (s->cs) { lss::coarse_destroy(s->cs); s->cs = nullptr; }
    // the flat lists of the row-sharded step follow the coarse space: rebuilt at the next solve
    if (s->fr_key) { cudaFreeAsync(s->fr_key, s->stream); s->fr_key = nullptr; }
    if (s->fe_key) { cudaFreeAsync(s->fe_key, s->stream); s->fe_key = nullptr; }
    if (s->tgt) { cudaFreeAsync(s->tgt, s->stream); s->tgt = nullptr; }
    LSS_TRY(coarse_create(&s->cs, s->n_cols, n_coarse, cp_idx, cp_w, pt_ptr, pt_col, pt_w, nnz_p, (int)(s->e1 - s->e0), s->stream));
    int per_sm = 0;
    LSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_ml_kernel_for(s->bp->bs), VEC_BLOCK, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > STEP3_PER_SM) per_sm = STEP3_PER_SM;
    s->step_ml_grid = kSMs * per_sm;
    if (s->shard) s->step_ml_grid = s->step_sh_grid;
    s->cs_refresh = refresh;
    s->cs_built = false;
    s->have_weights = false;
    return LSS_OK;
}


// ------------------------------------------------------------------------------------------------ row-sharded solver
namespace {

// z = M^-1 r on this rank's columns (own blocks exactly, coarse level through an all-reduce of the partial restriction);
// every other entry of z is zero.  Set-up / solve start only: the iterations use the fused step kernel.
int shard_precond(lss_solver *s, const double *r, double *z) {
    cudaStream_t st = s->stream;
    LSS_CUDA(cudaMemsetAsync(z, 0, sizeof(double) * s->n_cols, st));
    LSS_TRY(block_precond_apply(s->bp, s->n_cols, r, s->minv, z, st, s->blk_lo, s->blk_hi));
    if (s->cs && s->cs->ready) LSS_TRY(coarse_apply_add(s->cs, r, z, s->pc));
    mask_own_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(s->n_cols, s->owned, s->e0, s->e1, 0, z);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

// v_E = N_EE^-1 (c_E - N_Eg v_g) with the column sums completed over the ranks; the raw sums (N_Eg v_g) stay in s->eraw
int shard_slave(lss_solver *s, double *v, const double *c) {
    cudaStream_t st = s->stream;
    const int n_e = (int)(s->e1 - s->e0);
    if (n_e <= 0) return LSS_OK;
    elim_solve_kernel<<<n_e, ELIM_BLOCK, 0, st>>>(s->e0, s->e1, s->nm->indptr, s->nm->indices, s->nm->data, v, c, s->einv, v, nullptr, s->eraw);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(s->pc, s->eraw, n_e, 0, 0, st));
    elim_finish_kernel<<<ceil_div(n_e, 256), 256, 0, st>>>(s->e0, s->e1, s->eraw, c, s->einv, v, nullptr);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int shard_set_weights(lss_solver *s, const double *w) {
    cudaStream_t st = s->stream;
    const int64_t n = s->n_cols;
    const int n_e = (int)(s->e1 - s->e0);
    if (!s->bp || !s->z) return fail(LSS_EINVAL, "solver_set_weights: the row-sharded solver needs lss_solver_set_structure with blocks first");
    if (s->n_rows > 0) {
        LSS_CUDA(cudaMemcpyAsync(s->w, w, sizeof(double) * s->n_rows, cudaMemcpyDeviceToDevice, st));
        mul_kernel<<<stride_grid(s->n_rows, VEC_BLOCK), VEC_BLOCK, 0, st>>>(s->n_rows, s->w, s->rhs, s->u);
        LSS_LAUNCH_CHECK();
    }
    // right-hand side and N at the own columns: every row touching them is held by this rank (own points + halo copies)
    LSS_CUDA(cudaMemsetAsync(s->g, 0, sizeof(double) * n, st));
    LSS_TRY(ell_normal_launch(1, s->n_data, s->ell_pad, s->ell_k, s->ell_col, s->ell_val, nullptr, s->w, s->rhs, s->g, nullptr, nullptr, st));
    LSS_TRY(normal_update(s->nm, s->w));
    LSS_TRY(normal_diag(s->nm, s->diag));
    if (n_e > 0) {
        // eliminated block: diagonal and right-hand side from OWN rows only (+ the constraint rows on rank 0), summed over ranks
        LSS_CUDA(cudaMemsetAsync(s->esum, 0, sizeof(double) * 2 * n_e, st));
        if (s->own_row_hi > s->own_row_lo) {
            own_bias_sums_kernel<<<stride_grid(s->own_row_hi - s->own_row_lo, 256), 256, 0, st>>>(s->own_row_lo, s->own_row_hi, s->ell_pad, s->ell_k,
                                                                                              s->ell_col, s->ell_val, s->w, s->rhs, s->e0, n_e, s->esum);
            LSS_LAUNCH_CHECK();
        }
        if (s->pc->rank == 0 && s->n_rows > s->con_row_lo) {
            own_bias_sums_kernel<<<stride_grid(s->n_rows - s->con_row_lo, 256), 256, 0, st>>>(s->con_row_lo, s->n_rows, s->ell_pad, s->ell_k, s->ell_col,
                                                                                          s->ell_val, s->w, s->rhs, s->e0, n_e, s->esum);
            LSS_LAUNCH_CHECK();
        }
        LSS_TRY(comm_allreduce(s->pc, s->esum, 2 * n_e, 0, 0, st));
        set_block_kernel<<<ceil_div(n_e, 256), 256, 0, st>>>(s->e0, n_e, s->esum, s->diag, s->g);
        LSS_LAUNCH_CHECK();
    }
    mask_own_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->owned, s->e0, s->e1, 1, s->g);
    LSS_LAUNCH_CHECK();
    jacobi_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->diag, s->fixed, s->g, s->minv, s->part2);
    LSS_LAUNCH_CHECK();
    LSS_TRY(block_precond_factor(s->bp, s->nm, nullptr, s->blk_lo, s->blk_hi));
    if (n_e > 0) {
        elim_setup_kernel<<<ceil_div(n_e, 256), 256, 0, st>>>(s->e0, s->e1, s->diag, s->einv, s->minv);
        LSS_LAUNCH_CHECK();
    }
    const int64_t rows = (n_e > 0 && s->e1 == n) ? s->e0 : n;
    if (s->cs) {
        if (s->cs_refresh || !s->cs_built) {
            LSS_TRY(coarse_factor(s->cs, rows, s->nm->indptr, s->nm->indices, s->nm->data, s->e0, s->e1, s->einv, s->pc));
            s->cs_built = true;
        }
        // T^T P for the CURRENT weights (the iteration carries T^T p along by linearity and needs it exactly)
        LSS_TRY(coarse_update_b(s->cs, rows, s->nm->indptr, s->nm->indices, s->nm->data, s->e0, s->e1, s->pc));
    }
    // right-hand side of the Schur system on the own rows: g' = g_g - N_gE N_EE^-1 g_E
    const double *gq = nullptr;
    if (n_e > 0) {
        LSS_CUDA(cudaMemsetAsync(s->p, 0, sizeof(double) * n, st));
        elim_scale_kernel<<<ceil_div(n_e, 256), 256, 0, st>>>(s->e0, n_e, s->einv, s->g + s->e0, 1.0, s->p);
        LSS_LAUNCH_CHECK();
        LSS_TRY(normal_apply(s->nm, s->p, s->buf, nullptr, nullptr));
        gq = s->buf;
    }
    resid_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->g, gq, s->fixed, s->e0, s->e1, nullptr, s->r);
    LSS_LAUNCH_CHECK();
    mask_own_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->owned, s->e0, s->e1, 0, s->r);
    LSS_LAUNCH_CHECK();
    LSS_TRY(shard_precond(s, s->r, s->z));
    init_p_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->r, s->z, nullptr, s->part2);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(s->pc, s->part2, VEC_GRID, 0, 0, st));
    cg_scalar_kernel<<<1, VEC_BLOCK, 0, st>>>(s->part2, VEC_GRID, s->scal, S_BNORM, 0.0, 0);
    LSS_LAUNCH_CHECK();
    s->have_weights = true;
    return LSS_OK;
}

int shard_solve(lss_solver *s, double *x, double rtol, int32_t max_iter, int32_t *iters_h, double *relres_h) {
    cudaStream_t st = s->stream;
    const int64_t n = s->n_cols;
    const int n_e = (int)(s->e1 - s->e0);
    lss_comm *pc = s->pc;
    const int P = pc->n_ranks;
    s->launches = 0;
    s->n_apply = 0;
    s->ms_a = s->ms_at = s->ms_step = 0.0;
    s->n_prof = 0;
    // mailbox layout (8-byte words)
    const int n_c = (s->cs && s->cs->ready) ? s->cs->n_c : 0;
    const int bs = s->bp->bs;
    ShardArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.rank = pc->rank; sa.n_ranks = P; sa.blk_lo = s->blk_lo; sa.blk_hi = s->blk_hi;
    sa.LB = (1 + n_c + n_e + 3) & ~3;
    sa.ZP = (long long)s->bp->n_blocks * bs;
    sa.off_lla = 0; sa.off_zf = 4 * P; sa.off_llx = (6 * P + 7) & ~7;
    sa.off_z = (sa.off_llx + 4ll * P * sa.LB + 7) & ~7ll;
    const size_t need = sizeof(double) * (size_t)(sa.off_z + 4 * sa.ZP);
    if (need > pc->box_bytes) return fail(LSS_EINVAL, "solver_solve: the mailbox holds %zu bytes, the exchange needs %zu", pc->box_bytes, need);
    for (int r = 0; r < P; ++r) sa.box[r] = reinterpret_cast<unsigned long long *>(pc->box_peer[r]);
    sa.seq = pc->seq;
    pc->seq += 2ull * (unsigned long long)(max_iter > 0 ? max_iter : 0) + 8ull;      // the same on every rank: all call solve with the same arguments
    sa.n_e = n_e; sa.e0 = s->e0;
    sa.ech = 1;
    if (n_e > 0) {
        const int n_warps = s->step_sh_grid * (VEC_BLOCK / 32);
        sa.ech = (2 * n_warps + n_e - 1) / n_e;
        if (sa.ech < 1) sa.ech = 1;
        if (sa.ech > SHARD_ECH_MAX) sa.ech = SHARD_ECH_MAX;
    }
    sa.n_indptr = s->nm->indptr; sa.n_indices = s->nm->indices; sa.n_data = s->nm->data;
    sa.einv = s->einv; sa.Bt = s->cs ? s->cs->Bt : nullptr; sa.u_e = s->u_e; sa.tk = s->tk;
    Scratch trace_buf;
    if (getenv("LSS_SHARD_TRACE")) {
        LSS_CUDA(trace_buf.alloc(sizeof(unsigned long long) * 1024 * 16, st));
        LSS_CUDA(cudaMemsetAsync(trace_buf.p, 0, sizeof(unsigned long long) * 1024 * 16, st));
        sa.trace = reinterpret_cast<unsigned long long *>(trace_buf.p);
    }
    if (!s->dlist && P > 1 && !getenv("LSS_SHARD_FULL_P")) {
        // columns referenced by the local rows of N -> slots of the packed z buffer this rank reads in phase D; per own block the
        // ranks that reference it (they get its z in phase B).  Once per solver: the pattern is fixed.
        const int64_t slots_all = (int64_t)s->bp->n_blocks * bs;
        const int n_own = s->blk_hi - s->blk_lo;
        if (s->blk_lo != (int)(((long long)pc->rank * s->bp->n_blocks) / P) || s->blk_hi != (int)(((long long)(pc->rank + 1) * s->bp->n_blocks) / P))
            return fail(LSS_EINVAL, "solver_solve: the row-sharded step expects the blocks split evenly in rank order");
        Scratch need, mask32, counter;
        LSS_CUDA(need.alloc((size_t)n, st));
        LSS_CUDA(mask32.alloc(sizeof(unsigned int) * (size_t)(n_own + 1), st));
        LSS_CUDA(counter.alloc(sizeof(int), st));
        LSS_CUDA(cudaMemsetAsync(need.p, 0, (size_t)n, st));
        LSS_CUDA(cudaMemsetAsync(mask32.p, 0, sizeof(unsigned int) * (size_t)(n_own + 1), st));
        LSS_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), st));
        shard_need_kernel<<<ceil_div(n * 32, 256), 256, 0, st>>>(n, s->nm->indptr, s->nm->indices, s->bp->col_block, s->bp->n_blocks, P, s->blk_lo,
                                                               s->blk_hi, need.as<unsigned char>(), mask32.as<unsigned int>());
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaMallocAsync((void **)&s->dlist, sizeof(int32_t) * (size_t)(slots_all + 1), st));
        shard_dlist_kernel<<<ceil_div(slots_all, 256), 256, 0, st>>>(slots_all, s->bp->cols_pad, need.as<unsigned char>(), s->dlist, counter.as<int>());
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaMallocAsync((void **)&s->blk_peers, (size_t)(n_own + 1), st));
        narrow_mask_kernel<<<ceil_div(n_own, 256), 256, 0, st>>>(n_own, mask32.as<unsigned int>(), s->blk_peers);
        LSS_LAUNCH_CHECK();
        LSS_CUDA(cudaMemcpyAsync(&s->n_dlist, counter.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaStreamSynchronize(st));
    }
    sa.blk_peers = s->dlist ? s->blk_peers : nullptr;
    sa.dlist = s->dlist; sa.n_dlist = s->n_dlist;
    if (!s->tgt) {
        // flat entry lists of phase C1: a row key per entry of P^T (own columns) and of the rows of N_E (once per solver)
        LSS_CUDA(cudaMallocAsync((void **)&s->tgt, sizeof(double) * (size_t)(n_c + n_e + 1), st));
        if (n_c > 0) {
            s->n_fr = (int)s->cs->nnz_p;
            LSS_CUDA(cudaMallocAsync((void **)&s->fr_key, sizeof(int32_t) * (size_t)(s->n_fr + 1), st));
            expand_rows_kernel<<<ceil_div((int64_t)n_c * 32, 256), 256, 0, st>>>(n_c, s->cs->pt_ptr, 0, 0, s->fr_key);
            LSS_LAUNCH_CHECK();
        }
        if (n_e > 0) {
            int32_t be[2] = {0, 0};
            LSS_CUDA(cudaMemcpyAsync(&be[0], s->nm->indptr + s->e0, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            LSS_CUDA(cudaMemcpyAsync(&be[1], s->nm->indptr + s->e1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            LSS_CUDA(cudaStreamSynchronize(st));
            s->kE0 = be[0];
            s->n_fe = be[1] - be[0];
            LSS_CUDA(cudaMallocAsync((void **)&s->fe_key, sizeof(int32_t) * (size_t)(s->n_fe + 1), st));
            expand_rows_kernel<<<ceil_div((int64_t)n_e * 32, 256), 256, 0, st>>>(n_e, s->nm->indptr + s->e0, be[0], n_c, s->fe_key);
            LSS_LAUNCH_CHECK();
        }
    }
    LSS_CUDA(cudaMemsetAsync(s->tgt, 0, sizeof(double) * (size_t)(n_c + n_e + 1), st));
    sa.fr_key = s->fr_key; sa.fe_key = s->fe_key; sa.n_fr = n_c > 0 ? s->n_fr : 0; sa.n_fe = n_e > 0 ? s->n_fe : 0; sa.kE0 = s->kE0; sa.tgt = s->tgt;
    CoarseArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.n_c = n_c;
    if (n_c > 0) {
        lss::CoarseSpace *c = s->cs;
        ca.chunks = (s->step_sh_grid * (VEC_BLOCK / 32)) / n_c;
        if (ca.chunks < 1) ca.chunks = 1;
        if (ca.chunks > CS_CHUNKS_MAX) ca.chunks = CS_CHUNKS_MAX;
        ca.pt_ptr = c->pt_ptr; ca.pt_col = c->pt_col; ca.pt_w = c->pt_w; ca.inv = c->inv;
        ca.cp_idx = reinterpret_cast<const int4 *>(c->cp_idx); ca.cp_w = reinterpret_cast<const float4 *>(c->cp_w);
        ca.rc_part = c->rc_part; ca.zc_part = c->zc_part;
    } else {
        ca.chunks = 1;
    }
    LSS_CUDA(cudaEventRecord(s->ev0, st));
    s->q_clean = false;
    LSS_CUDA(cudaMemsetAsync(s->scal + S_ERR, 0, sizeof(double), st));
    // ---- start: r = g' - S x_g on the own rows (x is complete on every rank), z, p = z everywhere, r.z, T^T p
    zero_fixed_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->fixed, x);
    LSS_LAUNCH_CHECK();
    if (n_e > 0) LSS_TRY(shard_slave(s, x, s->g));
    LSS_TRY(apply_normal(s, x, false, nullptr));
    resid_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->g, s->buf, s->fixed, s->e0, s->e1, x, s->r);
    LSS_LAUNCH_CHECK();
    mask_own_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->owned, s->e0, s->e1, 0, s->r);
    LSS_LAUNCH_CHECK();
    LSS_CUDA(cudaMemcpyAsync(s->r2, s->r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    LSS_TRY(shard_precond(s, s->r, s->z));
    init_p_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->r, s->z, s->p, s->part2);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(pc, s->part2, VEC_GRID, 0, 0, st));
    LSS_TRY(comm_allreduce(pc, s->p, n, 0, 0, st));                  // zero-padded own parts: the sum is the complete direction
    if (n_e > 0) {
        LSS_TRY(shard_slave(s, s->p, nullptr));                      // p_E = -N_EE^-1 N_Eg p_g
        LSS_CUDA(cudaMemcpyAsync(s->u_e, s->eraw, sizeof(double) * n_e, cudaMemcpyDeviceToDevice, st));
    }
    cg_scalar_kernel<<<1, VEC_BLOCK, 0, st>>>(s->part2, VEC_GRID, s->scal, S_RZ0, rtol, max_iter);
    LSS_LAUNCH_CHECK();
    s->launches += 8;
    const void *kern = step_shard_kernel_for(bs);
    int issued = 0, parity = 0, slot = 0;
    bool done = false;
    bool pending[2] = {false, false};
    int in_slot[2] = {0, 0}, base_it[2] = {0, 0};
    auto harvest = [&](int sl) {     // as in lss_solver_solve: event pairs around every kProfEvery-th real iteration
        const int real = (int)s->host_scal[sl * S_COUNT + S_ITERS] - base_it[sl];
        for (int k = 0; k < in_slot[sl] && k < real && s->profile; k += kProfEvery) {
            float a = 0.f, b = 0.f, c = 0.f;
            const int e = 4 * (sl * kBatch + k);
            cudaEventElapsedTime(&a, (*s->prof_events)[e], (*s->prof_events)[e + 1]);
            cudaEventElapsedTime(&b, (*s->prof_events)[e + 1], (*s->prof_events)[e + 2]);
            cudaEventElapsedTime(&c, (*s->prof_events)[e + 2], (*s->prof_events)[e + 3]);
            s->ms_a += a;
            s->ms_at += b;
            s->ms_step += c;
            s->n_prof += 1;
        }
        in_slot[sl] = 0;
    };
    while (!done) {
        const int todo = (max_iter - issued) < kBatch ? (max_iter - issued) : kBatch;
        for (int k = 0; k < todo; ++k) {
            const bool prof_k = s->profile && (k % kProfEvery) == 0;
            LSS_TRY(apply_normal(s, s->p, true, s->scal + S_DONE0 + parity, prof_k ? slot * kBatch + k : -1));
            double *qv = s->buf, *minv = s->minv, *pv = s->p, *zv = s->z, *part = s->part2, *scal = s->scal, *slots = s->buf + n;
            int64_t nn = n;
            int par = parity, nb = s->bp->n_blocks;
            const int *bcols = s->bp->cols_pad;
            const float *binv = s->bp->inv;
            const double *r_in = parity ? s->r2 : s->r;
            double *r_out = parity ? s->r : s->r2;
            void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &r_in, &r_out, &zv, &part, &scal, &par, &nb, &bcols, &binv, &ca, &sa};
            LSS_CUDA(cudaLaunchCooperativeKernel(kern, dim3(s->step_sh_grid), dim3(VEC_BLOCK), args, 0, st));
            ++lss::g_launches;
            s->launches += 1;
            s->q_clean = true;
            if (prof_k) LSS_CUDA(cudaEventRecord((*s->prof_events)[4 * (slot * kBatch + k) + 3], st));   // after the step + exchange
            parity ^= 1;
        }
        base_it[slot] = issued;
        issued += todo;
        in_slot[slot] = todo;
        LSS_CUDA(cudaMemcpyAsync(s->host_scal + slot * S_COUNT, s->scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaEventRecord(s->ev_batch[slot], st));
        pending[slot] = true;
        const int prev = slot ^ 1;
        if (pending[prev]) {
            LSS_CUDA(cudaEventSynchronize(s->ev_batch[prev]));
            pending[prev] = false;
            harvest(prev);
            if (s->host_scal[prev * S_COUNT + S_DONE0] != 0.0 || s->host_scal[prev * S_COUNT + S_DONE1] != 0.0) done = true;
        }
        if (!done && (issued >= max_iter || todo == 0)) {
            LSS_CUDA(cudaEventSynchronize(s->ev_batch[slot]));
            pending[slot] = false;
            done = true;
        }
        slot ^= 1;
    }
    // ---- end: back-substitution of the eliminated block, then the own parts of x become the complete x on every rank
    if (n_e > 0) LSS_TRY(shard_slave(s, x, s->g));
    {
        Scratch keep;
        if (n_e > 0) {
            LSS_CUDA(keep.alloc(sizeof(double) * n_e, st));
            LSS_CUDA(cudaMemcpyAsync(keep.p, x + s->e0, sizeof(double) * n_e, cudaMemcpyDeviceToDevice, st));
        }
        mask_own_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->owned, s->e0, s->e1, 0, x);
        LSS_LAUNCH_CHECK();
        LSS_TRY(comm_allreduce(pc, x, n, 0, 0, st));
        if (n_e > 0) LSS_CUDA(cudaMemcpyAsync(x + s->e0, keep.p, sizeof(double) * n_e, cudaMemcpyDeviceToDevice, st));     // identical on every rank
    }
    LSS_CUDA(cudaMemcpyAsync(s->host_scal, s->scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaEventRecord(s->ev1, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    s->ms_total = ms;
    harvest(0);
    harvest(1);
    const double bn = s->host_scal[S_BNORM], rzl = s->host_scal[S_RZ_LAST];
    if (iters_h) *iters_h = (int32_t)s->host_scal[S_ITERS];
    if (relres_h) *relres_h = bn > 0.0 ? sqrt(rzl / bn) : 0.0;
    if (sa.trace) {            // diagnostics: mean time between the phase stamps of CTA 0 over the recorded iterations
        std::vector<unsigned long long> h(1024 * 16);
        cudaMemcpy(h.data(), sa.trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost);
        const int n_it = (int)s->host_scal[S_ITERS] < 1024 ? (int)s->host_scal[S_ITERS] : 1024;
        double acc[13] = {0};
        int cnt = 0;
        for (int i = 8; i + 1 < n_it; ++i) {
            const unsigned long long *a = &h[(size_t)i * 16], *b = &h[(size_t)(i + 1) * 16];
            if (!a[0] || !b[0] || !a[10]) continue;
            for (int k = 0; k < 10; ++k) acc[k] += (double)(a[k + 1] - a[k]) * 1e-3;
            acc[11] += (double)a[11] * 1e-3;
            acc[12] += (double)a[12] * 1e-3;
            acc[10] += (double)(b[0] - a[10]) * 1e-3;        // end of this step -> start of the next: the product and the launch gaps
            ++cnt;
        }
        if (cnt) {
            fprintf(stderr, "[lss shard trace rank %d] %d its, us:", pc->rank, cnt);
            for (int k = 0; k < 13; ++k) fprintf(stderr, " %d:%.1f", k, acc[k] / cnt);
            fprintf(stderr, "\n");
        }
    }
    if (s->host_scal[S_ERR] != 0.0) return fail(LSS_ENCCL, "solver_solve: a peer did not answer within the exchange time-out (row-sharded CG)");
    return LSS_OK;
}

}  // namespace

extern "C" int lss_solver_set_weights(lss_solver_t *s, const double *w) {
    if (!s || !w) return fail(LSS_EINVAL, "solver_set_weights: null argument");
    if (s->shard) return shard_set_weights(s, w);
    cudaStream_t st = s->stream;
    if (s->n_rows > 0) LSS_CUDA(cudaMemcpyAsync(s->w, w, sizeof(double) * s->n_rows, cudaMemcpyDeviceToDevice, st));
    // g = A^T (w .* b)  and the Jacobi diagonal diag(A^T W A)
    if (s->n_rows > 0) {
        mul_kernel<<<stride_grid(s->n_rows, VEC_BLOCK), VEC_BLOCK, 0, st>>>(s->n_rows, s->w, s->rhs, s->u);
        LSS_LAUNCH_CHECK();
    }
    if (s->blocks) {
        if (s->n_con > 0) {
            LSS_TRY(spmv_launch(s->n_cols, s->ct_indptr, s->ct_indices, s->ct_data, s->u + s->n_data, s->g, LSS_SPMV_PLAIN, nullptr,
                                s->c_nnz, &s->plan_ct, nullptr, st, nullptr));
            diag_kernel<<<stride_grid(s->n_cols * 32, 256, 16), 256, 0, st>>>(s->n_cols, s->ct_indptr, s->ct_indices, s->ct_data,
                                                                              s->w + s->n_data, s->diag);
            LSS_LAUNCH_CHECK();
        } else {
            LSS_CUDA(cudaMemsetAsync(s->g, 0, sizeof(double) * s->n_cols, st));
            LSS_CUDA(cudaMemsetAsync(s->diag, 0, sizeof(double) * s->n_cols, st));
        }
        LSS_TRY(ell_normal_launch(1, s->n_data, s->ell_pad, s->ell_k, s->ell_col, s->ell_val, nullptr, s->w, s->rhs, s->g, nullptr, nullptr, st));
        if (s->nm) {
            LSS_TRY(normal_update(s->nm, s->w));       // N += A^T diag(w - w_prev) A : only rows whose weight changed
            if (s->sn) LSS_TRY(lss::structured_refresh(s->sn, s->nm->data));
            LSS_TRY(normal_diag(s->nm, s->diag));
        } else {
            LSS_TRY(ell_normal_launch(2, s->n_data, s->ell_pad, s->ell_k, s->ell_col, s->ell_val, nullptr, s->w, nullptr, s->diag, nullptr, nullptr, st));
        }
    } else {
        LSS_TRY(spmv_launch(s->n_cols, s->t_indptr, s->t_indices, s->t_data, s->u, s->g, LSS_SPMV_PLAIN, nullptr, s->nnz, &s->plan_at, nullptr, st, nullptr));
        diag_kernel<<<stride_grid(s->n_cols * 32, 256, 16), 256, 0, st>>>(s->n_cols, s->t_indptr, s->t_indices, s->t_data, s->w, s->diag);
        LSS_LAUNCH_CHECK();
    }
    LSS_TRY(comm_allreduce(s->comm, s->g, s->n_cols, 0, 0, st));
    LSS_TRY(comm_allreduce(s->comm, s->diag, s->n_cols, 0, 0, st));
    jacobi_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(s->n_cols, s->diag, s->fixed, s->g, s->minv, s->part2);
    LSS_LAUNCH_CHECK();
    if (s->z) {
        // structured form: ||b||^2 = g'^T M^-1 g' with g' the right-hand side of the reduced (Schur) system
        const int64_t n = s->n_cols;
        if (s->bp) LSS_TRY(block_precond_factor(s->bp, s->nm, s->comm));
        const double *gq = nullptr;
        if (s->e1 > s->e0) {
            elim_setup_kernel<<<ceil_div(s->e1 - s->e0, 256), 256, 0, st>>>(s->e0, s->e1, s->diag, s->einv, s->minv);
            LSS_LAUNCH_CHECK();
        }
        if (s->cs && s->bp && (s->cs_refresh || !s->cs_built)) {
            // Galerkin coarse operator of the Schur complement for the current weights, inverted densely
            const int64_t rows = (s->e1 > s->e0 && s->e1 == s->n_cols) ? s->e0 : s->n_cols;
            LSS_TRY(coarse_factor(s->cs, rows, s->nm->indptr, s->nm->indices, s->nm->data, s->e0, s->e1, s->einv, s->comm));
            s->cs_built = true;
        }
        if (s->e1 > s->e0) {
            LSS_CUDA(cudaMemsetAsync(s->p, 0, sizeof(double) * n, st));
            LSS_TRY(elim_slave(s, s->p, s->g, nullptr));
            LSS_TRY(normal_apply(s->nm, s->p, s->buf, nullptr, nullptr));
            LSS_TRY(comm_allreduce(s->comm, s->buf, n, 0, 0, st));
            gq = s->buf;
        }
        resid_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->g, gq, s->fixed, s->e0, s->e1, nullptr, s->r);
        LSS_LAUNCH_CHECK();
        LSS_TRY(apply_precond(s, s->r, s->z));
        init_p_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->r, s->z, nullptr, s->part2);
        LSS_LAUNCH_CHECK();
    }
    cg_scalar_kernel<<<1, VEC_BLOCK, 0, st>>>(s->part2, VEC_GRID, s->scal, S_BNORM, 0.0, 0);
    LSS_LAUNCH_CHECK();
    s->have_weights = true;
    return LSS_OK;
}

extern "C" int lss_solver_solve(lss_solver_t *s, double *x, double rtol, int32_t max_iter, int32_t *iters_h, double *relres_h) {
    if (!s || !x) return fail(LSS_EINVAL, "solver_solve: null argument");
    if (!s->have_weights) return fail(LSS_EINVAL, "solver_solve: call lss_solver_set_weights first");
    if (s->shard) return shard_solve(s, x, rtol, max_iter, iters_h, relres_h);
    cudaStream_t st = s->stream;
    const int64_t n = s->n_cols;
    s->launches = 0;
    s->n_apply = 0;
    s->ms_a = s->ms_at = s->ms_step = 0.0;
    s->n_prof = 0;
    LSS_CUDA(cudaEventRecord(s->ev0, st));
    s->q_clean = false;
    // r = g - N x0
    zero_fixed_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->fixed, x);
    LSS_LAUNCH_CHECK();
    const bool structured = s->z != nullptr;
    const bool elim = s->e1 > s->e0;
    auto slave = [&](double *v, const double *c, const double *stop) -> int { return elim_slave(s, v, c, stop); };
    if (structured) {
        if (elim) LSS_TRY(slave(x, s->g, nullptr));
        LSS_TRY(apply_normal(s, x, false, nullptr));
        resid_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->g, s->buf, s->fixed, s->e0, s->e1, x, s->r);
        LSS_LAUNCH_CHECK();
        LSS_TRY(apply_precond(s, s->r, s->z));
        init_p_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->r, s->z, s->p, s->part2);
        LSS_LAUNCH_CHECK();
        if (elim) LSS_TRY(slave(s->p, nullptr, nullptr));
        s->launches += 2;
    } else {
        LSS_TRY(apply_normal(s, x, false, nullptr));
        cg_init_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->g, s->buf, s->minv, s->fixed, x, s->r, s->p, s->part2);
        LSS_LAUNCH_CHECK();
    }
    cg_scalar_kernel<<<1, VEC_BLOCK, 0, st>>>(s->part2, VEC_GRID, s->scal, S_RZ0, rtol, max_iter);
    LSS_LAUNCH_CHECK();
    s->launches += 3;
    const int batch = kBatch;
    int issued = 0, parity = 0, slot = 0;
    bool done = false;
    bool pending[2] = {false, false};
    int in_slot[2] = {0, 0}, base_it[2] = {0, 0};
    auto harvest = [&](int sl) {     // the batch's completion event has been waited for: its timing events are final
        // iterations enqueued after convergence are no-ops on the device: leave them out of the averages
        const int real = (int)s->host_scal[sl * S_COUNT + S_ITERS] - base_it[sl];
        for (int k = 0; k < in_slot[sl] && k < real && s->profile; k += kProfEvery) {
            float a = 0.f, b = 0.f;
            const int e = 4 * (sl * kBatch + k);
            float c = 0.f;
            cudaEventElapsedTime(&a, (*s->prof_events)[e], (*s->prof_events)[e + 1]);
            cudaEventElapsedTime(&b, (*s->prof_events)[e + 1], (*s->prof_events)[e + 2]);
            cudaEventElapsedTime(&c, (*s->prof_events)[e + 2], (*s->prof_events)[e + 3]);
            s->ms_a += a;
            s->ms_at += b;
            s->ms_step += c;
            s->n_prof += 1;
        }
        in_slot[sl] = 0;
    };
    while (!done) {
        // enqueue one batch, then look at the flag the PREVIOUS batch left behind
        const int todo = (max_iter - issued) < batch ? (max_iter - issued) : batch;
        for (int k = 0; k < todo; ++k) {
            const bool prof_k = s->profile && (k % kProfEvery) == 0;      // event pairs around every kProfEvery-th iteration only
            LSS_TRY(apply_normal(s, s->p, true, s->scal + S_DONE0 + parity, prof_k ? slot * kBatch + k : -1));
            if (structured) {
                double *qv = s->buf, *minv = s->minv, *pv = s->p, *rv = s->r, *zv = s->z, *part = s->part2, *scal = s->scal;
                int64_t nn = n;
                int par = parity, nb = s->bp ? s->bp->n_blocks : 0, bs = s->bp ? s->bp->bs : 0;
                const int *bcols = s->bp ? s->bp->cols_pad : nullptr, *cblk = s->bp ? s->bp->col_block : nullptr;
                const float *binv = s->bp ? s->bp->inv : nullptr;
                double *slots = s->buf + n;
                int n_stage = s->step_stages;
                if (s->step_rows == 2 && s->cs && s->cs->ready) {
                    const double *r_in = parity ? s->r2 : s->r;
                    double *r_out = parity ? s->r : s->r2;
                    lss::CoarseSpace *c = s->cs;
                    CoarseArgs ca;
                    ca.n_c = c->n_c;
                    ca.chunks = (s->step_ml_grid * (VEC_BLOCK / 32)) / c->n_c;
                    if (ca.chunks < 1) ca.chunks = 1;
                    if (ca.chunks > CS_CHUNKS_MAX) ca.chunks = CS_CHUNKS_MAX;
                    ca.pt_ptr = c->pt_ptr; ca.pt_col = c->pt_col; ca.pt_w = c->pt_w; ca.inv = c->inv;
                    ca.cp_idx = reinterpret_cast<const int4 *>(c->cp_idx); ca.cp_w = reinterpret_cast<const float4 *>(c->cp_w);
                    ca.rc_part = c->rc_part; ca.zc_part = c->zc_part;
                    { static int dbg = -1; if (dbg < 0) { const char *e = getenv("LSS_ML_DEBUG"); dbg = e ? atoi(e) : 0; } ca.dbg = dbg; }
                    void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &r_in, &r_out, &zv, &part, &scal, &par, &nb, &bcols, &cblk, &binv, &ca};
                    LSS_CUDA(cudaLaunchCooperativeKernel(step_ml_kernel_for(bs), dim3(s->step_ml_grid), dim3(VEC_BLOCK), args, 0, st));
                } else if (s->step_rows == 2) {
                    const double *r_in = parity ? s->r2 : s->r;
                    double *r_out = parity ? s->r : s->r2;
                    void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &r_in, &r_out, &zv, &part, &scal, &par, &nb, &bcols, &cblk, &binv};
                    LSS_CUDA(cudaLaunchCooperativeKernel(step_tile_kernel_for(bs), dim3(s->step_grid), dim3(VEC_BLOCK), args, 0, st));
                } else if (s->step_rows) {
                    const double *r_in = parity ? s->r2 : s->r;
                    double *r_out = parity ? s->r : s->r2;
                    void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &r_in, &r_out, &zv, &part, &scal, &par, &nb, &bs, &bcols, &cblk, &binv};
                    LSS_CUDA(cudaLaunchCooperativeKernel((void *)cg_step_rows_kernel, dim3(s->step_grid), dim3(VEC_BLOCK), args, 0, st));
                } else {
                    void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &rv, &zv, &part, &scal, &par, &nb, &bs, &bcols, &cblk, &binv, &n_stage};
                    LSS_CUDA(cudaLaunchCooperativeKernel((void *)cg_step_p_kernel, dim3(s->step_grid), dim3(VEC_BLOCK), args, s->step_smem, st));
                }
                ++lss::g_launches;
                s->launches += 1;
                s->q_clean = true;
                if (elim) LSS_TRY(slave(s->p, nullptr, s->scal + S_DONE0 + (parity ^ 1)));
            } else if (s->nm) {
                double *qv = s->buf, *minv = s->minv, *pv = s->p, *rv = s->r, *part = s->part2, *scal = s->scal;
                int64_t nn = n;
                int par = parity;
                void *args[] = {&nn, &qv, &minv, &pv, &x, &rv, &part, &scal, &par};
                LSS_CUDA(cudaLaunchCooperativeKernel((void *)cg_step_n_kernel, dim3(VEC_GRID), dim3(VEC_BLOCK), args, 0, st));
                ++lss::g_launches;
                s->launches += 1;
                s->q_clean = true;
            } else if (s->blocks && s->n_con == 0) {
                double *qv = s->buf, *slots = s->buf + n, *minv = s->minv, *pv = s->p, *rv = s->r, *part = s->part2, *scal = s->scal;
                int64_t nn = n;
                int par = parity;
                void *args[] = {&nn, &qv, &slots, &minv, &pv, &x, &rv, &part, &scal, &par};
                LSS_CUDA(cudaLaunchCooperativeKernel((void *)cg_step_kernel, dim3(VEC_GRID), dim3(VEC_BLOCK), args, 0, st));
                ++lss::g_launches;
                s->launches += 1;
                s->q_clean = true;
            } else {
                cg_update_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->buf, s->buf + n, s->blocks ? PQ_SLOTS : VEC_GRID, s->minv, s->p, x,
                                                                 s->r, s->part2, s->scal, parity);
                LSS_LAUNCH_CHECK();
                cg_direction_kernel<<<VEC_GRID, VEC_BLOCK, 0, st>>>(n, s->r, s->minv, s->p, s->part2, VEC_GRID, s->scal, parity);
                LSS_LAUNCH_CHECK();
                s->launches += 2;
            }
            if (prof_k) LSS_CUDA(cudaEventRecord((*s->prof_events)[4 * (slot * kBatch + k) + 3], st));   // after the vector step
            parity ^= 1;
        }
        base_it[slot] = issued;
        issued += todo;
        in_slot[slot] = todo;
        LSS_CUDA(cudaMemcpyAsync(s->host_scal + slot * S_COUNT, s->scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        LSS_CUDA(cudaEventRecord(s->ev_batch[slot], st));
        pending[slot] = true;
        const int prev = slot ^ 1;
        if (pending[prev]) {
            LSS_CUDA(cudaEventSynchronize(s->ev_batch[prev]));
            pending[prev] = false;
            harvest(prev);
            if (s->host_scal[prev * S_COUNT + S_DONE0] != 0.0 || s->host_scal[prev * S_COUNT + S_DONE1] != 0.0) done = true;
        }
        if (!done && (issued >= max_iter || todo == 0)) {
            LSS_CUDA(cudaEventSynchronize(s->ev_batch[slot]));
            pending[slot] = false;
            done = true;
        }
        slot ^= 1;
    }
    if (structured && elim) LSS_TRY(slave(x, s->g, nullptr));       // back-substitution from the final grid part
    LSS_CUDA(cudaMemcpyAsync(s->host_scal, s->scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
    LSS_CUDA(cudaEventRecord(s->ev1, st));
    LSS_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    s->ms_total = ms;
    harvest(0);
    harvest(1);
    const double bn = s->host_scal[S_BNORM], rzl = s->host_scal[S_RZ_LAST];
    if (iters_h) *iters_h = (int32_t)s->host_scal[S_ITERS];
    if (relres_h) *relres_h = bn > 0.0 ? sqrt(rzl / bn) : 0.0;
    return LSS_OK;
}

extern "C" int lss_solver_export_pattern(lss_solver_t *s, int32_t *indptr, int32_t *indices) {
    if (!s || !indptr || !indices) return fail(LSS_EINVAL, "solver_export_pattern: null argument");
    if (!s->nm) return fail(LSS_EINVAL, "solver_export_pattern: the solver holds no explicit normal matrix");
    LSS_CUDA(cudaMemcpyAsync(indptr, s->nm->indptr, sizeof(int32_t) * (s->n_cols + 1), cudaMemcpyDeviceToDevice, s->stream));
    LSS_CUDA(cudaMemcpyAsync(indices, s->nm->indices, sizeof(int32_t) * s->nm->nnz, cudaMemcpyDeviceToDevice, s->stream));
    return LSS_OK;
}

extern "C" int lss_solver_set_structured(lss_solver_t *s, const lss_structured_t *desc) {
    if (!s || !desc) return fail(LSS_EINVAL, "solver_set_structured: null argument");
    if (!s->nm || s->shard) return fail(LSS_EINVAL, "solver_set_structured: needs the explicit-normal form on one GPU");
    if (desc->rows > s->n_cols || desc->nnz > s->nm->nnz) return fail(LSS_EINVAL, "solver_set_structured: the tables do not match N");
    if (s->sn) { lss::structured_destroy(s->sn); s->sn = nullptr; }
    LSS_TRY(lss::structured_create(&s->sn, desc, s->stream));
    if (s->have_weights) LSS_TRY(lss::structured_refresh(s->sn, s->nm->data));
    return LSS_OK;
}

extern "C" int lss_solver_set_profile(lss_solver_t *s, int32_t on) {
    if (!s) return fail(LSS_EINVAL, "solver_set_profile: null solver");
    s->profile = on != 0;
    return LSS_OK;
}

extern "C" int lss_solver_info(lss_solver_t *s, int64_t *info_h) {
    if (!s || !info_h) return fail(LSS_EINVAL, "solver_info: null argument");
    info_h[0] = s->nm ? 2 : (s->blocks ? 1 : 0);
    info_h[1] = s->nm ? s->nm->nnz : 0;
    info_h[2] = s->nm ? s->nm->n_pairs : 0;
    info_h[3] = s->blocks ? s->n_data : 0;
    info_h[4] = s->bp ? s->bp->n_blocks : 0;
    info_h[5] = s->e1 - s->e0;
    info_h[6] = 0;
    // non-zeros the CG product streams per application (the rows of an eliminated trailing block are left out)
    info_h[7] = 0;
    if (s->nm) {
        const bool part = s->e1 > s->e0 && s->e1 == s->n_cols;
        if (s->sn) info_h[7] = s->sn->d.nnz;
        else if (s->nm->owned && s->nm->cp_for_rows >= 0) info_h[7] = s->nm->cp_nnz;
        else if (part && lss::normal_product_variant() >= 10 && s->nm->tma_plan.tile_row && s->nm->tma_plan.n_rows == s->e0) info_h[7] = s->nm->tma_nnz;
        else if (part && s->nm->c8_part_rows == s->e0 && lss::normal_product_variant() >= 7) info_h[7] = s->nm->c8_part_nnz;
        else if (part && s->nm->part_rows == s->e0) info_h[7] = s->nm->part_nnz;
        else info_h[7] = s->nm->nnz;
    }
    if (s->bp) {
        unsigned int nb = 0;
        cudaMemcpyAsync(&nb, s->bp->n_bad, sizeof(nb), cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
        info_h[6] = nb;
    }
    return LSS_OK;
}

extern "C" int lss_solver_stats(lss_solver_t *s, double *stats_h) {
    if (!s || !stats_h) return fail(LSS_EINVAL, "solver_stats: null argument");
    stats_h[0] = s->ms_total;
    stats_h[1] = s->ms_a;
    stats_h[2] = s->ms_at;
    stats_h[3] = (double)(s->profile ? s->n_prof : s->n_apply);   // applications behind stats_h[1], [2] when profiling
    stats_h[4] = (double)s->launches;
    stats_h[5] = (double)s->nnz;
    stats_h[6] = s->ms_step;
    stats_h[7] = 0.0;
    return LSS_OK;
}
