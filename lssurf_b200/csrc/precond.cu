// Block-Jacobi preconditioner: set-up (gather diagonal blocks of N, invert) and the stand-alone apply kernel.
#include "precond.cuh"

namespace lss {

__global__ void __launch_bounds__(256)
bp_index_kernel(int n_blocks, int bs, const int *__restrict__ ptr, const int *__restrict__ cols, int *__restrict__ col_block,
                int *__restrict__ col_local, int *__restrict__ cols_pad) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_blocks) return;
    const int b0 = ptr[warp], b1 = ptr[warp + 1];
    for (int k = b0 + (int)lane_id(); k < b1; k += 32) {
        col_block[cols[k]] = warp;
        col_local[cols[k]] = k - b0;
    }
    for (int l = (int)lane_id(); l < bs; l += 32) cols_pad[(size_t)warp * bs + l] = b0 + l < b1 ? cols[b0 + l] : -1;
}

// dense[b][li][lj] = N[cols[li], cols[lj]]   (one warp per block; the rest of the block stays zero)
__global__ void __launch_bounds__(256)
bp_gather_kernel(int n_blocks, int bs, const int *__restrict__ ptr, const int *__restrict__ cols, const int *__restrict__ col_block,
                 const int *__restrict__ col_local, const int *__restrict__ n_indptr, const int *__restrict__ n_indices,
                 const double *__restrict__ n_data, double *__restrict__ dense, int b_first) {
    const int b = b_first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= n_blocks) return;
    const unsigned lane = lane_id();
    double *D = dense + (size_t)b * bs * bs;
    for (int i = (int)lane; i < bs * bs; i += 32) D[i] = 0.0;
    __syncwarp();
    const int b0 = ptr[b], m = ptr[b + 1] - b0;
    for (int li = 0; li < m; ++li) {
        const int row = cols[b0 + li];
        const int e = n_indptr[row + 1];
        for (int k = n_indptr[row] + (int)lane; k < e; k += 32) {
            const int j = n_indices[k];
            if (col_block[j] == b) D[li * bs + col_local[j]] = n_data[k];
        }
    }
}

// In-place Gauss-Jordan inverse of the SPD block (no pivoting), one warp per block, block staged in shared memory.
// Not positive definite (or non-finite) -> inverse of the diagonal.  The result is symmetrised and rounded to fp32.
__global__ void __launch_bounds__(32)
bp_invert_kernel(int n_blocks, int bs, const int *__restrict__ ptr, const double *__restrict__ dense, float *__restrict__ inv,
                 unsigned int *__restrict__ n_bad, int b_first) {
    extern __shared__ double sh[];      // [bs][bs + 1] work + [bs] original diagonal
    const int b = b_first + blockIdx.x;
    if (b >= n_blocks) return;
    const int ld = bs + 1;
    double *B = sh, *dg = sh + bs * ld;
    const unsigned lane = lane_id();
    const int m = ptr[b + 1] - ptr[b];
    const double *D = dense + (size_t)b * bs * bs;
    for (int i = (int)lane; i < m * m; i += 32) {
        const int r = i / m, c = i - r * m;
        B[r * ld + c] = D[r * bs + c];
    }
    __syncwarp();
    for (int i = (int)lane; i < m; i += 32) dg[i] = B[i * ld + i];
    __syncwarp();
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        const double piv = B[k * ld + k];
        if (!(piv > 0.0) || !(piv < 1e300)) { ok = false; break; }    // uniform: every lane reads the same value
        const double ip = 1.0 / piv;
        __syncwarp();
        for (int j = (int)lane; j < m; j += 32) B[k * ld + j] = (j == k) ? ip : B[k * ld + j] * ip;
        __syncwarp();
        for (int i = 0; i < m; ++i) {
            if (i == k) continue;
            const double f = B[i * ld + k];
            __syncwarp();
            for (int j = (int)lane; j < m; j += 32) B[i * ld + j] = (j == k) ? -f * ip : B[i * ld + j] - f * B[k * ld + j];
        }
        __syncwarp();
    }
    __syncwarp();
    float *O = inv + (size_t)b * bs * bs;
    if (ok) {
        for (int i = (int)lane; i < bs * bs; i += 32) {
            const int r = i / bs, c = i - r * bs;
            O[i] = (r < m && c < m) ? (float)(0.5 * (B[r * ld + c] + B[c * ld + r])) : 0.f;
        }
    } else {
        for (int i = (int)lane; i < bs * bs; i += 32) {
            const int r = i / bs, c = i - r * bs;
            O[i] = (r == c && r < m && dg[r] > 0.0) ? (float)(1.0 / dg[r]) : 0.f;
        }
        if (lane == 0) atomicAdd(n_bad, 1u);
    }
}

// z = M^-1 r: warps take blocks round robin, then every thread sweeps the columns outside blocks
__global__ void __launch_bounds__(256)
bp_apply_kernel(int n_blocks, int bs, const int *__restrict__ ptr, const int *__restrict__ cols, const int *__restrict__ col_block,
                const float *__restrict__ inv, int64_t n_cols, const double *__restrict__ r, const double *__restrict__ minv,
                double *__restrict__ z, int b_first) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const unsigned lane = lane_id();
    const uint64_t keep = l2_policy_evict_last();
    for (int b = b_first + gw; b < n_blocks; b += nw) {
        const int b0 = ptr[b], m = ptr[b + 1] - b0;
        const int c0 = (int)lane < m ? cols[b0 + lane] : -1, c1 = (int)lane + 32 < m ? cols[b0 + 32 + lane] : -1;
        const double r0 = c0 >= 0 ? r[c0] : 0.0, r1 = c1 >= 0 ? r[c1] : 0.0;
        double z0, z1;
        BlockRegs R;
        block_regs_load(R, inv + (size_t)b * bs * bs, bs, keep);
        block_apply_warp(inv + (size_t)b * bs * bs, bs, R, r0, r1, z0, z1, keep);
        if (c0 >= 0) z[c0] = z0;
        if (c1 >= 0) z[c1] = z1;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += (int64_t)gridDim.x * blockDim.x)
        if (col_block[i] < 0) z[i] = minv[i] * r[i];
}

__global__ void __launch_bounds__(256) bp_fill_i32_kernel(int *p, int64_t n, int v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int block_precond_destroy(BlockPrecond *bp) {
    if (!bp) return LSS_OK;
    cudaStream_t st = bp->stream;
    void *ptrs[] = {bp->ptr, bp->cols, bp->cols_pad, bp->col_block, bp->col_local, bp->dense, bp->inv, bp->n_bad};
    for (void *p : ptrs)
        if (p) cudaFreeAsync(p, st);
    delete bp;
    return LSS_OK;
}

int block_precond_create(BlockPrecond **out, int64_t n_cols, int n_blocks, const int32_t *ptr_dev, const int32_t *cols_dev,
                         int max_size, cudaStream_t st) {
    if (n_blocks <= 0 || max_size < 1 || max_size > BP_MAX) return fail(LSS_EINVAL, "block preconditioner: block sizes must be in 1..%d", BP_MAX);
    BlockPrecond *bp = new BlockPrecond();
    memset(bp, 0, sizeof(*bp));
    bp->n_blocks = n_blocks; bp->bs = (max_size + 3) / 4 * 4; bp->n_cols = n_cols; bp->stream = st;
    Scratch::configure_pool();
    int n_members = 0;
    if (cudaMemcpyAsync(&n_members, ptr_dev + n_blocks, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { delete bp; return fail(LSS_ECUDA, "block preconditioner: cannot read the block pointers"); }
    const size_t blk = (size_t)n_blocks * bp->bs * bp->bs;
    bool ok = cudaMallocAsync((void **)&bp->ptr, sizeof(int) * (n_blocks + 1), st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->cols, sizeof(int) * (size_t)(n_members > 0 ? n_members : 1), st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->cols_pad, sizeof(int) * (size_t)n_blocks * bp->bs, st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->col_block, sizeof(int) * n_cols, st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->col_local, sizeof(int) * n_cols, st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->dense, sizeof(double) * blk, st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->inv, sizeof(float) * blk, st) == cudaSuccess &&
              cudaMallocAsync((void **)&bp->n_bad, sizeof(unsigned int), st) == cudaSuccess;
    if (!ok) { block_precond_destroy(bp); return fail(LSS_ECUDA, "block preconditioner: cudaMallocAsync failed"); }
    cudaMemcpyAsync(bp->ptr, ptr_dev, sizeof(int) * (n_blocks + 1), cudaMemcpyDeviceToDevice, st);
    if (n_members > 0) cudaMemcpyAsync(bp->cols, cols_dev, sizeof(int) * (size_t)n_members, cudaMemcpyDeviceToDevice, st);
    cudaMemsetAsync(bp->n_bad, 0, sizeof(unsigned int), st);
    bp_fill_i32_kernel<<<stride_grid(n_cols, 256), 256, 0, st>>>(bp->col_block, n_cols, -1);
    bp_fill_i32_kernel<<<stride_grid(n_cols, 256), 256, 0, st>>>(bp->col_local, n_cols, 0);
    bp_index_kernel<<<ceil_div((int64_t)n_blocks * 32, 256), 256, 0, st>>>(n_blocks, bp->bs, bp->ptr, bp->cols, bp->col_block, bp->col_local,
                                                                           bp->cols_pad);
    g_launches += 3;
    if (cudaGetLastError() != cudaSuccess) { block_precond_destroy(bp); return fail(LSS_ECUDA, "block preconditioner: launch failed"); }
    *out = bp;
    return LSS_OK;
}

int block_precond_factor(BlockPrecond *bp, const NormalMatrix *nm, lss_comm *comm, int b_lo, int b_hi) {
    cudaStream_t st = bp->stream;
    if (b_hi < 0) b_hi = bp->n_blocks;          // blocks [b_lo, b_hi): a rank of a row-sharded N factors its own blocks only
    const int nb = b_hi - b_lo;
    if (nb <= 0) return LSS_OK;
    const int grid = ceil_div((int64_t)nb * 32, 256);
    bp_gather_kernel<<<grid, 256, 0, st>>>(b_hi, bp->bs, bp->ptr, bp->cols, bp->col_block, bp->col_local, nm->indptr, nm->indices,
                                           nm->data, bp->dense, b_lo);
    LSS_LAUNCH_CHECK();
    LSS_TRY(comm_allreduce(comm, bp->dense, (int64_t)bp->n_blocks * bp->bs * bp->bs, 0, 0, st));
    const size_t smem = sizeof(double) * ((size_t)bp->bs * (bp->bs + 1) + bp->bs);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        LSS_CUDA(cudaFuncSetAttribute(bp_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    bp_invert_kernel<<<nb, 32, smem, st>>>(b_hi, bp->bs, bp->ptr, bp->dense, bp->inv, bp->n_bad, b_lo);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

int block_precond_apply(const BlockPrecond *bp, int64_t n_cols, const double *r, const double *minv, double *z, cudaStream_t st, int b_lo,
                        int b_hi) {
    if (b_hi < 0) b_hi = bp->n_blocks;
    bp_apply_kernel<<<2 * kSMs, 256, 0, st>>>(b_hi, bp->bs, bp->ptr, bp->cols, bp->col_block, bp->inv, n_cols, r, minv, z, b_lo);
    LSS_LAUNCH_CHECK();
    return LSS_OK;
}

}  // namespace lss
