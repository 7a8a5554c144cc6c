// Block-Jacobi preconditioner on the explicit normal matrix: disjoint groups of <= 64 unknowns (one dz node column
// over all epochs plus the z0 nodes around it) whose diagonal blocks of N are inverted exactly.  Inside a block sits
// the strongest coupling of the problem -- every data point ties z0 to dz(t) at the same place -- which scalar
// Jacobi ignores.  Inverses are stored in fp32 (a preconditioner only has to be symmetric positive definite and
// fixed during a solve), applied in fp64.
#pragma once
#include "common.cuh"

namespace lss {

constexpr int BP_MAX = 64;   // largest block: two values per lane

struct BlockPrecond {
    int n_blocks, bs;            // bs: row stride of the stored blocks (max block size rounded up to 4)
    int64_t n_cols;
    int32_t *ptr, *cols;         // block b holds columns cols[ptr[b] .. ptr[b+1])   (owned copies)
    int32_t *cols_pad;           // [n_blocks][bs] the same members, padded with -1
    int32_t *col_block, *col_local;   // per column: block id (-1: scalar Jacobi) and position inside the block
    double *dense;               // [n_blocks][bs][bs] gathered blocks (summed over ranks before inversion)
    float *inv;                  // [n_blocks][bs][bs] symmetric inverses
    unsigned int *n_bad;         // blocks that were not positive definite (fell back to their diagonal)
    cudaStream_t stream;
};

int block_precond_create(BlockPrecond **out, int64_t n_cols, int n_blocks, const int32_t *ptr_dev, const int32_t *cols_dev,
                         int max_size, cudaStream_t stream);
int block_precond_destroy(BlockPrecond *bp);
// gather the diagonal blocks from N, sum over ranks, invert
int block_precond_factor(BlockPrecond *bp, const NormalMatrix *nm, lss_comm *comm);
// z = M^-1 r for all columns (blocks: exact block inverse; others: z = minv .* r)
int block_precond_apply(const BlockPrecond *bp, int64_t n_cols, const double *r, const double *minv, double *z, cudaStream_t stream);

// Warp-cooperative z = Binv r over one block.  Lane l holds members l and l + 32 (r0, r1) and receives z for the same
// members.  The stored block is bs x bs with zero padding, so the loop bound does not depend on the member count and
// every load address is known up front: BP_CH rows of the (symmetric) inverse are requested at a time, the first
// group by the caller before it touches anything else.
constexpr int BP_CH = 12;
struct BlockRows { float a0[BP_CH], a1[BP_CH]; };

__device__ __forceinline__ void block_rows_load(BlockRows &R, const float *__restrict__ inv, int bs, int j0, uint64_t keep) {
    const unsigned lane = threadIdx.x & 31u;
    const bool h0 = (int)lane < bs, h1 = (int)lane + 32 < bs;
#pragma unroll
    for (int t = 0; t < BP_CH; ++t) {
        const bool in = j0 + t < bs;
        R.a0[t] = (in && h0) ? ldg_keep(inv + (j0 + t) * bs + lane, keep) : 0.f;
        R.a1[t] = (in && h1) ? ldg_keep(inv + (j0 + t) * bs + 32 + lane, keep) : 0.f;
    }
}

__device__ __forceinline__ void block_apply_warp(const float *__restrict__ inv, int bs, BlockRows &R, double r0, double r1, double &z0,
                                                 double &z1, uint64_t keep) {
    z0 = 0.0;
    z1 = 0.0;
    for (int j0 = 0; j0 < bs; j0 += BP_CH) {
        BlockRows nxt;
        if (j0 + BP_CH < bs) block_rows_load(nxt, inv, bs, j0 + BP_CH, keep);      // in flight during the sums below
#pragma unroll
        for (int t = 0; t < BP_CH; ++t) {
            const int j = j0 + t;
            const double rj = __shfl_sync(0xffffffffu, (j < 32) ? r0 : r1, j & 31);
            z0 += (double)R.a0[t] * rj;
            z1 += (double)R.a1[t] * rj;
        }
        if (j0 + BP_CH < bs) R = nxt;
    }
}

}  // namespace lss
