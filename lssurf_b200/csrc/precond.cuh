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
int block_precond_factor(BlockPrecond *bp, const NormalMatrix *nm, lss_comm *comm, int b_lo = 0, int b_hi = -1);
// z = M^-1 r for all columns (blocks: exact block inverse; others: z = minv .* r)
int block_precond_apply(const BlockPrecond *bp, int64_t n_cols, const double *r, const double *minv, double *z, cudaStream_t stream,
                        int b_lo = 0, int b_hi = -1);

// Warp-cooperative z = Binv r over one block.  Lane l holds members l and l + 32 (r0, r1) and receives z for the same
// members.  The stored block is bs x bs with zero padding and symmetric, so nothing depends on the member count and
// every address is known before any vector entry is: the caller requests the block (BlockRegs) first, then gathers
// the vector entries, then calls block_apply_warp -- one memory round trip per block.
//   members 0..31 : z_l = sum_j inv[j][l] r_j           lane l keeps column l of rows 0..BP_ROWS-1 in registers
//   members >= 32 : z_i = <inv[i][:], r>                  a warp reduction per member, BP_TAIL rows in registers
constexpr int BP_ROWS = 40, BP_TAIL = 4;
struct BlockRegs { float a[BP_ROWS], t0[BP_TAIL], t1[BP_TAIL]; };

__device__ __forceinline__ void block_regs_load(BlockRegs &R, const float *__restrict__ inv, int bs, uint64_t keep) {
    const unsigned lane = threadIdx.x & 31u;
    const bool h0 = (int)lane < bs, h1 = (int)lane + 32 < bs;
#pragma unroll
    for (int t = 0; t < BP_ROWS; ++t) R.a[t] = (t < bs && h0) ? ldg_keep(inv + t * bs + lane, keep) : 0.f;
#pragma unroll
    for (int t = 0; t < BP_TAIL; ++t) {
        R.t0[t] = (32 + t < bs && h0) ? ldg_keep(inv + (32 + t) * bs + lane, keep) : 0.f;
        R.t1[t] = (32 + t < bs && h1) ? ldg_keep(inv + (32 + t) * bs + 32 + lane, keep) : 0.f;
    }
}

__device__ __forceinline__ void block_apply_warp(const float *__restrict__ inv, int bs, const BlockRegs &R, double r0, double r1,
                                                 double &z0, double &z1, uint64_t keep) {
    const unsigned lane = threadIdx.x & 31u;
    z0 = 0.0;
    z1 = 0.0;
#pragma unroll
    for (int j = 0; j < BP_ROWS; ++j) {
        const double rj = __shfl_sync(0xffffffffu, (j < 32) ? r0 : r1, j & 31);
        z0 += (double)R.a[j] * rj;
    }
    for (int j = BP_ROWS; j < bs; ++j) {            // blocks beyond BP_ROWS members: remaining rows straight from memory
        const double rj = __shfl_sync(0xffffffffu, r1, j & 31);
        z0 += (double)((int)lane < bs ? ldg_keep(inv + j * bs + lane, keep) : 0.f) * rj;
    }
#pragma unroll
    for (int t = 0; t < BP_TAIL; ++t) {
        const double d = warp_sum((double)R.t0[t] * r0 + (double)R.t1[t] * r1);
        if ((int)lane == t) z1 = d;
    }
    for (int i = 32 + BP_TAIL; i < bs; ++i) {
        const float b0 = (int)lane < bs ? ldg_keep(inv + i * bs + lane, keep) : 0.f;
        const float b1 = (int)lane + 32 < bs ? ldg_keep(inv + i * bs + 32 + lane, keep) : 0.f;
        const double d = warp_sum((double)b0 * r0 + (double)b1 * r1);
        if ((int)lane == i - 32) z1 = d;
    }
}

}  // namespace lss
