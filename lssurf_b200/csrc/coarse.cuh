// Coarse level of the two-level preconditioner of the explicit-normal PCG solver.
//
// Block-Jacobi (precond.cuh) resolves the strong local coupling z0 <-> dz(t) but leaves the global, bias-coupled low
// end of the Schur complement untouched: its iteration count grows with the number of grid cells across the domain.
// An exact solve on a very coarse bilinear space supplies exactly that part:
//
//     M^-1 = M_BJ^-1 + P S_c^-1 P^T,      S_c = P^T (N_gg - N_gE N_EE^-1 N_Eg) P      (Galerkin, Schur complement)
//
// P interpolates every fine unknown from at most 4 coarse unknowns (bilinear in space, identity in time); the caller
// hands P in two forms: per fine column the <= 4 (parent, weight) pairs, and P^T as CSR by coarse row (so that the
// restriction is a deterministic gather, not a scatter).  S_c is a dense n_c x n_c matrix (n_c <= 1024: a few MB that
// stay in L2); it is formed with RED atomics over the non-zeros of N, the rank-n_E bias term is subtracted densely,
// and the inverse comes from a cooperative blocked Gauss-Jordan sweep (32 x 32 pivot blocks, one grid barrier per block).
#pragma once
#include "common.cuh"

namespace lss {

constexpr int CS_MAX = 1024;      // largest coarse space (the step kernel keeps r_c / z_c in shared memory); 2048 was measured: no gain

struct CoarseSpace {
    int n_c;                      // coarse unknowns
    int64_t n_cols, nnz_p;
    int32_t *cp_idx;              // [n_cols][4] coarse parents of every fine column (-1: none)          (owned copies)
    float *cp_w;                  // [n_cols][4] their weights
    int32_t *pt_ptr, *pt_col;     // P^T as CSR by coarse row: fine columns ...
    float *pt_w;                  // ... and weights
    double *A, *A2;               // [n_c][n_c] Galerkin matrix / Gauss-Jordan ping-pong; the inverse ends in `inv`
    double *inv;                  // -> A or A2
    double *B, *Bt;               // [n_c][n_e] P^T N_gE and its transpose [n_e][n_c]
    double *rc_part, *zc_part;    // [CS_CHUNKS_MAX][n_c] partial restrictions / partial coarse solutions of the step kernel
    double *rc, *zc, *rczc;       // [n_c]
    int *bad;                     // device flag: a pivot was not positive
    int n_e;
    bool ready;                   // inverse valid for the current weights
    cudaStream_t stream;
};

constexpr int CS_CHUNKS_MAX = 8;

int coarse_create(CoarseSpace **out, int64_t n_cols, int n_c, const int32_t *cp_idx, const float *cp_w, const int32_t *pt_ptr,
                  const int32_t *pt_col, const float *pt_w, int64_t nnz_p, int n_e, cudaStream_t stream);
int coarse_destroy(CoarseSpace *cs);
// S_c from the current values of N (rows [0, n_rows) of the CSR arrays; columns [e0, e1) are the eliminated block with
// einv = 1 / N_EE), summed over ranks, then inverted.  Leaves cs->ready false when S_c is not positive definite.
int coarse_factor(CoarseSpace *cs, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, int64_t e0,
                  int64_t e1, const double *einv, lss_comm *comm);
int coarse_update_b(CoarseSpace *cs, int64_t n_rows, const int32_t *indptr, const int32_t *indices, const double *data, int64_t e0,
                    int64_t e1, lss_comm *comm);
// z += P S_c^-1 P^T r   (stand-alone kernels for the set-up path; the CG step kernel has its own fused phases)
int coarse_apply_add(CoarseSpace *cs, const double *r, double *z, lss_comm *comm = nullptr);

}  // namespace lss
