/*
 * lssurf_b200 -- C ABI of the B200-native smooth_fit hot path.
 *
 * The reference (SmithB/LSsurf) is pure Python with no FFI layer; its boundary for this path is
 * its Python API.  These entry points are what a ctypes binding inside the reference's
 * lin_op.py / smooth_fit.py / calc_sigma_extra.py / RDE.py would call (INTEGRATION.md shows the
 * stubs).  Each function cites the reference lines it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _h (host pointer);
 *   - index arrays are int32, values fp64, sizes int64; `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success, a negative LSS_E* code otherwise; lss_last_error()
 *     returns a thread-local message for the last failure;
 *   - functions with *_h outputs synchronise `stream` before returning, all others are
 *     asynchronous on `stream`;
 *   - nothing here falls back to the CPU: without a CUDA device every compute call fails.
 */
#ifndef LSSURF_B200_H
#define LSSURF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSS_OK 0
#define LSS_EINVAL -1   /* bad argument */
#define LSS_ECUDA -2    /* CUDA runtime error */
#define LSS_ENCCL -3    /* NCCL error / NCCL not loadable */
#define LSS_EBOUNDS -4  /* points out of bounds with bounds_error set (lin_op.py:202-203) */
#define LSS_ERANGE -5   /* size exceeds the int32 index range of the CSR format */

#define LSS_MAX_DIMS 3

/* Regular node grid: the hot-path subset of fd_grid (LSsurf/fd_grid.py:61-80). */
typedef struct lss_grid {
    int32_t ndim;                 /* 1..3, coordinate order (y, x, t)            fd_grid.py:25 */
    int64_t shape[LSS_MAX_DIMS];  /* nodes per dimension                         fd_grid.py:61 */
    int64_t stride[LSS_MAX_DIMS]; /* C-order strides in columns                  fd_grid.py:67 */
    double lo[LSS_MAX_DIMS];      /* bds[d][0] = first node centre               fd_grid.py:63 */
    double hi[LSS_MAX_DIMS];      /* bds[d][1] = last node centre                fd_grid.py:63 */
    double delta[LSS_MAX_DIMS];   /* node spacing                                fd_grid.py:64 */
    int64_t col_0;                /* first global column of the grid             fd_grid.py:68 */
} lss_grid_t;

int lss_version(void);
const char *lss_last_error(void);
/* Number of CUDA devices visible (0 => every compute entry point fails with LSS_ECUDA). */
int lss_device_count(void);
/* Kernels launched by this library in the calling process so far (for bench.py's gpu_launches). */
int64_t lss_launch_count(void);

/* ------------------------------------------------------------------ design-matrix assembly */

/* lin_op.interp_mtx (lin_op.py:163-247) with fd_grid.validate_pts / float_sub / cell_sub_for_pts /
 * global_ind (fd_grid.py:90-145).  One row per input point i (row index = row_0 + i); 2^ndim
 * triplets per row in np.mgrid corner order (dimension 0 slowest).  Weights are formed as
 * 1.0 * prod_d (frac_d | 1 - frac_d) for ascending d with individually rounded fp64 operations
 * (no FMA contraction), so they are bit-identical to NumPy's.  Points outside the closed grid
 * bounds (or with a non-finite first coordinate) get valid[i] = 0, zero weights and column col_0.
 * pts_h: host array of ndim device pointers (y, x[, t]).  n_invalid_h: host out. */
int lss_interp_coo(const lss_grid_t *grid, const double *const *pts_h, int64_t n, int64_t row_0,
                   int32_t *rows, int32_t *cols, double *vals, uint8_t *valid, int64_t *n_invalid_h,
                   void *stream);

/* Number of equations lin_op.diff_op produces (lin_op.py:88-103, 109).  Host-only helper.
 * offsets_h: [ndim][n_off] int32 (row-major).  first_h/count_h (optional, [ndim]) receive the first
 * valid centre subscript and the number of valid centres per dimension. */
int lss_stencil_count(const lss_grid_t *grid, int32_t n_off, const int32_t *offsets_h, int32_t valid_only,
                      int64_t *n_eq_h, int64_t *first_h, int64_t *count_h);

/* lin_op.diff_op (lin_op.py:80-132): one row per valid centre node in C order, n_off triplets per
 * row; ind0[i] = global column of the centre (lin_op.py:128).  With valid_only == 0 every node is a
 * centre and out-of-grid offsets get column 0 and value 0 (lin_op.py:118-121). */
int lss_stencil_coo(const lss_grid_t *grid, int32_t n_off, const int32_t *offsets_h, const double *coeffs_h,
                    int64_t row_0, int32_t valid_only, int64_t n_eq, int32_t *rows, int32_t *cols,
                    double *vals, int64_t *ind0, void *stream);

/* max(r[v != 0]) (lin_op.py:750-752), -1 if there is none; and the count of v != 0. */
int lss_coo_extent(int64_t nnz, const int32_t *rows, const double *vals, int64_t *max_row_h,
                   int64_t *n_nonzero_h, void *stream);

/* lin_op.toCSR (lin_op.py:745-753) == scipy coo_tocsr + sum_duplicates: drop v == 0, stable sort by
 * (row, col), sum duplicates left to right, build indptr.  Output is canonical CSR (sorted
 * indices, no duplicates).  indices/data must hold nnz entries; *nnz_out_h receives the CSR nnz. */
int lss_coo_to_csr(int64_t nnz, const int32_t *rows, const int32_t *cols, const double *vals,
                   int64_t n_rows, int64_t n_cols, int32_t *indptr, int32_t *indices, double *data,
                   int64_t *nnz_out_h, void *stream);

/* CSR of the transpose (scipy csr.T.tocsr(); used for G^T products, smooth_fit.py:142, 339-345).
 * Entries of each output row keep ascending source-row order. */
int lss_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *indptr,
                      const int32_t *indices, const double *data, int32_t *t_indptr, int32_t *t_indices,
                      double *t_data, void *stream);

/* Spatial sort keys (coarse (y, x) cell, fine (y, x) cell, coarse time cell; row-major) for the solver's INTERNAL
 * ordering of the points: neighbours on the ground become neighbours in memory.  pts_h: host array of device
 * pointers (y, x[, t]).  *key_bits_h receives the number of significant key bits (for lss_argsort_u64). */
int lss_cell_keys(const lss_grid_t *fine, const lss_grid_t *coarse, const double *const *pts_h, int64_t n,
                  uint64_t *keys, int32_t *key_bits_h, void *stream);

/* Stable argsort of int64 keys in [0, 2^key_bits) (used for the internal spatial ordering of points). */
int lss_argsort_u64(int64_t n, const uint64_t *keys, int32_t key_bits, int32_t *order, void *stream);

/* Dense rank of every value among the sorted distinct values of x (what np.unique(return_inverse=True)
 * returns; NaN is first replaced by nan_value, as assign_bias_ID does with -9999, bias_functions.py:55-60).
 * unique_h (host, optional, capacity unique_cap) receives the distinct values in ascending order. */
int lss_dense_rank_f64(const double *x, int64_t n, double nan_value, int32_t *rank, double *unique_h,
                       int64_t unique_cap, int64_t *n_unique_h, void *stream);

/* out[g] = np.nanmedian(vals[ids == g]) for g in [0, n_groups) (bias_functions.py:79); NaN for an empty group. */
int lss_group_median(const int32_t *ids, const double *vals, int64_t n, int32_t n_groups, double *out,
                     void *stream);

/* ------------------------------------------------------------------ CSR products */

#define LSS_SPMV_PLAIN 0     /* y = A x                                                        */
#define LSS_SPMV_SCALE 1     /* y = s .* (A x)        (row weights, smooth_fit.py:129, 142)     */
#define LSS_SPMV_RESIDUAL 2  /* y = s - A x           (r_data = z - G m0, smooth_fit.py:146)    */
#define LSS_SPMV_ACCUM 3     /* y += A x with atomics (internal: products that share y)         */

/* scipy csr_matvec replacement (smooth_fit.py:142, 146, 326-345).  nnz-balanced tiles, deterministic
 * (no atomics).  `s` is read only for modes 1 and 2.  rows [row_begin, row_end) only. */
int lss_spmv(int64_t n_rows, int64_t n_cols, const int32_t *indptr, const int32_t *indices, const double *data,
             const double *x, double *y, int32_t mode, const double *s, void *stream);

/* q = A^T u computed from the CSR rows of A by scattering (fp64 RED atomics; summation order is run dependent).
 * For occasional transpose products (parse_model's count / misfit maps, smooth_fit.py:339-345) where building
 * A^T with lss_csr_transpose is not worth it. */
int lss_csr_rmatvec(int64_t n_rows, int64_t n_cols, const int32_t *indptr, const int32_t *indices, const double *data,
                    const double *u, double *q, void *stream);

/* Same product with the structure-dependent set-up (first row of every non-zero tile, carry buffers) done
 * once: for matrices that are multiplied many times (the edit loop's residual and G^T products). */
typedef struct lss_spmv_plan lss_spmv_plan_t;
int lss_spmv_plan_create(lss_spmv_plan_t **plan, int64_t n_rows, const int32_t *indptr, void *stream);
int lss_spmv_plan_destroy(lss_spmv_plan_t *plan);
int lss_spmv_planned(lss_spmv_plan_t *plan, int64_t n_rows, const int32_t *indptr, const int32_t *indices,
                     const double *data, const double *x, double *y, int32_t mode, const double *s, void *stream);

/* Tuning knobs.  "spmv_variant": CTA shape of the CSR product kernel (index into a fixed table of
 * (threads, non-zeros per thread) pairs); affects speed only, never results.  "explicit_normal" (default 1):
 * lss_solver_create_blocks forms N = A^T W A explicitly (see lss_normal_create) instead of applying A twice.
 * "normal_product" (default 0, 0..11): kernel of the CG product q = N p -- 0 nnz-tiled, 1/5/6 register rows, 7..9 padded
 * chunks of 8, 10 TMA bulk-copy ring, 11 TMA ring + per-tile column dictionary (profiles/product_variants_r01.md).
 * "step_kernel" (default 2): CG vector step -- 0 warp per preconditioner block, 1 thread per block row, 2 CTA tile.
 * Both change speed and the order of floating-point sums only. */
int lss_set_option(const char *name, int32_t value);

/* lin_op.toCSR (lin_op.py:745-753) for operators whose triplets come in row-regular blocks -- the sum of interpolation
 * operators and per-row bias entries smooth_fit builds for its data rows (smooth_fit.py:461-513): block b holds k_h[b]
 * entries for EVERY row (row-major, cols_h[b] / vals_h[b] are host arrays of device pointers, 1..8 blocks).  Row r of the
 * result takes its entries block after block with zero values dropped (toCSR drops them, lin_op.py:748); that is the
 * canonical CSR whenever the kept columns of each row ascend strictly, which holds for np.mgrid corner order on grids
 * with ascending column ranges.  No sort: two streaming passes.  *unordered_h receives the number of rows that are
 * not ascending (or hold a column >= n_cols); when it is not 0 nothing is written and the caller uses lss_coo_to_csr.
 * indices / data need room for n_rows * sum(k_h) entries. */
int lss_blocks_to_csr(int64_t n_rows, int64_t n_cols, int32_t n_blocks, const int32_t *k_h, const int32_t *const *cols_h,
                      const double *const *vals_h, int32_t *indptr, int32_t *indices, double *data, int64_t *nnz_out_h,
                      int64_t *unordered_h, void *stream);

/* ------------------------------------------------------------------ explicit normal matrix */

/* N = A^T diag(w) A for a matrix A whose rows hold at most row_len <= 24 entries (canonical CSR in; the interpolation
 * rows of lin_op.interp_mtx and the stencil rows of lin_op.diff_op all qualify).  This is the matrix SuiteSparseQR
 * factorises implicitly in sparseqr.solve (smooth_fit.py:142); the PCG solver multiplies by it instead.
 * lss_normal_create computes the sparsity pattern (row i of N = union of the column sets of the rows of A holding
 * column i; rows / columns listed in fixed_cols are left empty) with all values zero; *nnz_h receives nnz(N).
 * lss_normal_update(w) adds A^T diag(w - w_prev) A, where w_prev is the weight vector of the previous call (zero at
 * first): only rows whose weight changed are touched.  Values are accumulated with fp64 RED atomics (summation
 * order, hence the last bits, are run dependent).  lss_normal_export copies canonical CSR arrays (sorted indices)
 * out: indptr[n_cols + 1], indices[nnz], data[nnz].  lss_normal_apply: y = N x with the tiled CSR kernel. */
typedef struct lss_normal lss_normal_t;
int lss_normal_create(lss_normal_t **normal, int64_t n_rows, int64_t n_cols, int32_t row_len, const int32_t *indptr,
                      const int32_t *indices, const double *data, const int32_t *fixed_cols, int64_t n_fixed,
                      int64_t *nnz_h, void *stream);
int lss_normal_update(lss_normal_t *normal, const double *w);
int lss_normal_export(lss_normal_t *normal, int32_t *indptr, int32_t *indices, double *data);
int lss_normal_apply(lss_normal_t *normal, const double *x, double *y);
int lss_normal_destroy(lss_normal_t *normal);

/* ------------------------------------------------------------------ multi-GPU communicator */

typedef struct lss_comm lss_comm_t;
#define LSS_UNIQUE_ID_BYTES 128
/* ncclGetUniqueId: rank 0 calls this and ships the bytes to the other ranks (any host transport). */
int lss_comm_unique_id(char *id_h);
int lss_comm_create(lss_comm_t **comm, const char *id_h, int32_t rank, int32_t n_ranks);
int lss_comm_destroy(lss_comm_t *comm);
int lss_comm_allreduce_f64(lss_comm_t *comm, double *buf, int64_t n, void *stream); /* in-place sum */
/* Peer-to-peer mailbox for the row-sharded PCG (lss_solver_create_sharded): every rank allocates `bytes` of device memory
 * (zeroed) and gets a 64-byte CUDA IPC handle for it; the handles of all ranks, concatenated in rank order (any host
 * transport), are then opened so that kernels of this rank can store straight into the peers' boxes over NVLink.
 * With one rank the box is local and handles_h may be NULL.  lss_comm_box_bytes: size of the open box (0: none). */
#define LSS_IPC_HANDLE_BYTES 64
int lss_comm_box_alloc(lss_comm_t *comm, int64_t bytes, char *handle_h);
int lss_comm_box_open(lss_comm_t *comm, const char *handles_h);
int lss_comm_box_bytes(lss_comm_t *comm, int64_t *bytes_h);

/* ------------------------------------------------------------------ normal-equation PCG solve */

typedef struct lss_solver lss_solver_t;

/* Replaces sparseqr.solve on the row-selected, 1/sigma-scaled system (smooth_fit.py:123-142) by a
 * fp64 preconditioned CG on (A^T W A) x = A^T W b, where A = [G; C] is this rank's block
 * of data rows stacked on its share of the constraint rows, W = diag(w), b = [z; prior].  This entry point is the GENERIC
 * form (rows of any length): A and A^T as CSR, two products per iteration.  smooth_fit uses lss_solver_create_blocks /
 * lss_solver_create_sharded below, which form the normal matrix explicitly and multiply by it once per iteration.
 * Columns listed in fixed_cols (the reference-epoch columns removed by Ip_c,
 * constraint_functions.py:128-151) are held at exactly 0.  With a communicator (this form only) the products
 * A^T W (A p) and the dot products are summed over ranks with one NCCL all-reduce per iteration; the row-sharded
 * normal-matrix form (lss_solver_create_sharded) has no collective inside the iteration.
 * The solver borrows every array (the caller keeps them alive until lss_solver_destroy). */
int lss_solver_create(lss_solver_t **solver, int64_t n_rows, int64_t n_cols, const int32_t *indptr,
                      const int32_t *indices, const double *data, const int32_t *t_indptr,
                      const int32_t *t_indices, const double *t_data, const double *rhs,
                      const int32_t *fixed_cols, int64_t n_fixed, lss_comm_t *comm, void *stream);
int lss_solver_destroy(lss_solver_t *solver);
/* Block form of the same solver, the one smooth_fit uses: A = [G; C] is given as its two row blocks.
 *   G (n_data rows): every row holds at most row_len <= 24 entries (4 z0 + 8 dz corners + bias, slope and sensor-grid columns).  It is copied
 *     column-major (ELL) and applied by ONE fused kernel q += G^T (w .* (G p)) -- a single pass over G per CG
 *     iteration, contributions of neighbouring rows combined with warp shuffles before fp64 RED atomics -- so no
 *     transpose of G is ever built.  The CSR arrays of G are only read during this call.
 *   C (n_con rows) and its transpose C^T (n_cols rows), both canonical CSR (borrowed): applied with the tiled
 *     CSR kernel.
 * rhs has n_data + n_con entries.  The atomics make the summation order of q run-dependent (last-bit differences). */
int lss_solver_create_blocks(lss_solver_t **solver, int64_t n_data, int64_t n_con, int64_t n_cols, int32_t row_len,
                             const int32_t *g_indptr, const int32_t *g_indices, const double *g_data,
                             const int32_t *c_indptr, const int32_t *c_indices, const double *c_data,
                             const int32_t *ct_indptr, const int32_t *ct_indices, const double *ct_data,
                             const double *rhs, const int32_t *fixed_cols, int64_t n_fixed, lss_comm_t *comm,
                             void *stream);
/* Row-sharded form for the GPUs of one node (one process per GPU): rank k owns a contiguous range of preconditioner
 * blocks [block_lo, block_hi) of the structure given later to lss_solver_set_structure -- the rows of A^T W A at their
 * columns, the block inverses, the matching entries of x and of the residual -- while the search direction p is
 * replicated.  A = [G_ext; C] holds every row that touches an own column: the rank's own points
 * (rows [own_row_lo, own_row_hi)), halo copies of neighbouring ranks' points around them, and all constraint rows
 * (rows [con_row_lo, n_rows)); owned_cols[n_cols] flags the own columns (device).  Per CG iteration there is NO
 * collective: the local product q_own = N_own p is followed by one kernel that exchanges the p.q partials, the own
 * slice of z = M^-1 r, the partial coarse restriction and the partial bias coupling with plain stores into the peers'
 * mailboxes over NVLink (lss_comm_box_open) behind sequence-numbered flags.  NCCL (the communicator's all-reduce) is used
 * only in lss_solver_set_weights and at the start / end of a solve.  [elim_begin, elim_end) must be the range later
 * given to lss_solver_set_structure.  x handed to lss_solver_solve must be complete on every rank and comes back
 * complete and bit-identical on every rank.  Works with one rank too (the mailbox is then local). */
int lss_solver_create_sharded(lss_solver_t **solver, int64_t n_rows, int64_t n_cols, int32_t row_len, const int32_t *indptr,
                              const int32_t *indices, const double *data, const double *rhs, const int32_t *fixed_cols,
                              int64_t n_fixed, const uint8_t *owned_cols, int32_t block_lo, int32_t block_hi,
                              int64_t own_row_lo, int64_t own_row_hi, int64_t con_row_lo, int64_t elim_begin, int64_t elim_end,
                              lss_comm_t *comm, void *stream);
/* Which operator form the solver uses and its size: info_h[0] = 0 generic CSR pair (A, A^T), 1 fused ELL pass,
 * 2 explicit normal matrix; info_h[1] = nnz of the explicit normal matrix (0 otherwise); info_h[2] = column pairs
 * sorted by its symbolic phase; info_h[3] = rows of A held in ELL form; info_h[4] = preconditioner blocks;
 * info_h[5] = eliminated columns; info_h[6] = blocks that were not positive definite in the last set_weights
 * (they fall back to their diagonal); info_h[7] = non-zeros of N streamed per CG product (rows of an eliminated
 * trailing block are left out; valid after the first solve).  info_h has 8 entries. */
int lss_solver_info(lss_solver_t *solver, int64_t *info_h);
/* Optional structure of the unknowns for the explicit-normal form (ignored by the other forms); call once, before
 * lss_solver_set_weights.
 *   blocks: n_blocks disjoint groups of at most 64 columns, block b = block_cols[block_ptr[b] .. block_ptr[b+1])
 *     (device arrays, copied): block-Jacobi preconditioner -- the diagonal blocks of A^T W A are inverted exactly.
 *     smooth_fit groups every dz node's epochs with the z0 nodes around it (each data point couples z0 to dz(t) at
 *     one place; scalar Jacobi ignores exactly that).  Columns outside all blocks keep scalar Jacobi.
 *   [elim_begin, elim_end): a column range whose diagonal block of A^T W A is diagonal (every row of A holds at most
 *     one of them: the per-(rgt, cycle) bias columns of assign_bias_ID, bias_functions.py:25-82).  They are eliminated
 *     exactly (CG runs on the Schur complement and back-substitutes); LSS_EINVAL if the block is not diagonal.
 *     With row-sharded ranks the slaved entries cost one extra small all-reduce (elim_end - elim_begin doubles)
 *     per CG iteration. */
int lss_solver_set_structure(lss_solver_t *solver, int32_t n_blocks, const int32_t *block_ptr, const int32_t *block_cols,
                             int32_t max_block, int64_t elim_begin, int64_t elim_end);
/* Optional coarse level for the block-Jacobi preconditioner (two-level, additive): M^-1 = M_BJ^-1 + P S_c^-1 P^T with
 * S_c = P^T S P the Galerkin projection of the (Schur-complement) operator the CG runs on, formed densely on the device
 * and inverted exactly (n_coarse <= 1024).  Call after lss_solver_set_structure with blocks, before
 * lss_solver_set_weights; ignored by the forms without a block preconditioner.  P interpolates every fine column
 * from at most 4 coarse unknowns: cp_idx[n_cols][4] (-1 = none) / cp_w[n_cols][4] (float); the same P transposed as
 * CSR by coarse row: pt_ptr[n_coarse + 1], pt_col[nnz_p], pt_w[nnz_p] (entries of one row in a fixed order: the
 * restriction is a deterministic gather).  All device arrays, copied.  refresh != 0: S_c is rebuilt with every
 * lss_solver_set_weights; 0: built once from the first weights (still a valid preconditioner afterwards).
 * The reference's direct QR solve (smooth_fit.py:142) has no counterpart; this only changes the iteration count. */
int lss_solver_set_coarse(lss_solver_t *solver, int32_t n_coarse, const int32_t *cp_idx, const float *cp_w,
                          const int32_t *pt_ptr, const int32_t *pt_col, const float *pt_w, int64_t nnz_p, int32_t refresh);
/* Grid-structured storage of N for the CG product (explicit-normal form on one GPU; optional).  The reference's solve
 * (sparseqr.solve, smooth_fit.py:142) has no counterpart: this only changes how q = N p is computed.  On the z0(y, x) +
 * dz(y, x, t) unknowns that setup_grids builds (grid_functions.py:26-134) almost all of N is regular: z0 node x dz node
 * pairs hold one value per epoch, dz node x dz node pairs are banded in time, z0 x z0 is a fixed set of (dy, dx) offsets.
 * These families are stored without column indices (z0-dz pairs once for both triangles) and the rest as a CSR remainder.
 * lss_solver_export_pattern copies the pattern of N out (indptr[n_cols + 1], indices[nnz], device arrays);
 * lssurf_b200/structured.py builds the tables from it; lss_solver_set_structured hands them over.  All pointers are device
 * arrays that must stay valid for the life of the solver (borrowed).  Values are taken from N through `map` after every
 * lss_solver_set_weights.  Layout of the value buffer (doubles; nt = epochs per dz node, nn = dz nodes, n0 = z0 nodes); node
 * blocks and pair records are multiples of 16 bytes so that each travels as one bulk copy:
 *   node block a at [a * node_len]:   [((band + 1) * dd_stride + j) * nt + t]   N[(a, t), (b_j, t + band)], band = -1..1
 *                                     [3 * dd_stride * nt ...]                  the neighbour ids b_j = dd_b[a * dd_stride + j]
 *                                                                               as int32, two per double (-1: slot not used)
 *                                     [self_off + (band + 2) * nt + t]          N[(a, t), (a, t + band)], band = -2..2,
 *                                                                               self_off = 3 * dd_stride * nt + (dd_stride + 1) / 2
 *   pair record at [zd_base + pair * rec_len]:  [t] N[k, (a, t)], [nt] = k = zd_k[pair] as int64; the pairs of dz node a
 *                                     are [zd_ptr[a], zd_ptr[a + 1]) (at most max_pairs <= 8 * (256 / nt))
 *   [zz_base + (d * n0 + i)]          N[z0 i, z0 i + zz_off[d]] */
typedef struct {
    int32_t nt, nn, n0, n_diag, dd_stride, node_len, rec_len, max_pairs, n_zd, rem_n_rows;
    int64_t z0_col0, dz_col0, rows, nnz, n_vals, zd_base, zz_base, rem_nnz;
    const int32_t *map;                       /* [nnz] slot of every CSR entry of the leading rows, -1: not stored here */
    const int32_t *dd_b, *zd_ptr, *zd_k, *zz_off;
    const int32_t *rem_src, *rem_rows, *rem_indptr, *rem_indices;   /* remainder: CSR over rem_n_rows non-empty rows */
} lss_structured_t;
int lss_solver_export_pattern(lss_solver_t *solver, int32_t *indptr, int32_t *indices);
int lss_solver_set_structured(lss_solver_t *solver, const lss_structured_t *desc);
/* Row weights w (length n_rows): 1/sigma^2 for selected data rows, 0 for edited rows, 1/E_c^2 for
 * constraint rows.  Recomputes the right-hand side A^T W b and the Jacobi diagonal. */
int lss_solver_set_weights(lss_solver_t *solver, const double *w);
/* Solve from the initial guess in x (length n_cols).  Stops when sqrt(r^T M^-1 r / b^T M^-1 b) <= rtol
 * or after max_iter iterations.  iters_h / relres_h: host outputs. */
int lss_solver_solve(lss_solver_t *solver, double *x, double rtol, int32_t max_iter, int32_t *iters_h,
                     double *relres_h);
/* Record CUDA-event pairs around every A and A^T product of the following solves (small overhead). */
int lss_solver_set_profile(lss_solver_t *solver, int32_t on);
/* Statistics of the last solve, timed with CUDA events on the solver's stream:
 * stats_h[0] = total ms; [1], [2] = ms in the two halves of the operator (generic form: y = w.*(A p), then A^T u;
 * block form: the constraint products C, C^T, then the fused data-block kernel), 0 unless profiling is on;
 * [3] = number of operator applications behind [1], [2]; [4] = kernels launched; [5] = nnz(A); [6] = ms in the vector
 * step of the iterations (+ collective, + elimination) when profiling; [7] reserved.  stats_h has 8 entries. */
int lss_solver_stats(lss_solver_t *solver, double *stats_h);

/* ------------------------------------------------------------------ robust statistics / editing */

/* RDE(r / sqrt(s1^2 + sigma^2)) over points with mask != 0 and a finite value (RDE.py:10-18 inside
 * calc_sigma_extra.py:36-39): exact selection of the four order statistics np.interp touches, then
 * the same interpolation arithmetic.  sigma == NULL => RDE(r).  mask == NULL => all points.
 * n_used_h receives the (global) number of points used; the result is NaN when it is < 2. */
int lss_rde(const double *r, const double *sigma, const uint8_t *mask, double s1, int64_t n, lss_comm_t *comm,
            double *rde_h, int64_t *n_used_h, void *stream);

/* calc_sigma_extra for one mask group (calc_sigma_extra.py:34-41): bounded Brent minimisation of
 * (RDE(r/sqrt(s^2+sigma^2)) - 1)^2 on [0, RDE(r)] with scipy's constants (xatol 1e-5, 500 calls).
 * Returns 0 in *sigma_extra_h when fewer than 10 points are selected. */
int lss_sigma_extra(const double *r, const double *sigma, const uint8_t *mask, int64_t n, lss_comm_t *comm,
                    double *sigma_extra_h, int32_t *nfev_h, void *stream);

/* in_TSE = |r / sqrt(sigma^2 + sigma_extra^2)| < n_sigma (smooth_fit.py:159-163) plus the statistics
 * the convergence tests need: stats_h[0] = count(in_TSE), stats_h[1] = min sigma over in_TSE
 * (smooth_fit.py:203), stats_h[2] = number of points with ||r/sigma_aug| - n_sigma| < 1e-6. */
int lss_edit_mask(const double *r, const double *sigma, const double *sigma_extra, int64_t n, double n_sigma,
                  uint8_t *in_TSE, double *stats_h, lss_comm_t *comm, void *stream);

/* w[i] = mask[i] ? 1 / (sigma[i]^2 (+ sigma_extra[i]^2)) : 0   (smooth_fit.py:123-135). */
int lss_data_weights(const double *sigma, const double *sigma_extra, const uint8_t *mask, int64_t n, double *w,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSSURF_B200_H */
