// Grid-structured storage of the normal matrix for the CG product (structured.cu; tables built by lssurf_b200/structured.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lssurf_b200.h"
#include "common.cuh"

namespace lss {

struct StructuredN {
    lss_structured_t d;          // sizes and borrowed device tables (the caller keeps them alive)
    double *vals;                // [n_vals] stored values of the four families, zero where N has no entry
    double *rem_data;            // [rem_nnz] values of the remainder CSR
    int *rem_row;                // [rem_nnz] row of every remainder entry (the product folds equal rows inside a warp)
    int stage_len, grid;         // doubles per ring stage, persistent CTAs of the product kernel
    size_t smem;
    cudaStream_t stream;
};

int structured_create(StructuredN **out, const lss_structured_t *desc, cudaStream_t stream);
void structured_destroy(StructuredN *s);
// values from the CSR values of N (after every numeric update)
int structured_refresh(StructuredN *s, const double *csr_data);
// q += N p over the leading rows (q holds zeros on entry); dot_slots[256] += partial sums of p . (N p)
int structured_apply(StructuredN *s, const double *p, double *q, const double *stop, int64_t *launches, double *dot_slots);

}  // namespace lss
