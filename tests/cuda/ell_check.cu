// Stand-alone check of the fused ELL kernel q += G^T (w .* (G x)) against a host loop (run on the GPU box):
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tests/cuda/ell_check.cu -Llssurf_b200 -llssurf_b200 -o /tmp/ell_check
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace lss {
int ell_build(int64_t n_rows, int K, const int32_t *indptr, const int32_t *indices, const double *data, int32_t *ell_col,
              double *ell_val, int64_t n_pad, cudaStream_t stream);
int ell_normal_launch(int mode, int64_t n_rows, int64_t n_pad, int K, const int32_t *ell_col, const double *ell_val, const double *x,
                      const double *w, const double *b, double *q, double *pq_slots, const double *stop, cudaStream_t stream);
}
extern "C" const char *lss_last_error(void);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 6000;
    const int K = argc > 2 ? atoi(argv[2]) : 12;
    const int ncols = argc > 3 ? atoi(argv[3]) : 4174;
    srand(1);
    std::vector<int> indptr(n + 1), idx;
    std::vector<double> val, x(ncols), w(n), b(n);
    for (int64_t r = 0; r < n; ++r) {
        indptr[r] = (int)idx.size();
        int len = (r % 7 == 0) ? K - 2 : K;                 // some shorter rows (dropped zero weights)
        int base = (int)((r / 5) % (ncols - 2 * K));        // neighbouring rows share columns -> runs
        for (int k = 0; k < len; ++k) { idx.push_back(base + k); val.push_back(0.1 + (rand() % 1000) / 1000.0); }
    }
    indptr[n] = (int)idx.size();
    for (auto &v : x) v = (rand() % 2000) / 1000.0 - 1.0;
    for (int64_t r = 0; r < n; ++r) { w[r] = (r % 11 == 0) ? 0.0 : 1.0 + (rand() % 100) / 100.0; b[r] = (rand() % 2000) / 1000.0 - 1.0; }
    int *d_ip, *d_ix, *d_ec; double *d_v, *d_ev, *d_x, *d_w, *d_b, *d_q;
    const int64_t n_pad = (n + 31) / 32 * 32;
    CK(cudaMalloc(&d_ip, 4 * (n + 1))); CK(cudaMalloc(&d_ix, 4 * idx.size())); CK(cudaMalloc(&d_v, 8 * val.size()));
    CK(cudaMalloc(&d_ec, 4 * n_pad * K)); CK(cudaMalloc(&d_ev, 8 * n_pad * K));
    CK(cudaMalloc(&d_x, 8 * ncols)); CK(cudaMalloc(&d_w, 8 * n)); CK(cudaMalloc(&d_b, 8 * n)); CK(cudaMalloc(&d_q, 8 * ncols));
    CK(cudaMemcpy(d_ip, indptr.data(), 4 * (n + 1), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_ix, idx.data(), 4 * idx.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_v, val.data(), 8 * val.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_x, x.data(), 8 * ncols, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w.data(), 8 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_b, b.data(), 8 * n, cudaMemcpyHostToDevice));
    if (lss::ell_build(n, K, d_ip, d_ix, d_v, d_ec, d_ev, n_pad, 0)) { printf("build: %s\n", lss_last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    double *d_pq; CK(cudaMalloc(&d_pq, 8 * 256)); CK(cudaMemset(d_pq, 0, 8 * 256));
    int bad = 0;
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(d_q, 0, 8 * ncols));
        if (lss::ell_normal_launch(mode, n, n_pad, K, d_ec, d_ev, d_x, d_w, d_b, d_q, mode == 0 ? d_pq : nullptr, nullptr, 0)) { printf("launch: %s\n", lss_last_error()); return 1; }
        CK(cudaDeviceSynchronize());
        std::vector<double> q(ncols), ref(ncols, 0.0);
        CK(cudaMemcpy(q.data(), d_q, 8 * ncols, cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r < n; ++r) {
            double t = 0.0;
            for (int k = indptr[r]; k < indptr[r + 1]; ++k) t += val[k] * x[idx[k]];
            double u = mode == 0 ? w[r] * t : (mode == 1 ? w[r] * b[r] : w[r]);
            for (int k = indptr[r]; k < indptr[r + 1]; ++k) ref[idx[k]] += (mode == 2 ? val[k] * val[k] : val[k]) * u;
        }
        double err = 0, mx = 0;
        for (int j = 0; j < ncols; ++j) { err = fmax(err, fabs(q[j] - ref[j])); mx = fmax(mx, fabs(ref[j])); }
        printf("mode %d: max |q - ref| = %.3e (max |ref| %.3e)\n", mode, err, mx);
        if (mode == 0) {          // x . q must equal the sum of the 64 p.q slots
            double slots[256], pq = 0, want = 0;
            CK(cudaMemcpy(slots, d_pq, sizeof(slots), cudaMemcpyDeviceToHost));
            for (double v : slots) pq += v;
            for (int j = 0; j < ncols; ++j) want += x[j] * ref[j];
            printf("        x.q from slots %.12e, from q %.12e\n", pq, want);
            if (!(fabs(pq - want) <= 1e-11 * fabs(want))) bad = 1;
        }
        if (!(err <= 1e-11 * mx)) bad = 1;
    }
    printf(bad ? "FAIL\n" : "PASS\n");
    return bad;
}
