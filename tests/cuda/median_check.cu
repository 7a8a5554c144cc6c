// Stand-alone check of lss_group_median / lss_dense_rank_f64 against host loops (run on the GPU box).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../../include/lssurf_b200.h"
int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 20;
    const int ng = argc > 2 ? atoi(argv[2]) : 3;
    srand(3);
    std::vector<int> ids(n);
    std::vector<double> v(n);
    for (int64_t i = 0; i < n; ++i) { ids[i] = rand() % ng; v[i] = (rand() % 2001) / 1000.0 - 1.0; if (rand() % 50 == 0) v[i] = NAN; }
    int *d_ids; double *d_v, *d_out;
    cudaMalloc(&d_ids, 4 * n); cudaMalloc(&d_v, 8 * n); cudaMalloc(&d_out, 8 * ng);
    cudaMemcpy(d_ids, ids.data(), 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(d_v, v.data(), 8 * n, cudaMemcpyHostToDevice);
    int rc = lss_group_median(d_ids, d_v, n, ng, d_out, 0);
    cudaDeviceSynchronize();
    std::vector<double> out(ng);
    cudaMemcpy(out.data(), d_out, 8 * ng, cudaMemcpyDeviceToHost);
    int bad = rc != 0;
    for (int g = 0; g < ng; ++g) {
        std::vector<double> s;
        for (int64_t i = 0; i < n; ++i) if (ids[i] == g && v[i] == v[i]) s.push_back(v[i]);
        std::sort(s.begin(), s.end());
        double want = s.empty() ? NAN : (s.size() & 1 ? s[s.size() / 2] : (s[s.size() / 2 - 1] + s[s.size() / 2]) / 2.0);
        bool ok = (want != want && out[g] != out[g]) || want == out[g];
        if (!ok || ng <= 8) printf("group %d: got %.6f want %.6f (%zu values)%s\n", g, out[g], want, s.size(), ok ? "" : "  <-- MISMATCH");
        bad |= !ok;
    }
    printf(bad ? "FAIL rc=%d %s\n" : "PASS\n", rc, lss_last_error());
    return bad;
}
