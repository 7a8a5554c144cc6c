"""Micro-benchmark of the CSR product kernel on the cfg-4 operator (A = [G; C] and A^T), per CTA-shape variant.

    python tools/spmv_bench.py [--points 10000000] [--variants 0,1,2,3,4,5,6] [--sorted 0|1]
Times lss_spmv-family launches with CUDA events (20 launches after 3 warm-ups; matrices are 1.7 GB each, far
larger than L2) and prints achieved GB/s against the algorithmic byte count of SURVEY.md section 8(d).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lssurf_b200 as L                                  # noqa: E402
from lssurf_b200 import synthetic                        # noqa: E402
from lssurf_b200.lin_op import DeviceCSR                 # noqa: E402
from lssurf_b200.solver import NormalSolver              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=10_000_000)
    ap.add_argument('--width', type=float, default=1.e5)
    ap.add_argument('--variants', default='0,1,2,3,4,5,6')
    ap.add_argument('--reps', type=int, default=20)
    a = ap.parse_args()
    pts, kw = synthetic.icesat2_tracks(n=a.points, width=a.width, seed=0)
    fo = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, return_fit_objects=True, **kw)
    G = fo['G_data']._csr_device(row_N=int(fo['G_data'].N_eq))
    C = fo['Gc']._csr_device(col_N=G.shape[1], row_N=int(fo['Gc'].N_eq))
    A = DeviceCSR.vstack([G, C])
    At = A.transpose()
    n_rows, n_cols, nnz = A.shape[0], A.shape[1], A.nnz
    rng = np.random.default_rng(0)
    x = L._cabi.to_dev(rng.normal(size=n_cols)); w = L._cabi.to_dev(rng.uniform(1, 2, n_rows))
    u = torch.empty(n_rows, dtype=torch.float64, device='cuda'); q = torch.empty(n_cols, dtype=torch.float64, device='cuda')
    b_a = 12 * nnz + 4 * (n_rows + 1) + 8 * n_rows + 8 * n_cols + 8 * n_rows
    b_at = 12 * nnz + 4 * (n_cols + 1) + 8 * n_cols + 8 * n_rows
    ref_u = ref_q = None
    for v in [int(t) for t in a.variants.split(',')]:
        L._cabi.set_option('spmv_variant', v)
        A._plan = At._plan = None          # plans are per tile size
        res = {}
        for name, M, xin, out, mode, s, nb in (('A', A, x, u, 1, w, b_a), ('At', At, u, q, 0, None, b_at)):
            for _ in range(3):
                M.dot(xin, mode=mode, s=s, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.reps):
                M.dot(xin, mode=mode, s=s, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            res[name] = (ms, nb / ms / 1e6)
        if ref_u is None:
            ref_u, ref_q = u.clone(), q.clone()
        du = float((u - ref_u).abs().max() / ref_u.abs().max()); dq = float((q - ref_q).abs().max() / ref_q.abs().max())
        print(json.dumps({'variant': v, 'A_ms': round(res['A'][0], 4), 'A_GBs': round(res['A'][1], 1), 'At_ms': round(res['At'][0], 4),
                          'At_GBs': round(res['At'][1], 1), 'rel_diff_vs_first': [du, dq], 'nnz': nnz}), flush=True)


if __name__ == '__main__':
    main()
