"""Multi-GPU parity check (run under torchrun on N GPUs of one box; not collected by pytest).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank runs the multi-GPU smooth_fit in both solve modes (replicated solve: rows all-gathered once, no collective
inside CG; row-sharded solve, LSS_SHARDED_SOLVE=1: one all-reduce per CG iteration); rank 0 then repeats the fit on one
GPU and requires the same edit masks after every outer iteration and the same model to 1e-9 relative L2 (summation
order differs from the single-GPU order).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import lssurf_b200 as L                       # noqa: E402
from lssurf_b200 import parallel, synthetic   # noqa: E402


def main():
    ctx = parallel.init_from_env()
    cases = [('uniform_3d', synthetic.uniform_3d(n=200_000, width=3.e4, seed=4)),
             ('tracks_bias', synthetic.icesat2_tracks(n=300_000, width=2.e4, seed=4))]
    ok = True
    for name, (pts, kw), mode in [(c[0], c[1], m) for c in cases for m in ('replicated', 'sharded')]:
        hist = []
        if mode == 'sharded':
            os.environ['LSS_SHARDED_SOLVE'] = '1'
        else:
            os.environ.pop('LSS_SHARDED_SOLVE', None)
        S = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, dist=ctx if ctx.world > 1 else None,
                         history=hist, **kw)
        name = name + '/' + mode
        if ctx.rank == 0:
            hist1 = []
            S1 = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, history=hist1, **kw)
            rel = float(np.linalg.norm(S['m']['all'] - S1['m']['all']) / np.linalg.norm(S1['m']['all']))
            same_final = bool(np.array_equal(S['data'].three_sigma_edit, S1['data'].three_sigma_edit))
            same_count = [h['count'] for h in hist] == [h['count'] for h in hist1]
            dz_est = float(np.max(np.abs(S['data'].z_est - S1['data'].z_est)))
            cnt = float(np.nanmax(np.abs(S['m']['z0'].count - S1['m']['z0'].count)))
            print(f'[dist_check] {name}: world={ctx.world} rel-L2(m0)={rel:.2e} masks identical={same_final} '
                  f'per-iteration counts equal={same_count} max|dz_est|={dz_est:.2e} count-map diff={cnt} '
                  f'its={[h["pcg"]["iterations"] for h in hist]} vs {[h["pcg"]["iterations"] for h in hist1]}', flush=True)
            ok = ok and rel < 1e-9 and same_final and same_count and cnt < 1e-9
    import torch.distributed as td
    if ctx.world > 1:
        td.barrier()
    ctx.destroy()
    if ctx.rank == 0:
        print('[dist_check] PASS' if ok else '[dist_check] FAIL', flush=True)
        sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
