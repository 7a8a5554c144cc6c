"""Multi-GPU parity check: one rank of an N-rank run (torchrun, or spawned by tests/test_gpu_multi.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank runs the multi-GPU smooth_fit in its solve modes:
  p2p         (default) row-sharded normal matrix, p exchanged inside the CG step kernel through the peer-to-peer
              mailbox over NVLink, no collective per iteration;
  replicated  (LSS_REPLICATED_SOLVE=1) rows all-gathered once, every rank runs the whole CG;
  allreduce   (LSS_REPLICATED_SOLVE=1 LSS_SHARDED_SOLVE=1) row-sharded products summed by one NCCL all-reduce per iteration.
Rank 0 then repeats the fit on one GPU and requires the same edit mask after every outer iteration and the same model
to 1e-9 relative L2 (summation order differs from the single-GPU order).
"""
import faulthandler
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import lssurf_b200 as L                       # noqa: E402
from lssurf_b200 import parallel, synthetic   # noqa: E402

MODES = {'p2p': {}, 'replicated': {'LSS_REPLICATED_SOLVE': '1'},
         'allreduce': {'LSS_REPLICATED_SOLVE': '1', 'LSS_SHARDED_SOLVE': '1'}}


def main():
    faulthandler.dump_traceback_later(int(os.environ.get('LSS_DIST_CHECK_TIMEOUT', '600')), exit=True)
    ctx = parallel.init_from_env()
    modes = [m for m in os.environ.get('LSS_DIST_CHECK_MODES', 'p2p,replicated,allreduce').split(',') if m]
    scale = float(os.environ.get('LSS_DIST_CHECK_SCALE', '1'))
    cases = [('uniform_3d', synthetic.uniform_3d(n=int(200_000 * scale), width=3.e4 * scale ** 0.5, seed=4)),
             ('tracks_bias', synthetic.icesat2_tracks(n=int(300_000 * scale), width=2.e4 * scale ** 0.5, seed=4)),
             # wide enough for the two-level preconditioner (>= 2000 dz nodes) on top of the bias elimination
             ('tracks_bias_coarse', synthetic.icesat2_tracks(n=int(400_000 * scale), width=5.e4 * scale ** 0.5, seed=5))]
    # spatially varying sigma_extra: the bins are global, every rank holds a slice of each (points stay in the caller's order)
    pb, kb = synthetic.uniform_3d(n=int(120_000 * scale), width=2.e4 * scale ** 0.5, seed=6)
    pb['z'] = pb['z'] + np.random.default_rng(6).normal(0, 1., pb['x'].size) * 0.3 * (pb['x'] > 0)
    cases.append(('uniform_sigma_bins', (pb, dict(kb, sigma_extra_bin_spacing=5.e3 * scale ** 0.5, sigma_extra_max=5.))))
    only = [c for c in os.environ.get('LSS_DIST_CHECK_CASES', '').split(',') if c]
    if only:
        cases = [c for c in cases if c[0] in only]
    ok = True
    single = {}
    for name, (pts, kw) in cases:
        for mode in modes:
            for k in ('LSS_REPLICATED_SOLVE', 'LSS_SHARDED_SOLVE'):
                os.environ.pop(k, None)
            os.environ.update(MODES[mode])
            hist = []
            S = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, dist=ctx if ctx.world > 1 else None,
                             history=hist, **kw)
            if mode == 'p2p' and name == 'tracks_bias' and ctx.world > 1:
                # gather_outputs='root': the same fit, per-point outputs copied to the host on rank 0 only
                Sr = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, dist=ctx, gather_outputs='root', **kw)
                if ctx.rank == 0:
                    rel_r = float(np.linalg.norm(Sr['m']['all'] - S['m']['all']) / np.linalg.norm(S['m']['all']))
                    same_r = bool(np.array_equal(Sr['data'].three_sigma_edit, S['data'].three_sigma_edit))
                    dz_r = float(np.max(np.abs(Sr['data'].z_est - S['data'].z_est)))
                    print(f'[dist_check] {name}/{mode}: gather_outputs=root rel-L2(m0)={rel_r:.2e} final mask identical={same_r} '
                          f'max|dz_est|={dz_r:.2e}', flush=True)
                    ok = ok and rel_r < 1e-9 and same_r and dz_r < 1e-6
                else:
                    assert not any(f in Sr['data'].fields for f in ('z_est', 'sigma_extra', 'bias_ID')), 'outputs are rank 0\'s'
                del Sr
            if ctx.rank != 0:
                continue
            if name not in single:
                hist1 = []
                S1 = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, history=hist1, **kw)
                single[name] = (S1, hist1)
            S1, hist1 = single[name]
            rel = float(np.linalg.norm(S['m']['all'] - S1['m']['all']) / np.linalg.norm(S1['m']['all']))
            same_final = bool(np.array_equal(S['data'].three_sigma_edit, S1['data'].three_sigma_edit))
            same_count = [h['count'] for h in hist] == [h['count'] for h in hist1]
            # the row-sharded mode records global masks; the others (and fits that cannot be row-sharded, e.g. with gridded
            # sigma_extra) record rank-local ones, whose global counts are compared instead
            global_masks = mode == 'p2p' and len(hist) == len(hist1) and all(a['mask_out'].size == b['mask_out'].size for a, b in zip(hist, hist1))
            same_masks = all(np.array_equal(a['mask_out'], b['mask_out']) for a, b in zip(hist, hist1)) if global_masks else same_count
            dz_est = float(np.max(np.abs(S['data'].z_est - S1['data'].z_est)))
            cnt = float(np.nanmax(np.abs(S['m']['z0'].count - S1['m']['z0'].count)))
            mis = float(np.nanmax(np.abs(S['m']['dz'].misfit_rms - S1['m']['dz'].misfit_rms)))
            print(f'[dist_check] {name}/{mode}: world={ctx.world} rel-L2(m0)={rel:.2e} final mask identical={same_final} '
                  f'masks identical per iteration={same_masks} counts equal={same_count} max|dz_est|={dz_est:.2e} '
                  f'count-map diff={cnt} misfit-map diff={mis:.1e} coarse={hist[0]["pcg"].get("coarse")} '
                  f'its={[h["pcg"]["iterations"] for h in hist]} vs {[h["pcg"]["iterations"] for h in hist1]}', flush=True)
            ok = ok and rel < 1e-9 and same_final and same_count and same_masks and cnt < 1e-9 and mis < 1e-8
    import torch.distributed as td
    if ctx.world > 1:
        td.barrier()
    ctx.destroy()
    if ctx.rank == 0:
        print('[dist_check] PASS' if ok else '[dist_check] FAIL', flush=True)
        sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
