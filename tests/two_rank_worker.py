"""One rank of the world-size-2 host-logic test (tests/test_host_double.py): gloo process group, C-ABI wrappers replaced
by the test double, the package's own smooth_fit with this rank's DistContext.  Usage: rank world port case out.npz [gather_outputs]
(case = a fixture name, or the path of a pickled (pts, kw) pair written by tools/dist_fuzz.py)"""
import faulthandler
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port, name, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    gather = sys.argv[6] if len(sys.argv) > 6 else 'all'
    if gather == 'root':
        os.environ['LSS_SHARED_UPLOAD_MIN'] = '0'       # also exercise the shared upload (each rank copies 1 / P of every field)
    faulthandler.dump_traceback_later(240, exit=True)          # a stuck collective must not hang the suite
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import lssurf_b200 as L
    from lssurf_b200.parallel import DistContext
    from tests import host_double
    from tests.util import load_case
    if name.endswith('.pkl'):           # tools/dist_fuzz.py: a random case instead of a fixture
        import pickle
        with open(name, 'rb') as f:
            pts, kw = pickle.load(f)
    else:
        fx, pts, kw = load_case(name)
    hist = []
    with host_double.installed():
        S = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, history=hist, sort_points=False,
                         dist=DistContext(rank, world, None), gather_outputs=gather, **kw)
    if gather == 'root' and rank != 0:
        # per-point outputs are rank 0's; this rank returns the model, the grids and the all-reduced statistics
        assert not any(f in S['data'].fields for f in ('z_est', 'sigma_extra'))
        np.savez(out, m0=S['m']['all'], count=S['m']['z0'].count, masks=np.array([h['mask_out'] for h in hist]), n_solves=len(S['timing']['pcg']))
        dist.destroy_process_group()
        return
    n_edited = S['m']['bias']['N_edited'] if 'bias' in S['m'] else []
    np.savez(out, m0=S['m']['all'], in_TSE=np.packbits(S['data'].three_sigma_edit), masks=np.array([h['mask_out'] for h in hist]),
             z_est=np.asarray(S['data'].z_est), count=S['m']['z0'].count, n_edited=np.asarray(n_edited, dtype=float),
             n_solves=len(S['timing']['pcg']))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
