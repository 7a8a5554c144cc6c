"""cProfile of one warm smooth_fit call on the cfg-4 workload: where the host-side (Python / NumPy / sync) time goes."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts, kw = synthetic.icesat2_tracks(n=n, width=1.e5, seed=0)
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}
def fit():
    D = L.point_data.data().from_dict(dict(dev))
    S = L.smooth_fit(data=D, VERBOSE=False, **kw)
    torch.cuda.synchronize()
    return S
fit(); fit()
pr = cProfile.Profile(); pr.enable(); fit(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25); print(s.getvalue()[:5000])
