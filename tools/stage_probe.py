"""Wall time per stage of warm smooth_fit calls on the cfg-4 workload (device-resident inputs), three fits in a row."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic
pts, kw = synthetic.icesat2_tracks(n=10_000_000, width=1.e5, seed=0)
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}
def fit(stages):
    D = L.point_data.data().from_dict(dict(dev))
    t0 = time.perf_counter()
    S = L.smooth_fit(data=D, VERBOSE=False, time_stages=stages, **kw)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, S
fit(False); fit(False)
for k in range(3):
    dt, S = fit(False)
    print('plain fit %.3f s' % dt)
for k in range(2):
    dt, S = fit(True)
    st = S['timing']['stages']
    print('staged fit %.3f s' % dt, {k: round(v, 4) for k, v in st.items() if v > 0.008})
