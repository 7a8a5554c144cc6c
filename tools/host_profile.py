"""GPU probe: cProfile of one warm smooth_fit on the bench workload (device-resident inputs) -- where the HOST time of the
non-CG stages goes (blocking .item() / .cpu() calls, Python bookkeeping).  Usage: python tools/host_profile.py [points]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pts, kw = synthetic.icesat2_tracks(n=n, width=1.e5 * (n / 1e7) ** 0.5, seed=0)
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}


def fit(**extra):
    D = L.point_data.data().from_dict({f: dev[f] for f in fields})
    return L.smooth_fit(data=D, VERBOSE=False, **dict(kw, **extra))


S = None
for _ in range(3):
    S = fit()
torch.cuda.synchronize()
t0 = time.perf_counter()
S = fit()
torch.cuda.synchronize()
print('warm fit %.1f ms' % (1e3 * (time.perf_counter() - t0)))
S2 = fit(time_stages=True)
print('stages', {k: round(v * 1e3, 1) for k, v in S2['timing']['stages'].items() if v >= 0.0005})
pr = cProfile.Profile()
pr.enable()
S = fit()
torch.cuda.synchronize()
pr.disable()
for key in ('tottime', 'cumulative'):
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).strip_dirs().sort_stats(key).print_stats(45)
    print(buf.getvalue()[:9000])
