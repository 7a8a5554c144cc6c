"""GPU probe: host time per source line of smooth_fit / parse_model / iterate_fit_device for one warm fit of the bench
workload (sys.settrace on those code objects only; no device synchronisation added, so a line's time is how long the HOST
stayed on it, blocking copies and allocator stalls included).  Usage: python tools/line_probe.py [points]"""
import collections
import linecache
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic
import lssurf_b200.smooth_fit          # noqa: F401

SF = sys.modules['lssurf_b200.smooth_fit']
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pts, kw = synthetic.icesat2_tracks(n=n, width=1.e5 * (n / 1e7) ** 0.5, seed=0)
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}


def fit():
    D = L.point_data.data().from_dict({f: dev[f] for f in fields})
    return L.smooth_fit(data=D, VERBOSE=False, **kw)


S = None
for _ in range(3):
    S = fit()
torch.cuda.synchronize()
from lssurf_b200 import solver
CF, GF, LO = (sys.modules['lssurf_b200.' + k] for k in ('constraint_functions', 'grid_functions', 'lin_op'))
watch = {f.__code__ for f in (SF.smooth_fit, SF.parse_model, SF.iterate_fit_device, SF._bias_statistics, SF._assign_bias_ID_device,
                              CF.setup_smoothness_constraints, GF.setup_grids, solver.NormalSolver.__init__,
                              LO.lin_op.vstack, LO.lin_op.diff_op)}
acc = collections.defaultdict(float)
state = {}


def local(frame, event, arg):
    now = time.perf_counter()
    key = id(frame)
    prev = state.get(key)
    if prev is not None:
        acc[prev[0]] += now - prev[1]
    if event == 'return':
        state.pop(key, None)
    else:
        state[key] = ((frame.f_code.co_filename, frame.f_lineno), time.perf_counter())
    return local


def tracer(frame, event, arg):
    if event == 'call' and frame.f_code in watch:
        state[id(frame)] = ((frame.f_code.co_filename, frame.f_lineno), time.perf_counter())
        return local
    return None


t0 = time.perf_counter()
sys.settrace(tracer)
S = fit()
sys.settrace(None)
torch.cuda.synchronize()
print('traced fit %.1f ms' % (1e3 * (time.perf_counter() - t0)))
for (fn, ln), t in sorted(acc.items(), key=lambda kv: -kv[1])[:70]:
    print('%7.2f ms  %s:%d  %s' % (t * 1e3, os.path.basename(fn), ln, linecache.getline(fn, ln).strip()[:110]))
