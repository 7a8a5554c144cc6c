"""GPU probe: torch.profiler (CUPTI) timeline of one warm smooth_fit on the bench workload, reduced per stage to wall time,
GPU-busy time (union of kernel / memcpy / memset intervals) and the top kernels.  Usage: python tools/timeline_probe.py [points]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import lssurf_b200 as L
from lssurf_b200 import synthetic, parallel
import lssurf_b200.smooth_fit          # noqa: F401

ctx = parallel.init_from_env()          # under torchrun: the row-sharded multi-GPU fit, rank 0 reports

SF = sys.modules['lssurf_b200.smooth_fit']
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pts, kw = synthetic.icesat2_tracks(n=n, width=1.e5 * (n / 1e7) ** 0.5, seed=0)
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}
marks = []


class Stages(object):          # stage boundaries as host timestamps WITHOUT device synchronisation
    def __init__(self, on, sink):
        marks.append(('start', time.perf_counter_ns()))

    def mark(self, name):
        marks.append((name, time.perf_counter_ns()))


def fit():
    D = L.point_data.data().from_dict({f: dev[f] for f in fields})
    return L.smooth_fit(data=D, VERBOSE=False, dist=ctx if ctx.world > 1 else None, **kw)


S = None
for _ in range(3):
    S = fit()
torch.cuda.synchronize()
SF._Stages = Stages
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t_host0 = time.perf_counter_ns()
    S = fit()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter_ns()
if ctx.rank != 0:
    import torch.distributed as td
    td.barrier()
    ctx.destroy()
    sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
print('fit wall %.1f ms, %d device events' % ((t_host1 - t_host0) / 1e6, len(ev)))
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
t0 = iv[0][0]
busy = 0.0
cur_s, cur_e = iv[0][0], iv[0][1]
gaps = []
for s, e, name in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0, name))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
span = iv[-1][1] - t0
print('device span %.1f ms, busy %.1f ms, idle %.1f ms in %d gaps' % (span / 1e3, busy / 1e3, (span - busy) / 1e3, len(gaps)))
print('largest gaps (us, at ms, next kernel):')
for g, at, name in sorted(gaps, reverse=True)[:40]:
    print('  %8.0f  at %7.1f  %s' % (g, at / 1e3, name[:90]))
# gap histogram
h = collections.Counter()
for g, _, _ in gaps:
    h['<5us' if g < 5 else '<20us' if g < 20 else '<100us' if g < 100 else '<1ms' if g < 1000 else '>=1ms'] += g
print('idle by gap size (ms):', {k: round(v / 1e3, 1) for k, v in h.items()})
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, name in iv:
    agg[name][0] += 1
    agg[name][1] += e - s
print('top device activities:')
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print('  %8.2f ms %6d  %s' % (t / 1e3, c, name[:100]))
# host-side stage boundaries -> which stage each gap falls in (device timestamps and perf_counter share no epoch:
# align the first mark with the first device event of the fit, good to a few 100 us)
print('host stage marks (ms since start):', [(nm, round((t - marks[0][1]) / 1e6, 1)) for nm, t in marks])

if ctx.world > 1:
    import torch.distributed as td
    td.barrier()
    ctx.destroy()
