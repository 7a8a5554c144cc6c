"""GPU probe: time of the symbolic part of the structured normal-matrix layout (structured.build) on the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic, structured, solver as S

pts, kw = synthetic.icesat2_tracks(n=int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, width=1.e5, seed=0)
os.environ['LSS_STRUCTURED_N'] = '0'
captured = {}
orig = S.NormalSolver.__init__
def hook(self, *a, **k):
    orig(self, *a, **k)
    captured['solver'] = self
    captured['structure'] = k.get('structure')
S.NormalSolver.__init__ = hook
orig_close = S.NormalSolver.close
S.NormalSolver.close = lambda self: None
kw['max_iterations'] = 1
L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, **kw)
sv = captured['solver']
gd = captured['structure']['grids_desc']
dev = sv.x.device
indptr = torch.empty(sv.n_cols + 1, dtype=torch.int32, device=dev)
indices = torch.empty(sv.nnz_N, dtype=torch.int32, device=dev)
L._cabi.check(L._cabi.lib().lss_solver_export_pattern(sv.handle, L._cabi.ptr(indptr), L._cabi.ptr(indices)))
rows = sv.n_cols - sv.n_elim
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    st = structured.build(indptr, indices, rows, *gd)
    torch.cuda.synchronize(); print('build %.1f ms' % (1e3 * (time.time() - t0)), st['counts'], 'bytes %.1f MB' % (structured.product_bytes(st) / 1e6))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st = structured.build(indptr, indices, rows, *gd)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=18, max_name_column_width=60))
