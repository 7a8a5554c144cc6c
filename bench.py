#!/usr/bin/env python
"""Benchmark of the smooth_fit solve hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload = "cfg4"): BASELINE.json configs[3] -- ICESat-2-shaped synthetic tracks,
1e7 points on 100 x 100 km, z0 250 m (401^2), dz 1 km x 0.25 yr (101^2 x 21), per-(rgt, cycle) bias
columns, curvature / slope / d2/dt2 constraints, up to 10 robust edit iterations; it fits one B200.
One "step" = one complete smooth_fit(): design-matrix assembly (interp + stencil COO -> CSR), the explicit normal
matrix N = A^T W A, the edit loop of fp64 PCG solves on it (two-level preconditioner, bias columns eliminated),
residuals, sigma_extra (RDE + Brent) and 3-sigma masks, parse_model.

  value   points/s with the per-point inputs already resident in HBM (CUDA tensors handed to smooth_fit),
          timed with CUDA events between barrier + synchronize, max over ranks; no profiling events inside;
  e2e     the same call with HOST NumPy inputs (pinned): host->device copies of the inputs and the
          device->host copies of the model and per-point outputs are inside the timed region;
  roofline  achieved HBM GB/s of the dominant kernel (q = N p of the CG iteration) from CUDA event pairs the solver
          records on its launch stream in a SEPARATE untimed pass, against MEASURED_PEAKS.json; the CG iteration counts,
          the per-stage wall times and the bench-scale parity record ride in the same object (keys the driver keeps);
  cpu_baseline  the reference path on the CPU on a bounded sample of the same data set (rank 0, N = 1 only).

With N > 1 the normal matrix is row-sharded: every rank owns a range of dz-node blocks (its rows of N, built from its own
points plus halo copies), and the ranks exchange p inside the CG step kernel with stores into peer memory over NVLink --
no collective inside the iteration.  Strong scaling (total work fixed).  An untimed pass compares the N-GPU fit with a one-GPU fit of
the same points on rank 0 (roofline.parity).

`--impl reference` runs only the CPU arm (rank 0; other ranks exit 0): the reference's own smooth_fit under stub
dependencies when /root/reference exists (build container), else the oracle port (GPU box); SuiteSparseQR is replaced by
SuperLU on the normal equations (bias columns ordered last) + 3 refinement steps in both.  Each step is one fit of the
points of the cfg-4 data set inside a centred 10 x 10 km window (~1e5 points; the cost of the direct solve grows faster
than the window: a 31.6 km window with 1e6 points takes ~280 s per fit, profiles/README.md), and the line says so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'smooth_fit_points_per_s'
UNIT = 'points/s'
SAMPLE_WIDTH = 1.0e4     # CPU arms: the points of the cfg-4 data set inside a centred 10 x 10 km window


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--points', type=int, default=10_000_000)
    ap.add_argument('--width', type=float, default=1.0e5)
    ap.add_argument('--cg-rtol', type=float, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--max-outer', type=int, default=None, help='profiling only: cap the edit iterations')
    ap.add_argument('--cg-maxiter', type=int, default=None, help='profiling only: cap the CG iterations per solve')
    ap.add_argument('--skip-e2e', action='store_true', help='profiling only: no host-input pass')
    ap.add_argument('--stages', action='store_true', help='kept for old command lines (the stage split is always reported)')
    ap.add_argument('--no-parity', action='store_true', help='skip the untimed bench-scale parity record')
    ap.add_argument('--grid-bias', action='store_true',
                    help='add the setup_grid_bias term of SURVEY 8(d) cfg 4 (10 km grid x a per-point covariate) to the timed workload; '
                         'one GPU only (its columns belong to no preconditioner block, which the row-sharded solver needs)')
    ap.add_argument('--no-variant', action='store_true', help='skip the untimed pass that measures the fit with the grid-bias term')
    ap.add_argument('--ref-window', type=float, default=SAMPLE_WIDTH, help='CPU arms: width of the sample window in metres')
    return ap.parse_args()


def workload_config(a):
    if a.gpus == 1:
        par = 'single GPU'
    else:
        par = ('normal matrix row-sharded over %d GPUs by dz-node blocks (own points + halo copies per rank); p, p.q and the '
               'coarse-level packet exchanged inside the CG step kernel through peer memory over NVLink, no collective per '
               'iteration; per-point outputs gathered over NVLink and copied to the host on rank 0 (gather_outputs="root")' % a.gpus)
    return {'workload': 'cfg4: ICESat-2-shaped tracks, %d points, %.0f x %.0f km, z0 250 m, dz 1 km x 0.25 yr, '
                        "bias_params=['rgt','cycle']%s, max_iterations=10" % (a.points, a.width / 1e3, a.width / 1e3,
                                                                               ' + setup_grid_bias term (10 km grid)' if a.grid_bias else ''),
            'points': a.points, 'extent_km': a.width / 1e3, 'parallelism': par,
            'grid_bias_term': ('in the timed workload (covariate x 10 km grid)' if a.grid_bias else
                               'not in the timed workload (no call site in the reference smooth_fit; its columns cannot be owned by one '
                               'rank of the row-sharded solver, and the workload is the same at every GPU count); measured untimed in '
                               'roofline.grid_bias_variant at N = 1, timed with --grid-bias'),
            'l2_policy': 'inputs larger than L2 (N streamed per CG iteration is 0.37 GB at N=1, A is 1.7 GB)'}


def grid_bias_term(pts, width, seed):
    """SURVEY 8(d) cfg 4's `setup_grid_bias` term: a per-point covariate (seeded N(0, 1), e.g. a signal-strength anomaly), a
    smooth bias field proportional to it added to z, and the keyword arguments of the fit (10 km grid, expected_rms 0.1,
    expected_rms_grad 1e-3).  Returns (cov, z_with_bias, grid bounds)."""
    rng = np.random.default_rng(seed + 1)
    cov = rng.standard_normal(pts['x'].size)
    z = pts['z'] + 0.05 * np.cos(2 * np.pi * pts['x'] / 6e4) * np.sin(2 * np.pi * pts['y'] / 8e4) * cov
    return cov, z, (-width / 2, width / 2, -width / 2, width / 2)


GRID_BIAS = dict(spacing=1.0e4, expected_rms=0.1, expected_rms_grad=1e-3)


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(a, pts=None, kw=None):
    """The points of the bench data set inside a centred window, to be fitted on that window with the same spacings."""
    from lssurf_b200 import synthetic
    if pts is None:
        pts, kw = synthetic.icesat2_tracks(n=a.points, width=a.width, seed=a.seed)
    ws = min(a.ref_window, a.width)
    sel = (np.abs(pts['x']) < ws / 2) & (np.abs(pts['y']) < ws / 2)
    if a.grid_bias and 'cov' not in pts:
        cov, z_gb, _ = grid_bias_term(pts, a.width, a.seed)
        pts = dict(pts, z=z_gb, cov=cov)
    sub = {k: np.ascontiguousarray(np.asarray(v)[sel]) for k, v in pts.items()}
    kw = dict(kw)
    kw.pop('grid_bias_model_args', None)
    if a.grid_bias:           # the same 10 km grid, cut to the window
        kw['_grid_bias'] = dict(values=sub['cov'], bounds=(-ws / 2, ws / 2, -ws / 2, ws / 2), **GRID_BIAS)
    kw['W'] = dict(kw['W'], x=ws, y=ws)
    n_bias = int(np.unique(sub['rgt'] * 1000 + sub['cycle']).size)
    desc = ('the %d points of the bench data set inside the centred %.1f x %.1f km window, fitted on that window with the same '
            'spacings (%d bias columns), full edit loop' % (sub['x'].size, ws / 1e3, ws / 1e3, n_bias))
    return sub, kw, desc, n_bias


def cpu_kind(a=None):
    """'reference': the unmodified reference under stubs (build container); 'port': the oracle port (GPU box, and always for the
    --grid-bias workload: the live reference driver has no call site for setup_grid_bias, the oracle restates the composition and
    is pinned against the reference's own functions in oracle/make_golden.py)."""
    from oracle import ref_harness
    if a is not None and a.grid_bias:
        return 'port'
    return 'reference' if ref_harness.available() else 'port'


def cpu_step(pts, kw, kind, n_bias):
    """One pass of the reference path on the CPU.  kind 'reference': the unmodified reference smooth_fit under stub
    dependencies (oracle/ref_harness.py; build container only); 'port': the oracle port (oracle/lssurf_oracle.py).
    The only places bench.py executes oracle/ are this baseline leg and --impl reference."""
    kw = dict(kw)
    gbias = kw.pop('_grid_bias', None)
    t0 = time.perf_counter()
    if kind == 'reference' and gbias is not None:
        kind = 'port'             # the live reference driver has no call site for setup_grid_bias; the oracle restates the composition
    if kind == 'reference':
        from oracle import ref_harness as H
        L = H.load()
        import sparseqr
        sparseqr.TRAILING_DENSE = n_bias
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            S = L.smooth_fit(data=H.point_data(**pts), VERBOSE=False, **kw)
        return time.perf_counter() - t0, int(S['data'].size if hasattr(S['data'], 'size') else pts['x'].size)
    from oracle import lssurf_oracle as O
    res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'],
                       kw['E_RMS'], reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                       bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'], grid_bias=gbias)
    return time.perf_counter() - t0, res['sys']['n_data']


CPU_NOTE = ('SuiteSparseQR (not installable here) replaced by SuperLU on the normal equations, bias columns ordered last, + 3 '
            'refinement steps; SciPy sparse kernels and SuperLU are sequential apart from the BLAS-3 supernode updates, '
            'which may use `cores` threads')


def cpu_threads():
    """Let the BLAS pools use every host core (torchrun exports OMP_NUM_THREADS=1) and report the limit that is in force."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        pools = [p.get('num_threads', 1) for p in threadpoolctl.threadpool_info()]
        return int(max(pools)) if pools else 1
    except Exception:
        return 1


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = cpu_threads()
    pts, kw, desc, n_bias = cpu_sample(a)
    kind = cpu_kind(a)
    for _ in range(a.warmup):
        cpu_step(pts, kw, kind, n_bias)
    times = []
    for _ in range(a.steps):
        dt, n = cpu_step(pts, kw, kind, n_bias)
        times.append(dt)
    t = float(np.mean(times))
    val = n / t
    ws = min(a.ref_window, a.width)
    cfg = {'workload': 'cfg4 window: %s' % desc, 'points': int(n), 'extent_km': ws / 1e3, 'parallelism': 'host CPU, %d thread(s)' % cores,
           'sample_of': workload_config(a)['workload'],
           'note': 'value = points of the sample / seconds per fit of the sample; NOT a time for the full 1e7-point workload '
                   '(the direct solve grows faster than linearly with the window)'}
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': 1e3 * t, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': desc, 'note': CPU_NOTE},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------------- GPU arm
def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def spmv_bytes(n_rows, n_cols, nnz, with_scale):
    """Algorithmic bytes of one CSR product y = A x: 12 B per non-zero (fp64 value +
    int32 column), the row pointers, y written, x read once, plus the row-weight vector when it is applied."""
    return 12 * nnz + 4 * (n_rows + 1) + 8 * n_rows + 8 * n_cols + (8 * n_rows if with_scale else 0)


def run_b200(a):
    import torch
    import lssurf_b200 as L
    from lssurf_b200 import synthetic, parallel
    dist = parallel.init_from_env()
    rank, world = dist.rank, dist.world
    assert world == a.gpus or world == 1, f'--gpus {a.gpus} but WORLD_SIZE={world}'
    import torch.distributed as td

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    pts, kw = synthetic.icesat2_tracks(n=a.points, width=a.width, seed=a.seed)
    if a.cg_rtol is not None:
        kw['cg_rtol'] = a.cg_rtol
    if a.max_outer is not None:
        kw['max_iterations'] = a.max_outer
    if a.cg_maxiter is not None:
        kw['cg_maxiter'] = a.cg_maxiter
    n_pts = int(pts['x'].size)
    fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle')
    cov, z_gb, gb_bounds = grid_bias_term(pts, a.width, a.seed)

    def grid_bias_kw():
        g = L.fd_grid([[gb_bounds[0], gb_bounds[1]], [gb_bounds[2], gb_bounds[3]]], [GRID_BIAS['spacing']] * 2, name='cov_bias')
        return [dict(param='cov', grid=g, expected_rms=GRID_BIAS['expected_rms'], expected_value=0.,
                     expected_rms_grad=GRID_BIAS['expected_rms_grad'])]
    if a.grid_bias:
        if world > 1:
            raise SystemExit('bench.py --grid-bias runs on one GPU')
        pts = dict(pts, z=z_gb, cov=cov)
        fields = fields + ('cov',)
        kw['grid_bias_model_args'] = grid_bias_kw()
    # pinned host copies (e2e input) and device-resident copies (value input)
    host = {f: torch.from_numpy(pts[f]).pin_memory() for f in fields}
    dev = {f: host[f].cuda(non_blocking=True) for f in fields}
    h2d = sum(host[f].numel() * 8 for f in fields)
    dist_arg = dist if world > 1 else None
    if world == 1 and os.environ.get('LSS_SHARD_SINGLE'):       # diagnostics: the row-sharded solver with a local mailbox
        dist_arg = parallel.single_rank_context()

    def fit(inputs, profile=False, stages=False, history=None):
        D = L.point_data.data().from_dict({f: (inputs[f] if isinstance(inputs[f], torch.Tensor) and inputs[f].is_cuda
                                               else inputs[f].numpy()) for f in fields})
        # one job, one result: with several ranks the per-point outputs are gathered and copied to the host on rank 0 only
        return L.smooth_fit(data=D, VERBOSE=False, dist=dist_arg, profile_spmv=profile, time_stages=stages, history=history,
                            gather_outputs='root' if world > 1 else 'all', **kw)

    # ---- warm-up (device-resident steps, plus one host-input step)
    S = None
    for _ in range(a.warmup):
        S = fit(dev)                  # held like in the timed loops: the previous result is still alive while the next fit runs
    if not a.skip_e2e:
        S = fit(host)
    barrier()

    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    # ---- value: inputs resident in HBM, device time by CUDA events, no profiling events inside the solver
    launches0 = L._cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pcg = []
    barrier()
    ev0.record()
    for _ in range(a.steps):
        S = fit(dev)
        pcg += S['timing']['pcg']
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1e3
    launches = L._cabi.launch_count() - launches0
    # ---- e2e: host inputs through the public API, H2D + D2H inside the timed region
    barrier()
    w0 = time.perf_counter()
    for _ in range(0 if a.skip_e2e else a.steps):
        S = fit(host)
        _ = float(S['m']['all'].sum())             # the model vector is already host memory (D2H done inside)
    torch.cuda.synchronize()
    t_e2e = max(time.perf_counter() - w0, 1e-9)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    n_cols = int(S['m']['all'].size)
    n_in_tse = int(np.sum(S['data'].three_sigma_edit)) if rank == 0 else 0
    S = None                          # results of the timed loops are released: the untimed passes below must not grow the pinned pool
    # ---- untimed passes: kernel timings (event pairs around every 4th CG iteration), wall time per stage with device syncs
    Sp = fit(dev, profile=True)
    prof = Sp['timing']['pcg']
    Sp = None
    barrier()
    Sd = fit(host, stages=True)               # twice: the first pass after the buffers above were released re-grows the allocator pools
    Sd = None
    Sd = fit(host, stages=True)
    stage_times = {k: round(v, 4) for k, v in Sd['timing']['stages'].items()}
    Sd = None
    barrier()
    # ---- untimed multi-GPU parity record: the row-sharded fit of all ranks against a one-GPU fit of the same points on rank 0
    # (what tests/dist_check.py checks at small sizes): masks after every outer iteration, final mask, model
    parity_multi = None
    if world > 1 and not a.no_parity:
        try:
            hm = []
            Sm = fit(dev, history=hm)
            if rank == 0:
                h1 = []
                D1 = L.point_data.data().from_dict({f: dev[f] for f in fields})
                S1 = L.smooth_fit(data=D1, VERBOSE=False, history=h1, **kw)
                mm, m1 = Sm['m']['all'], S1['m']['all']
                parity_multi = {
                    'against': 'the same fit on one GPU (rank 0, dist=None) on the same 1e7 points',
                    'masks_identical_per_iteration': bool(len(hm) == len(h1) and all(x['mask_out'].size == y['mask_out'].size and
                                                                                     np.array_equal(x['mask_out'], y['mask_out']) for x, y in zip(hm, h1))),
                    'final_mask_identical': bool(np.array_equal(Sm['data'].three_sigma_edit, S1['data'].three_sigma_edit)),
                    'rel_l2_m0': float(np.linalg.norm(mm - m1) / np.linalg.norm(m1)),
                    'max_abs_z_est': float(np.max(np.abs(np.asarray(Sm['data'].z_est) - np.asarray(S1['data'].z_est)))),
                    'cg_iterations': [[h['pcg']['iterations'] for h in hh] for hh in (hm, h1)]}
                S1 = D1 = None
            Sm = None
        except Exception as e:       # noqa: BLE001 -- a record, not a gate: the timed numbers above stand either way
            parity_multi = {'error': repr(e)[:300]} if rank == 0 else None
        barrier()
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device='cuda')
        td.all_reduce(t, op=td.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
    # m0 + three_sigma_edit, sigma_extra, z_est, bias_ID (rank 0; with several ranks the others copy m0 only)
    d2h = 8 * n_cols + n_pts * (1 + 8 + 8 + (8 if kw.get('bias_params') else 0))

    # ---- roofline of the dominant kernel (q = N p) from the solver's own event pairs (this rank's launches)
    n_apply = sum(p['n_apply'] for p in prof)
    ms_prod = sum(p['ms_At'] for p in prof) / max(n_apply, 1)
    ms_vec = sum(p.get('ms_step', 0.) for p in prof) / max(n_apply, 1)
    p0 = prof[0]
    peak, peak_src = measured_peak()
    form = p0.get('form', '')
    nnzN = p0.get('nnz_N', 0)
    nnzP = p0.get('nnz_product') or nnzN          # the rows of the eliminated bias block are not part of the CG product
    rows_P = n_cols - p0.get('eliminated', 0)
    traffic = None
    tfile = os.path.join(ROOT, 'profiles', 'spmv_traffic.json')
    if form != 'explicit_normal':
        raise SystemExit('bench.py: the workload is expected to run on the explicit normal matrix, got form %r' % form)
    if p0.get('structured_bytes'):
        kern = ('lss::sn_product_kernel<21>: q = N p from the grid-structured layout of N (TMA-pipelined node blocks and z0-dz pair '
                'records, z0-z0 diagonals, flat remainder; one launch)')
        b_dom = int(p0['structured_bytes'])
        b_note = 'bytes of the structured layout one product streams (lssurf_b200/structured.py product_bytes): values 8 B, tables, flat remainder 16 B / entry, q and p once'
        tkey = 'structured_normal'
    else:
        if world > 1:
            rows_P = max(rows_P // world, 1)       # this rank's rows of N (dz-node blocks split evenly)
        kern = ('lss::spmv_persistent_kernel<128,8,false,true> on the explicit normal matrix (q = N p, N = A^T W A; one launch, '
                'atomic carries, p.q by-product)' + ('; this rank\'s rows only' if world > 1 else ''))
        b_dom = spmv_bytes(rows_P, n_cols, nnzP, False)
        b_note = 'SURVEY 8(d) CSR formula applied to the rows of N one launch covers: 12 nnz + 4 (rows + 1) + 8 rows + 8 cols'
        tkey = 'explicit_normal'
    if os.path.exists(tfile):       # ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by kernel
        traffic = json.load(open(tfile)).get(tkey if world == 1 else '', {}).get('dram_bytes_per_launch')
    ach = b_dom / (ms_prod / 1e3) / 1e9 if ms_prod > 0 else 0.0
    outer = len(pcg) / a.steps
    iters = sum(p['iterations'] for p in pcg) / a.steps
    solve_ms = sum(p['ms'] for p in pcg) / a.steps
    solve = {'outer_iterations': outer, 'cg_iterations_per_step': iters, 'pcg_ms_per_step': solve_ms,
             'ms_per_cg_iteration': solve_ms / max(iters, 1), 'ms_product_per_iteration': ms_prod,
             'ms_vector_step_per_iteration': ms_vec,
             'preconditioner': {'blocks': p0.get('precond_blocks', 0), 'eliminated_columns': p0.get('eliminated', 0),
                                'coarse_unknowns': p0.get('coarse', 0)},
             'cg_rtol': kw.get('cg_rtol', 1e-12), 'in_TSE': n_in_tse, 'unknowns': n_cols,
             'nnz_A_per_rank': p0['nnz'], 'cg_iterations': [p['iterations'] for p in pcg[:int(outer)]],
             'converged': all(p.get('converged', True) for p in pcg)}
    roofline = {'bound': 'hbm', 'kernel': kern, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_source': peak_src,
                'traffic': traffic, 'bytes_per_launch': b_dom, 'bytes_definition': b_note, 'launch_ms': ms_prod,
                'launch_ms_samples': n_apply,
                'launch_ms_sampling': 'CUDA event pairs on the solver stream around every 4th CG iteration of one untimed profiling step',
                'nnz_N': nnzN, 'nnz_streamed_per_product': nnzP, 'frac_of_nominal_8TBs': ach / 8000.0,
                'csr_equivalent': {'bytes': spmv_bytes(rows_P, n_cols, nnzP, False),
                                   'gbs': spmv_bytes(rows_P, n_cols, nnzP, False) / (ms_prod / 1e3) / 1e9 if ms_prod > 0 else None,
                                   'note': 'bytes of the same rows of N as CSR divided by the launch time (equals achieved for the CSR kernel)'},
                'solve': solve, 'stages_s': stage_times}

    # ---- untimed bench-scale parity record: default form vs the deterministic CSR-pair form (A and A^T as CSR, Jacobi) on the
    # same points -- masks after every outer iteration, model, points close to the 3-sigma threshold
    parity = None
    if rank == 0 and world == 1 and not a.no_parity:
        h0, h1 = [], []
        Sa = fit(dev, history=h0)
        os.environ['LSS_NO_FUSED'] = '1'
        try:
            Sb = fit(dev, history=h1)
        finally:
            os.environ.pop('LSS_NO_FUSED', None)
        ma, mb = Sa['m']['all'], Sb['m']['all']
        parity = {'against': 'generic CSR-pair solver form (LSS_NO_FUSED=1: two deterministic CSR products per iteration, Jacobi) on the same 1e7 points',
                  'form_b': Sb['timing']['pcg'][0].get('form'),
                  'masks_identical_per_iteration': bool(len(h0) == len(h1) and all(np.array_equal(x['mask_out'], y['mask_out']) for x, y in zip(h0, h1))),
                  'final_mask_identical': bool(np.array_equal(Sa['data'].three_sigma_edit, Sb['data'].three_sigma_edit)),
                  'rel_l2_m0': float(np.linalg.norm(ma - mb) / np.linalg.norm(mb)),
                  'near_threshold': [int(v) for v in Sa['timing'].get('near_threshold', [])],
                  'near_threshold_definition': 'points whose scaled residual lies within 1e-6 of the 3-sigma threshold, per outer iteration',
                  'cg_iterations': [[h['pcg']['iterations'] for h in hh] for hh in (h0, h1)]}
        roofline['parity'] = parity
    elif parity_multi is not None:
        parity = parity_multi
        roofline['parity'] = parity

    # ---- untimed variant: the same fit WITH the setup_grid_bias term of cfg 4 (the timed workload leaves it out: the reference's
    # live smooth_fit has no call site for it and the row-sharded solver cannot own its columns), so its cost is on record
    if rank == 0 and world == 1 and not a.grid_bias and not a.no_variant:
        vdev = dict(dev, z=torch.from_numpy(z_gb).cuda(), cov=torch.from_numpy(cov).cuda())
        vfields = fields + ('cov',)

        def vfit():
            D = L.point_data.data().from_dict({f: vdev[f] for f in vfields})
            return L.smooth_fit(data=D, VERBOSE=False, **dict(kw, grid_bias_model_args=grid_bias_kw()))
        Sv = vfit()
        Sv = None
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        Sv = vfit()
        torch.cuda.synchronize()
        t_v = time.perf_counter() - w0
        gb = Sv['m']['grid_biases']['cov_bias']
        truth = 0.05 * np.cos(2 * np.pi * gb.x[None, :] / 6e4) * np.sin(2 * np.pi * gb.y[:, None] / 8e4)
        roofline['grid_bias_variant'] = {
            'what': 'one untimed fit of the same points with the setup_grid_bias term (covariate x 10 km grid, 121 columns, expected_rms 0.1, '
                    'expected_rms_grad 1e-3; rows of 17 entries); bench.py --grid-bias times this workload',
            'ms_per_step': 1e3 * t_v, 'cg_iterations': [p['iterations'] for p in Sv['timing']['pcg']],
            'form': Sv['timing']['pcg'][0].get('form'), 'nnz_N': Sv['timing']['pcg'][0].get('nnz_N'),
            'bias_field_correlation_with_truth': float(np.corrcoef(gb.cov_bias.ravel(), truth.ravel())[0, 1])}
        Sv = vdev = None

    cpu = None
    if rank == 0 and a.gpus == 1 and not a.no_cpu_baseline:
        cores = cpu_threads()
        cp, ckw, desc, n_bias = cpu_sample(a, pts, kw)
        kind = cpu_kind(a)
        dt, n = cpu_step(cp, ckw, kind, n_bias)
        cpu = {'value': n / dt, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': desc, 'seconds': dt,
               'host_cores_available': os.cpu_count(), 'note': CPU_NOTE}
    if rank == 0:
        line = {'metric': METRIC, 'value': n_pts * a.steps / t_dev, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': 1e3 * t_dev / a.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(a),
                'e2e': {'value': n_pts * a.steps / t_e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': 1e3 * t_e2e / a.steps},
                'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
                'solve': solve, 'stages_s': stage_times, 'parity': parity}
        print(json.dumps(line), flush=True)
    dist.destroy()
    if world > 1:
        td.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)


if __name__ == '__main__':
    main()
