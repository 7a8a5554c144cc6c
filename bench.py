#!/usr/bin/env python
"""Benchmark of the smooth_fit solve hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--points P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload = "cfg4"): BASELINE.json configs[3] -- ICESat-2-shaped synthetic tracks,
1e7 points on 100 x 100 km, z0 250 m (401^2), dz 1 km x 0.25 yr (101^2 x 21), per-(rgt, cycle) bias
columns, curvature / slope / d2/dt2 constraints, up to 10 robust edit iterations; it fits one B200.
One "step" = one complete smooth_fit(): design-matrix assembly (interp + stencil COO -> CSR), the
edit loop of matrix-free PCG solves, residuals, sigma_extra (RDE + Brent) and 3-sigma masks.

  value   points/s with the per-point inputs already resident in HBM (CUDA tensors handed to smooth_fit),
          timed with CUDA events between barrier + synchronize, max over ranks;
  e2e     the same call with HOST NumPy inputs (pinned): host->device copies of the inputs and the
          device->host copies of the model and per-point outputs are inside the timed region;
  roofline  achieved HBM GB/s of the dominant kernel (the nnz-tiled CSR product, used for A p and A^T u)
          from CUDA events recorded by the solver around every product on its launch stream, against
          MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle port of the reference path (NumPy/SciPy, SuperLU stand-in for SPQR) on a
          bounded sample of the same generator, timed on this host (rank 0, N = 1 only).

`--impl reference` runs only that CPU arm (rank 0; other ranks exit 0).  With N > 1 the data rows (and
the constraint rows) are sharded over the ranks: strong scaling, one NCCL all-reduce per CG iteration.
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
SAMPLE_WIDTH = 5.0e3     # CPU arms: a 5 x 5 km window of the same generator at the same point density


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
    ap.add_argument('--stages', action='store_true', help='add an untimed pass reporting wall time per stage')
    return ap.parse_args()


def workload_config(a):
    return {'workload': 'cfg4: ICESat-2-shaped tracks, %d points, %.0f x %.0f km, z0 250 m, dz 1 km x 0.25 yr, '
                        "bias_params=['rgt','cycle'], max_iterations=10" % (a.points, a.width / 1e3, a.width / 1e3),
            'points': a.points, 'parallelism': 'row-sharded dp%d' % a.gpus,
            'l2_policy': 'inputs larger than L2 (the normal matrix streamed per CG iteration is ~0.75 GB per rank at N=1, A is 1.7 GB)'}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(a):
    from lssurf_b200 import synthetic
    n = max(2000, int(round(a.points * (SAMPLE_WIDTH / a.width) ** 2)))
    pts, kw = synthetic.icesat2_tracks(n=n, width=SAMPLE_WIDTH, seed=a.seed)
    desc = ('%d-point %.0f x %.0f km window of the same track generator and spacings (same point density), full '
            'edit loop' % (pts['x'].size, SAMPLE_WIDTH / 1e3, SAMPLE_WIDTH / 1e3))
    return pts, kw, desc


def cpu_step(pts, kw):
    """One pass of the reference path on the CPU: the oracle port (oracle/lssurf_oracle.py).  The only
    places bench.py executes oracle/ are this baseline leg and --impl reference."""
    from oracle import lssurf_oracle as O
    t0 = time.perf_counter()
    res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'],
                       kw['E_RMS'], reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                       bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
    return time.perf_counter() - t0, res['sys']['n_data']


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pts, kw, desc = cpu_sample(a)
    for _ in range(min(a.warmup, 1)):          # the CPU path has no warm-up effects worth 3 x 12 s
        cpu_step(pts, kw)
    times = []
    for _ in range(a.steps):
        dt, n = cpu_step(pts, kw)
        times.append(dt)
    t = float(np.mean(times))
    val = n / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': min(a.warmup, 1), 'ms_per_step': 1e3 * t, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(a),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': desc,
                             'note': 'oracle port of the reference python (SciPy is single-threaded); SuiteSparseQR '
                                     'replaced by SuperLU on the normal equations + 3 refinement steps'},
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
    # pinned host copies (e2e input) and device-resident copies (value input)
    host = {f: torch.from_numpy(pts[f]).pin_memory() for f in fields}
    dev = {f: host[f].cuda(non_blocking=True) for f in fields}
    h2d = sum(host[f].numel() * 8 for f in fields)
    dist_arg = dist if world > 1 else None
    if world == 1 and os.environ.get('LSS_SHARD_SINGLE'):       # diagnostics: the row-sharded solver with a local mailbox
        dist_arg = parallel.single_rank_context()

    def fit(inputs, profile, stages=False):
        D = L.point_data.data().from_dict({f: (inputs[f] if isinstance(inputs[f], torch.Tensor) and inputs[f].is_cuda
                                               else inputs[f].numpy()) for f in fields})
        return L.smooth_fit(data=D, VERBOSE=False, dist=dist_arg, profile_spmv=profile, time_stages=stages, **kw)

    # ---- warm-up (device-resident steps, plus one host-input step)
    S = None
    for _ in range(a.warmup):
        S = fit(dev, False)           # held like in the timed loops: the previous result is still alive while the next fit runs
    if not a.skip_e2e:
        S = fit(host, False)
    barrier()

    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    # ---- value: inputs resident in HBM, device time by CUDA events
    launches0 = L._cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pcg = []
    barrier()
    ev0.record()
    for _ in range(a.steps):
        S = fit(dev, True)
        pcg += S['timing']['pcg']
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1e3
    launches = L._cabi.launch_count() - launches0
    # ---- e2e: host inputs through the public API, H2D + D2H inside the timed region
    barrier()
    w0 = time.perf_counter()
    for _ in range(0 if a.skip_e2e else a.steps):
        S = fit(host, False)
        _ = float(S['m']['all'].sum())             # the model vector is already host memory (D2H done inside)
    torch.cuda.synchronize()
    t_e2e = max(time.perf_counter() - w0, 1e-9)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    stage_times = None
    if a.stages:                                   # untimed diagnostic pass: wall time per stage with device syncs
        Sd = fit(host, False, stages=True)
        stage_times = {k: round(v, 4) for k, v in Sd['timing']['stages'].items()}
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device='cuda')
        td.all_reduce(t, op=td.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
    d2h = 8 * int(S['m']['all'].size) + n_pts * (1 + 8 + 8)      # m0 + three_sigma_edit, sigma_extra, z_est

    # ---- roofline of the dominant kernel from the solver's own event pairs (this rank's launches)
    n_apply = sum(p['n_apply'] for p in pcg)
    ms_a, ms_at = sum(p['ms_A'] for p in pcg), sum(p['ms_At'] for p in pcg)
    p0 = pcg[0]
    nnz = p0['nnz']
    n_cols = int(S['m']['all'].size)
    peak, peak_src = measured_peak()
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'spmv_traffic.json')
    if os.path.exists(prof):       # ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by operator form
        traffic = json.load(open(prof)).get(p0.get('form', ''), {}).get('dram_bytes_per_launch')
    if p0.get('form') == 'explicit_normal':
        # explicit form: q = N p with the tiled CSR kernel over N = A^T W A: 12 B per non-zero of N, row pointers,
        # q written and p read once (SURVEY 8(d) formula applied to N).  The two-product CSR form of the same
        # operator (what SURVEY 8(d) budgets per CG iteration) is reported alongside for comparison.
        nnzN = p0['nnz_N']
        nnzP = p0.get('nnz_product') or nnzN          # the rows of the eliminated bias block are not part of the CG product
        b_dom = spmv_bytes(n_cols - p0.get('eliminated', 0), n_cols, nnzP, False)
        ms_dom = ms_at / max(n_apply, 1)
        ach = b_dom / (ms_dom / 1e3) / 1e9 if ms_dom > 0 else 0.0
        b_two = spmv_bytes(p0['rows'], n_cols, nnz, True) + spmv_bytes(n_cols, p0['rows'], nnz, False)
        roofline = {'bound': 'hbm', 'kernel': 'lss::spmv_persistent_kernel<128,8,false,true> on the explicit normal matrix (q = N p, N = A^T W A; one launch, '
                              'atomic carries, p.q by-product)',
                    'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_source': peak_src,
                    'traffic': traffic, 'bytes_per_launch': {'N': b_dom}, 'launch_ms': {'N': ms_dom},
                    'launch_ms_samples': n_apply, 'launch_ms_sampling': 'CUDA event pairs on the solver stream around every 4th CG iteration of the timed steps',
                    'nnz_N': nnzN, 'nnz_streamed_per_product': nnzP, 'nnz_A': nnz,
                    'csr_pair_equivalent': {'bytes': b_two, 'gbs': b_two / (ms_dom / 1e3) / 1e9 if ms_dom > 0 else None,
                                            'note': 'bytes the A / A^T two-product form of the same operator would stream, '
                                                    'divided by the time of the N product (not a bandwidth claim)'},
                    'frac_of_nominal_8TBs': ach / 8000.0}
    elif p0.get('fused'):
        # block form: ONE pass over the ELL copy of G does q += G^T (w .* (G p)): 12 B per entry, the row weights,
        # x read once and q read-modify-written once (compulsory); the constraint pair (C, C^T) is the small rest
        if p0.get('all_ell'):      # constraint rows ride in the same ELL block: one kernel per operator application
            b_dom = 12 * (p0['nnz_G'] + p0['nnz_C']) + 8 * (p0['n_data'] + p0['n_con']) + 24 * n_cols
            b_con = 0
        else:
            b_dom = 12 * p0['nnz_G'] + 8 * p0['n_data'] + 24 * n_cols
            b_con = spmv_bytes(p0['n_con'], n_cols, p0['nnz_C'], True) + spmv_bytes(n_cols, p0['n_con'], p0['nnz_C'], False)
        ms_dom, ms_con = ms_at / n_apply, ms_a / n_apply
        ach = b_dom / (ms_dom / 1e3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': 'lss::ell_normal_kernel<13,0> (fused q += A^T (w .* (A p)), one pass over the ELL rows)',
                    'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_source': peak_src,
                    'traffic': traffic, 'bytes_per_launch': {'fused_G': b_dom, 'C_and_Ct': b_con},
                    'launch_ms': {'fused_G': ms_dom, 'C_and_Ct': ms_con},
                    'gbs': {'fused_G': ach, 'C_and_Ct': b_con / (ms_con / 1e3) / 1e9 if ms_con > 0 else None},
                    'frac_of_nominal_8TBs': ach / 8000.0}
    else:
        rows_loc = p0['rows']
        b_a = spmv_bytes(rows_loc, n_cols, nnz, True)
        b_at = spmv_bytes(n_cols, rows_loc, nnz, False)
        ach = (b_a + b_at) * n_apply / ((ms_a + ms_at) / 1e3) / 1e9 if (ms_a + ms_at) > 0 else None
        roofline = {'bound': 'hbm', 'kernel': 'lss::spmv_persistent_kernel (y = w.*(A p) and q = A^T u)', 'achieved': ach, 'peak': peak,
                    'unit': 'GB/s', 'frac': None if ach is None else ach / peak, 'peak_source': peak_src, 'traffic': traffic,
                    'bytes_per_launch': {'A': b_a, 'At': b_at}, 'launch_ms': {'A': ms_a / n_apply, 'At': ms_at / n_apply},
                    'gbs': {'A': b_a / (ms_a / n_apply / 1e3) / 1e9, 'At': b_at / (ms_at / n_apply / 1e3) / 1e9},
                    'frac_of_nominal_8TBs': None if ach is None else ach / 8000.0}
    outer = len(pcg) / a.steps
    iters = sum(p['iterations'] for p in pcg) / a.steps
    solve_ms = sum(p['ms'] for p in pcg) / a.steps

    cpu = None
    if rank == 0 and a.gpus == 1 and not a.no_cpu_baseline:
        cp, ckw, desc = cpu_sample(a)
        dt, n = cpu_step(cp, ckw)
        cpu = {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': desc, 'seconds': dt,
               'host_cores_available': os.cpu_count(),
               'note': 'oracle port of the reference python; SciPy sparse + SuperLU are single-threaded; SuiteSparseQR '
                       'replaced by SuperLU on the normal equations + 3 refinement steps'}
    if rank == 0:
        line = {'metric': METRIC, 'value': n_pts * a.steps / t_dev, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': 1e3 * t_dev / a.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(a),
                'e2e': {'value': n_pts * a.steps / t_e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': 1e3 * t_e2e / a.steps},
                'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
                'solve': {'outer_iterations': outer, 'cg_iterations_per_step': iters, 'pcg_ms_per_step': solve_ms,
                          'ms_per_cg_iteration': solve_ms / max(iters, 1),
                          'ms_product_per_iteration': ms_at / max(n_apply, 1), 'ms_vector_step_per_iteration': sum(p.get('ms_step', 0.) for p in pcg) / max(n_apply, 1),
                          'preconditioner': {'blocks': p0.get('precond_blocks', 0), 'eliminated_columns': p0.get('eliminated', 0)},
                          'cg_rtol': kw.get('cg_rtol', 1e-12),
                          'in_TSE': int(np.sum(S['data'].three_sigma_edit)), 'unknowns': n_cols, 'nnz_A_per_rank': nnz,
                          'cg_iterations': [p['iterations'] for p in pcg[:int(outer)]]},
                'stages_s': stage_times}
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
