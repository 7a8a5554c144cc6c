"""Pins the oracle (oracle/lssurf_oracle.py) against the unmodified reference on random configurations, beyond the four
fixed cases of oracle/make_golden.py (build container only).  Only the keywords the oracle restates are drawn: the E_RMS
mix, reference_epoch, max_iterations, min_iterations, converge_tol_dz, sigma_extra_relax, per-(rgt, cycle) biases, NaN heights
and points outside the grids.  Required: G_data / Gc as canonical CSR and Ec bit-identical, the same final edit mask,
rel-L2(m0) < 1e-9.
    python tools/oracle_fuzz.py [n_cases=12] [seed=0]"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lssurf_oracle as O
from oracle import ref_harness as H
from tests.util import csr_digest, digest


def one_case(rng):
    n = int(rng.integers(300, 2500))
    W = float(rng.choice([6e3, 8e3, 1.2e4]))
    t_span = float(rng.choice([1., 2., 3.]))
    x, y, t = rng.uniform(-W / 2 * 1.05, W / 2 * 1.05, n), rng.uniform(-W / 2, W / 2, n), rng.uniform(-t_span / 2 * 1.1, t_span / 2 * 1.1, n)
    z = 30 * np.sin(x / 3000.) * np.cos(y / 2500.) + t * (1 + x / W) + rng.normal(0, 0.1, n)
    bad = rng.random(n) < 0.05
    z[bad] += rng.normal(0, 10, int(bad.sum()))
    if rng.random() < 0.3:
        z[rng.integers(0, n, 3)] = np.nan
    pts = dict(x=x, y=y, time=t, z=z, sigma=rng.choice([0.1, 0.2], n))
    E = {'d2z0_dx2': 0.02, 'dz0_dx': 50, 'd3z_dx2dt': 3e-4, 'd2z_dxdt': 0.75, 'd2z_dt2': 5000}
    for key in ('dz0_dx', 'd3z_dx2dt', 'd2z_dxdt', 'd2z_dt2'):
        if rng.random() < 0.2:
            E[key] = E[key] * 3
    if rng.random() < 0.15:
        E['d2z_dxdt'] = None
    dt = float(rng.choice([0.25, 0.5]))
    n_ep = int(round(t_span / dt))
    kw = dict(W={'x': W, 'y': W, 't': t_span}, ctr={'x': 0., 'y': 0., 't': 0.},
              spacing={'z0': float(rng.choice([250., 500.])), 'dz': 1000., 'dt': dt}, E_RMS=E,
              reference_epoch=int(rng.integers(0, n_ep + 1)), max_iterations=int(rng.choice([1, 2, 4, 10])))
    if rng.random() < 0.5:
        pts.update(rgt=rng.integers(0, 6, n).astype(float), cycle=rng.integers(0, 3, n).astype(float), sigma_corr=np.full(n, 0.05))
        kw['bias_params'] = ['rgt', 'cycle']
    if rng.random() < 0.3:
        kw['sigma_extra_relax'] = True
    if rng.random() < 0.3:
        kw['min_iterations'] = int(rng.integers(0, 4))
        kw['converge_tol_dz'] = float(rng.choice([0.01, 0.05, 1.0]))
    return pts, kw


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    R = H.load()
    bad = 0
    for k in range(n_cases):
        pts, kw = one_case(rng)
        tag = 'case %d: n=%d %s' % (k, pts['x'].size, {a: b for a, b in kw.items() if a not in ('W', 'ctr')})
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            fo = R.smooth_fit(data=H.point_data(**pts), VERBOSE=False, return_fit_objects=True, **kw)
            out = R.smooth_fit(data=H.point_data(**pts), VERBOSE=False, **kw)
            extra = {a: kw[a] for a in ('min_iterations', 'converge_tol_dz', 'sigma_extra_relax') if a in kw}
            res = O.smooth_fit(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
                               reference_epoch=kw['reference_epoch'], max_iterations=kw['max_iterations'],
                               bias_params=[pts[p] for p in kw['bias_params']] if 'bias_params' in kw else None,
                               sigma_corr=pts.get('sigma_corr'), **extra)
        s = res['sys']
        m_ref = out['m']['all']
        checks = dict(G=digest(*s['G'][:3]) == csr_digest(fo['G_data'].toCSR()), Gc=digest(*s['C'][:3]) == csr_digest(fo['Gc'].toCSR()),
                      Ec=digest(s['Ec']) == digest(np.asarray(fo['Ec'], dtype=float)),
                      mask=np.array_equal(np.asarray(out['data'].three_sigma_edit, dtype=bool), res['in_TSE']),
                      m0=np.linalg.norm(res['m0'] - m_ref) <= 1e-9 * np.linalg.norm(m_ref),
                      sigma_extra=np.max(np.abs(np.asarray(out['data'].sigma_extra, dtype=float) - res['sigma_extra'])) < 2e-5)
        failed = [c for c, v in checks.items() if not v]
        print(('ok   ' if not failed else 'DIFF ') + tag + ('' if not failed else '  FAILED: %s' % failed))
        bad += bool(failed)
    print('%d of %d cases differ' % (bad, n_cases))
    return bad


if __name__ == '__main__':
    sys.exit(1 if main() else 0)
