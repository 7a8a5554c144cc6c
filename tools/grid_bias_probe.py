"""GPU probe: the cfg-4 fit with and without the setup_grid_bias term (10 km grid x a per-point covariate)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lssurf_b200 as L
from lssurf_b200 import synthetic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
width = 1.e5 * (n / 1e7) ** 0.5
pts, kw = synthetic.icesat2_tracks(n=n, width=width, seed=0)
rng = np.random.default_rng(1)
pts['cov'] = rng.standard_normal(pts['x'].size)
pts['z'] = pts['z'] + 0.05 * np.cos(2 * np.pi * pts['x'] / 6e4) * np.sin(2 * np.pi * pts['y'] / 8e4) * pts['cov']
fields = ('x', 'y', 'time', 'z', 'sigma', 'sigma_corr', 'rgt', 'cycle', 'cov')
dev = {f: torch.from_numpy(pts[f]).cuda() for f in fields}


def fit(with_term, **extra):
    D = L.point_data.data().from_dict({f: dev[f] for f in fields})
    k = dict(kw, **extra)
    if with_term:
        g = L.fd_grid([[-width / 2, width / 2], [-width / 2, width / 2]], [1.e4, 1.e4], name='cov_bias')
        k['grid_bias_model_args'] = [dict(param='cov', grid=g, expected_rms=0.1, expected_value=0., expected_rms_grad=1e-3)]
    return L.smooth_fit(data=D, VERBOSE=False, **k)


res = {}
for with_term in (False, True):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        S = fit(with_term)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    p = S['timing']['pcg']
    print('grid bias' if with_term else 'plain    ', '%.1f ms' % (dt * 1e3), 'form', p[0].get('form'), 'iterations', [q['iterations'] for q in p],
          'ms/it %.4f' % (sum(q['ms'] for q in p) / sum(q['iterations'] for q in p)), 'nnz_N', p[0].get('nnz_N'), 'in_TSE', int(S['data'].three_sigma_edit.sum()),
          'converged', all(q.get('converged', True) for q in p))
    res[with_term] = S
    if with_term:
        gb = S['m']['grid_biases']['cov_bias'].cov_bias
        print('recovered bias field: rms %.4f, corr with truth %.3f' % (float(np.sqrt(np.mean(gb ** 2))),
              float(np.corrcoef(gb.ravel(), (0.05 * np.cos(2 * np.pi * S['m']['grid_biases']['cov_bias'].x[None, :] / 6e4) * np.sin(2 * np.pi * S['m']['grid_biases']['cov_bias'].y[:, None] / 8e4)).ravel())[0, 1])))
St = fit(True, time_stages=True)
print('stages', {k: round(v * 1e3, 1) for k, v in St['timing']['stages'].items() if v >= 0.001})
