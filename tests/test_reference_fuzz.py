"""CPU tier, build container only: short runs of the differential fuzzers in tools/ against the UNMODIFIED reference (under the
stub dependencies of oracle/ref_stubs), side by side in three processes:
  diff_fuzz.py            the package's host logic (kernel wrappers replaced by tests/host_double.py) on random keyword mixes --
  diff_fuzz.py widened    operators, edit masks, models, parse_model outputs, error types and messages must agree;
  oracle_fuzz.py          the oracle: canonical CSR of G_data / Gc and Ec bit-identical, same final mask, rel-L2(m0) < 1e-9.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/LSsurf'), reason='needs the reference checkout')


def test_differential_fuzz_against_the_reference():
    env = dict(os.environ, OMP_NUM_THREADS='2', OPENBLAS_NUM_THREADS='2')
    runs = {'diff': ['diff_fuzz.py', '3', '4242'], 'diff widened': ['diff_fuzz.py', '3', '4242', 'widened'],
            'oracle': ['oracle_fuzz.py', '3', '4242']}
    procs = {k: subprocess.Popen([sys.executable, os.path.join(ROOT, 'tools', v[0]), *v[1:]], cwd=ROOT, env=env,
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for k, v in runs.items()}
    for k, p in procs.items():
        try:
            out = p.communicate(timeout=900)[0]
        except subprocess.TimeoutExpired:
            for q in procs.values():
                q.kill()
            raise
        assert p.returncode == 0 and '0 of 3 cases differ' in out, k + '\n' + out[-3000:]


def test_error_propagation_without_a_fit_like_the_reference():
    """compute_E with max_iterations = 0 (smooth_fit.py:689-712 runs outside the iteration branch): the reference propagates errors
    for the initial edit mask with the sigma_extra the data carry, and fails on data.z_est when they carry none; the package
    (host double) must do the same, every entry of E within 1e-4.  Found by tools/diff_fuzz.py (widened more compute_E)."""
    import warnings

    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import diff_fuzz as F
    import lssurf_b200 as L
    from oracle import ref_harness as H
    from tests import host_double
    from tests.util import load_case
    R = H.load()
    fx, pts, kw = load_case('compute_E')
    kw = dict(kw, max_iterations=0)
    rng = np.random.default_rng(0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with pytest.raises(AttributeError, match="has no attribute 'z_est'"):
            R.smooth_fit(data=H.point_data(**pts), VERBOSE=False, **kw)
        with host_double.installed(), pytest.raises(AttributeError, match="has no attribute 'z_est'"):
            L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, sort_points=False, **kw)
        p = dict(pts, sigma_extra=rng.choice([0., 0.05], pts['x'].size), three_sigma_edit=(rng.random(pts['x'].size) < 0.9).astype(float))
        S_r = R.smooth_fit(data=H.point_data(**p), VERBOSE=False, **kw)
        with host_double.installed():
            S_p = L.smooth_fit(data=L.point_data.data().from_dict(p), VERBOSE=False, sort_points=False, **kw)
    assert S_r['m'] == {} and S_p['m'] == {} and len(S_r['E']) == 9 and set(S_r['E']) == set(S_p['E'])
    assert all(F._same_E(key, g, S_p['E']) for key, g in S_r['E'].items())
