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
