"""CPU tier, build container only: a short run of tools/diff_fuzz.py -- the UNMODIFIED reference (under the stub dependencies of
oracle/ref_stubs) and the package's host logic (kernel wrappers replaced by tests/host_double.py) side by side on random keyword
mixes, plain and widened.  Operators, edit masks, models, parse_model outputs, error types and messages must agree.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/LSsurf'), reason='needs the reference checkout')


@pytest.mark.parametrize('mode', [(), ('widened',)])
def test_differential_fuzz_against_the_reference(mode):
    env = dict(os.environ, OMP_NUM_THREADS='4', OPENBLAS_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'diff_fuzz.py'), '3', '4242', *mode], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and '0 of 3 cases differ' in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


def test_oracle_fuzz_against_the_reference():
    """The oracle against the unmodified reference on three random configurations (tools/oracle_fuzz.py): canonical CSR of G_data
    and Gc and Ec bit-identical, the same final edit mask, rel-L2(m0) < 1e-9."""
    env = dict(os.environ, OMP_NUM_THREADS='4', OPENBLAS_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'oracle_fuzz.py'), '3', '4242'], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and '0 of 3 cases differ' in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
