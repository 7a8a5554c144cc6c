"""GPU tier, multi-GPU: a 2-rank run of tests/dist_check.py (skipped where fewer than two GPUs are visible).

The row-sharded solve must give the masks of the one-GPU fit after every outer iteration and
the same model to 1e-9 relative L2; the per-node count / misfit maps of parse_model must agree as well."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('modes', ['p2p', 'replicated,allreduce'])
def test_two_rank_fit_matches_one_gpu(modes):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    env = dict(os.environ, LSS_DIST_CHECK_MODES=modes, LSS_DIST_CHECK_TIMEOUT='420')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dist_check.py')]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(out.stdout[-4000:])
    print(out.stderr[-4000:])
    assert out.returncode == 0 and '[dist_check] PASS' in out.stdout
