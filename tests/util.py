"""Helpers shared by the CPU and GPU test tiers."""
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def digest(*arrays):
    """sha256 over dtype, shape and bytes of each array (same recipe as oracle/make_golden.py)."""
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def csr_digest(M):
    M = M.tocsr()
    return digest(M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float64))


def load_case(name):
    fx = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    pts = {k[3:]: fx[k] for k in fx.files if k.startswith('in_') and k != 'in_TSE'}
    kw = eval(str(fx['kw']), {'np': np, 'array': np.array, 'float64': np.float64})   # repr() of plain dicts
    return fx, pts, kw


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


FIT_CASES = ['strip_nb', 'strip_outliers', 'uniform3d_small', 'tracks_bias_small']
# fixtures taken from the reference run alone (oracle/make_golden.py::case_widened)
WIDENED_CASES = ['uniform_avg_scales', 'uniform_groups_editable', 'tracks_bias_edit', 'uniform_sigma_grid', 'tracks_sensor_slope']
