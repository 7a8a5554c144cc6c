"""CPU tier: the index tables of the grid-structured normal-matrix layout (lssurf_b200/structured.py) against SciPy.

N = A^T W A is formed with SciPy from the oracle's operators for a small tracks-with-bias problem; the structured product
(NumPy restatement of what the CUDA kernel computes from the tables) must reproduce N p on the rows the CG product covers."""
import numpy as np
import scipy.sparse as sp
import torch

from oracle import lssurf_oracle as O
from lssurf_b200 import structured, synthetic


def _normal_matrix(n=40000, width=8.e3, seed=3):
    pts, kw = synthetic.icesat2_tracks(n=n, width=width, seed=seed)
    s = O.assemble(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
                   kw['reference_epoch'], bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
    G = O.csr_matrix(*s['G'][:3], shape=(s['n_data'], s['col_N']))
    C = O.csr_matrix(*s['C'][:3], shape=(s['Ec'].size, s['col_N']))
    A = sp.vstack([sp.diags(1. / s['Ed']) @ G, sp.diags(1. / s['Ec']) @ C]).tocsr()
    fixed = np.setdiff1d(np.arange(s['col_N']), s['keep_cols'])
    keep = sp.diags(np.isin(np.arange(s['col_N']), s['keep_cols']).astype(float))
    N = (keep @ (A.T @ A) @ keep).tocsr()
    N.eliminate_zeros()
    N.sort_indices()
    return N, s, fixed


def test_structured_layout_reproduces_the_csr_product():
    N, s, fixed = _normal_matrix()
    gz, gd = s['grids']['z0'], s['grids']['dz']
    rows = s['bias']['col_0']                      # the bias block is eliminated: its rows are not part of the product
    st = structured.build(torch.from_numpy(N.indptr.astype(np.int32)), torch.from_numpy(N.indices.astype(np.int32)), rows,
                          gz['col_0'], gz['shape'], gd['col_0'], gd['shape'])
    assert st is not None
    nnz_rows = int(N.indptr[rows])
    c = st['counts']
    assert c['self'] + c['dropped'] + c['remainder'] == nnz_rows
    assert c['self'] + c['dropped'] > 0.6 * nnz_rows        # most of N is regular (the rest: bias columns, dense on this small domain)
    # every kept entry has its own slot
    m = st['map'].numpy()
    assert np.unique(m[m >= 0]).size == (m >= 0).sum()
    rng = np.random.default_rng(0)
    p = rng.standard_normal(N.shape[0])
    p[fixed] = 0.
    want = (N @ p)[:rows]
    got = structured.emulate(st, N.data, p)
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)
    assert structured.product_bytes(st) < 12 * nnz_rows                          # fewer bytes than CSR even with 38 % of N in the remainder


def test_structured_layout_declines_what_it_cannot_hold():
    ind = torch.zeros(5, dtype=torch.int32)
    assert structured.build(ind, ind[:0], 4, 0, (2, 2), 4, (1, 1, 40)) is None    # more than 32 epochs
