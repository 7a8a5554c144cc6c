"""Pattern statistics of the explicit normal matrix for the cfg-4 workload (CPU, scipy): row lengths, padding of a
sliced-ELL layout with C rows per slice, and cache lines touched per warp gather.  Design aid for the SELL product."""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, '.')
from oracle import lssurf_oracle as O
from lssurf_b200 import synthetic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
width = float(sys.argv[2]) if len(sys.argv) > 2 else 1.e5
t0 = time.time()
pts, kw = synthetic.icesat2_tracks(n=n, width=width)
print('generated', pts['x'].size, time.time() - t0)
S = O.assemble(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
               reference_epoch=kw['reference_epoch'], bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
print('assembled', time.time() - t0)
(g_indptr, g_indices, g_data, _), (c_indptr, c_indices, c_data, _) = S['G'], S['C']
col_N = S['col_N']
# unique column sets of the data rows (pattern only; zero-valued corners were dropped by the CSR conversion, pad with -1)
row_len = np.diff(g_indptr)
K = int(row_len.max())
idx = np.full((S['n_data'], K), -1, dtype=np.int64)
pos = np.arange(g_indices.size) - np.repeat(g_indptr[:-1], row_len)
idx[np.repeat(np.arange(S['n_data']), row_len), pos] = g_indices
u = np.unique(idx, axis=0)
print('unique data rows', u.shape)
uu = u.ravel(); ok = uu >= 0
Au = sp.csr_matrix((np.ones(int(ok.sum()), dtype=np.float32), (np.repeat(np.arange(u.shape[0]), K)[ok], uu[ok])), shape=(u.shape[0], col_N))
Cp = sp.csr_matrix((np.ones(c_indices.size, dtype=np.float32), c_indices, c_indptr), shape=(c_indptr.size - 1, col_N))
P = (Au.T @ Au + Cp.T @ Cp).tocsr()
keep = np.zeros(col_N, bool); keep[S['keep_cols']] = True
# drop fixed columns / rows entirely
D = sp.diags(keep.astype(np.float32))
P = (D @ P @ D).tocsr(); P.eliminate_zeros(); P.sort_indices()
e0 = S['bias']['col_0']
Pg = P[:e0]
lens = np.diff(Pg.indptr)
print('nnz N', P.nnz, 'grid rows nnz', Pg.nnz, 'rows', e0, 'len mean/max', lens.mean(), lens.max())
np.savez_compressed('/tmp/npattern.npz', indptr=Pg.indptr, indices=Pg.indices, e0=e0)
