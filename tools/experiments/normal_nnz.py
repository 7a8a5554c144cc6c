"""CPU experiment: size of the explicit normal matrix N = A^T W A against A for a cut of cfg 4,
and PCG iteration counts of candidate preconditioners built from N.  (tools/ only; uses the oracle.)"""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lssurf_oracle as O
from lssurf_b200 import synthetic

width = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0e4
n = int(1e7 * (width / 1e5) ** 2)
pts, kw = synthetic.icesat2_tracks(n=n, width=width, seed=0)
t0 = time.time()
s = O.assemble(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
               kw['reference_epoch'], bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
G = O.csr_matrix(*s['G'][:3], shape=(s['n_data'], s['col_N']))
C = O.csr_matrix(*s['C'][:3], shape=(s['Ec'].size, s['col_N']))
keep = s['keep_cols']
A = sp.vstack([sp.diags(1. / s['Ed']) @ G[:, keep], sp.diags(1. / s['Ec']) @ C[:, keep]]).tocsr()
b = np.concatenate([s['z'] / s['Ed'], s['prior'] / s['Ec']])
print('assemble %.1fs  A %s nnz %d' % (time.time() - t0, A.shape, A.nnz))
t0 = time.time()
N = (A.T @ A).tocsr()
print('N nnz %d (%.3f of A), %.1f per row, %.1fs' % (N.nnz, N.nnz / A.nnz, N.nnz / N.shape[0], time.time() - t0))
nz0 = s['grids']['z0']['n_nodes']
ncols = np.diff(N.indptr)
kept_is_z0 = keep < nz0
kept_is_b = keep >= s['grids']['dz']['col_N']
print('rows: z0 %.1f  dz %.1f  bias %.1f nnz/row' % (ncols[kept_is_z0].mean(), ncols[~kept_is_z0 & ~kept_is_b].mean(), ncols[kept_is_b].mean()))
np.save('/tmp/N_shape.npy', np.array([N.shape[0], N.nnz]))
sp.save_npz('/tmp/N_cut.npz', N)
np.save('/tmp/rhs_cut.npy', A.T @ b)
np.savez('/tmp/cut_meta.npz', keep=keep, nz0=nz0, dz_shape=np.array(s['grids']['dz']['shape']), z0_shape=np.array(s['grids']['z0']['shape']),
         dz_col0=s['grids']['dz']['col_0'], dz_colN=s['grids']['dz']['col_N'], ref=kw['reference_epoch'])
print('row length percentiles', np.percentile(ncols, [0, 5, 25, 50, 75, 95, 99, 100]))
