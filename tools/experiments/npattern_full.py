"""CPU experiment (tools/ only; uses the oracle): pattern of the normal matrix N = A^T W A of the full cfg-4 problem from
the distinct column sets of the data rows + the constraint rows; saved to /tmp/np/npattern.npz for the format studies."""
import sys, os, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lssurf_oracle as O
from lssurf_b200 import synthetic

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
width = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0e5
t0 = time.time()
pts, kw = synthetic.icesat2_tracks(n=n, width=width, seed=0)
print('points %.1fs' % (time.time() - t0), flush=True)
# constraint rows and column layout from a thin subset (they do not depend on the data)
sub = slice(0, n, max(n // 20000, 1))
s = O.assemble(pts['x'][sub], pts['y'][sub], pts['time'][sub], pts['z'][sub], pts['sigma'][sub], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
               kw['reference_epoch'])
gz, gd = s['grids']['z0'], s['grids']['dz']
print({k: gz[k] for k in ('shape', 'col_0', 'col_N')}, {k: gd[k] for k in ('shape', 'col_0', 'col_N')})
C = O.csr_matrix(*s['C'][:3], shape=(s['Ec'].size, s['col_N']))
# distinct (z0 cell, dz time cell) of the data rows
def cell(g, d, v):
    return np.floor((v - g['lo'][d]) / g['delta'][d]).astype(np.int64)
iy0, ix0 = cell(gz, 0, pts['y']), cell(gz, 1, pts['x'])
it = cell(gd, 2, pts['time'])
key = (iy0 * gz['shape'][1] + ix0) * gd['shape'][2] + it
uk = np.unique(key)
print('distinct (z0 cell, epoch cell):', uk.size, flush=True)
c0 = uk // gd['shape'][2]; itu = uk % gd['shape'][2]
iy0, ix0 = c0 // gz['shape'][1], c0 % gz['shape'][1]
yy = gz['lo'][0] + (iy0 + 0.5) * gz['delta'][0]; xx = gz['lo'][1] + (ix0 + 0.5) * gz['delta'][1]
tt = gd['lo'][2] + (itu + 0.5) * gd['delta'][2]
r0, c0_, v0 = O.interp_triplets(gz, [yy, xx])[:3]
r1, c1_, v1 = O.interp_triplets(gd, [yy, xx, tt])[:3]
Gu = sp.csr_matrix((np.ones(r0.size + r1.size), (np.r_[r0.ravel(), r1.ravel()], np.r_[c0_.ravel(), c1_.ravel()])), shape=(uk.size, s['col_N']))
Cb = C.copy(); Cb.data[:] = 1
N = (Gu.T @ Gu + Cb.T @ Cb).tocsr()
N.sort_indices()
keep = s['keep_cols']
print('N (all grid columns) nnz', N.nnz, 'free cols', keep.size, '%.1fs' % (time.time() - t0))
np.savez('/tmp/np/npattern.npz', indptr=N.indptr, indices=N.indices, keep=keep, z0_shape=np.array(gz['shape']), dz_shape=np.array(gd['shape']),
         dz_col0=gd['col_0'], ref=kw['reference_epoch'])
