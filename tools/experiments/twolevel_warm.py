"""CPU experiment 7 (planning for round 2): the WHOLE edit loop (warm-started solves, weights changing with the mask) with
block-Jacobi alone and with block-Jacobi + one exact coarse solve (Galerkin Schur complement on a bilinear coarse space).
Counts PCG iterations per outer iteration, to rtol 1e-12 like the GPU solver.  Uses the oracle's assemble / iterate.
    python tools/experiments/twolevel_warm.py [width=3e4] [coarsening=8]"""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lssurf_oracle as O
from lssurf_b200 import synthetic

width = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0e4
coarsen = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(1e7 * (width / 1e5) ** 2)
pts, kw = synthetic.icesat2_tracks(n=n, width=width, seed=0)
s = O.assemble(pts['x'], pts['y'], pts['time'], pts['z'], pts['sigma'], kw['ctr'], kw['W'], kw['spacing'], kw['E_RMS'],
               kw['reference_epoch'], bias_params=[pts['rgt'], pts['cycle']], sigma_corr=pts['sigma_corr'])
keep = s['keep_cols']; g = s['grids']; nz0 = g['z0']['n_nodes']; z0s = g['z0']['shape']; dzs = g['dz']['shape']
c0, cN = g['dz']['col_0'], g['dz']['col_N']
is_b = keep >= cN
ib, ig = np.flatnonzero(is_b), np.flatnonzero(~is_b)
kg = keep[ig]; gz0 = kg < nz0
ratio = (z0s[0] - 1) // (dzs[0] - 1)
yy = np.empty(ig.size); xx = np.empty(ig.size)
k = kg[gz0]; yy[gz0], xx[gz0] = np.divmod(k, z0s[1]); yy[gz0] /= ratio; xx[gz0] /= ratio
k = kg[~gz0] - c0; yy[~gz0] = k // (dzs[1] * dzs[2]); xx[~gz0] = (k // dzs[2]) % dzs[1]
own = np.floor(yy + 0.5).astype(int) * 100000 + np.floor(xx + 0.5).astype(int)
order = np.argsort(own, kind='stable'); so = own[order]
st = np.flatnonzero(np.r_[True, so[1:] != so[:-1]]); en = np.r_[st[1:], so.size]
bs = int((en - st).max()); idx = np.full((st.size, bs), -1, dtype=np.int64)
for b, (a, e) in enumerate(zip(st, en)): idx[b, :e - a] = order[a:e]
safe = np.where(idx >= 0, idx, 0); bmask = idx >= 0


def interp1d(nf, c):
    coarse = np.unique(np.r_[np.arange(0, nf, c), nf - 1]); rows, cols, vals = [], [], []
    for i in range(nf):
        j = np.searchsorted(coarse, i, side='right') - 1
        if coarse[j] == i: rows.append(i); cols.append(j); vals.append(1.)
        else:
            w = (i - coarse[j]) / (coarse[j + 1] - coarse[j]); rows += [i, i]; cols += [j, j + 1]; vals += [1 - w, w]
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, coarse.size))


Pz = sp.kron(interp1d(int(z0s[0]), coarsen * ratio), interp1d(int(z0s[1]), coarsen * ratio))
Pd = sp.kron(sp.kron(interp1d(int(dzs[0]), coarsen), interp1d(int(dzs[1]), coarsen)), sp.identity(int(dzs[2])))
P = sp.block_diag([Pz, Pd]).tocsr()[kg]
P = P[:, np.flatnonzero(np.asarray(abs(P).sum(axis=0)).ravel() > 0)].tocsr(); PT = P.T.tocsr()
print('points %d  grid unknowns %d  bias %d  blocks %d (max %d)  coarse unknowns %d (z0 and dz coarsened %dx in dz cells)'
      % (s['n_data'], ig.size, ib.size, st.size, bs, P.shape[1], coarsen), flush=True)


class PCG(object):
    def __init__(self, two_level):
        self.two_level, self.x, self.log = two_level, None, []

    def __call__(self, A, b):
        t0 = time.time()
        At = A.T.tocsr(); N = (At @ A).tocsr(); rhs = At @ b
        Ngg = N[ig][:, ig].tocsr(); Ngb = N[ig][:, ib].tocsr(); Dbb = N[ib][:, ib].diagonal(); NbgT = Ngb.T.tocsr()
        rg = rhs[ig] - Ngb @ (rhs[ib] / Dbb)
        S = lambda v: Ngg @ v - Ngb @ ((NbgT @ v) / Dbb)
        inv = np.zeros((st.size, bs, bs))
        for bI in range(st.size):
            m = idx[bI][idx[bI] >= 0]; inv[bI, :m.size, :m.size] = np.linalg.inv(Ngg[m][:, m].toarray())

        def bj(r):
            zb = np.einsum('bij,bj->bi', inv, np.where(bmask, r[safe], 0.0)); z = np.zeros_like(r); z[idx[bmask]] = zb[bmask]; return z
        if self.two_level:
            PB = (PT @ Ngb).toarray()
            Sc = (PT @ Ngg @ P).toarray() - (PB / Dbb) @ PB.T
            cf = sla.cho_factor(Sc)
            M = lambda r: bj(r) + P @ sla.cho_solve(cf, PT @ r)
        else:
            M = bj
        x = np.zeros(ig.size) if self.x is None else self.x[ig].copy()          # warm start, like the GPU solver
        r = rg - S(x); z = M(r); p = z.copy(); rz = r @ z; bn = rg @ M(rg); it = 0
        while np.sqrt(rz / bn) > 1e-12 and it < 5000:
            q = S(p); a = rz / (p @ q); x += a * p; r -= a * q; z = M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn; it += 1
        full = np.zeros(keep.size); full[ig] = x; full[ib] = (rhs[ib] - NbgT @ x) / Dbb
        self.x = full; self.log.append(it)
        print('   outer %d: %d PCG iterations (%.0f s)' % (len(self.log) - 1, it, time.time() - t0), flush=True)
        return full


for two in (False, True):
    print('block-Jacobi + exact coarse solve' if two else 'block-Jacobi alone', flush=True)
    solver = PCG(two)
    m0, se, tse, rs = O.iterate(s, max_iterations=kw['max_iterations'], solver=solver)
    print('   total %d iterations over %d solves, %d points kept' % (sum(solver.log), len(solver.log), int(tse.sum())), flush=True)
