"""CPU experiment 5 (planning for round 2): ADDITIVE multilevel (BPX-type) preconditioner -- one fine product per CG
iteration.  z = M_0^-1 r + sum_l P_l M_l^-1 P_l^T r with bilinear prolongations P_l to grids coarsened 2^l x in space
(epochs kept), M_l = block-Jacobi on the DIAGONAL BLOCKS of the Galerkin operator P_l^T N_gg P_l (one coarse dz cell +
the coarse z0 nodes around it: only those blocks are ever formed), optionally an exact solve on the coarsest level.
Run normal_nnz.py first."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, time
exec(open(__file__.replace('bpx.py', 'twogrid.py')).read().split("# block-Jacobi smoother")[0])   # set-up + pcg + interp1d + prolongation pieces


def level(c):
    """prolongation to the c-times coarsened grids and the coarse coordinates (in coarse dz-cell units) of its columns"""
    Iy0, Ix0 = interp1d(int(z0s[0]), c), interp1d(int(z0s[1]), c)
    Iyd, Ixd = interp1d(int(dzs[0]), c), interp1d(int(dzs[1]), c)
    nt = int(dzs[2])
    Pz = sp.kron(Iy0, Ix0).tocsr(); Pd = sp.kron(sp.kron(Iyd, Ixd), sp.identity(nt)).tocsr()
    P = sp.block_diag([Pz, Pd]).tocsr()[kg]
    ncz = Pz.shape[1]
    cy = np.r_[np.repeat(np.arange(Iy0.shape[1]), Ix0.shape[1]) / ratio, np.repeat(np.arange(Iyd.shape[1]), Ixd.shape[1] * nt)]
    cx = np.r_[np.tile(np.arange(Ix0.shape[1]), Iy0.shape[1]) / ratio, np.tile(np.repeat(np.arange(Ixd.shape[1]), nt), Iyd.shape[1])]
    used = np.flatnonzero(np.asarray(abs(P).sum(axis=0)).ravel() > 0)
    return P[:, used].tocsr(), cy[used], cx[used]


def block_jacobi_of(Sm, cy, cx):
    own = np.floor(cy + 0.5).astype(int) * 1000 + np.floor(cx + 0.5).astype(int)
    order = np.argsort(own, kind='stable'); so = own[order]
    st = np.flatnonzero(np.r_[True, so[1:] != so[:-1]]); en = np.r_[st[1:], so.size]
    Sm = Sm.tocsr()
    blocks = [(order[s:e], np.linalg.inv(Sm[order[s:e]][:, order[s:e]].toarray())) for s, e in zip(st, en)]
    def apply(r):
        z = np.zeros_like(r)
        for idx, inv in blocks: z[idx] = inv @ r[idx]
        return z
    return apply, sum(len(b[0]) ** 2 for b in blocks)

bj0, e0 = block_jacobi_of(Ngg, yy, xx)
x, it0 = pcg(Sop, rg, bj0); print('block-Jacobi:', it0, 'iterations; inverse entries %.2e' % e0, flush=True)
levels = {}
for c in (2, 4, 8, 16):
    P, cy, cx = level(c)
    Nc = (P.T @ Ngg @ P).tocsr()
    bjc, ec = block_jacobi_of(Nc, cy, cx)
    levels[c] = (P, bjc, ec, Nc)
    print('level %2dx: %d coarse unknowns, block inverse entries %.2e' % (c, P.shape[1], ec), flush=True)
for use, exact in (((2,), None), ((2, 4), None), ((2, 4, 8), None), ((2, 4, 8), 16), ((2, 4), 8), ((4,), None), ((4, 8), 16)):
    lus = None
    if exact:
        Pe, _, _, Ne = levels[exact]
        lus = spl.splu((Pe.T @ S @ Pe).tocsc())
    def prec(r):
        z = bj0(r)
        for c in use:
            P, bjc, _, _ = levels[c]
            z = z + P @ bjc(P.T @ r)
        if lus is not None: z = z + Pe @ lus.solve(Pe.T @ r)
        return z
    x, it = pcg(Sop, rg, prec, maxit=1000)
    extra = sum(levels[c][2] for c in use) / e0
    print('additive levels %s%s: %d iterations (block-Jacobi alone %d); preconditioner data %.2f x block-Jacobi' %
          (use, ' + exact %dx' % exact if exact else '', it, it0, 1 + extra), flush=True)
