"""CPU experiment 4 (planning for round 2): two-grid preconditioner with BILINEAR prolongation (z0 and dz grids coarsened
2x in space, epochs kept), symmetric multiplicative cycle: damped block-Jacobi pre-smoothing, exact Galerkin coarse
solve, post-smoothing.  Run normal_nnz.py first.  If a two-grid cycle needs ~10x fewer iterations than block-Jacobi,
a geometric multigrid V-cycle is the round-2 direction for the 56 % of a fit that is CG."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, time
N = sp.load_npz('/tmp/N_cut.npz').tocsr(); rhs = np.load('/tmp/rhs_cut.npy'); M = np.load('/tmp/cut_meta.npz')
keep = M['keep']; nz0 = int(M['nz0']); dzs = M['dz_shape']; z0s = M['z0_shape']; c0 = int(M['dz_col0']); cN = int(M['dz_colN'])
is_z0 = keep < nz0; is_b = keep >= cN
ib = np.flatnonzero(is_b); ig = np.flatnonzero(~is_b)
Ngg = N[ig][:, ig].tocsr(); Ngb = N[ig][:, ib].tocsr(); Dbb = N[ib][:, ib].diagonal()
S = (Ngg - Ngb @ sp.diags(1. / Dbb) @ Ngb.T).tocsr()
rg = rhs[ig] - Ngb @ (rhs[ib] / Dbb)
Sop = lambda v: Ngg @ v - Ngb @ ((Ngb.T @ v) / Dbb)
n = ig.size
kg = keep[ig]; gz0 = kg < nz0
ratio = (z0s[0] - 1) // (dzs[0] - 1)
yy = np.empty(n); xx = np.empty(n)
k = kg[gz0]; yy[gz0], xx[gz0] = np.divmod(k, z0s[1]); yy[gz0] /= ratio; xx[gz0] /= ratio
k = kg[~gz0] - c0; yy[~gz0] = k // (dzs[1] * dzs[2]); xx[~gz0] = (k // dzs[2]) % dzs[1]


def pcg(Aop, b, Minv, rtol=1e-12, maxit=3000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = Aop(p); a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rzn = r @ z
        if np.sqrt(rzn / rz0) <= rtol: return x, k + 1
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def interp1d(nf, c=2):
    """linear interpolation from the coarse nodes 0, c, 2c, ... (+ the last fine node) to nf fine nodes"""
    coarse = np.unique(np.r_[np.arange(0, nf, c), nf - 1])
    rows, cols, vals = [], [], []
    for i in range(nf):
        j = np.searchsorted(coarse, i, side='right') - 1
        if coarse[j] == i:
            rows.append(i); cols.append(j); vals.append(1.)
        else:
            w = (i - coarse[j]) / (coarse[j + 1] - coarse[j])
            rows += [i, i]; cols += [j, j + 1]; vals += [1 - w, w]
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, coarse.size))


def prolongation(c):
    Pz = sp.kron(interp1d(int(z0s[0]), c), interp1d(int(z0s[1]), c)).tocsr()                      # z0: (y, x)
    Pd = sp.kron(sp.kron(interp1d(int(dzs[0]), c), interp1d(int(dzs[1]), c)), sp.identity(int(dzs[2]))).tocsr()   # dz: (y, x, t)
    P = sp.block_diag([Pz, Pd]).tocsr()              # all grid columns (fine) x all coarse columns
    Pk = P[kg]                                       # fine free columns only
    # coarse columns that only interpolate to the fixed (reference-epoch) fine columns drop out: keep coarse columns with support
    used = np.flatnonzero(np.asarray(abs(Pk).sum(axis=0)).ravel() > 0)
    # remove the coarse reference-epoch dz columns (they would be a null space): those whose fine images are all fixed were
    # dropped above; the others interpolate only within one epoch, so the reference epoch never appears
    return Pk[:, used].tocsr()


# block-Jacobi smoother (1 dz cell + z0 nodes), damped
by = np.floor(yy + 0.5).astype(int); bx = np.floor(xx + 0.5).astype(int); own = by * 1000 + bx
order = np.argsort(own, kind='stable'); so = own[order]
st = np.flatnonzero(np.r_[True, so[1:] != so[:-1]]); en = np.r_[st[1:], so.size]
blocks = [(order[s:e], np.linalg.inv(Ngg[order[s:e]][:, order[s:e]].toarray())) for s, e in zip(st, en)]
def bj(r):
    z = np.zeros_like(r)
    for idx, inv in blocks: z[idx] = inv @ r[idx]
    return z
x, it0 = pcg(Sop, rg, bj); print('block-Jacobi:', it0, 'iterations', flush=True)

for c in (2, 4):
    t0 = time.time()
    P = prolongation(c)
    Sc = (P.T @ S @ P).tocsc()
    lu = spl.splu(Sc)
    for omega, nu in ((0.7, 1), (0.5, 1), (0.7, 2)):
        def cycle(r):
            z = np.zeros_like(r)
            for _ in range(nu): z = z + omega * bj(r - Sop(z)) if _ else omega * bj(r)
            z = z + P @ lu.solve(P.T @ (r - Sop(z)))
            for _ in range(nu): z = z + omega * bj(r - Sop(z))
            return z
        x, it = pcg(Sop, rg, cycle, maxit=400)
        prods = 1 + 2 * nu            # fine products per CG iteration: 1 (CG) + nu-1 + 1 + nu residuals inside the cycle
        print('two-grid, coarsening %dx, omega %.1f, %d+%d smoothing: %d iterations, %d fine products each -> %d products (block-Jacobi PCG: %d); coarse unknowns %d, nnz %d  %.0fs'
              % (c, omega, nu, nu, it, prods, it * prods, it0, P.shape[1], Sc.nnz, time.time() - t0), flush=True)
    # additive variant: one product per iteration
    add = lambda r: bj(r) + P @ lu.solve(P.T @ r)
    x, it = pcg(Sop, rg, add, maxit=1000)
    print('two-level ADDITIVE with bilinear prolongation, coarsening %dx: %d iterations (1 fine product each)' % (c, it), flush=True)
