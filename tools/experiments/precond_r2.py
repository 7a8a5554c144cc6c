"""CPU experiment 3 (planning for round 2): iteration counts of PCG on the bias-eliminated (Schur) operator of the
cfg-4 cut (run normal_nnz.py first) for preconditioners beyond the 1-cell block-Jacobi the GPU solver uses today:
larger / overlapping blocks, two-level corrections with aggregate coarse spaces, deflation with Ritz vectors, and a
multicolour block Gauss-Seidel sweep.  Cost model next to each: bytes of preconditioner data streamed per application."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, time, sys
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
yy = np.empty(n); xx = np.empty(n); tt = np.full(n, -1)
k = kg[gz0]; yy[gz0], xx[gz0] = np.divmod(k, z0s[1]); yy[gz0] /= ratio; xx[gz0] /= ratio
k = kg[~gz0] - c0; yy[~gz0] = k // (dzs[1] * dzs[2]); xx[~gz0] = (k // dzs[2]) % dzs[1]; tt[~gz0] = k % dzs[2]
print('unknowns', n, 'nnz N_gg', Ngg.nnz, flush=True)


def pcg(Aop, b, Minv, rtol=1e-12, maxit=3000, x0=None):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - Aop(x) if x0 is not None else b.copy()
    z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = Aop(p); a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rzn = r @ z
        if np.sqrt(rzn / rz0) <= rtol: return x, k + 1
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def blocks_of(bs, overlap, Smat):
    by = np.floor((yy + 0.5) / bs).astype(int); bx = np.floor((xx + 0.5) / bs).astype(int)
    own = by * 1000 + bx
    out = []
    for b in np.unique(own):
        cy, cx = divmod(b, 1000)
        idx = np.flatnonzero((yy >= cy * bs - 0.5 - overlap) & (yy < (cy + 1) * bs - 0.5 + overlap) &
                             (xx >= cx * bs - 0.5 - overlap) & (xx < (cx + 1) * bs - 0.5 + overlap))
        out.append((idx, np.linalg.inv(Smat[idx][:, idx].toarray()), own[idx] == b, (cy, cx)))
    return out


def schwarz_apply(blocks, restricted):
    def apply(r):
        z = np.zeros_like(r)
        for idx, inv, mine, _ in blocks:
            t = inv @ r[idx]
            if restricted: z[idx[mine]] += t[mine]
            else: z[idx] += t
        return z
    return apply


t0 = time.time()
B1 = blocks_of(1, 0, Ngg)
bj = schwarz_apply(B1, False)
x_ref, it = pcg(Sop, rg, bj)
bytes1 = sum(len(b[0]) ** 2 for b in B1) * 4
print('block-Jacobi, 1 dz cell (today): its %d, inverse %.1f MB fp32 (%.2f x nnz N_gg x 12 B)  %.0fs' % (it, bytes1 / 1e6, bytes1 / (Ngg.nnz * 12), time.time() - t0), flush=True)

# ---- larger and overlapping blocks (additive needs symmetric weights; use plain additive Schwarz, which is SPD)
for bs, ov in ((2, 0), (1, 0.5), (2, 0.5), (3, 0)):
    t0 = time.time()
    B = blocks_of(bs, ov, Ngg)
    x, it = pcg(Sop, rg, schwarz_apply(B, False))
    by = sum(len(b[0]) ** 2 for b in B) * 4
    print('additive Schwarz block %d overlap %.1f: its %d  max block %d  inverse %.1f MB (%.2f x N_gg)  %.0fs' % (bs, ov, it, max(len(b[0]) for b in B), by / 1e6, by / (Ngg.nnz * 12), time.time() - t0), flush=True)

# ---- two-level: block-Jacobi + Galerkin coarse correction on aggregates of a x a dz cells, additive
def coarse_space(a, per_epoch=True):
    ay = np.floor((yy + 0.5) / a).astype(int); ax = np.floor((xx + 0.5) / a).astype(int)
    agg = ay * 1000 + ax
    _, agg = np.unique(agg, return_inverse=True)
    n_agg = agg.max() + 1
    nt = int(tt.max()) + 1
    # one coarse unknown per aggregate for z0, one per aggregate (and epoch) for dz
    col = np.where(tt < 0, agg, n_agg + (agg * nt + tt if per_epoch else agg))
    _, col = np.unique(col, return_inverse=True)
    return sp.csr_matrix((np.ones(n), (np.arange(n), col)), shape=(n, col.max() + 1))

for a, per_epoch in ((2, True), (4, True), (2, False)):
    t0 = time.time()
    P = coarse_space(a, per_epoch)
    Sc = (P.T @ S @ P).tocsc()
    lu = spl.splu(Sc)
    two = lambda r: bj(r) + P @ lu.solve(P.T @ r)
    x, it = pcg(Sop, rg, two)
    print('two-level additive: block-Jacobi + aggregates of %d x %d dz cells (%s): its %d  coarse unknowns %d, coarse nnz %d  %.0fs' % (a, a, 'per epoch' if per_epoch else 'all epochs together', it, P.shape[1], Sc.nnz, time.time() - t0), flush=True)

# ---- deflation with the k lowest eigenvectors of the block-Jacobi-preconditioned operator (what recycling Ritz vectors
#      of the previous outer iteration could approach)
t0 = time.time()
Sl = spl.LinearOperator((n, n), matvec=Sop, dtype=float)
Ml = spl.LinearOperator((n, n), matvec=bj, dtype=float)
try:
    w, V = spl.lobpcg(Sl, np.random.default_rng(0).normal(size=(n, 40)), M=Ml, B=None, largest=False, tol=1e-6, maxiter=200)
    print('lowest eigenvalues (lobpcg, %.0fs)' % (time.time() - t0), np.sort(w)[:8], flush=True)
    for kdef in (10, 20, 40):
        W = V[:, np.argsort(w)[:kdef]]
        SW = np.column_stack([Sop(W[:, j]) for j in range(kdef)])
        E = np.linalg.inv(W.T @ SW)
        # deflated start + projected preconditioner (A-DEF2-like): z = M^-1 r + W E W^T (r - S M^-1 r)
        def defl(r):
            z = bj(r)
            return z + W @ (E @ (W.T @ (r - Sop(z))))
        x0 = W @ (E @ (W.T @ rg))
        x, it = pcg(Sop, rg, defl, x0=x0)
        print('deflation with %d vectors (adapted deflation, 2 products per iteration): its %d' % (kdef, it), flush=True)
except Exception as e:
    print('deflation experiment failed:', e)

# ---- multicolour symmetric block Gauss-Seidel on the 1-cell blocks (4 colours), as a preconditioner
t0 = time.time()
colour = np.array([2 * (b[3][0] % 2) + (b[3][1] % 2) for b in B1])
def sgs(r):
    z = np.zeros_like(r)
    for sweep in (range(4), range(3, -1, -1)):
        for c in sweep:
            res = r - S @ z
            for (idx, inv, _, _), cc in zip(B1, colour):
                if cc == c: z[idx] += inv @ res[idx]
    return z
x, it = pcg(Sop, rg, sgs, maxit=600)
print('4-colour symmetric block Gauss-Seidel (8 products per application): its %d  %.0fs' % (it, time.time() - t0), flush=True)
