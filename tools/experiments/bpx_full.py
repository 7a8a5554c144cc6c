"""CPU experiment 6 (planning for round 2): the additive multilevel preconditioner of bpx.py on the FULL cfg-4 problem
(run normal_nnz.py 1e5 first; ~10 GB).  Block applications are batched (padded dense blocks) so that a PCG iteration
costs ~0.3 s on the CPU.  Reports cold-start iteration counts to rtol 1e-12 (the GPU solver needs 789 here)."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, time, sys
N = sp.load_npz('/tmp/N_cut.npz').tocsr(); rhs = np.load('/tmp/rhs_cut.npy'); M = np.load('/tmp/cut_meta.npz')
keep = M['keep']; nz0 = int(M['nz0']); dzs = M['dz_shape']; z0s = M['z0_shape']; c0 = int(M['dz_col0']); cN = int(M['dz_colN'])
is_z0 = keep < nz0; is_b = keep >= cN
ib = np.flatnonzero(is_b); ig = np.flatnonzero(~is_b)
Ngg = N[ig][:, ig].tocsr(); Ngb = N[ig][:, ib].tocsr(); Dbb = N[ib][:, ib].diagonal()
NgbT = Ngb.T.tocsr()
rg = rhs[ig] - Ngb @ (rhs[ib] / Dbb)
Sop = lambda v: Ngg @ v - Ngb @ ((NgbT @ v) / Dbb)
n = ig.size
kg = keep[ig]; gz0 = kg < nz0
ratio = (z0s[0] - 1) // (dzs[0] - 1)
yy = np.empty(n); xx = np.empty(n)
k = kg[gz0]; yy[gz0], xx[gz0] = np.divmod(k, z0s[1]); yy[gz0] /= ratio; xx[gz0] /= ratio
k = kg[~gz0] - c0; yy[~gz0] = k // (dzs[1] * dzs[2]); xx[~gz0] = (k // dzs[2]) % dzs[1]
print('unknowns', n, 'nnz N_gg', Ngg.nnz, flush=True)
maxit = int(sys.argv[1]) if len(sys.argv) > 1 else 1500


def pcg(Aop, b, Minv, rtol=1e-12, maxit=maxit, tag=''):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    t0 = time.time()
    for k in range(maxit):
        q = Aop(p); a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rzn = r @ z
        rel = np.sqrt(rzn / rz0)
        if k % 100 == 99: print('   %s it %d rel %.2e (%.0fs)' % (tag, k + 1, rel, time.time() - t0), flush=True)
        if rel <= rtol: return x, k + 1
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def interp1d(nf, c=2):
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


def level(c):
    Iy0, Ix0 = interp1d(int(z0s[0]), c), interp1d(int(z0s[1]), c)
    Iyd, Ixd = interp1d(int(dzs[0]), c), interp1d(int(dzs[1]), c)
    nt = int(dzs[2])
    Pz = sp.kron(Iy0, Ix0).tocsr(); Pd = sp.kron(sp.kron(Iyd, Ixd), sp.identity(nt)).tocsr()
    P = sp.block_diag([Pz, Pd]).tocsr()[kg]
    cy = np.r_[np.repeat(np.arange(Iy0.shape[1]), Ix0.shape[1]) / ratio, np.repeat(np.arange(Iyd.shape[1]), Ixd.shape[1] * nt)]
    cx = np.r_[np.tile(np.arange(Ix0.shape[1]), Iy0.shape[1]) / ratio, np.tile(np.repeat(np.arange(Ixd.shape[1]), nt), Iyd.shape[1])]
    used = np.flatnonzero(np.asarray(abs(P).sum(axis=0)).ravel() > 0)
    return P[:, used].tocsr(), cy[used], cx[used]


def batched_block_jacobi(Sm, cy, cx):
    """block-Jacobi with all blocks padded to the largest size: one batched matmul per application"""
    own = np.floor(cy + 0.5).astype(int) * 100000 + np.floor(cx + 0.5).astype(int)
    order = np.argsort(own, kind='stable'); so = own[order]
    st = np.flatnonzero(np.r_[True, so[1:] != so[:-1]]); en = np.r_[st[1:], so.size]
    nb = st.size; bs = int((en - st).max())
    idx = np.full((nb, bs), -1, dtype=np.int64)
    for b, (s, e) in enumerate(zip(st, en)): idx[b, :e - s] = order[s:e]
    Sm = Sm.tocsr()
    inv = np.zeros((nb, bs, bs))
    for b in range(nb):
        m = idx[b][idx[b] >= 0]
        inv[b, :m.size, :m.size] = np.linalg.inv(Sm[m][:, m].toarray())
    safe = np.where(idx >= 0, idx, 0); mask = idx >= 0
    def apply(r):
        rb = np.where(mask, r[safe], 0.0)
        zb = np.einsum('bij,bj->bi', inv, rb)
        z = np.zeros_like(r)
        z[idx[mask]] = zb[mask]
        return z
    return apply

t0 = time.time()
bj0 = batched_block_jacobi(Ngg, yy, xx)
print('block inverses built %.0fs' % (time.time() - t0), flush=True)
if len(sys.argv) > 2 and sys.argv[2] == 'skip':
    it0 = 789
else:
    x, it0 = pcg(Sop, rg, bj0, tag='block-Jacobi'); print('block-Jacobi: %d iterations (GPU solver: 789)' % it0, flush=True)
levels = {}
for c in (2, 4, 8, 16, 32):
    t0 = time.time()
    P, cy, cx = level(c)
    PT = P.T.tocsr()
    Nc = (PT @ Ngg @ P).tocsr()
    levels[c] = (P, PT, batched_block_jacobi(Nc, cy, cx), Nc)
    print('level %2dx: %d coarse unknowns (%.0fs)' % (c, P.shape[1], time.time() - t0), flush=True)
configs = (((2, 4, 8, 16), 32), ((2, 4, 8), 16), ((2, 4, 8, 16), None))
if len(sys.argv) > 3:       # e.g. "-:16;-:32;-:8;2,4:8"  (levels with block-Jacobi : level solved exactly)
    configs = [(tuple(int(v) for v in a.split(',') if v != '-'), int(b) if b != '-' else None)
               for a, b in (c.split(':') for c in sys.argv[3].split(';'))]
for use, exact in configs:
    lus = None
    if exact:
        Pe, PeT = levels[exact][0], levels[exact][1]
        PB = (PeT @ Ngb).tocsr()              # never form the fine Schur complement: every bias column couples ~3 700 nodes
        Se = (PeT @ Ngg @ Pe - PB @ sp.diags(1. / Dbb) @ PB.T).tocsc()
        lus = spl.splu(Se)
        print('exact level %dx: %d unknowns, nnz %d' % (exact, Se.shape[0], Se.nnz), flush=True)
    def prec(r):
        z = bj0(r)
        for c in use:
            P, PT, bjc, _ = levels[c]
            z = z + P @ bjc(PT @ r)
        if lus is not None: z = z + Pe @ lus.solve(PeT @ r)
        return z
    x, it = pcg(Sop, rg, prec, tag=str(use))
    print('additive levels %s%s: %d iterations (block-Jacobi alone %d)' % (use, ' + exact %dx' % exact if exact else '', it, it0), flush=True)
