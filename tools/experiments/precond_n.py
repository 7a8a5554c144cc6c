"""CPU experiment: PCG iteration counts on the explicit normal matrix of a cfg-4 cut for preconditioners
that a GPU can apply cheaply (run normal_nnz.py first)."""
import numpy as np, scipy.sparse as sp, scipy.linalg as sl, time
N = sp.load_npz('/tmp/N_cut.npz').tocsr(); rhs = np.load('/tmp/rhs_cut.npy'); M = np.load('/tmp/cut_meta.npz')
keep = M['keep']; nz0 = int(M['nz0']); dzs = M['dz_shape']; z0s = M['z0_shape']; c0 = int(M['dz_col0']); cN = int(M['dz_colN'])
n = N.shape[0]

def pcg(Aop, b, Minv, rtol=1e-12, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = Aop(p); a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rzn = r @ z
        if np.sqrt(rzn / rz0) <= rtol: return x, k + 1
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit

d = N.diagonal()
x0, it = pcg(lambda v: N @ v, rhs, lambda r: r / d); print('jacobi', it)

# classify kept columns
is_z0 = keep < nz0; is_b = keep >= cN; is_dz = ~is_z0 & ~is_b
ib = np.flatnonzero(is_b); ig = np.flatnonzero(~is_b)
# --- bias Schur complement (N_bb diagonal?)
Nbb = N[ib][:, ib]; print('N_bb offdiag nnz', Nbb.nnz - ib.size)
Ngg = N[ig][:, ig].tocsr(); Ngb = N[ig][:, ib].tocsr(); Dbb = Nbb.diagonal()
Sop = lambda v: Ngg @ v - Ngb @ ((Ngb.T @ v) / Dbb)
rg = rhs[ig] - Ngb @ (rhs[ib] / Dbb)
dS = Ngg.diagonal() - np.asarray(Ngb.multiply(Ngb).dot(1. / Dbb)).ravel()
xg, it = pcg(Sop, rg, lambda r: r / dS); print('bias-schur + jacobi', it)

# --- block jacobi: block id per kept column: dz node (iy,ix) -> block; z0 node -> nearest dz node
def block_ids(ratio_shift=0):
    bid = np.empty(n, dtype=np.int64)
    k = keep[is_z0]; iy, ix = np.divmod(k, z0s[1]); r = (z0s[0] - 1) // (dzs[0] - 1)
    bid[is_z0] = ((iy + r // 2) // r) * dzs[1] + (ix + r // 2) // r
    k = keep[is_dz] - c0; iy = k // (dzs[1] * dzs[2]); ix = (k // dzs[2]) % dzs[1]
    bid[is_dz] = iy * dzs[1] + ix
    bid[is_b] = dzs[0] * dzs[1] + np.arange(ib.size)
    return bid

def block_jacobi(Nm, bid):
    order = np.argsort(bid, kind='stable'); sb = bid[order]
    starts = np.flatnonzero(np.r_[True, sb[1:] != sb[:-1]]); ends = np.r_[starts[1:], sb.size]
    Nc = Nm.tocsr(); invs = []
    for s, e in zip(starts, ends):
        idx = order[s:e]; B = Nc[idx][:, idx].toarray(); invs.append((idx, np.linalg.inv(B)))
    def apply(r):
        z = np.empty_like(r)
        for idx, Bi in invs: z[idx] = Bi @ r[idx]
        return z
    return apply, max(e - s for s, e in zip(starts, ends))

bid = block_ids()
t0 = time.time(); Mb, mx = block_jacobi(N, bid); print('block build %.1fs max block %d' % (time.time() - t0, mx))
x1, it = pcg(lambda v: N @ v, rhs, Mb); print('block-jacobi (dz column + 4x4 z0)', it, np.linalg.norm(x1 - x0) / np.linalg.norm(x0))
# dz-only column blocks, z0 jacobi
bid2 = bid.copy(); bid2[is_z0] = bid.max() + 1 + np.arange(is_z0.sum())
Mb2, mx = block_jacobi(N, bid2); x1, it = pcg(lambda v: N @ v, rhs, Mb2); print('block-jacobi (dz time columns only)', it)
# block + schur
bidg = bid[ig]
S_explicit = (Ngg - Ngb @ sp.diags(1. / Dbb) @ Ngb.T).tocsr()
Mb3, mx = block_jacobi(S_explicit, bidg); xg, it = pcg(Sop, rg, Mb3); print('bias-schur + block-jacobi', it)
