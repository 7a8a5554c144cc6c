"""CPU experiment 2: with the bias block eliminated (Schur complement), how far do grid preconditioners go?"""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, time
N = sp.load_npz('/tmp/N_cut.npz').tocsr(); rhs = np.load('/tmp/rhs_cut.npy'); M = np.load('/tmp/cut_meta.npz')
keep = M['keep']; nz0 = int(M['nz0']); dzs = M['dz_shape']; z0s = M['z0_shape']; c0 = int(M['dz_col0']); cN = int(M['dz_colN'])
n = N.shape[0]
is_z0 = keep < nz0; is_b = keep >= cN; is_dz = ~is_z0 & ~is_b
ib = np.flatnonzero(is_b); ig = np.flatnonzero(~is_b)
Ngg = N[ig][:, ig].tocsr(); Ngb = N[ig][:, ib].tocsr(); Dbb = N[ib][:, ib].diagonal()
S = (Ngg - Ngb @ sp.diags(1. / Dbb) @ Ngb.T).tocsr()
rg = rhs[ig] - Ngb @ (rhs[ib] / Dbb)
print('S nnz', S.nnz, 'Ngg nnz', Ngg.nnz)
Sop = lambda v: Ngg @ v - Ngb @ ((Ngb.T @ v) / Dbb)

def pcg(Aop, b, Minv, rtol=1e-12, maxit=3000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for k in range(maxit):
        q = Aop(p); a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rzn = r @ z
        if np.sqrt(rzn / rz0) <= rtol: return x, k + 1
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit

kg = keep[ig]; gz0 = kg < nz0
# grid coordinates (in dz-cell units) of every grid unknown
ratio = (z0s[0] - 1) // (dzs[0] - 1)
yy = np.empty(kg.size); xx = np.empty(kg.size)
k = kg[gz0]; yy[gz0], xx[gz0] = np.divmod(k, z0s[1]); yy[gz0] /= ratio; xx[gz0] /= ratio
k = kg[~gz0] - c0; yy[~gz0] = k // (dzs[1] * dzs[2]); xx[~gz0] = (k // dzs[2]) % dzs[1]

def schwarz(bs, overlap, restricted=True, Smat=S):
    """blocks of bs x bs dz cells (+ their z0 nodes); overlap in dz-cell units; RAS: only owned entries are written back"""
    by = np.floor((yy + 0.5) / bs).astype(int); bx = np.floor((xx + 0.5) / bs).astype(int)
    own = by * 1000 + bx
    blocks = []
    for b in np.unique(own):
        cy, cx = divmod(b, 1000)
        lo_y, hi_y = cy * bs - 0.5 - overlap, (cy + 1) * bs - 0.5 + overlap
        lo_x, hi_x = cx * bs - 0.5 - overlap, (cx + 1) * bs - 0.5 + overlap
        idx = np.flatnonzero((yy >= lo_y) & (yy < hi_y) & (xx >= lo_x) & (xx < hi_x))
        B = Smat[idx][:, idx].toarray()
        inv = np.linalg.inv(B)
        mine = own[idx] == b
        blocks.append((idx, inv, mine))
    def apply(r):
        z = np.zeros_like(r)
        for idx, inv, mine in blocks:
            t = inv @ r[idx]
            if restricted: z[idx[mine]] += t[mine]
            else: z[idx] += t
        return z
    return apply, max(len(b[0]) for b in blocks), sum(len(b[0]) ** 2 for b in blocks)

for bs, ov in ():
    t0 = time.time()
    Mi, mx, nn = schwarz(bs, ov, restricted=False)
    x, it = pcg(Sop, rg, Mi)
    print('additive schwarz block %d overlap %.1f: its %d  max block %d  inverse entries %.2e (%.2f x nnz S)  %.0fs' % (bs, ov, it, mx, nn, nn / S.nnz, time.time() - t0), flush=True)

print('--- blocks from N_gg instead of S, fp32 inverse')
Mi, mx, nn = schwarz(1, 0, restricted=False, Smat=Ngg)
x, it = pcg(Sop, rg, Mi); print('schur operator + block-jacobi from N_gg blocks:', it)
def schwarz32(Smat):
    by = np.floor(yy + 0.5).astype(int); bx = np.floor(xx + 0.5).astype(int); own = by * 1000 + bx
    order = np.argsort(own, kind='stable'); so = own[order]
    st = np.flatnonzero(np.r_[True, so[1:] != so[:-1]]); en = np.r_[st[1:], so.size]
    blocks = [(order[s:e], np.linalg.inv(Smat[order[s:e]][:, order[s:e]].toarray()).astype(np.float32).astype(np.float64)) for s, e in zip(st, en)]
    def apply(r):
        z = np.zeros_like(r)
        for idx, inv in blocks: z[idx] = 0.5 * (inv + inv.T) @ r[idx]
        return z
    return apply
x, it = pcg(Sop, rg, schwarz32(S)); print('schur + block-jacobi(S blocks) stored fp32:', it)
x, it = pcg(Sop, rg, schwarz32(Ngg)); print('schur + block-jacobi(N_gg blocks) stored fp32:', it)
