import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
N = sp.load_npz('/tmp/N_cut.npz').tocsc(); M = np.load('/tmp/cut_meta.npz')
keep = M['keep']; nz0 = int(M['nz0']); dzs = M['dz_shape']; z0s = M['z0_shape']; c0 = int(M['dz_col0']); cN = int(M['dz_colN'])
d = N.diagonal(); S = sp.diags(1 / np.sqrt(d)); P = (S @ N @ S).tocsc()
w, V = spl.eigsh(P, k=12, sigma=0, which='LM')
wl = spl.eigsh(P, k=3, which='LA', return_eigenvectors=False)
print('smallest', w); print('largest', wl, 'cond', wl.max() / w.min())
is_z0 = keep < nz0; is_b = keep >= cN; is_dz = ~is_z0 & ~is_b
for j in range(6):
    v = V[:, j]
    e = [np.sum(v[m] ** 2) for m in (is_z0, is_dz, is_b)]
    # spatial roughness of z0 part
    z = np.zeros(z0s[0] * z0s[1]); z[keep[is_z0]] = (S @ v)[is_z0]; z = z.reshape(z0s)
    rough = np.sum(np.diff(z, axis=0) ** 2) + np.sum(np.diff(z, axis=1) ** 2)
    print(j, 'lam %.3e' % w[j], 'energy z0/dz/bias %.3f %.3f %.3f' % tuple(e), 'z0 roughness/norm %.3e' % (rough / max(np.sum(z ** 2), 1e-300)))
