"""CPU emulation of the generic solver form (Jacobi-preconditioned CG on N = A^T W A, rtol 1e-12, reference-epoch columns
fixed) on the widened fixtures, with the operators the package assembles under the test double: does plain CG reach the
parity tolerance (m within 1e-6) on fits with slope / sensor-grid / edited-bias columns?  Build container only.
Result (round 1): tracks_sensor_slope 2 683 iterations, 5e-9 from the direct solution; uniform_sigma_grid 515, 3e-11;
tracks_bias_edit 1 346, 1e-9."""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0, '/root/repo')
import lssurf_b200 as L
from tests import host_double
from tests.util import load_case
for name in ['tracks_sensor_slope', 'uniform_sigma_grid', 'tracks_bias_edit']:
    fx, pts, kw = load_case(name)
    with host_double.installed():
        fo = L.smooth_fit(data=L.point_data.data().from_dict(pts), VERBOSE=False, return_fit_objects=True, **kw)
        G, C = fo['G_data'].toCSR(), fo['Gc'].toCSR()
    n = max(G.shape[1], C.shape[1])
    G = sp.csr_matrix((G.data, G.indices, G.indptr), shape=(G.shape[0], n)); C = sp.csr_matrix((C.data, C.indices, C.indptr), shape=(C.shape[0], n))
    Ed, Ec = np.asarray(fo['Ed']), np.asarray(fo['Ec'])
    A = sp.vstack([G, C]).tocsr(); w = np.r_[1/Ed**2, 1/Ec**2]
    z = np.asarray(fo['data'].z, float)
    b = np.r_[z, np.zeros(C.shape[0])]
    N = (A.T @ sp.diags(w) @ A).tocsr(); g = A.T @ (w*b)
    d = N.diagonal(); free = d > 0
    dz = fo['grids']['dz']; fixed = (dz.col_0 + np.arange(dz.N_nodes).reshape(dz.shape)[:, :, kw['reference_epoch']]).ravel(); free[fixed] = False
    # fixed columns (reference epoch) have their own constraint? emulate: restrict to columns with nonzero diag
    idx = np.flatnonzero(free); Nf = N[idx][:, idx]; gf = g[idx]
    xd = spl.spsolve(Nf.tocsc(), gf)
    M = 1/Nf.diagonal()
    x = np.zeros_like(gf); r = gf.copy(); zz = M*r; p = zz.copy(); rz = r@zz; r0 = np.sqrt(r@r); it = 0
    while it < 50000:
        q = Nf@p; a = rz/(p@q); x += a*p; r -= a*q; it += 1
        if np.sqrt(r@r) < 1e-12*r0: break
        zz = M*r; rz2 = r@zz; p = zz + (rz2/rz)*p; rz = rz2
    print(name, 'n', idx.size, 'rowlen', int(np.diff(G.indptr).max()), 'jacobi-PCG its', it, 'rel err vs direct', np.linalg.norm(x-xd)/np.linalg.norm(xd), 'cond-ish', d[free].max()/d[free].min())
