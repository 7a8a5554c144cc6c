"""The low-level fits of BASELINE.json configs[0..1] written against the reference's public API (fd_grid / lin_op /
sparse matrices handed to a solve(A, b)): the SAME code builds the systems with the reference package (fixture generator,
oracle/make_golden.py) and with lssurf_b200 (tests).  Follows notebooks/smooth_fit_demo_aniso.ipynb cells 3-6, 11-13, 18-20
(the direction field of cell 12 is constant, so its interpolation is a constant) and SURVEY.md section 8(d) cfg 1b."""
import numpy as np
import scipy.sparse as sp


def _scaled_system(G_data, Gc, sigma, expected, z):
    """cells 6 / 13 / 20: stack, collect the expected deviations, scale both sides by their inverse."""
    Gcoo = sp.vstack([G_data.toCSR(), Gc.toCSR()]).tocoo()
    E = np.concatenate([np.ravel(sigma)] + [np.ravel(e) for e in expected])
    rhs = np.concatenate([np.ravel(z), np.zeros([Gc.shape[0]])], axis=0)
    TCinv = sp.dia_matrix((1 / E, 0), shape=(rhs.size, rhs.size))
    A = sp.csr_matrix(TCinv.dot(Gcoo))
    A.sum_duplicates()              # canonical form (what the digests are taken over); the solvers accept any format
    A.sort_indices()
    return A, TCinv.dot(rhs)


def build(L, make_data):
    """name -> (A, b, grid shape).  ``L``: the package (reference LSsurf or lssurf_b200); ``make_data(dict)``: its point container."""
    out = {}
    # ---- cfg 1b: 10 000 random points, 1 km z0 grid, curvature constraint only, one solve
    rng = np.random.default_rng(0)
    n = 10_000
    x, y = rng.uniform(-5e4, 5e4, n), rng.uniform(-5e4, 5e4, n)
    z = 50 * np.sin(2 * np.pi * x / 4e4) * np.cos(2 * np.pi * y / 3e4) + rng.normal(0, 1, n)
    data = make_data({'x': x, 'y': y, 'z': z, 'sigma': np.ones(n)})
    g = L.fd_grid([[-5e4, 5e4], [-5e4, 5e4]], 1000. * np.ones(2), name='z0')
    G_data = L.lin_op(g, name='interp_z').interp_mtx(data.coords()[0:2])
    grad2 = L.lin_op(g, name='grad2_z0').grad2(DOF='z0')
    grad2.expected = 1e-3 / np.sqrt(np.prod(g.delta)) * np.ones(grad2.N_eq)
    Gc = L.lin_op(None, name='constraints').vstack([grad2])
    out['cfg1b'] = _scaled_system(G_data, Gc, data.sigma, [grad2.expected], data.z) + (tuple(g.shape),)
    # ---- cfg 2: the anisotropic-smoothing notebook on a 401 x 401 grid
    data = make_data({'x': np.array([0.]), 'y': np.array([0.]), 'z': np.array([1.]), 'sigma': np.array([0.1])})
    g = L.fd_grid([[-10, 10], [-10, 10]], 0.05 * np.ones(2))
    G_data = L.lin_op(g, name='interp_z').interp_mtx(data.coords()[0:2])
    root_dA = np.sqrt(np.prod(g.delta))
    grad2 = L.lin_op(g, name='grad2_z0').grad2(DOF='z0')
    grad2.expected = 0.5 / root_dA * np.ones(grad2.N_eq)                                     # cell 4
    mag = L.lin_op(g, name='mag_z0').one(DOF='z0')
    mag.expected = 0.5 / root_dA * np.ones_like(mag.v.ravel())                               # cell 5
    Gc = L.lin_op(None, name='constraints').vstack([grad2, mag])
    out['cfg2_iso'] = _scaled_system(G_data, Gc, data.sigma, [grad2.expected, mag.expected], data.z) + (tuple(g.shape),)
    # cells 8-13: directional second derivative along theta, stencils padded to a common set of offsets
    coeffs = np.array([-1., 2., -1.]) / (g.delta[0] ** 2)
    stencils = {'d2zdx2': [([0, 0, 0], [-1, 0, 1]), coeffs], 'd2zdy2': [([-1, 0, 1], [0, 0, 0]), coeffs],
                'd2zdxdy': [([-1, -1, 1, 1], [-1, 1, -1, 1]), np.array([-1., 1., 1., -1]) / (4 * g.delta[0] ** 2)]}
    offsets = set()
    for st in stencils.values():
        offsets.update(set(tuple(jj) for jj in zip(*st[0])))
    for st in stencils.values():
        st[0] = tuple(list(o) for o in st[0])
        st[1] = list(st[1])
        have = set(tuple(jj) for jj in zip(*st[0]))
        for off in sorted(offsets):
            if off not in have:
                for dim, oo in enumerate(off):
                    st[0][dim].append(oo)
                st[1].append(0)
        st[1] = np.array(st[1], dtype=float)
    theta = 65 * np.pi / 180
    u, v = np.cos(theta), np.sin(theta)
    Axy = L.lin_op(grid=g).diff_op(*stencils['d2zdx2'])
    Axy.v *= u ** 2
    temp = L.lin_op(grid=g).diff_op(*stencils['d2zdxdy'])
    temp.v *= v
    temp.v *= u
    temp.v *= 2
    Axy.add(temp)
    temp = L.lin_op(grid=g).diff_op(*stencils['d2zdy2'])
    temp.v *= v ** 2
    Axy.add(temp)
    Axy.expected = np.zeros(Axy.shape[0]) + 1.0                                              # cell 13
    Gc = L.lin_op(None, name='constraints').vstack([Axy, mag])
    out['cfg2_aniso'] = _scaled_system(G_data, Gc, data.sigma, [Axy.expected, mag.expected], data.z) + (tuple(g.shape),)
    # cells 18-20: eight points on a circle, isotropic constraint with a smaller expected curvature
    pts = 5 * np.exp(1j * np.arange(0, 2 * np.pi, np.pi / 4))
    data_c = make_data({'x': np.real(pts), 'y': np.imag(pts), 'z': np.ones(pts.shape), 'sigma': 0.1 * np.ones(pts.shape)})
    G_data_c = L.lin_op(g, name='interp_z').interp_mtx(data_c.coords()[0:2])
    mag.expected[:] = 2                                                                      # cell 16
    grad2.expected = 0.25 / root_dA * np.ones(grad2.N_eq)
    Gc = L.lin_op(None, name='constraints').vstack([grad2, mag])
    out['cfg2_circle'] = _scaled_system(G_data_c, Gc, data_c.sigma, [grad2.expected, mag.expected], data_c.z) + (tuple(g.shape),)
    return out
