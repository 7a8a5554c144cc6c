"""Seeded synthetic workloads for tests and bench.py.

All generators use ``np.random.default_rng(seed)`` and return plain dicts of fp64 numpy arrays
plus the smooth_fit keyword arguments of the configuration, so the same inputs can be handed to
the CUDA product, the CPU oracle and (in the build container) the unmodified reference.
Point schema follows the reference's ICESat-2 reader (LSsurf/examples/read_ICESat2.py:110-132):
x, y, z, time, sigma, sigma_corr, rgt, cycle, sensor.
"""
import numpy as np


def _surface(x, y, t):
    return 100. * np.cos(2 * np.pi * x / 2.e4) * np.sin(2 * np.pi * y / 3.e4) + t * (1. + np.cos(2 * np.pi * x / 4.e4))


def notebook_strip(outliers=False, seed=0):
    """cfg 1a: the demo-notebook strip (smooth_fit_demo.ipynb cells 4-6, 17, 24-28)."""
    W = {'x': 1.e4, 'y': 200, 't': 2}
    ctr = {'x': 0., 'y': 0., 't': 0.}
    spacing = {'z0': 50, 'dz': 100, 'dt': 0.25}
    x = np.arange(-W['x'] / 2, W['x'] / 2, 100)
    amp, lam = 100, 2000
    if not outliers:
        z = -amp * np.cos(2 * np.pi * x / lam)
        pts = dict(x=np.r_[x, x], y=np.zeros(2 * x.size), z=np.r_[z, z],
                   time=np.r_[np.zeros_like(x) - 0.5, np.zeros_like(x) + 0.5], sigma=np.ones(2 * x.size))
        E = {'d2z0_dx2': 0.006, 'dz0_dx': 0.06 * 2500, 'd3z_dx2dt': 0.0001, 'd2z_dxdt': 0.0001 * 2500,
             'd2z_dt2': 5000}
        kw = dict(W=W, ctr=ctr, spacing=spacing, E_RMS=E, reference_epoch=2, max_iterations=1)
        return pts, kw
    rng = np.random.default_rng(seed)
    z = np.r_[np.zeros_like(x), amp - amp * np.cos(2 * np.pi * x / lam)]
    z = z + rng.standard_normal(z.size)
    bad = np.flatnonzero(rng.random(z.size) > 0.90)
    z[bad] += (rng.random(bad.size) - 0.5) * 200
    pts = dict(x=np.r_[x, x], y=np.zeros(2 * x.size), z=z,
               time=np.r_[np.zeros_like(x) - 0.99, np.zeros_like(x) + 0.99], sigma=np.ones(2 * x.size))
    E = {'d2z0_dx2': 0.03, 'dz0_dx': 0.03 * 2500, 'd3z_dx2dt': 0.006, 'd2z_dxdt': 0.006 * 2500, 'd2z_dt2': 5000}
    kw = dict(W=W, ctr=ctr, spacing=spacing, E_RMS=E, reference_epoch=4, max_iterations=10)
    return pts, kw


def uniform_3d(n=1_000_000, width=1.e5, seed=0, z0_spacing=500., dz_spacing=1000., dt=0.25, t_span=5.,
               outlier_frac=0.03, sigma=0.1):
    """cfg 3: n points uniform in (x, y, t); z0+dz dh/dt fit with 3 % outliers."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-width / 2, width / 2, n)
    y = rng.uniform(-width / 2, width / 2, n)
    t = rng.uniform(-t_span / 2, t_span / 2, n)
    z = _surface(x, y, t) + rng.normal(0, sigma, n)
    bad = rng.random(n) < outlier_frac
    z[bad] += rng.normal(0, 20, int(bad.sum()))
    pts = dict(x=x, y=y, z=z, time=t, sigma=np.full(n, sigma))
    n_ep = int(round(t_span / dt))
    kw = dict(W={'x': width, 'y': width, 't': t_span}, ctr={'x': 0., 'y': 0., 't': 0.},
              spacing={'z0': z0_spacing, 'dz': dz_spacing, 'dt': dt},
              E_RMS={'d2z0_dx2': 0.02, 'dz0_dx': 50, 'd3z_dx2dt': 3e-4, 'd2z_dxdt': 0.75, 'd2z_dt2': 5000},
              reference_epoch=n_ep // 2, max_iterations=10)
    return pts, kw


def icesat2_tracks(n=10_000_000, width=1.e5, seed=0, z0_spacing=250., dz_spacing=1000., dt=0.25, t_span=5.,
                   n_cycles=20, track_sep=None, sigma=0.05, sigma_corr=0.03, outlier_frac=0.01,
                   bias_params=True):
    """cfg 4: ICESat-2-shaped repeat tracks (3 beam pairs x 2 beams, ascending + descending).

    Reference ground tracks alternate headings of +-(75 +- 3) degrees; each has beam pairs at
    0 / +-3.3 km and beams at +-45 m; every cycle (91 days apart) re-flies each track with N(0, 10 m)
    cross-track jitter; a per-(rgt, cycle) bias N(0, 0.03 m) and 1 % outliers are added.  The
    along-track spacing is chosen so the point count is >= n, then a seeded subset of exactly n
    points is kept (order preserved).
    """
    rng = np.random.default_rng(seed)
    half = width / 2
    if track_sep is None:
        track_sep = width / 16.5            # ~33 reference tracks over the two headings
    offs = np.arange(-half * 1.3, half * 1.3, track_sep)
    tracks = []
    for k, o in enumerate(offs):
        for sgn in (1., -1.):
            tracks.append((o + (0.25 * track_sep if sgn < 0 else 0.), sgn * np.deg2rad(75. + rng.uniform(-3, 3))))
    n_rgt = len(tracks)
    beams = [(p, b) for p in (-3300., 0., 3300.) for b in (-45., 45.)]
    cyc_t = (np.arange(n_cycles) - (n_cycles - 1) / 2.) * (91. / 365.25)
    L = 1.6 * half / np.sin(np.deg2rad(72.))
    jitter = rng.normal(0, 10., (n_rgt, n_cycles))

    def fly(ds, count_only=False):
        s = np.arange(-L, L, ds)
        xs, ys, ts, rg, cy, bp, cnt = [], [], [], [], [], [], 0
        for r, (x0, hd) in enumerate(tracks):
            ux, uy = np.cos(hd), np.sin(hd)          # along-track unit vector
            for c in range(n_cycles):
                for ib, (po, bo) in enumerate(beams):
                    xt = po + bo + jitter[r, c]
                    x = x0 + s * ux - xt * uy
                    y = s * uy + xt * ux
                    ok = (np.abs(x) < half) & (np.abs(y) < half)
                    m = int(ok.sum())
                    cnt += m
                    if m == 0 or count_only:
                        continue
                    xs.append(x[ok]); ys.append(y[ok])
                    ts.append(np.full(m, cyc_t[c]) + s[ok] / 7000. / 86400. / 365.25)
                    rg.append(np.full(m, float(r))); cy.append(np.full(m, float(c))); bp.append(np.full(m, float(ib)))
        return cnt, xs, ys, ts, rg, cy, bp

    ds0 = 2000. if width >= 2.e4 else 200.
    cnt0 = fly(ds0, count_only=True)[0]
    ds = ds0 * cnt0 / (1.05 * n)               # along-track spacing giving ~5 % more than n points
    _, xs, ys, ts, rg, cy, bp = fly(ds)
    x = np.concatenate(xs); y = np.concatenate(ys); t = np.concatenate(ts)
    rgt = np.concatenate(rg); cyc = np.concatenate(cy); beam = np.concatenate(bp)
    if x.size > n:
        keep = np.sort(rng.choice(x.size, n, replace=False))
        x, y, t, rgt, cyc, beam = x[keep], y[keep], t[keep], rgt[keep], cyc[keep], beam[keep]
    m = x.size
    bias = rng.normal(0, 0.03, (n_rgt, n_cycles))
    z = _surface(x, y, t) + bias[rgt.astype(int), cyc.astype(int)] + rng.normal(0, sigma, m)
    bad = rng.random(m) < outlier_frac
    z[bad] += rng.normal(0, 20, int(bad.sum()))
    pts = dict(x=x, y=y, z=z, time=t, sigma=np.full(m, sigma), sigma_corr=np.full(m, sigma_corr),
               rgt=rgt, cycle=cyc, sensor=np.zeros(m), BP=beam // 2, LR=beam % 2)
    n_ep = int(round(t_span / dt))
    kw = dict(W={'x': width, 'y': width, 't': t_span}, ctr={'x': 0., 'y': 0., 't': 0.},
              spacing={'z0': z0_spacing, 'dz': dz_spacing, 'dt': dt},
              E_RMS={'d2z0_dx2': 0.02, 'dz0_dx': 50, 'd3z_dx2dt': 3e-4, 'd2z_dxdt': 0.75, 'd2z_dt2': 5000},
              reference_epoch=n_ep // 2, max_iterations=10)
    if bias_params:
        kw['bias_params'] = ['rgt', 'cycle']
    return pts, kw
