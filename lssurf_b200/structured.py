"""Grid-structured storage of the normal matrix N = A^T W A for the CG product (symbolic part; ``lss_solver_set_structured``).

The CSR form of N costs 12 bytes per entry and gathers p entry by entry.  On the z0(y, x) + dz(y, x, t) grids of
smooth_fit (LSsurf/smooth_fit.py:354-662 builds exactly these unknowns) almost all of N falls into four regular families,
which are stored without column indices and read p in contiguous pieces:

  zd    z0 node k  x  dz node a : one value per epoch (the data rows tie z0 to dz(t) at the same place).  Stored ONCE
        per (a, k) pair, grouped by dz node, and used for both triangles: q_dz[a, t] += v[t] p_z0[k], q_z0[k] += v . p_dz[a, :]
  dd    dz node a  x  dz node b != a : tridiagonal in time (interpolation in t and the d/dt stencils couple t with t +- 1)
  self  dz node a  x  itself: five diagonals in time (d2z/dt2 constraint rows)
  zz    z0 node  x  z0 node : a fixed set of (dy, dx) offsets inside 5 x 5 (interpolation + second-derivative stencils),
        stored by diagonal
Everything else (bias columns, longer-range couplings, rows outside the two grids) stays in a remainder CSR.

This module only builds index tables with tensor operations (on whatever device the pattern lives on; the CPU tier checks
them against SciPy).  Values are filled on the device from the CSR values through ``map`` after every numeric update.
"""
import numpy as np
import torch

MAX_ROUNDS = 8     # z0-dz pairs per dz node <= MAX_ROUNDS * (256 // nt): what the product kernel keeps in registers


def build(indptr, indices, rows, z0_col0, z0_shape, dz_col0, dz_shape, min_diag_fill=0.25):
    """indptr / indices: CSR pattern of N (int32 tensors); rows: leading rows the product covers.
    Returns a dict of tensors and sizes, or None when the layout does not apply."""
    dev = indptr.device
    ny0, nx0 = (int(v) for v in z0_shape)
    ny, nx, nt = (int(v) for v in dz_shape)
    n0, nn = ny0 * nx0, ny * nx
    if nt > 32 or nt < 2 or rows < max(z0_col0 + n0, dz_col0 + nn * nt):
        return None
    i64 = torch.int64
    ptr = indptr[:rows + 1].to(i64)
    nnz = int(ptr[-1].item())
    if nnz == 0 or nnz >= 2 ** 31:
        return None
    col = indices[:nnz].to(i64)
    row = torch.repeat_interleave(torch.arange(rows, device=dev), ptr[1:] - ptr[:-1])
    r_z0 = (row >= z0_col0) & (row < z0_col0 + n0)
    c_z0 = (col >= z0_col0) & (col < z0_col0 + n0)
    r_dz = (row >= dz_col0) & (row < dz_col0 + nn * nt)
    c_dz = (col >= dz_col0) & (col < dz_col0 + nn * nt)
    slot = torch.full((nnz,), -1, dtype=i64, device=dev)        # -1: remainder, -2: transposed copy of a zd entry (dropped)
    T = nt
    S = 256 // T                # (slot, t) threads of one CTA of the product kernel

    # ---- self / dd: dz rows x dz columns
    m = r_dz & c_dz
    e = torch.nonzero(m).reshape(-1)
    ra, rt = torch.div(row[e] - dz_col0, T, rounding_mode='floor'), (row[e] - dz_col0) % T
    ca, ct = torch.div(col[e] - dz_col0, T, rounding_mode='floor'), (col[e] - dz_col0) % T
    dt = ct - rt
    same = (ra == ca) & (dt.abs() <= 2)
    nb = (ra != ca) & (dt.abs() <= 1)
    key = ra[nb] * nn + ca[nb]
    ukey, inv = torch.unique(key, return_inverse=True)
    n_dd = int(ukey.numel())
    dd_a = torch.div(ukey, nn, rounding_mode='floor')
    dd_first = torch.searchsorted(dd_a, torch.arange(nn + 1, device=dev))
    # every node gets the same number of neighbour slots (the largest count): addresses follow from the node id alone
    dd_stride = int((dd_first[1:] - dd_first[:-1]).max().item()) if n_dd else 0
    if dd_stride > 2 * S:
        return None
    j_in = torch.arange(n_dd, device=dev) - dd_first[:-1][dd_a]                # rank of the pair among its node's neighbours
    dd_b = torch.full((nn * max(dd_stride, 1),), -1, dtype=torch.int32, device=dev)
    dd_b[dd_a * dd_stride + j_in] = (ukey % nn).to(torch.int32)
    # one block per node, a multiple of 16 bytes (one bulk copy): [3 bands][dd_stride slots][T] neighbour values, the
    # neighbour ids (int32, two per double), [5 bands][T] self values
    ids_len = (dd_stride + 1) // 2
    self_off = 3 * dd_stride * T + ids_len
    node_len = (self_off + 5 * T + 1) // 2 * 2
    slot[e[nb]] = ra[nb] * node_len + ((dt[nb] + 1) * dd_stride + j_in[inv]) * T + rt[nb]
    slot[e[same]] = ra[same] * node_len + self_off + (dt[same] + 2) * T + rt[same]
    del e, ra, rt, ca, ct, dt, same, nb, key, inv

    # ---- zd: z0 rows x dz columns (kept), dz rows x z0 columns (dropped: the same values, used transposed)
    m = r_z0 & c_dz
    e = torch.nonzero(m).reshape(-1)
    a = torch.div(col[e] - dz_col0, T, rounding_mode='floor')
    t = (col[e] - dz_col0) % T
    k0 = row[e] - z0_col0
    key = a * n0 + k0
    ukey, inv = torch.unique(key, return_inverse=True)
    n_zd = int(ukey.numel())
    # one record per pair, a multiple of 16 bytes: T values, then the z0 column (int64 in a double's place)
    rec_len = (T + 2) // 2 * 2
    zd_base = nn * node_len
    slot[e] = zd_base + inv * rec_len + t
    mt = r_dz & c_z0
    et = torch.nonzero(mt).reshape(-1)
    key_t = torch.div(row[et] - dz_col0, T, rounding_mode='floor') * n0 + (col[et] - z0_col0)
    # the transposed copies may only be dropped when N's pattern is symmetric there
    if et.numel() != e.numel() or not torch.equal(torch.unique(key_t * T + (row[et] - dz_col0) % T), torch.unique(key * T + t)):
        return None
    slot[et] = -2
    zd_a = torch.div(ukey, n0, rounding_mode='floor')
    zd_k = (ukey % n0 + z0_col0).to(torch.int32)
    zd_ptr = torch.searchsorted(zd_a, torch.arange(nn + 1, device=dev))
    if n_zd and int((zd_ptr[1:] - zd_ptr[:-1]).max().item()) > S * MAX_ROUNDS:
        return None
    del e, et, a, t, k0, key, key_t, inv, mt

    # ---- zz: z0 rows x z0 columns inside the 5 x 5 neighbourhood, by diagonal
    m = r_z0 & c_z0
    e = torch.nonzero(m).reshape(-1)
    ri, cj = row[e] - z0_col0, col[e] - z0_col0
    dy = torch.div(cj, nx0, rounding_mode='floor') - torch.div(ri, nx0, rounding_mode='floor')
    dx = cj % nx0 - ri % nx0
    near = (dy.abs() <= 2) & (dx.abs() <= 2)
    d = (dy + 2) * 5 + (dx + 2)
    fill = torch.bincount(d[near], minlength=25)
    used = fill >= max(1, int(min_diag_fill * n0))
    rank = torch.cumsum(used.to(i64), 0) - 1
    n_diag = int(used.sum().item())
    zz_base = zd_base + n_zd * rec_len
    ok = near & used[d.clamp(0, 24)]
    slot[e[ok]] = zz_base + rank[d[ok]] * n0 + ri[ok]
    dd_ = torch.nonzero(used).reshape(-1)
    zz_off = ((torch.div(dd_, 5, rounding_mode='floor') - 2) * nx0 + (dd_ % 5 - 2)).to(torch.int32)
    n_vals = zz_base + n_diag * n0
    del e, ri, cj, dy, dx, near, d, ok
    if n_vals >= 2 ** 31:
        return None

    # ---- remainder CSR over the non-empty rows
    rem = torch.nonzero(slot == -1).reshape(-1)
    rem_rows_all = row[rem]
    rr, cnt = torch.unique_consecutive(rem_rows_all, return_counts=True)
    rem_indptr = torch.zeros(rr.numel() + 1, dtype=torch.int32, device=dev)
    rem_indptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)

    i32 = torch.int32
    return dict(nt=T, nn=nn, n0=n0, n_diag=n_diag, z0_col0=int(z0_col0), dz_col0=int(dz_col0), rows=int(rows), nnz=nnz,
                n_vals=int(n_vals), node_len=int(node_len), rec_len=int(rec_len), zd_base=int(zd_base), zz_base=int(zz_base),
                n_dd=n_dd, n_zd=n_zd, dd_stride=dd_stride, max_pairs=int((zd_ptr[1:] - zd_ptr[:-1]).max().item()) if n_zd else 0, map=slot.clamp(min=-1).to(i32).contiguous(),
                dd_b=dd_b.contiguous(), zd_ptr=zd_ptr.to(i32).contiguous(), zd_k=zd_k.contiguous(), zz_off=zz_off.contiguous(),
                rem_src=rem.to(i32).contiguous(), rem_rows=rr.to(i32).contiguous(), rem_indptr=rem_indptr.contiguous(),
                rem_indices=col[rem].to(i32).contiguous(),
                counts=dict(self=int((slot >= 0).sum().item()), dropped=int((slot == -2).sum().item()), remainder=int(rem.numel())))


def product_bytes(s):
    """Bytes one product streams from this layout: every stored value once (8 B), the index tables, the remainder as CSR
    (12 B per entry + compact row pointers and row ids), q written and p read once."""
    idx = 4 * (s['zd_ptr'].numel() + s['zz_off'].numel())          # neighbour ids and z0 columns travel inside the value buffer
    rem = 16 * s['rem_src'].numel()                                 # flat entries: row, column, value
    return 8 * s['n_vals'] + idx + rem + 16 * s['rows']


def emulate(s, data, p):
    """NumPy restatement of the structured product (what the CUDA kernel computes), for the CPU tier: data = CSR values of N
    in the order of the pattern handed to build(), p = vector; returns q over the leading rows."""
    g = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in s.items()}
    T, nn, n0 = g['nt'], g['nn'], g['n0']
    vals = np.zeros(g['n_vals'])
    m = g['map']
    vals[m[m >= 0]] = np.asarray(data)[:g['nnz']][m >= 0]
    q = np.zeros(g['rows'])
    pd = p[g['dz_col0']:g['dz_col0'] + nn * T].reshape(nn, T)
    qd = np.zeros((nn, T))
    ds, nl, rl = g['dd_stride'], g['node_len'], g['rec_len']
    blk = vals[:nn * nl].reshape(nn, nl)
    self_off = 3 * ds * T + (ds + 1) // 2
    sv = blk[:, self_off:self_off + 5 * T].reshape(nn, 5, T)
    for band in range(-2, 3):
        src = np.zeros((nn, T))
        lo, hi = max(0, -band), min(T, T - band)
        src[:, lo:hi] = pd[:, lo + band:hi + band]
        qd += sv[:, band + 2, :] * src
    dv = blk[:, :3 * ds * T].reshape(nn, 3, ds, T)
    nbr = g['dd_b'].reshape(nn, ds) if ds else np.zeros((nn, 0), dtype=np.int32)
    for j in range(ds):
        b = nbr[:, j]
        pb = np.where((b >= 0)[:, None], pd[np.maximum(b, 0)], 0.)
        for band in range(-1, 2):
            src = np.zeros((nn, T))
            lo, hi = max(0, -band), min(T, T - band)
            src[:, lo:hi] = pb[:, lo + band:hi + band]
            qd += dv[:, band + 1, j, :] * src
    zv = vals[g['zd_base']:g['zd_base'] + g['n_zd'] * rl].reshape(-1, rl)[:, :T]
    a_of = np.repeat(np.arange(nn), np.diff(g['zd_ptr']))
    np.add.at(qd, a_of, zv * p[g['zd_k']][:, None])
    np.add.at(q, g['zd_k'], np.einsum('pt,pt->p', zv, pd[a_of]))
    q[g['dz_col0']:g['dz_col0'] + nn * T] += qd.ravel()
    zz = vals[g['zz_base']:g['zz_base'] + g['n_diag'] * n0].reshape(-1, n0)
    pz = p[g['z0_col0']:g['z0_col0'] + n0]
    for d, off in enumerate(g['zz_off']):
        i = np.arange(max(0, -off), min(n0, n0 - off))
        q[g['z0_col0'] + i] += zz[d, i] * pz[i + off]
    rv = np.asarray(data)[g['rem_src']]
    for r in range(g['rem_rows'].size):
        b0, b1 = g['rem_indptr'][r], g['rem_indptr'][r + 1]
        q[g['rem_rows'][r]] += rv[b0:b1] @ p[g['rem_indices'][b0:b1]]
    return q
