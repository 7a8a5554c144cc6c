"""CPU analysis (planning for round 2): 128-byte lines touched by the gather of 32 consecutive entries of a row of the
explicit normal matrix (full cfg-4 pattern from sell_pattern.py) under different orderings of the unknowns."""
import numpy as np
d = np.load('/tmp/npattern.npz')
indptr, indices, e0 = d['indptr'].astype(np.int64), d['indices'].astype(np.int64), int(d['e0'])
n = indptr.size - 1
nz0 = 401 * 401
ny = nx = 101; nt = 21
col = np.arange(e0)
# coordinates of every grid column
iy0, ix0 = np.divmod(np.arange(nz0), 401)
k = np.arange(e0 - nz0); dy = k // (nx * nt); dx = (k // nt) % nx; dt = k % nt
# block (nearest dz node) of every column
by = np.r_[np.clip((iy0 + 2) // 4, 0, ny - 1), dy]; bx = np.r_[np.clip((ix0 + 2) // 4, 0, nx - 1), dx]


def lines_per_32(perm_of_col, rows_order=None, samples=20000, seed=0):
    """perm_of_col[c] = new index of column c (bias columns keep their place after the grid columns)"""
    newc = np.where(indices < e0, perm_of_col[np.minimum(indices, e0 - 1)], indices)
    rng = np.random.default_rng(seed)
    tot = 0; cnt = 0
    rows = rng.choice(n, 4000, replace=False)
    for r in rows:
        c = np.sort(newc[indptr[r]:indptr[r + 1]])          # a row is stored with ascending (new) columns
        for s in range(0, c.size, 32):
            tot += np.unique(c[s:s + 32] >> 4).size; cnt += 1
    return tot / cnt


ident = np.arange(e0)
print('current order (z0 row-major, then dz with time fastest): %.2f lines per 32 entries' % lines_per_32(ident))
# (a) block order: all unknowns of a preconditioner block contiguous (z0 nodes of the block row-major, then its epochs)
key = (by * nx + bx) * 64 + np.r_[(iy0 % 4) * 4 + (ix0 % 4), 16 + dt]
order = np.argsort(key, kind='stable'); perm = np.empty(e0, dtype=np.int64); perm[order] = np.arange(e0)
print('block order (16 z0 nodes + 21 epochs of a dz node contiguous, blocks row-major): %.2f' % lines_per_32(perm))
# (b) block order with blocks along a Morton (Z) curve
def morton(a, b):
    z = np.zeros_like(a)
    for i in range(8): z |= ((a >> i) & 1) << (2 * i + 1) | ((b >> i) & 1) << (2 * i)
    return z
key = morton(by, bx) * 64 + np.r_[(iy0 % 4) * 4 + (ix0 % 4), 16 + dt]
order = np.argsort(key, kind='stable'); perm = np.empty(e0, dtype=np.int64); perm[order] = np.arange(e0)
print('block order, blocks along a Morton curve: %.2f' % lines_per_32(perm))
# (c) z0 tiled 4 x 4 only (dz untouched)
key = np.r_[((iy0 // 4) * 101 + ix0 // 4) * 16 + (iy0 % 4) * 4 + ix0 % 4, nz0 + k]
order = np.argsort(key, kind='stable'); perm = np.empty(e0, dtype=np.int64); perm[order] = np.arange(e0)
print('z0 in 4 x 4 tiles, dz unchanged: %.2f' % lines_per_32(perm))
