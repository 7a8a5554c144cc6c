import numpy as np
d = np.load('/tmp/npattern.npz')
indptr, indices, e0 = d['indptr'].astype(np.int64), d['indices'].astype(np.int64), int(d['e0'])
n = indptr.size - 1
lens = np.diff(indptr)
z0n = 160801
print('z0 rows len mean/max/min', lens[:z0n].mean(), lens[:z0n].max(), lens[:z0n].min(), ' dz rows', lens[z0n:].mean(), lens[z0n:].max(), lens[z0n:].min())
print('hist', np.percentile(lens, [1, 10, 25, 50, 75, 90, 99]))
for C in (8, 16, 32):
    npad = (n + C - 1) // C * C
    L = np.zeros(npad, dtype=np.int64); L[:n] = lens
    w = L.reshape(-1, C).max(axis=1)
    print('C', C, 'padded nnz', int((w * C).sum()), 'overhead', (w * C).sum() / lens.sum() - 1)
# gather sectors for SELL-32, thread per row: for entry k, distinct 32-byte sectors (4 doubles) over the 32 rows
C = 32
tot_sec = 0; tot_inst = 0; tot_lines = 0
rng = np.random.default_rng(0)
slices = rng.choice(n // C, 600, replace=False)
for s in slices:
    rows = np.arange(s * C, s * C + C)
    w = lens[rows].max()
    M = np.full((C, w), -1, dtype=np.int64)
    for i, r in enumerate(rows):
        M[i, :lens[r]] = indices[indptr[r]:indptr[r + 1]]
    for k in range(w):
        c = M[:, k]; c = c[c >= 0]
        tot_sec += np.unique(c >> 2).size; tot_lines += np.unique(c >> 4).size; tot_inst += 1
print('SELL-32 thread/row: sectors per gather instr', tot_sec / tot_inst, 'lines(128B)', tot_lines / tot_inst)
# CSR nnz-tiled: 32 consecutive nnz
tot_sec = 0; tot_inst = 0; tot_lines = 0
for s in rng.choice(indices.size // 32, 20000, replace=False):
    c = indices[s * 32:(s + 1) * 32]
    tot_sec += np.unique(c >> 2).size; tot_lines += np.unique(c >> 4).size; tot_inst += 1
print('CSR 32 consecutive nnz: sectors per gather instr', tot_sec / tot_inst, 'lines', tot_lines / tot_inst)
