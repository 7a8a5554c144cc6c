"""CPU tier: host-side mirror of the reference interface (no kernels involved)."""
import os
import sys

import numpy as np
import pytest
import torch

from lssurf_b200 import fd_grid, unique_by_rows
from lssurf_b200.bias_functions import assign_bias_ID
from lssurf_b200.parallel import DistContext
from lssurf_b200.point_data import data as point_data
from lssurf_b200.smooth_fit import DEFAULTS, smooth_fit
from oracle import lssurf_oracle as O


@pytest.mark.parametrize('bounds,deltas', [([[-5e4, 5e4], [-5e4, 5e4]], [250., 250.]),
                                            ([[-3., 5.], [100., 131.]], [0.5, 1.7]),
                                            ([[-1e3, 1e3], [-2e3, 1e3], [-1., 1.]], [100., 100., 0.25]),
                                            ([[0., 0.3]], [0.1])])
def test_fd_grid_geometry(bounds, deltas):
    g = fd_grid(bounds, deltas, col_0=11)
    o = O.grid_geometry(bounds, deltas, col_0=11)
    assert np.array_equal(g.shape, o['shape']) and np.array_equal(g.stride, o['stride'])
    assert g.col_N == o['col_N'] and g.N_nodes == o['n_nodes']
    for d in range(len(bounds)):
        assert np.array_equal(g.ctrs[d], o['ctrs'][d]) and g.bds[d][0] == o['lo'][d] and g.bds[d][1] == o['hi'][d]
    s = g.as_struct()
    assert s.ndim == len(bounds) and s.col_0 == 11 and s.shape[0] == g.shape[0]


def test_fd_grid_point_helpers():
    g = fd_grid([[0., 10.], [0., 20.]], [1., 2.], col_0=3)
    pts = [np.array([0., 10., 10.0001, np.nan, 3.5]), np.array([0., 20., 1., 1., 7.])]
    assert g.validate_pts(pts).tolist() == [True, True, False, False, True]
    assert np.array_equal(g.global_ind([np.array([1.]), np.array([2.])]), [3 + 11 + 2])
    ind, ok = g.global_ind([np.array([1, 11]), np.array([2, 0])], return_valid=True)
    assert ok.tolist() == [True, False] and ind[1] == 0
    sub = g.cell_sub_for_pts(pts)
    assert sub[0][4] == 3 and sub[1][4] == 3 and np.isnan(sub[0][2])


def test_unique_by_rows_and_bias_ids():
    x = np.array([[2., 1.], [1., 5.], [2., 1.], [1., np.nan], [1., 5.]])
    x2 = np.nan_to_num(x, nan=-9999)
    u, d = unique_by_rows(x2, return_dict=True)
    assert u.tolist() == [[1., -9999.], [1., 5.], [2., 1.]]
    assert sorted(d[(2., 1.)].tolist()) == [0, 2]
    D = point_data().from_dict({'rgt': x[:, 0], 'cycle': x[:, 1], 'sigma_corr': np.array([.1, .2, .3, .4, .6]), 'z': np.zeros(5)})
    D, bm = assign_bias_ID(D, ['rgt', 'cycle'])
    ids, _ = O.bias_ids([x[:, 0], x[:, 1]])
    assert np.array_equal(D.bias_ID.astype(int), ids)
    assert bm['E_bias'][1] == pytest.approx(0.4) and np.isnan(bm['bias_ID_dict'][0]['cycle'])


def test_smooth_fit_argument_errors():
    with pytest.raises(ValueError):
        smooth_fit(data=point_data().from_dict({'z': np.zeros(3)}))
    with pytest.raises(ValueError):          # checked before any device work
        smooth_fit(data=point_data().from_dict({'z': np.zeros(3)}), W={}, ctr={}, spacing={}, E_RMS={}, gather_outputs='everyone')
    assert DEFAULTS['max_iterations'] == 10 and DEFAULTS['mask_scale'] == {0: 10, 1: 1} and DEFAULTS['gather_outputs'] == 'all'
    assert DEFAULTS['converge_tol_dz'] == 0.05 and DEFAULTS['min_iterations'] == 2


def test_shard_bounds():
    for n in (0, 1, 7, 10_000_001):
        for world in (1, 2, 4, 8):
            cuts = [DistContext(r, world).shard(n) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ctx = DistContext(rank, world, None)
    lo, hi = ctx.shard(11)
    out = ctx.gather_rows(np.arange(lo, hi, dtype=float) * 2)
    # the data-path reduction every rank must agree on (what ncclAllReduce does on the GPU)
    t = torch.tensor([float(hi - lo), float(rank)], dtype=torch.float64)
    dist.all_reduce(t)
    # replicated solve plumbing: per-rank CSR row blocks (different sizes) stacked in rank order on every rank
    import scipy.sparse as sp
    from lssurf_b200.lin_op import DeviceCSR
    M = sp.random(11, 6, density=0.4, random_state=3, format='csr')
    blk = M[lo:hi].tocsr()
    A = DeviceCSR(torch.from_numpy(blk.indptr.astype(np.int32)), torch.from_numpy(blk.indices.astype(np.int32)),
                  torch.from_numpy(blk.data), blk.shape)
    F = ctx.gather_csr(A)
    same = (F.shape == M.shape and np.array_equal(F.indptr.numpy(), M.indptr) and np.array_equal(F.indices.numpy(), M.indices)
            and np.array_equal(F.data.numpy(), M.data) and ctx.gather_counts(hi - lo) == [5, 6])
    x = torch.full((3,), float(rank + 1), dtype=torch.float64)
    ctx.broadcast(x, src=0)
    q.put((rank, out.tolist(), t.tolist() + [float(same), float(x.sum())]))
    dist.destroy_process_group()


def test_row_sharding_world2_gloo():
    """World-size-2 host logic on CPU: shards tile the rows, gathers restore the row order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, gathered, red in res:
        assert gathered == (np.arange(11.) * 2).tolist()
        assert red == [11.0, 1.0, 1.0, 3.0]


def test_grid_structure_partition():
    """Preconditioner blocks handed to lss_solver_set_structure: every free z0 / dz column sits in exactly one block,
    a block is one dz node's free epochs plus the z0 nodes nearest to it, the bias range is passed through."""
    from lssurf_b200.solver import grid_structure
    from lssurf_b200.constraint_functions import reference_epoch_columns
    by, bx, bt = [-3000., 3000.], [-4000., 4000.], [-1., 1.]
    z0 = fd_grid([by, bx], [250., 250.], name='z0')
    dz = fd_grid([by, bx, bt], [1000., 1000., 0.25], name='dz', col_0=z0.N_nodes)
    grids = {'z0': z0, 'dz': dz}
    fixed = reference_epoch_columns(grids, 3)
    bias = np.arange(dz.col_N, dz.col_N + 7)
    st = grid_structure(grids, fixed, bias_cols=bias)
    ptr, cols = st['block_ptr'].cpu().numpy(), st['block_cols'].cpu().numpy()
    assert st['elim'] == (dz.col_N, dz.col_N + 7)
    assert ptr[0] == 0 and ptr[-1] == cols.size and np.all(np.diff(ptr) > 0) and np.diff(ptr).max() == st['max_block'] <= 64
    free = np.setdiff1d(np.arange(dz.col_N), fixed)
    assert np.array_equal(np.sort(cols), free)                        # a partition of the free grid columns
    ny, nx, nt = dz.shape
    assert ptr.size - 1 == ny * nx
    for b in (0, nx + 1, ny * nx - 1):
        members = cols[ptr[b]:ptr[b + 1]]
        dz_part = members[members >= dz.col_0] - dz.col_0
        assert np.all(dz_part // nt == b) and dz_part.size == nt - 1   # all free epochs of one dz node
        z0_part = members[members < dz.col_0]
        iy, ix = np.divmod(z0_part, z0.shape[1])
        jy, jx = divmod(b, nx)
        assert np.all(np.abs(z0.ctrs[0][iy] - dz.ctrs[0][jy]) <= dz.delta[0] / 2)
        assert np.all(np.abs(z0.ctrs[1][ix] - dz.ctrs[1][jx]) <= dz.delta[1] / 2)
    assert grid_structure({'z0': z0}, fixed) is None                  # no dz grid: scalar Jacobi


def test_coarse_structure_is_a_bilinear_partition_of_unity():
    """Coarse space handed to lss_solver_set_coarse: every free grid column interpolates from <= 4 coarse unknowns with
    weights that sum to 1 (bilinear in space, identity in time), fixed / foreign columns have no parents, P^T (CSR by
    coarse row) holds exactly the same entries, and a fine node that is a coarse node has a single parent."""
    from lssurf_b200.solver import coarse_structure
    from lssurf_b200.constraint_functions import reference_epoch_columns
    by, bx, bt = [-20000., 20000.], [-30000., 30000.], [-1., 1.]
    z0 = fd_grid([by, bx], [250., 250.], name='z0')
    dz = fd_grid([by, bx, bt], [1000., 1000., 0.25], name='dz', col_0=z0.N_nodes)
    grids = {'z0': z0, 'dz': dz}
    fixed = reference_epoch_columns(grids, 4)
    n_cols = int(dz.col_N) + 5                                         # five bias columns behind the grids
    cs = coarse_structure(grids, fixed, n_cols)
    idx, w = cs['cp_idx'].cpu().numpy(), cs['cp_w'].cpu().numpy().astype(float)
    assert cs['n_c'] <= 1024 and idx.shape == (n_cols, 4) and idx.max() < cs['n_c']
    free = np.ones(n_cols, dtype=bool); free[fixed] = False; free[int(dz.col_N):] = False
    assert np.all(idx[~free] == -1)
    assert np.allclose(np.where(idx[free] >= 0, w[free], 0).sum(axis=1), 1.0, atol=1e-6)
    assert np.all((idx[free] >= 0).sum(axis=1) >= 1)
    # z0 and dz parents never mix; a dz column keeps its epoch
    nz0c = idx[:int(z0.N_nodes)].max() + 1
    assert np.all(idx[int(dz.col_0):int(dz.col_N)][idx[int(dz.col_0):int(dz.col_N)] >= 0] >= nz0c)
    nt = int(dz.shape[2])
    t_of = (np.arange(int(dz.N_nodes)) % nt)[:, None]
    pd = idx[int(dz.col_0):int(dz.col_N)]
    assert np.all(((pd - nz0c) % nt == t_of) | (pd < 0))
    # corner node of both grids: exactly one parent with weight 1
    assert (idx[0] >= 0).sum() == 1 and w[0][idx[0] >= 0][0] == 1.0
    # P^T as CSR
    ptr, col, pw = cs['pt_ptr'].cpu().numpy(), cs['pt_col'].cpu().numpy(), cs['pt_w'].cpu().numpy()
    assert ptr[0] == 0 and ptr[-1] == col.size == int((idx >= 0).sum())
    dense = {}
    for c in range(n_cols):
        for a in range(4):
            if idx[c, a] >= 0:
                dense[(int(idx[c, a]), c)] = np.float32(w[c, a])
    m = 7
    got = {(m, int(c)): np.float32(x) for c, x in zip(col[ptr[m]:ptr[m + 1]], pw[ptr[m]:ptr[m + 1]])}
    assert got == {k: v for k, v in dense.items() if k[0] == m}
    assert np.all(np.diff(col[ptr[m]:ptr[m + 1]]) >= 0)               # fixed entry order inside a coarse row


def test_mask_for_ind0_all_true_fast_path_matches_sampling():
    """The all-True shortcut of lin_op.mask_for_ind0 returns what sampling the mask node by node returns
    (LSsurf/lin_op.py:633-667), for the default mask_scale, a permuted one and keys written as booleans."""
    from lssurf_b200.lin_op import lin_op
    g = fd_grid([[0., 10.], [0., 20.]], [1., 2.], col_0=5)
    op = lin_op(g, name='probe')
    op.ind0 = g.col_0 + np.array([0, 3, 17, 120, 64])
    op.N_eq = 5
    partial = g.mask.copy()
    partial[0, 0] = False                                # forces the sampling path; node 0 is masked
    for ms in ({0: 10, 1: 1}, {1: 3.5, 0: 2.}, {True: 7., False: 9.}, {0: 4.}):
        fast = op.mask_for_ind0(ms)
        slow = op.mask_for_ind0(ms, mask=partial)
        assert fast.dtype == float and np.array_equal(fast[1:], slow[1:])
        assert slow[0] == ms.get(0, 0.0)
    assert op.mask_for_ind0(None).dtype == bool and op.mask_for_ind0(None).all()
