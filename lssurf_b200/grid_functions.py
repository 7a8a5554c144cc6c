"""Grid setup for the surface fit (API mirror of the no-mask-file branch of LSsurf/grid_functions.py:26-175)."""
import numpy as np

from .fd_grid import fd_grid
from .lin_op import lin_op


_warned_proj = [False]


def calc_cell_area(grid):
    """Cell area of each (y, x) node: delta_y * delta_x (grid_functions.py:169-175).

    The reference multiplies by the polar-stereographic scale factor when ``srs_proj4`` is set; that
    needs pointCollection / a projection library and is out of scope, so planar areas are used (a
    warning is printed once).  Cell areas never enter the solve, only the output grids.
    """
    if grid.srs_proj4 is not None and not _warned_proj[0]:
        print('lssurf_b200.calc_cell_area: srs_proj4 given, but projected scale factors are out of scope; '
              'using planar cell areas')
        _warned_proj[0] = True
    return np.ones(grid.shape[0:2]) * grid.delta[0] * grid.delta[1]


def _tapered_sum_op(grid, kernel_size, sub0s=None, dims=None):
    """Tapered box sum around coarse-node centres: the ``taper=True, valid_equations_only=False`` case of
    lin_op.sum_to_grid3 (lin_op.py:404-488), which is what sum_cell_area needs."""
    kernel_size = np.asarray(kernel_size)
    half = np.floor((kernel_size - 1) / 2).astype(int)
    if dims is None:
        dims = range(len(grid.shape))
    if sub0s is None:
        sub0s = np.meshgrid(*[np.arange(0, grid.shape[d] + 1, kernel_size[d] - 1, dtype=int) for d in dims], indexing='ij')
    centres = grid.global_ind(sub0s[0:len(dims)])
    odd = np.mod(kernel_size[0] / 2, 1) != 0
    di, dj = np.meshgrid(np.arange(-half[0], half[0] + (1 if odd else 0)),
                         np.arange(-half[1], half[1] + (1 if odd else 0)), indexing='ij')
    wt = np.ones(kernel_size[0:2], dtype=float)
    for edge in [0, -1]:
        wt[edge, :] /= 2
        wt[:, edge] /= 2
    offs = [di.ravel(), dj.ravel(), np.zeros_like(di.ravel())][:len(grid.shape)]
    return lin_op(grid=grid).diff_op(offs, wt.ravel().astype(float), which_nodes=centres, valid_equations_only=False)


def sum_cell_area(grid_f, grid_c, cell_area_f=None, return_op=False, sub0s=None, taper=True):
    """Area of the (masked) fine-grid cells inside each coarse cell (grid_functions.py:136-167)."""
    if cell_area_f is None:
        cell_area_f = calc_cell_area(grid_f) * grid_f.mask
    dims = list(range(cell_area_f.ndim))
    if not return_op:
        same = all(len(grid_f.ctrs[d]) == len(grid_c.ctrs[d]) and np.all(grid_f.ctrs[d] == grid_c.ctrs[d]) for d in (0, 1))
        if same:
            return cell_area_f.copy()
    n_k = (grid_c.delta[0:2] / grid_f.delta[0:2] + 1).astype(int)
    if cell_area_f.ndim == 3:
        n_k = np.array(list(n_k) + [1])
    tmp = fd_grid([grid_f.bds[d] for d in dims], deltas=grid_f.delta[dims])
    op = _tapered_sum_op(tmp, n_k, sub0s=sub0s, dims=dims)
    result = op.toCSR().dot(cell_area_f.ravel()).reshape(grid_c.shape[dims])
    if return_op:
        return result, op
    return result


def setup_grids(args):
    """z0 / dz / t grids for a fit (grid_functions.py:26-134): bounds = ctr +- W/2 (:46), the dz columns
    start right after the z0 columns (:80), cell areas from the grid spacing."""
    bds = {k: args['ctr'][k] + np.array([-0.5, 0.5]) * args['W'][k] for k in ('x', 'y', 't')}
    if args.get('mask_file') is not None:
        raise NotImplementedError('mask_file needs GDAL (out of scope); use mask_data with NumPy arrays')
    mask_data = args.get('mask_data')
    z0_mask = dz_mask = None
    if isinstance(mask_data, dict):
        z0_mask, dz_mask = mask_data.get('z0'), mask_data.get('dz')
    elif mask_data is not None:
        raise NotImplementedError('mask_data must be a dict {"z0": array, "dz": array} of NumPy masks')
    if args.get('lagrangian_coords') is not None:
        raise NotImplementedError('lagrangian_dz grids are outside the hot-path scope (SURVEY.md section 8(f))')
    srs = args.get('srs_proj4')
    erode = args.get('erode_source_mask', True)
    grids = dict()
    grids['z0'] = fd_grid([bds['y'], bds['x']], args['spacing']['z0'] * np.ones(2), name='z0', srs_proj4=srs,
                          mask_data=z0_mask, erode_source_mask=erode)
    grids['dz'] = fd_grid([bds['y'], bds['x'], bds['t']],
                          [args['spacing']['dz'], args['spacing']['dz'], args['spacing']['dt']],
                          name='dz', col_0=grids['z0'].N_nodes, srs_proj4=srs, mask_data=dz_mask,
                          mask_interp_threshold=0.5 if isinstance(mask_data, dict) else 0.95,
                          erode_source_mask=erode)
    grids['t'] = fd_grid([bds['t']], [args['spacing']['dt']], name='t')
    grids['z0'].cell_area = calc_cell_area(grids['z0'])
    if np.any(grids['dz'].delta[0:2] >= grids['z0'].delta):
        grids['dz'].cell_area = sum_cell_area(grids['z0'], grids['dz'])
    else:
        grids['dz'].cell_area = calc_cell_area(grids['dz'])
        if grids['dz'].mask is not None:
            grids['dz'].cell_area = grids['dz'].cell_area * grids['dz'].mask
    if grids['z0'].mask is not None:
        grids['z0'].cell_area = grids['z0'].cell_area * grids['z0'].mask
    return grids, bds


def setup_averaging_ops(grid, col_N, args, cell_area=None):
    """Un-averaged dz/dt operators per lag (grid_functions.py:187-214, no 3-D mask).  The block
    averages (``avg_scales``) are next-tier."""
    ops = {}
    if args.get('dzdt_lags') is not None:
        for lag in args['dzdt_lags']:
            name = 'dzdt_lag' + str(lag)
            op = lin_op(grid, name=name, col_N=col_N).dzdt(lag=lag).ravel()
            op.dst_grid.cell_area = grid.cell_area
            op.normalize_by_unit_product()
            ops[name] = op
    if args.get('avg_scales'):
        raise NotImplementedError('avg_scales block averages are next-tier (SURVEY.md section 8(f) #2)')
    return ops
