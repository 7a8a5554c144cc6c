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


def _tapered_box_sum_host(area, n_k, shape_c):
    """The tapered box sums of _tapered_sum_op applied to a 2-D array, on the host: the same products added in the same
    (ascending column) order as the CSR product of the operator, entries outside the array left out.  None when the
    kernel or the centres fall outside the plain case (the caller builds the operator instead)."""
    k0, k1 = int(n_k[0]), int(n_k[1])
    if k0 < 3 or k1 < 3 or k0 % 2 == 0 or k1 % 2 == 0:
        return None
    ny, nx = area.shape
    ci, cj = np.arange(0, ny + 1, k0 - 1), np.arange(0, nx + 1, k1 - 1)
    if ci[-1] >= ny or cj[-1] >= nx or (ci.size, cj.size) != tuple(shape_c):
        return None
    h0, h1 = (k0 - 1) // 2, (k1 - 1) // 2
    wt = np.ones((k0, k1))
    for edge in [0, -1]:
        wt[edge, :] /= 2
        wt[:, edge] /= 2
    pad = np.zeros((ny + 2 * h0, nx + 2 * h1))
    pad[h0:h0 + ny, h1:h1 + nx] = area
    out = np.zeros((ci.size, cj.size))
    for a in range(k0):
        for b in range(k1):
            out += wt[a, b] * pad[ci[:, None] + a, cj[None, :] + b]
    return out


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
    if not return_op and sub0s is None and cell_area_f.ndim == 2:
        # a few thousand coarse cells: 25 shifted adds on the host instead of an operator, a sort and a product on the device
        result = _tapered_box_sum_host(np.asarray(cell_area_f, dtype=float), n_k, grid_c.shape[0:2])
        if result is not None:
            return result
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
    if args.get('lagrangian_coords') is not None:
        # a 2-D grid of height offsets in the coordinates that move with the ice (grid_functions.py:89-93)
        grids['lagrangian_dz'] = fd_grid([bds['y'], bds['x']], [args['spacing']['lagrangian_dz'], args['spacing']['lagrangian_dz']],
                                         name='lagrangian_dz', col_0=grids['z0'].N_nodes + grids['dz'].N_nodes, srs_proj4=srs)
    if np.any(grids['dz'].delta[0:2] >= grids['z0'].delta):
        grids['dz'].cell_area = sum_cell_area(grids['z0'], grids['dz'])
    else:
        grids['dz'].cell_area = calc_cell_area(grids['dz'])
        if grids['dz'].mask is not None:
            grids['dz'].cell_area = grids['dz'].cell_area * grids['dz'].mask
    if grids['z0'].mask is not None:
        grids['z0'].cell_area = grids['z0'].cell_area * grids['z0'].mask
    return grids, bds


def sym_range(N, ni, offset=0.5):
    """Indices spaced ``ni`` apart, symmetric about the centre of an ``N``-node axis (grid_functions.py:177-185)."""
    out = np.arange((ni * offset), (N / 2), (ni))
    if offset == 0:
        out = np.r_[-out[-1:0:-1], out] + int(np.floor(N / 2))
    else:
        out = np.r_[-out[-1::-1], out] + int(np.floor(N / 2))
    return np.floor(out[np.abs(out - N / 2) <= N / 2 - ni / 2]).astype(int)


def setup_z0_avg(grids, col_N, args):
    """Operator that averages z0 to ``z0_average_scale`` (grid_functions.py:187-209): tapered box sums centred every
    kernel_N nodes (edge boxes truncated, not dropped), weighted by the z0 cell areas, rows normalised to unit sum."""
    if 'z0_average_scale' not in args or args['z0_average_scale'] is None:
        return {}
    scale = args['z0_average_scale']
    this_name = 'avg_z0_' + str(int(scale)) + 'm'
    kernel_N = np.floor(np.array([scale / dd for dd in grids['z0'].delta[0:2]])).astype(int)
    N_grid = [ctrs.size for ctrs in grids['z0'].ctrs]
    grid_ctr_subs = [np.arange(0, N_grid[dim] + 1, kernel_N[dim]) for dim in [0, 1]]
    sub0s = np.meshgrid(*grid_ctr_subs, indexing='ij')
    op = lin_op(grids['z0'], name=this_name, col_N=col_N).sum_to_grid3(kernel_N + 1, sub0s=sub0s, taper=True,
                                                                       valid_equations_only=False)
    op.apply_mask(grids['z0'].cell_area)
    op.normalize_by_unit_product()
    return {this_name: op}


def setup_averaging_ops(grid, col_N, args, cell_area=None):
    """dz/dt operators per lag and, with ``avg_scales``, block averages of dz and of dz/dt at each scale
    (grid_functions.py:187-313, branch without a 3-D mask).  Every operator is a stencil template applied at the coarse
    centres (``lss_stencil_coo``) scaled by the cell areas and normalised row by row (device gathers / one SpMV)."""
    ops = {}
    if args.get('dzdt_lags') is not None:
        for lag in args['dzdt_lags']:
            name = 'dzdt_lag' + str(lag)
            op = lin_op(grid, name=name, col_N=col_N).dzdt(lag=lag).ravel()
            op.dst_grid.cell_area = grid.cell_area
            op.normalize_by_unit_product()
            ops[name] = op
    if not args.get('avg_scales') or args.get('dzdt_lags') is None:      # grid_functions.py:217-220
        return ops
    if getattr(grid, 'mask_3d', None) is not None:
        raise NotImplementedError('averaging operators with a 3-D mask are outside the hot-path scope')
    N_grid = [ctrs.size for ctrs in grid.ctrs]
    for scale in args['avg_scales']:
        this_name = 'avg_dz_' + str(int(scale)) + 'm'
        kernel_N = np.floor(np.array([scale / dd for dd in grid.delta[0:2]] + [1])).astype(int)
        # the largest scale is centred on the middle of the grid, the others on odd multiples of half their size
        offset = 0 if scale == np.max(args['avg_scales']) else 0.5
        grid_ctr_subs = [sym_range(N_grid[0], kernel_N[0], offset=offset), sym_range(N_grid[1], kernel_N[1], offset=offset),
                         np.arange(grid.shape[2], dtype=int)]
        sub0s = np.meshgrid(*grid_ctr_subs, indexing='ij')
        op = lin_op(grid, name=this_name, col_N=col_N).sum_to_grid3(kernel_N + 1, sub0s=sub0s, taper=True)
        op.apply_mask(mask=cell_area)
        if cell_area is not None:
            op.normalize_by_unit_product()
        else:
            op.v /= (kernel_N[0] * kernel_N[1])
        op.dst_grid.cell_area = sum_cell_area(grid, op.dst_grid, sub0s=sub0s, cell_area_f=cell_area)
        ops[this_name] = op
        for lag in args['dzdt_lags']:
            dz_name = 'avg_dzdt_' + str(int(scale)) + 'm' + '_lag' + str(lag)
            op = lin_op(grid, name=this_name, col_N=col_N)
            op.sum_to_grid3(kernel_N + 1, sub0s=sub0s, lag=lag, taper=True).apply_2d_mask(mask=cell_area)
            if cell_area is not None:
                op.normalize_by_unit_product(wt=2 / (lag * grid.delta[2]))
            else:
                op.v /= (kernel_N[0] * kernel_N[1])
            ops[dz_name] = op
    return ops


def setup_avg_mask_ops(grid, col_N, avg_masks, dzdt_lags):
    """Mean of dz, and of dz/dt at every lag, over each named mask (grid_functions.py:314-325)."""
    if avg_masks is None:
        return {}
    avg_ops = {}
    for name, mask in avg_masks.items():
        this_name = name + '_avg_dz'
        avg_ops[this_name] = lin_op(grid, col_N=col_N, name=this_name).mean_of_mask(mask, dzdt_lag=None)
        for lag in dzdt_lags:
            this_name = name + f'_avg_dzdt_lag{lag}'
            avg_ops[this_name] = lin_op(grid, col_N=col_N, name=this_name).mean_of_mask(mask, dzdt_lag=lag)
    return avg_ops


def _drop_points(data, valid_data, keep):
    """Shared tail of setup_mask / validate_by_dz_mask (grid_functions.py:341-343, 372-374): subset ``data`` in place and
    fold the decision into the caller's valid-point mask."""
    if not np.all(keep):
        data.index(keep)
        valid_data[valid_data] = keep


def validate_by_dz_mask(data, grids, valid_data):
    """Mark the points under masked-out dz nodes invalid (grid_functions.py:346-374): the dz mask (2-D, or ``mask_3d``)
    interpolated to each point must exceed 0.5.  Same in-place behaviour as the reference: ``data`` is subset, ``valid_data``
    updated.  The interpolation runs on the GPU (``lss_interp_coo``)."""
    import torch
    from . import _cabi
    from .smooth_fit import _validate_by_dz_mask_device
    y, x = (_cabi.to_dev(np.asarray(c, dtype=float), torch.float64) for c in data.coords()[0:2])
    t = _cabi.to_dev(np.asarray(data.coords()[2], dtype=float), torch.float64) if len(data.coords()) > 2 else None
    keep = _validate_by_dz_mask_device(grids, y, x, t)
    if keep is not None:
        _drop_points(data, valid_data, _cabi.to_host(keep.to(torch.uint8)).astype(bool))


def setup_mask(data, grids, valid_data, bds, args):
    """Mark the points where the interpolated z0 mask is zero invalid (grid_functions.py:327-343).  The reference reaches
    this function only with ``mask_file`` (smooth_fit.py:437-438), which needs GDAL and is out of scope; with NumPy masks
    (``mask_data``) already on the z0 grid it works the same way."""
    import torch
    from . import _cabi
    if args.get('mask_file') is not None:
        raise NotImplementedError('mask_file needs GDAL (out of scope); use mask_data with NumPy arrays')
    z0 = grids['z0']
    tmp = fd_grid([bds['y'], bds['x']], [args['spacing']['z0'], args['spacing']['z0']], name='z0')
    y, x = (_cabi.to_dev(np.asarray(c, dtype=float), torch.float64) for c in data.coords()[0:2])
    _, cols, vals, _, _ = _cabi.interp_coo(tmp.as_struct(), [y, x])
    mask_d = _cabi.to_dev(np.asarray(z0.mask, dtype=np.float64).ravel(), torch.float64)
    cols, order = torch.sort(cols.long(), dim=1)              # ascending columns: the summation order of the CSR product
    vals = torch.gather(vals, 1, order)
    acc = torch.zeros(cols.shape[0], dtype=torch.float64, device=vals.device)
    for k in range(cols.shape[1]):
        acc = acc + vals[:, k] * mask_d[cols[:, k]]
    acc = torch.where(torch.isfinite(acc), acc, torch.zeros_like(acc))
    _drop_points(data, valid_data, _cabi.to_host((acc != 0).to(torch.uint8)).astype(bool))
