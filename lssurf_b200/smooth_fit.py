"""smooth_fit / iterate_fit: smooth least-squares surface fit with an iterative robust edit loop.

Drop-in for the reference driver (LSsurf/smooth_fit.py:100-210 ``iterate_fit``, :354-717
``smooth_fit``): same keyword arguments, defaults, ValueErrors, early returns and result dict.
The three expensive stages run on the GPU through the C ABI (include/lssurf_b200.h):

  assembly   lin_op.interp_mtx / diff_op / toCSR  ->  lss_interp_coo, lss_stencil_coo, lss_coo_to_csr
  solve      sparseqr.solve on the scaled system  ->  lss_solver_* : matrix-free fp64 PCG on
             (A^T W A) x = A^T W b, A = [G; C] resident in HBM, W = diag(in_TSE / sigma^2, 1 / E_c^2)
  edit loop  residual, calc_sigma_extra (RDE + Brent), 3-sigma mask  ->  lss_spmv, lss_sigma_extra,
             lss_edit_mask, all on device arrays between solves

The row selection ``Ip_r`` and scaling ``TCinv`` of the reference are never materialised (they are
the weight vector), the column selection ``Ip_c`` is a set of unknowns held at zero.  New optional
keywords (defaults keep existing calls unchanged): ``cg_rtol``, ``cg_maxiter``, ``dist`` (a
``parallel.DistContext`` for row-sharded multi-GPU solves), ``profile_spmv``.
"""
from time import time, ctime

import numpy as np
import torch

from . import _cabi
from .bias_functions import assign_bias_ID, setup_bias_fit, parse_biases
from .calc_sigma_extra import sigma_extra_device
from .constraint_functions import setup_smoothness_constraints, build_reference_epoch_matrix
from .grid_functions import setup_grids, setup_averaging_ops
from .lin_op import lin_op, DeviceCSR
from .point_data import data as point_data, grid_data
from .solver import NormalSolver

REQUIRED = ('data', 'W', 'ctr', 'spacing', 'E_RMS')

DEFAULTS = {
    'reference_epoch': 0, 'W_ctr': 1e4, 'return_fit_objects': False, 'mask_file': None, 'mask_data': None,
    'mask_update_function': None, 'mask_scale': {0: 10, 1: 1}, 'compute_E': False, 'max_iterations': 10,
    'min_iterations': 2, 'sigma_extra_relax': False, 'sigma_extra_bin_spacing': None, 'sigma_extra_max': None,
    'sigma_extra_keys': None, 'srs_proj4': None, 'N_subset': None, 'bias_params': None, 'bias_filter': None,
    'repeat_res': None, 'converge_tol_dz': 0.05, 'converge_tol_frac_TSE': 0., 'DEM_tol': None, 'repeat_dt': 1,
    'Edit_only': False, 'dzdt_lags': None, 'prior_args': None, 'prior_edge_args': None, 'avg_scales': [],
    'data_slope_sensors': None, 'E_slope_bias': 0.01, 'E_RMS_d2x_PS_bias': None, 'E_RMS_PS_bias': None,
    'constraint_scaling_maps': None, 'error_res_scale': None, 'avg_masks': None, 'sigma_extra_masks': None,
    'bias_nsigma_edit': None, 'bias_nsigma_iteration': 2, 'bias_edit_vals': None, 'sensor_grid_bias_params': None,
    'ancillary_data': None, 'lagrangian_coords': None, 'z0_average_scale': None, 'erode_source_mask': True,
    'VERBOSE': True, 'DEBUG': False,
    # solver controls (new; defaults chosen for 1e-6 relative-L2 parity with a direct solve)
    'cg_rtol': 1e-12, 'cg_maxiter': 50000, 'dist': None, 'profile_spmv': False, 'history': None,
}

_UNSUPPORTED = ('compute_E', 'mask_file', 'mask_update_function', 'sigma_extra_bin_spacing', 'prior_args',
                'prior_edge_args', 'data_slope_sensors', 'sensor_grid_bias_params', 'lagrangian_coords',
                'constraint_scaling_maps', 'avg_masks', 'z0_average_scale', 'DEM_tol')


class DiagonalOp(object):
    """The inverse-sigma scaling TCinv as a diagonal (what the reference builds with sp.dia_matrix,
    smooth_fit.py:613).  Only ``diagonal()`` / ``shape`` are needed by iterate_fit."""

    def __init__(self, d):
        self._d = np.asarray(d, dtype=float)
        self.shape = (self._d.size, self._d.size)

    def diagonal(self):
        return self._d


def _edit_by_bias(data, m0, in_TSE_d, iteration, bias_model, args):
    """Flag biases that are implausibly large and drop their data (smooth_fit.py:65-98).  Inactive
    unless ``bias_nsigma_edit`` is given; host logic on the few hundred bias values."""
    if args['bias_nsigma_edit'] is None:
        return False
    bpd = bias_model['bias_param_dict']
    if 'edited' not in bpd:
        bpd['edited'] = np.zeros_like(bpd['ID'], dtype=bool)
    b, _ = parse_biases(m0, bias_model, args['bias_params'])
    scaled = np.abs(b['val']) / np.array(b['expected'])
    before = bpd['edited'].copy()
    bad = before.copy()
    if iteration >= args['bias_nsigma_iteration']:
        extreme = np.minimum(50, 3 * np.percentile(scaled, 95))
        if np.any(scaled > extreme):
            bad[scaled == np.max(scaled)] = True
        else:
            bad[scaled > args['bias_nsigma_edit']] = True
    bad_ids = np.array(b['ID'])[bad]
    for ID in bad_ids:
        bpd['edited'][bpd['ID'].index(ID)] = True
    if len(bad_ids) > 0:
        hit = np.isin(data.bias_ID, bad_ids)
        print(f"\t have {len(bad_ids)} bad biases, with {np.sum(hit)} data.")
        in_TSE_d[_cabi.to_dev(hit)] = 0
    elif iteration >= args['bias_nsigma_iteration']:
        print("No bad biases found")
    return not np.all(bpd['edited'] == before)


def _value_set(mask):
    """Distinct values of a 0/1 mask (what np.setxor1d sees when handed boolean arrays)."""
    has_one, all_one = bool(mask.any().item()), bool(mask.all().item())
    return ({1} if has_one else set()) | (set() if all_one else {0})


def iterate_fit(data, Gcoo, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args, grids,
                bias_model=None, sigma_extra_masks=None):
    """The edit / reweight loop (smooth_fit.py:100-210) on device arrays.

    ``Gcoo`` is accepted for signature compatibility and ignored: the operator is taken from the
    ``G_data`` / ``Gc`` lin_ops (device CSR).  ``TCinv`` only needs ``diagonal()`` (1/sigma for the
    data rows then the constraint rows); ``Ip_c`` (col_N x n_free selection) only tells which columns
    are solved for.  Returns ``m0, sigma_extra, in_TSE, rs_data`` as NumPy arrays like the reference.
    With ``args['dist']`` the data rows of this rank are its shard; m0 is identical on every rank.
    """
    args = dict(DEFAULTS, **args) if not all(k in args for k in ('cg_rtol', 'cg_maxiter', 'dist')) else args
    dist = args.get('dist')
    comm = dist.comm if dist is not None else None
    n_data = int(G_data.N_eq)
    n_cols = int(Ip_c.shape[0])
    E_all = 1. / np.asarray(TCinv.diagonal())
    sigma_h, Ec_h = E_all[:n_data], E_all[n_data:]
    dev = _cabi.device()
    f64 = torch.float64
    sigma = _cabi.to_dev(sigma_h, f64)
    z = _cabi.to_dev(np.asarray(rhs[:n_data], dtype=float), f64)
    free_cols = np.unique(Ip_c.tocoo().row) if hasattr(Ip_c, 'tocoo') else np.asarray(Ip_c.free_cols)
    fixed_cols = np.setdiff1d(np.arange(n_cols), free_cols)

    # device operator: this rank's data rows + its share of the constraint rows
    G_csr = G_data._csr_device(col_N=n_cols, row_N=n_data)
    C_all = Gc._csr_device(col_N=n_cols, row_N=int(Gc.N_eq))
    c_lo, c_hi = (0, int(Gc.N_eq)) if dist is None else dist.shard(int(Gc.N_eq))
    C_loc = C_all if (c_lo, c_hi) == (0, int(Gc.N_eq)) else C_all.row_slice(c_lo, c_hi)
    prior = _cabi.to_dev(np.asarray(rhs[n_data + c_lo:n_data + c_hi], dtype=float), f64)
    Ec = _cabi.to_dev(Ec_h[c_lo:c_hi], f64)
    solver = NormalSolver(G_csr, C_loc, z, prior, sigma, Ec, fixed_cols, dist)
    solver.set_profile(bool(args.get('profile_spmv')))
    stats = timing.setdefault('pcg', [])

    in_TSE_d = _cabi.to_dev(np.asarray(in_TSE).astype(np.uint8), torch.uint8)
    if bias_model is None:
        bias_model = {}
    _edit_by_bias(data, np.zeros(n_cols), in_TSE_d, -1, bias_model, args)
    in_TSE_original = in_TSE_d.clone()
    editable = None
    if 'editable' in data.fields:
        editable = _cabi.to_dev(np.asarray(data.editable) != 0)
    groups = None
    if sigma_extra_masks is not None:
        groups = {k: _cabi.to_dev(np.asarray(v).astype(np.uint8), torch.uint8) for k, v in sigma_extra_masks.items()}

    sigma_extra = torch.zeros(n_data, dtype=f64, device=dev)
    r_data = torch.empty(n_data, dtype=f64, device=dev)
    dz_cols = np.asarray(Gc.TOC['cols']['dz']) if 'dz' in Gc.TOC['cols'] else np.zeros(0, dtype=int)
    last_iteration = (args['max_iterations'] <= 1)
    m0 = np.zeros(n_cols)
    se_values = {}
    for iteration in range(args['max_iterations']):
        relax = last_iteration and args['sigma_extra_relax']
        if relax and args['VERBOSE']:
            print('smooth_fit.iterate_fit: relaxing errors by sigma_extra')
        solver.set_rows(in_TSE_d, sigma_extra if relax else None)
        if args['VERBOSE']:
            print("starting qr solve for iteration %d at %s" % (iteration, ctime()), flush=True)
        tic = time()
        m0_last = m0
        x = solver.solve(rtol=args['cg_rtol'], max_iter=args['cg_maxiter'], warm_start=True)
        m0 = x.cpu().numpy().copy()
        timing['sparseqr_solve'] = time() - tic
        stats.append(dict(solver.last, iteration=iteration, wall_s=timing['sparseqr_solve']))
        if args['VERBOSE']:
            print('\tpcg: %d iterations, relres=%.2e, %.1f ms' % (solver.last['iterations'], solver.last['relres'],
                                                                  solver.last['ms']), flush=True)
        # residual over ALL data rows, edited or not (smooth_fit.py:146)
        solver.residual(x, z, out=r_data)
        if last_iteration:
            break
        sigma_extra, se_values = sigma_extra_device(r_data, sigma, in_TSE_d, groups, comm, out=sigma_extra)
        in_TSE_last = in_TSE_d
        in_TSE_d, est = _cabi.edit_mask(r_data, sigma, sigma_extra, 3.0, comm)
        timing.setdefault('near_threshold', []).append(est['near_threshold'])
        if args.get('history') is not None:      # per-iteration tap for parity tests
            args['history'].append(dict(iteration=iteration, m0=m0.copy(), sigma_extra=dict(se_values),
                                        mask_out=np.packbits(in_TSE_d.cpu().numpy()), count=est['count'],
                                        near_threshold=est['near_threshold'], pcg=dict(solver.last)))
        bias_editing_changed = _edit_by_bias(data, m0, in_TSE_d, iteration, bias_model, args)
        recount = bias_editing_changed
        if editable is not None:
            in_TSE_d = torch.where(editable, in_TSE_d, in_TSE_original)
            recount = True
        if recount:
            sel = in_TSE_d.bool()
            est['count'] = int(sel.sum().item())
            est['min_sigma'] = float(sigma[sel].min().item()) if est['count'] else np.inf
        if est['count'] == 0:
            if args['VERBOSE']:
                print("Edited data empty, returning")
            break
        if dz_cols.size and (np.max(np.abs((m0_last - m0)[dz_cols])) < args['converge_tol_dz']) \
                and (iteration > args['min_iterations']):
            if args['VERBOSE']:
                print("Solution identical to previous iteration with tolerance %3.1f, exiting after iteration %d"
                      % (args['converge_tol_dz'], iteration))
            last_iteration = True
        if args['VERBOSE']:
            print('found %d in TSE, dt=%3.0f' % (est['count'], timing['sparseqr_solve']), flush=True)
            if sigma_extra_masks is None:
                print(f"\t median(sigma_extra)={se_values.get('all', 0.):3.4f}")
            else:
                for key in sigma_extra_masks:
                    print(f'\t sigma_extra for {key} : {se_values.get(key, 0.):3.4f}', flush=True)
        if iteration > 0 and iteration > args['bias_nsigma_iteration']:
            # the reference applies setxor1d to the boolean masks as value sets (smooth_fit.py:196)
            vals_last = _value_set(in_TSE_last)
            vals_now = _value_set(in_TSE_d)
            n_editable = int(np.sum(data.editable)) if 'editable' in data.fields else data.size
            if len(vals_last ^ vals_now) / n_editable < args['converge_tol_frac_TSE']:
                if args['VERBOSE']:
                    print("filtering unchanged with tolerance %3.5f, will exit after iteration %d"
                          % (args['converge_tol_frac_TSE'], iteration + 1))
                last_iteration = True
        if iteration >= np.maximum(args['min_iterations'], args['bias_nsigma_iteration'] + 1):
            se_max = max(se_values.values()) if se_values else 0.
            if se_max < 0.5 * est['min_sigma'] and not bias_editing_changed:
                if args['VERBOSE']:
                    print("sigma_extra is small, performing one additional iteration", flush=True)
                last_iteration = True
        if iteration == args['max_iterations'] - 2:
            last_iteration = True
    rs_data = (r_data / sigma).cpu().numpy()
    out = (m0, sigma_extra.cpu().numpy(), in_TSE_d.cpu().numpy().astype(bool), rs_data)
    solver.close()
    return out


def _grid_sums(Gt, cols, shape, vec):
    """(G[:, cols]^T vec) reshaped to a grid -- the products parse_model takes (smooth_fit.py:339-345)."""
    full = Gt.dot(vec)
    return full[torch.as_tensor(cols, device=full.device)].cpu().numpy().reshape(shape)


def parse_model(m, m0, data, R, RMS, G_data, averaging_ops, Gc, Ec, grids, bias_model, args, dist=None):
    """Reshape the solution into grids and compute per-node data counts / misfits and the constraint
    residual summaries (smooth_fit.py:276-352).  The G^T products run on the GPU (lss_spmv on the
    transposed CSR); with row sharding they are summed over ranks on the host."""
    z0g, dzg = grids['z0'], grids['dz']
    m['z0'] = grid_data().from_dict({'x': z0g.ctrs[1], 'y': z0g.ctrs[0], 'cell_area': z0g.cell_area, 'mask': z0g.mask,
                                     'z0': np.reshape(m0[G_data.TOC['cols']['z0']], z0g.shape)})
    m['dz'] = grid_data().from_dict({'x': dzg.ctrs[1], 'y': dzg.ctrs[0], 'time': dzg.ctrs[2],
                                     'dz': np.reshape(m0[G_data.TOC['cols']['dz']], dzg.shape),
                                     'cell_area': dzg.cell_area, 'mask': dzg.mask})
    for key, op in averaging_ops.items():
        fields = {coord: ctr for coord, ctr in zip(op.dst_grid.coords, op.dst_grid.ctrs)}
        fields.update({'cell_area': op.dst_grid.cell_area, key: op.grid_prod(m0)})
        m[key] = grid_data().from_dict(fields)
    if len(bias_model.keys()) > 0:
        m['bias'], m['slope_bias'] = parse_biases(m0, bias_model, args['bias_params'], data=data)
    m['all'] = m0
    m['extent'] = np.concatenate((z0g.bds[1], z0g.bds[0]))
    m['sensor_bias_grids'] = {}
    m['jitter_bias_grids'] = {}
    # constraint residuals: scaled (chi^2) and unscaled (RMS) per family
    n_cols = m0.size
    m0_d = _cabi.to_dev(m0, torch.float64)
    ru = Gc._csr_device(col_N=n_cols, row_N=int(Gc.N_eq)).dot(m0_d).cpu().numpy()
    rc = ru / Ec
    for eq_type in ['d2z_dt2', 'grad2_z0', 'grad2_dzdt', 'grad2_PS']:
        if eq_type in Gc.TOC['rows']:
            R[eq_type] = np.sum(rc[Gc.TOC['rows'][eq_type]] ** 2)
            RMS[eq_type] = np.sqrt(np.mean(ru[Gc.TOC['rows'][eq_type]] ** 2))
    # data counts and misfit maps: G^T applied to the edited-in indicator, r^2 and scaled r^2
    tse = np.asarray(data.three_sigma_edit, dtype=bool)
    r = np.where(tse, data.z - data.z_est, 0.)
    if args['sigma_extra_relax']:
        rs = r / np.sqrt(data.sigma ** 2 + data.sigma_extra ** 2)
    else:
        rs = r / data.sigma
    Gt = G_data._csr_device(col_N=n_cols, row_N=int(data.size)).transpose()
    vecs = {'count': tse.astype(float), 'misfit_scaled_rms': rs ** 2, 'misfit_rms': r ** 2}
    for ff in ['dz', 'z0']:
        cols = np.asarray(G_data.TOC['cols'][ff])
        sums = {}
        for key, vec in vecs.items():
            part = _grid_sums(Gt, cols, grids[ff].shape, _cabi.to_dev(vec, torch.float64))
            if dist is not None and dist.world > 1:
                import torch.distributed as td
                t = torch.from_numpy(part.copy())
                t = t.cuda() if td.get_backend() == 'nccl' else t
                td.all_reduce(t)
                part = t.cpu().numpy()
            sums[key] = part
        count = sums['count']
        count[count == 0] = np.nan
        m[ff].assign({'count': count})
        with np.errstate(invalid='ignore'):
            m[ff].assign({'misfit_scaled_rms': np.sqrt(sums['misfit_scaled_rms'] / count)})
            m[ff].assign({'misfit_rms': np.sqrt(sums['misfit_rms'] / count)})


def _empty_result(m, E, data, grids, valid_data, timing, args):
    return {'m': m, 'E': E, 'data': data, 'grids': grids, 'valid_data': valid_data, 'TOC': {}, 'R': {}, 'RMS': {},
            'timing': timing, 'E_RMS': args['E_RMS']}


def smooth_fit(**kwargs):
    """Fit z0(x, y) + dz(x, y, t) to point data with smoothness constraints and robust editing.

    Same call as the reference (smooth_fit.py:354-717).  ``data`` is any object with the
    pointCollection.data interface (``lssurf_b200.point_data.data`` is provided).  Features outside
    the hot-path scope (error propagation, mask files, priors, sensor / jitter grids,
    averaging scales ...) raise NotImplementedError instead of being silently ignored.
    """
    args = dict(DEFAULTS)
    args.update(kwargs)
    for field in REQUIRED:
        if field not in kwargs:
            raise ValueError("%s must be defined", field)
    for key in _UNSUPPORTED:
        if args[key]:
            raise NotImplementedError(f'smooth_fit keyword {key!r} is outside the GPU hot-path scope of lssurf_b200')
    if args['avg_scales']:
        raise NotImplementedError("smooth_fit keyword 'avg_scales' is next-tier (SURVEY.md section 8(f))")
    dist = args['dist']
    verbose = args['VERBOSE'] and (dist is None or dist.rank == 0)
    args['VERBOSE'] = verbose
    data_in = args['data']
    valid_data = np.isfinite(data_in.z)
    if 'sensor' not in data_in.fields:
        data_in.assign({'sensor': np.zeros(data_in.shape)})
    timing = dict()
    m, E, R, RMS = {}, {}, {}, {}
    tic = time()
    grids, bds = setup_grids(args)

    # ---- host -> device: the point coordinates are uploaded once; validity comes from the
    # interpolation kernel itself (fd_grid.validate_pts semantics, smooth_fit.py:423-425)
    f64 = torch.float64
    coords = data_in.coords()
    if len(coords) < 3:
        raise ValueError('data must have x, y and time fields')
    y_d, x_d, t_d = (_cabi.to_dev(np.asarray(c, dtype=float).ravel(), f64) for c in coords)
    op_z0 = lin_op(grids['z0'], name='interp_z')
    op_dz = lin_op(grids['dz'], name='interp_dz')
    rz = _cabi.interp_coo(grids['z0'].as_struct(), [y_d, x_d])
    rd = _cabi.interp_coo(grids['dz'].as_struct(), [y_d, x_d, t_d])
    if rz[4] or rd[4] or not np.all(valid_data):
        valid_data = valid_data & rz[3].cpu().numpy().astype(bool) & rd[3].cpu().numpy().astype(bool)
        if not np.any(valid_data):
            if verbose:
                print("smooth_fit: no valid data")
            return _empty_result(m, E, None, grids, valid_data, timing, args)
        data = data_in.copy_subset(valid_data)
        keep = _cabi.to_dev(valid_data)
        y_d, x_d, t_d = y_d[keep].contiguous(), x_d[keep].contiguous(), t_d[keep].contiguous()
        rz = rd = None
    else:
        data = data_in.view() if hasattr(data_in, 'view') else data_in.copy_subset(valid_data)
    if data.size == 0:
        print("\tsmooth_fit.py: after masking, no data found")
        return _empty_result(m, E, data, grids, valid_data, timing, args)
    n_all = int(data.size)

    # sigma_extra groups (smooth_fit.py:443-453)
    if args['sigma_extra_keys'] is not None:
        args['sigma_extra_masks'] = {}
        for key, field_vals in args['sigma_extra_keys'].items():
            sel = np.zeros(n_all, dtype=bool)
            for field, vals in field_vals.items():
                sel |= np.isin(getattr(data, field), vals)
            args['sigma_extra_masks'][key] = sel
    elif args['sigma_extra_masks'] is not None:
        args['sigma_extra_masks'] = {k: np.asarray(v)[valid_data == 1] for k, v in args['sigma_extra_masks'].items()}

    # bias IDs are global (unique parameter combinations over ALL points), assigned before sharding
    bias_model = {}
    if args['bias_params'] is not None:
        data, bias_model = assign_bias_ID(data, args['bias_params'], bias_filter=args['bias_filter'])
        if args['bias_nsigma_edit']:
            bias_model['bias_param_dict']['edited'] = np.zeros_like(bias_model['bias_param_dict']['ID'], dtype=bool)

    # ---- this rank's block of points
    lo, hi = (0, n_all) if dist is None else dist.shard(n_all)
    local = data if (lo, hi) == (0, n_all) else data.copy_subset(slice(lo, hi))
    if rz is not None and (lo, hi) == (0, n_all):
        # all points valid: the validity pass already produced the interpolation triplets
        for op, res, g in ((op_z0, rz, grids['z0']), (op_dz, rd, grids['dz'])):
            op._set_device(res[0], res[1], res[2])
            op.N_eq = n_all
            op.ind0 = np.arange(0, n_all, dtype='int')
            op.TOC['rows'] = {op.name: np.arange(n_all, dtype='int')}
            op.TOC['cols'] = {g.name: np.arange(g.col_0, g.col_0 + g.N_nodes)}
            op._update_shape()
    else:
        rz = rd = None
        ys, xs, ts = y_d[lo:hi].contiguous(), x_d[lo:hi].contiguous(), t_d[lo:hi].contiguous()
        op_z0.interp_mtx([ys, xs])
        op_dz.interp_mtx([ys, xs, ts])
    G_data = op_z0
    G_data.add(op_dz)
    del rz, rd

    constraint_op_list = []
    if verbose:
        print(f"smooth_fit: E_RMS={args['E_RMS']}")
    setup_smoothness_constraints(grids, constraint_op_list, args['E_RMS'], args['mask_scale'])
    if args['bias_params'] is not None:
        setup_bias_fit(local, bias_model, G_data, constraint_op_list, bias_param_name='bias_ID',
                       n_bias=int(np.max(data.bias_ID)) + 1)
        if args['bias_edit_vals'] is not None:
            want = np.c_[[args['bias_edit_vals'][k] for k in args['bias_edit_vals']]].T.tolist()
            have = np.c_[[bias_model['bias_param_dict'][k] for k in args['bias_edit_vals']]].T.tolist()
            for row in want:
                bias_model['bias_param_dict']['edited'][have.index(row)] = True
            bad = [bias_model['bias_param_dict']['ID'][i] for i in np.flatnonzero(bias_model['bias_param_dict']['edited'])]
            local.three_sigma_edit[np.isin(local.bias_ID, bad)] = False
    for op in constraint_op_list:
        if op.prior is None:
            op.prior = np.zeros_like(op.expected)

    Gc = lin_op(None, name='constraints').vstack(constraint_op_list)
    Ec = np.zeros(Gc.N_eq)
    for op in constraint_op_list:
        Ec[Gc.TOC['rows'][op.name]] = op.expected
    Ed = np.asarray(local.sigma, dtype=float).ravel()
    if np.any(Ed == 0):
        raise ValueError('zero value found in data sigma')
    if np.any(Ec == 0):
        raise ValueError('zero value found in constraint sigma')
    TCinv = DiagonalOp(1. / np.concatenate((Ed, Ec)))
    rhs = np.concatenate((np.asarray(local.z, dtype=float).ravel(), np.concatenate([op.prior for op in constraint_op_list])))
    Ip_c = build_reference_epoch_matrix(G_data, Gc, grids, args['reference_epoch'])
    timing['setup'] = time() - tic

    if "three_sigma_edit" in local.fields:
        in_TSE = np.asarray(local.three_sigma_edit) > 0.01
    else:
        in_TSE = np.ones(G_data.N_eq, dtype=bool)
    if verbose:
        print("initial: %d:" % (n_all - 1), flush=True)
    if args['return_fit_objects']:
        return {'data': local, 'G_data': G_data, 'Gc': Gc, 'grids': grids, 'Ed': Ed, 'Ec': Ec}

    averaging_ops = {}
    if args['max_iterations'] > 0:
        tic_iteration = time()
        m0, sigma_extra, in_TSE, rs_data = iterate_fit(local, None, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args,
                                                       grids, bias_model=bias_model,
                                                       sigma_extra_masks=None if args['sigma_extra_masks'] is None else
                                                       {k: v[lo:hi] for k, v in args['sigma_extra_masks'].items()})
        timing['iteration'] = time() - tic_iteration
        # model estimate of every data point, G_data . m0 (smooth_fit.py:662); the CSR is already cached
        z_est = G_data._csr_device(col_N=m0.size, row_N=int(local.size)).dot(_cabi.to_dev(m0, f64)).cpu().numpy()
        if dist is not None and dist.world > 1:
            in_TSE_all, sigma_extra_all, z_est_all = (dist.gather_rows(a) for a in (in_TSE, sigma_extra, z_est))
        else:
            in_TSE_all, sigma_extra_all, z_est_all = in_TSE, sigma_extra, z_est
        valid_data[valid_data] = in_TSE_all
        data.assign({'three_sigma_edit': in_TSE_all})
        data.assign({'sigma_extra': sigma_extra_all})
        data.assign({'z_est': np.reshape(z_est_all, data.shape)})
        local.assign({'three_sigma_edit': in_TSE, 'sigma_extra': sigma_extra, 'z_est': z_est})
        averaging_ops = setup_averaging_ops(grids['dz'], grids['dz'].col_N, args, grids['dz'].cell_area)
        parse_model(m, m0, local, R, RMS, G_data, averaging_ops, Gc, Ec, grids, bias_model, args, dist=dist)
        if 'bias' in m and local is not data:
            m['bias'], m['slope_bias'] = parse_biases(m0, bias_model, args['bias_params'], data=data)
        sel = data.three_sigma_edit == 1
        r_sel = data.z_est[sel] - data.z[sel]
        R['data'] = np.sum((r_sel / data.sigma[sel]) ** 2)
        RMS['data'] = np.sqrt(np.mean(r_sel ** 2))
    else:
        averaging_ops = setup_averaging_ops(grids['dz'], grids['dz'].col_N, args, grids['dz'].cell_area)
    return {'m': m, 'E': E, 'data': data, 'grids': grids, 'valid_data': valid_data, 'TOC': Gc.TOC, 'R': R, 'RMS': RMS,
            'timing': timing, 'E_RMS': args['E_RMS'], 'dzdt_lags': args['dzdt_lags']}
