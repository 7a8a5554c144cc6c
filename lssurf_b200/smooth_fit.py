"""smooth_fit / iterate_fit: smooth least-squares surface fit with an iterative robust edit loop.

Drop-in for the reference driver (LSsurf/smooth_fit.py:100-210 ``iterate_fit``, :354-717
``smooth_fit``): same keyword arguments, defaults, ValueErrors, early returns and result dict.
The three expensive stages run on the GPU through the C ABI (include/lssurf_b200.h):

  assembly   lin_op.interp_mtx / diff_op / toCSR  ->  lss_interp_coo, lss_stencil_coo, lss_coo_to_csr
  solve      sparseqr.solve on the scaled system  ->  lss_solver_* : fp64 PCG on (A^T W A) x = A^T W b with the
             normal matrix formed explicitly on the device (one gather product + one fused vector step per iteration,
             two-level preconditioner, bias columns eliminated), W = diag(in_TSE / sigma^2, 1 / E_c^2)
  edit loop  residual, calc_sigma_extra (RDE + Brent), 3-sigma mask  ->  lss_spmv, lss_sigma_extra,
             lss_edit_mask, all on device arrays between solves

The row selection ``Ip_r`` and scaling ``TCinv`` of the reference are never materialised (they are
the weight vector), the column selection ``Ip_c`` is a set of unknowns held at zero.  New optional
keywords (defaults keep existing calls unchanged): ``cg_rtol``, ``cg_maxiter``, ``dist`` (a
``parallel.DistContext``: rows of the normal matrix sharded over the GPUs of one node, exchanges inside the CG step kernel
over NVLink), ``gather_outputs`` ('all': every rank returns the per-point outputs of all points; 'root': only rank 0 copies
them to the host), ``profile_spmv``; ``compute_E`` (error propagation) takes the dense-covariance path of propagate_errors.py.
"""
import os
import warnings
from time import time, ctime

import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from .bias_functions import assign_bias_ID, setup_bias_fit, parse_biases
from .calc_sigma_extra import sigma_extra_device, sigma_extra_on_grid_device
from .constraint_functions import setup_smoothness_constraints, build_reference_epoch_matrix
from .data_slope_bias import data_slope_bias
from .grid_functions import setup_grids, setup_averaging_ops, setup_z0_avg, setup_avg_mask_ops
from .lin_op import lin_op, DeviceCSR, rows_of
from .match_priors import match_prior_dz
from .point_data import data as point_data, grid_data
from .setup_grid_bias import setup_grid_bias
from .setup_sensor_grid_bias import setup_sensor_grid_bias, parse_sensor_bias_grids
from .solver import NormalSolver, grid_structure, coarse_structure

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
    'grid_bias_model_args': None, 'ancillary_data': None, 'lagrangian_coords': None, 'z0_average_scale': None, 'erode_source_mask': True,
    'VERBOSE': True, 'DEBUG': False,
    # solver controls (new; defaults chosen for 1e-6 relative-L2 parity with a direct solve)
    'cg_rtol': 1e-12, 'cg_maxiter': 50000, 'dist': None, 'profile_spmv': False, 'history': None,
    'time_stages': False, 'sort_points': True,
    # several ranks: 'all' = every rank returns the per-point outputs of all points (each rank a full drop-in result),
    # 'root' = only rank 0 copies them to the host; the other ranks return the model, grids and statistics only
    'gather_outputs': 'all',
}

_UNSUPPORTED = ('mask_file', 'prior_edge_args')


class _Stages(object):
    """Optional per-stage wall times with a device synchronise at each boundary (``time_stages=True``)."""

    def __init__(self, on, sink):
        self.on, self.sink, self.t = on, sink, time()
        if on:
            torch.cuda.synchronize()
            self.t = time()

    def mark(self, name):
        if not self.on:
            return
        torch.cuda.synchronize()
        now = time()
        self.sink[name] = self.sink.get(name, 0.) + (now - self.t)
        self.t = now


class DiagonalOp(object):
    """The inverse-sigma scaling TCinv as a diagonal (what the reference builds with sp.dia_matrix,
    smooth_fit.py:613).  Only ``diagonal()`` / ``shape`` are needed by iterate_fit."""

    def __init__(self, d):
        self._d = d if isinstance(d, torch.Tensor) else np.asarray(d, dtype=float)
        n = int(self._d.numel()) if isinstance(d, torch.Tensor) else self._d.size
        self.shape = (n, n)

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


def _value_set(mask, dist=None):
    """Distinct values of a 0/1 mask (what np.setxor1d sees when handed boolean arrays); over all ranks with ``dist``."""
    n_one, n_all = float(mask.sum().item()), float(mask.numel())
    if dist is not None:
        tot = dist.allreduce_sum(torch.tensor([n_one, n_all], dtype=torch.float64, device=mask.device))
        n_one, n_all = float(tot[0].item()), float(tot[1].item())
    return ({1} if n_one > 0 else set()) | (set() if n_one == n_all else {0})


def iterate_fit(data, Gcoo, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args, grids,
                bias_model=None, sigma_extra_masks=None):
    """Reference signature (smooth_fit.py:100-101, 210): returns ``m0, sigma_extra, in_TSE, rs_data`` as NumPy
    arrays.  The work is done by ``iterate_fit_device`` on device arrays."""
    m0, se, tse, r_data, sigma = iterate_fit_device(data, Gcoo, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args, grids,
                                                    bias_model=bias_model, sigma_extra_masks=sigma_extra_masks)
    return m0, _cabi.to_host(se), _cabi.to_host(tse).astype(bool), _cabi.to_host(r_data / sigma)


def iterate_fit_device(data, Gcoo, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args, grids,
                       bias_model=None, sigma_extra_masks=None):
    """The edit / reweight loop (smooth_fit.py:100-210) on device arrays.

    ``Gcoo`` is accepted for signature compatibility and ignored: the operator is taken from the
    ``G_data`` / ``Gc`` lin_ops (device CSR).  ``TCinv`` only needs ``diagonal()`` (1/sigma for the
    data rows then the constraint rows); ``Ip_c`` (col_N x n_free selection) only tells which columns
    are solved for.  Returns ``m0`` (NumPy) and the CUDA tensors ``sigma_extra, in_TSE (uint8), r_data, sigma``.
    With ``args['dist']`` the data rows of this rank are its shard; m0 is identical on every rank.
    """
    args = dict(DEFAULTS, **args) if not all(k in args for k in ('cg_rtol', 'cg_maxiter', 'dist')) else args
    dist = args.get('dist')
    comm = dist.comm if dist is not None else None
    n_data = int(G_data.N_eq)
    n_cols = int(Ip_c.shape[0])
    dev = _cabi.device()
    f64 = torch.float64
    E_all = 1. / _cabi.to_dev(TCinv.diagonal(), f64)          # row sigmas: data rows then constraint rows
    rhs_d = _cabi.to_dev(rhs, f64)
    sigma = E_all[:n_data].contiguous()
    z = rhs_d[:n_data].contiguous()
    # rows of the selection matrix that hold an entry are the free columns (boolean mask instead of unique / setdiff1d)
    is_free = np.zeros(n_cols, dtype=bool)
    if sp.issparse(Ip_c):
        Ipc = Ip_c if Ip_c.format in ('csc', 'coo') else Ip_c.tocoo()
        is_free[Ipc.indices if Ipc.format == 'csc' else Ipc.row] = True
    else:
        is_free[np.asarray(Ip_c.free_cols)] = True
    fixed_cols = np.flatnonzero(~is_free)

    # device operator
    st0 = _Stages(bool(args.get('time_stages')), timing.setdefault('stages', {}))
    st0.mark('it:entry')
    G_csr = G_data._csr_device(col_N=n_cols, row_N=n_data)
    st0.mark('it:G_csr(sort)')
    C_all = Gc._csr_device(col_N=n_cols, row_N=int(Gc.N_eq))
    st0.mark('it:C_csr(sort)')
    shard = args.get('_shard')
    multi = dist is not None and dist.world > 1
    replicate = multi and shard is None and not os.environ.get('LSS_SHARDED_SOLVE')
    st = _Stages(bool(args.get('time_stages')), timing.setdefault('stages', {}))
    st.mark('it:prep')
    structure = grid_structure(grids, fixed_cols, bias_cols=G_data.TOC['cols'].get('data_bias')) if isinstance(grids, dict) else None
    if structure is not None and 'block_ptr' in structure and os.environ.get('LSS_COARSE', '1') != '0':
        # two-level preconditioner: exact solve on a coarse bilinear space on top of block-Jacobi (pays off once the
        # domain is many dz cells across; LSS_COARSE_MIN_BLOCKS moves the threshold, LSS_COARSE=0 switches it off)
        if structure['block_ptr'].numel() - 1 >= int(os.environ.get('LSS_COARSE_MIN_BLOCKS', '2000')):
            cs = coarse_structure(grids, fixed_cols, n_cols)
            if cs is not None:
                structure['coarse'] = cs
                structure['coarse_refresh'] = int(os.environ.get('LSS_COARSE_REFRESH', '0'))
    if shard is not None:
        # row-sharded solve: this rank's own blocks of N (built from its own points + halo copies and all constraint rows),
        # p exchanged with the peers inside the CG step kernel through the P2P mailbox (no collective per iteration)
        if structure is None or 'block_ptr' not in structure or structure['block_ptr'].numel() - 1 != shard['n_blocks']:
            raise _cabi.LssError('row-sharded solve: the preconditioner blocks do not match the dz nodes (use LSS_REPLICATED_SOLVE=1)')
        bp_t, bc_t = structure['block_ptr'], structure['block_cols']
        owned = torch.zeros(n_cols, dtype=torch.uint8, device=dev)
        owned[bc_t[int(bp_t[shard['block_lo']].item()):int(bp_t[shard['block_hi']].item())].long()] = 1
        if 'coarse' in structure:
            structure['coarse'] = coarse_structure(grids, fixed_cols, n_cols, owned=owned)
        n_e = (structure['elim'][1] - structure['elim'][0]) if 'elim' in structure else 0
        n_c = structure['coarse']['n_c'] if structure.get('coarse') else 0
        P = dist.world
        bs = (int(structure['max_block']) + 3) // 4 * 4
        dist.ensure_box(8 * (32 * P + 64 + 4 * P * (n_c + n_e + 8) + 4 * bs * shard['n_blocks']))
        st.mark('it:csr')
        solver = NormalSolver(G_csr, C_all, z, rhs_d[n_data:].contiguous(), sigma, E_all[n_data:].contiguous(), fixed_cols, dist,
                              structure=structure, stages=st,
                              shard=dict(owned_cols=owned, block_lo=shard['block_lo'], block_hi=shard['block_hi'], own_rows=shard['own_rows']))
        o0, o1 = shard['own_rows']
        solver.G_local = G_csr.row_slice(o0, o1) if (o0, o1) != (0, n_data) else G_csr
    elif replicate:
        # replicated solve: every rank gets all assembled rows once and runs the same PCG with no collective inside the
        # iteration; per-point work (residuals, sigma_extra, edit masks) stays on the rank's own rows
        row_counts = dist.gather_counts(n_data)
        G_full = dist.gather_csr(G_csr)
        z_full, sigma_full = dist.gather_var_device(z, row_counts), dist.gather_var_device(sigma, row_counts)
        st.mark('it:csr')
        solver = NormalSolver(G_full, C_all, z_full, rhs_d[n_data:].contiguous(), sigma_full, E_all[n_data:].contiguous(),
                              fixed_cols, None, structure=structure, G_local=G_csr, stages=st)
        del G_full
    else:
        # generic fall-back (rows too long for the normal-matrix forms): this rank's data rows + its share of the constraint rows,
        # one NCCL all-reduce per CG iteration
        c_lo, c_hi = (0, int(Gc.N_eq)) if dist is None else dist.shard(int(Gc.N_eq))
        C_loc = C_all if (c_lo, c_hi) == (0, int(Gc.N_eq)) else C_all.row_slice(c_lo, c_hi)
        prior = rhs_d[n_data + c_lo:n_data + c_hi].contiguous()
        Ec = E_all[n_data + c_lo:n_data + c_hi].contiguous()
        st.mark('it:csr')
        solver = NormalSolver(G_csr, C_loc, z, prior, sigma, Ec, fixed_cols, dist, structure=structure, stages=st)
    solver.set_profile(bool(args.get('profile_spmv')))
    stats = timing.setdefault('pcg', [])
    st.mark('it:solver_structure')

    if isinstance(in_TSE, torch.Tensor):
        in_TSE_d = _cabi.to_dev(in_TSE, torch.uint8).clone()
    else:
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

    if shard is not None:
        # per-point work (residuals, order statistics, masks) runs on the own rows; the solver's weights also need the
        # masks of the halo copies, which come from their owners (all-gather of the own masks after every edit)
        o0, o1 = shard['own_rows']
        in_TSE_ext = in_TSE_d
        in_TSE_d = in_TSE_ext[o0:o1].contiguous()
        in_TSE_original = in_TSE_d.clone()
        sigma_ext, sigma, z = sigma, sigma[o0:o1].contiguous(), z[o0:o1].contiguous()
        n_data = o1 - o0
    sigma_extra = torch.zeros(n_data, dtype=f64, device=dev)
    r_data = torch.empty(n_data, dtype=f64, device=dev)
    dem_d = None
    dz_cols = np.asarray(Gc.TOC['cols']['dz']) if 'dz' in Gc.TOC['cols'] else np.zeros(0, dtype=int)
    last_iteration = (args['max_iterations'] <= 1)
    m0 = np.zeros(n_cols)
    se_values = {}
    for iteration in range(args['max_iterations']):
        relax = last_iteration and args['sigma_extra_relax']
        if relax and args['VERBOSE']:
            print('smooth_fit.iterate_fit: relaxing errors by sigma_extra')
        if shard is not None:
            se_ext = None
            if relax:
                se_ext = torch.full_like(sigma_ext, float(se_values.get('all', 0.)))      # one global value (no groups in this mode)
            solver.set_rows(in_TSE_ext, se_ext)
        elif replicate:
            solver.set_rows(dist.gather_var_device(in_TSE_d, row_counts),
                            dist.gather_var_device(sigma_extra, row_counts) if relax else None)
        else:
            solver.set_rows(in_TSE_d, sigma_extra if relax else None)
        st.mark('it:set_weights')
        if args['VERBOSE']:
            print("starting qr solve for iteration %d at %s" % (iteration, ctime()), flush=True)
        tic = time()
        m0_last = m0
        x = solver.solve(rtol=args['cg_rtol'], max_iter=args['cg_maxiter'], warm_start=True)
        if replicate:
            dist.broadcast(x, src=0)          # every rank continues from rank 0's bits
        m0 = x.cpu().numpy().copy()
        timing['sparseqr_solve'] = time() - tic
        last = solver.last
        last['converged'] = bool(np.isfinite(last['relres']) and last['relres'] <= args['cg_rtol'] and last['iterations'] < args['cg_maxiter'])
        if not np.isfinite(last['relres']):
            raise _cabi.LssError('smooth_fit: the PCG solve of outer iteration %d broke down (residual is not finite)' % iteration)
        if not last['converged']:
            # the reference's direct QR solve cannot stall; say so instead of feeding an unconverged model into the edit loop silently
            warnings.warn('smooth_fit: PCG stopped at %d iterations with relative residual %.2e > cg_rtol %.1e (outer iteration %d)'
                          % (last['iterations'], last['relres'], args['cg_rtol'], iteration), RuntimeWarning)
        stats.append(dict(last, iteration=iteration, wall_s=timing['sparseqr_solve']))
        if args['VERBOSE']:
            print('\tpcg: %d iterations, relres=%.2e, %.1f ms' % (solver.last['iterations'], solver.last['relres'],
                                                                  solver.last['ms']), flush=True)
        # residual over ALL data rows, edited or not (smooth_fit.py:146)
        st.mark('it:pcg')
        solver.residual(x, z, out=r_data)
        st.mark('it:residual')
        if last_iteration:
            break
        if args.get('sigma_extra_bin_spacing') is None:
            sigma_extra, se_values = sigma_extra_device(r_data, sigma, in_TSE_d, groups, comm, out=sigma_extra)
        else:
            # spatially varying sigma_extra on overlapping bins (smooth_fit.py:154-158)
            sigma_extra, se_values = sigma_extra_on_grid_device(
                _field_dev(data, 'x'), _field_dev(data, 'y'), r_data, sigma, in_TSE_d, groups,
                spacing=args['sigma_extra_bin_spacing'], sigma_extra_max=args.get('sigma_extra_max'), comm=comm,
                dist=dist if multi else None)
        st.mark('it:sigma_extra')
        in_TSE_last = in_TSE_d
        in_TSE_d, est = _cabi.edit_mask(r_data, sigma, sigma_extra, 3.0, comm)
        if shard is not None:
            full_mask = dist.gather_var_device(in_TSE_d, shard['counts'])
            in_TSE_ext = full_mask[shard['lo_e']:shard['hi']].contiguous()
        st.mark('it:edit_mask')
        timing.setdefault('near_threshold', []).append(est['near_threshold'])
        if args.get('history') is not None:      # per-iteration tap for parity tests
            args['history'].append(dict(iteration=iteration, m0=m0.copy(), sigma_extra=dict(se_values),
                                        mask_out=np.packbits((full_mask if shard is not None else in_TSE_d).cpu().numpy()), count=est['count'],
                                        near_threshold=est['near_threshold'], pcg=dict(solver.last)))
        bias_editing_changed = _edit_by_bias(data, m0, in_TSE_d, iteration, bias_model, args)
        recount = args['bias_nsigma_edit'] is not None      # flagged biases drop their data from the mask on every pass
        if editable is not None:
            in_TSE_d = torch.where(editable, in_TSE_d, in_TSE_original)
            recount = True
        if args['DEM_tol'] is not None:
            # check_data_against_DEM (smooth_fit.py:31-37): selected points must stay within DEM_tol of the DEM once the
            # fitted height change (everything but z0) is taken out
            if dem_d is None:
                dem_d = _field_dev(data, 'DEM')
            m1 = m0.copy()
            m1[np.asarray(G_data.TOC['cols']['z0'])] = 0
            r_dem = solver.residual(_cabi.to_dev(m1, f64), z) - dem_d
            in_TSE_d = (in_TSE_d.bool() & (r_dem.abs() < args['DEM_tol'])).to(torch.uint8)
            recount = True
        if recount:
            # counted over ALL ranks: the loop control below must take the same decisions everywhere
            sel = in_TSE_d.bool()
            cnt = sel.sum().to(f64).reshape(1)
            mn = (sigma[sel].min() if bool(sel.any().item()) else torch.tensor(np.inf, dtype=f64, device=dev)).reshape(1)
            if multi:
                cnt, mn = dist.allreduce_sum(cnt), dist.allreduce_min(mn)
            est['count'] = int(cnt.item())
            est['min_sigma'] = float(mn.item()) if est['count'] else np.inf
            if args.get('history') is not None:  # the mask that goes into the next solve (after bias editing / `editable`)
                args['history'][-1]['mask_next'] = np.packbits(in_TSE_d.cpu().numpy())
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
            vals_last = _value_set(in_TSE_last, dist if multi else None)
            vals_now = _value_set(in_TSE_d, dist if multi else None)
            n_editable = int(np.sum(data.editable)) if 'editable' in data.fields else data.size
            if multi:
                n_editable = int(dist.allreduce_sum(torch.tensor([float(n_editable)], dtype=f64, device=dev)).item())
            if len(vals_last ^ vals_now) / n_editable < args['converge_tol_frac_TSE']:
                if args['VERBOSE']:
                    print("filtering unchanged with tolerance %3.5f, will exit after iteration %d"
                          % (args['converge_tol_frac_TSE'], iteration + 1))
                last_iteration = True
        if iteration >= np.maximum(args['min_iterations'], args['bias_nsigma_iteration'] + 1):
            # the reference tests every point: np.all(sigma_extra < 0.5 * min(sigma[in_TSE])) (smooth_fit.py:203)
            se_top = sigma_extra.max().reshape(1) if sigma_extra.numel() else torch.zeros(1, dtype=f64, device=dev)
            se_max = -float(dist.allreduce_min(-se_top).item()) if multi else float(se_top.item())
            if se_max < 0.5 * est['min_sigma'] and not bias_editing_changed:
                if args['VERBOSE']:
                    print("sigma_extra is small, performing one additional iteration", flush=True)
                last_iteration = True
        if iteration == args['max_iterations'] - 2:
            last_iteration = True
    st.mark('it:other')
    solver.close()
    st.mark('it:close')
    return m0, sigma_extra, in_TSE_d, r_data, sigma


def _bias_statistics(b_dict, bias_id, r, tse, n_bias, dist):
    """Per-bias residual statistics of parse_biases (bias_functions.py:140-155: nanstd of the residuals of
    each bias group, raw and edited) as grouped two-pass reductions on the device instead of one masked pass
    over all points per bias."""
    ids = bias_id.to(torch.int64)
    fin = torch.isfinite(r)
    rr = torch.where(fin, r, torch.zeros_like(r))

    def reduce(w):
        t = torch.bincount(ids, weights=w, minlength=n_bias)[:n_bias]
        return t if dist is None else dist.allreduce_sum(t)

    for key_rms, key_n, wt in (('rms_data_raw', 'N_raw', torch.ones_like(r)), ('rms_data_edited', 'N_edited', tse.to(r.dtype))):
        n_all = reduce(wt)                               # len(r[jj]) counts NaNs too
        n_fin = reduce(wt * fin)
        mean = reduce(wt * rr) / n_fin
        dev = torch.where(fin, r - mean[ids], torch.zeros_like(r))
        std = torch.sqrt(reduce(wt * dev * dev) / n_fin)
        b_dict[key_n] = [int(v) for v in n_all.cpu().numpy()]
        b_dict[key_rms] = [float(v) if c > 0 else np.nan for v, c in zip(std.cpu().numpy(), n_all.cpu().numpy())]


def parse_model(m, m0, dev_fields, R, RMS, G_data, averaging_ops, Gc, Ec, grids, bias_model, args, dist=None, stages=None):
    """Reshape the solution into grids and compute per-node data counts / misfits, bias statistics and the
    constraint residual summaries (smooth_fit.py:276-352).  ``dev_fields`` holds this rank's per-point
    device arrays (z, sigma, z_est, three_sigma_edit, sigma_extra, bias_ID).  The G^T products run on the GPU
    (lss_spmv on the transposed CSR); with row sharding the per-node sums are summed over ranks."""
    z0g, dzg = grids['z0'], grids['dz']
    m['z0'] = grid_data().from_dict({'x': z0g.ctrs[1], 'y': z0g.ctrs[0], 'cell_area': z0g.cell_area, 'mask': z0g.mask,
                                     'z0': np.reshape(m0[G_data.TOC['cols']['z0']], z0g.shape)})
    m['dz'] = grid_data().from_dict({'x': dzg.ctrs[1], 'y': dzg.ctrs[0], 'time': dzg.ctrs[2],
                                     'dz': np.reshape(m0[G_data.TOC['cols']['dz']], dzg.shape),
                                     'cell_area': dzg.cell_area, 'mask': dzg.mask})
    if 'lagrangian_dz' in grids:
        lg = grids['lagrangian_dz']
        m['lagrangian_dz'] = grid_data().from_dict({'x': lg.ctrs[1], 'y': lg.ctrs[0],
                                                    'dz': np.reshape(m0[G_data.TOC['cols']['lagrangian_dz']], lg.shape),
                                                    'cell_area': lg.cell_area, 'mask': lg.mask})
    for key, op in averaging_ops.items():
        fields = {coord: ctr for coord, ctr in zip(op.dst_grid.coords, op.dst_grid.ctrs)}
        fields.update({'cell_area': op.dst_grid.cell_area, key: op.grid_prod(m0)})
        m[key] = grid_data().from_dict(fields)
    mark = stages.mark if stages is not None else (lambda name: None)
    mark('pm:grids')
    f = dev_fields
    tse = f['three_sigma_edit'].bool()
    r = f['z'] - f['z_est']
    if len(bias_model.keys()) > 0:
        m['bias'], m['slope_bias'] = parse_biases(m0, bias_model, args['bias_params'])
        _bias_statistics(m['bias'], f['bias_ID'], r, tse, len(bias_model['bias_ID_dict']), dist)
    mark('pm:bias_stats')
    m['all'] = m0
    m['extent'] = np.concatenate((z0g.bds[1], z0g.bds[0]))
    m['sensor_bias_grids'] = parse_sensor_bias_grids(m0, G_data, grids)
    m['jitter_bias_grids'] = {}
    if args.get('grid_bias_model_args') is not None:            # deprecated/smooth_xytb_fit.py:303-310
        m['grid_biases'] = {}
        for bm in args['grid_bias_model_args']:
            g = bm['grid']
            m['grid_biases'][g.name] = grid_data().from_dict({'x': g.ctrs[1], 'y': g.ctrs[0],
                                                              g.name: np.reshape(m0[g.col_0:g.col_N], g.shape)})
    # constraint residuals: scaled (chi^2) and unscaled (RMS) per family
    n_cols = m0.size
    m0_d = _cabi.to_dev(m0, torch.float64)
    ru = Gc._csr_device(col_N=n_cols, row_N=int(Gc.N_eq)).dot(m0_d)
    Ec_d = f['Ec'] if f.get('Ec') is not None else _cabi.to_dev(Ec, torch.float64)
    rc = ru / Ec_d
    kinds = [k for k in ['d2z_dt2', 'grad2_z0', 'grad2_dzdt', 'grad2_PS'] if k in Gc.TOC['rows']]
    if kinds:
        # sums stay on the device (1.5e6 constraint rows at cfg 4): one small copy for all families
        parts = []
        for eq_type in kinds:
            sel = rows_of(Gc.TOC['rows'][eq_type])
            if not isinstance(sel, slice):
                sel = _cabi.to_dev(np.asarray(sel), torch.int64)
            parts += [(rc[sel] ** 2).sum(), (ru[sel] ** 2).mean()]
        parts = torch.stack(parts).cpu().numpy()
        for k, eq_type in enumerate(kinds):
            R[eq_type] = parts[2 * k]
            RMS[eq_type] = np.sqrt(parts[2 * k + 1])
    mark('pm:constraint_residuals')
    # data counts and misfit maps: G^T applied to the edited-in indicator, r^2 and scaled r^2
    r_in = torch.where(tse, r, torch.zeros_like(r))
    if args['sigma_extra_relax']:
        rs = r_in / torch.sqrt(f['sigma'] ** 2 + f['sigma_extra'] ** 2)
    else:
        rs = r_in / f['sigma']
    Gd = f.get('G_rows') or G_data._csr_device(col_N=n_cols, row_N=int(r.numel()))
    sums = {}
    for key, vec in (('count', tse.to(torch.float64)), ('misfit_scaled_rms', rs * rs), ('misfit_rms', r_in * r_in)):
        full = Gd.rdot(vec.contiguous())                 # G^T vec by scatter: three products do not pay for a transpose
        sums[key] = (full if dist is None else dist.allreduce_sum(full)).cpu().numpy()
    for ff in ['dz', 'z0']:
        cols = np.asarray(G_data.TOC['cols'][ff])
        count = sums['count'][cols].reshape(grids[ff].shape).copy()
        count[count == 0] = np.nan
        m[ff].assign({'count': count})
        with np.errstate(invalid='ignore'):
            for key in ('misfit_scaled_rms', 'misfit_rms'):
                m[ff].assign({key: np.sqrt(sums[key][cols].reshape(grids[ff].shape) / count)})
    mark('pm:count_misfit')
    # chi^2 and RMS of the edited-in data residuals (smooth_fit.py:676-678)
    tot = torch.stack([(rs * rs).sum(), (r_in * r_in).sum(), tse.sum().to(torch.float64)])
    tot = (tot if dist is None else dist.allreduce_sum(tot)).cpu().numpy()
    if args['sigma_extra_relax']:        # chi^2 against the unrelaxed errors (smooth_fit.py:676)
        raw = ((r_in / f['sigma']) ** 2).sum().reshape(1)
        R['data'] = float((raw if dist is None else dist.allreduce_sum(raw)).item())
    else:
        R['data'] = tot[0]
    RMS['data'] = np.sqrt(tot[1] / tot[2]) if tot[2] > 0 else np.nan


def _empty_result(m, E, data, grids, valid_data, timing, args):
    return {'m': m, 'E': E, 'data': data, 'grids': grids, 'valid_data': valid_data, 'TOC': {}, 'R': {}, 'RMS': {},
            'timing': timing, 'E_RMS': args['E_RMS']}


def _field_dev(obj, name, dist=None):
    """A per-point field as a contiguous fp64 CUDA tensor (NumPy arrays are uploaded, CUDA tensors used as is).
    With several ranks (``dist``) every rank holds the same host array: each uploads its 1 / P of the rows and the blocks are
    all-gathered over NVLink, so a byte crosses PCIe once per node instead of once per rank."""
    v = getattr(obj, name)
    if isinstance(v, torch.Tensor):
        return _cabi.to_dev(v.reshape(-1), torch.float64)
    v = np.asarray(v, dtype=float).ravel()
    if dist is not None and dist.world > 1 and v.size >= int(os.environ.get('LSS_SHARED_UPLOAD_MIN', _SHARED_UPLOAD_MIN)):
        lo, hi = dist.shard(v.size)
        return dist.gather_rows_device(_cabi.to_dev(v[lo:hi], torch.float64), v.size)
    return _cabi.to_dev(v, torch.float64)


_SHARED_UPLOAD_MIN = 1 << 16      # rows; below this one copy per rank is cheaper than a collective


def _assign_bias_ID_device(fields_d, bias_params, sigma_corr_d):
    """assign_bias_ID (bias_functions.py:25-82) on device arrays.

    Dense ranks of each parameter column (``lss_dense_rank_f64``: radix sort + head flags + scan, NaN -> -9999),
    combined in mixed radix so that ascending combined rank == lexicographic order of the unique parameter rows,
    ranked again to get consecutive IDs; E_bias = per-group nanmedian(sigma_corr) (``lss_group_median``).
    Returns (bias_ID int32 CUDA tensor, bias_model dict as the reference builds it).
    """
    ranks, uniq = [], []
    for p in bias_params:
        r, u = _cabi.dense_rank(fields_d[p], nan_value=-9999.0)
        ranks.append(r)
        uniq.append(u)
    comb = ranks[0].to(torch.int64)
    for k in range(1, len(ranks)):
        comb = comb * int(uniq[k].size) + ranks[k]
    ids, ucomb = _cabi.dense_rank(comb.to(torch.float64), nan_value=-9999.0)
    n_bias = int(ucomb.size)
    e_bias = _cabi.group_median(ids, sigma_corr_d, n_bias).cpu().numpy()
    model = {'E_bias': dict(), 'bias_param_dict': {p: list() for p in bias_params}, 'bias_ID_dict': dict()}
    model['bias_param_dict']['ID'] = list()
    for k, code in enumerate(ucomb.astype(np.int64)):
        vals = []
        for j in range(len(bias_params) - 1, -1, -1):
            vals.append(uniq[j][code % uniq[j].size])
            code //= uniq[j].size
        vals = vals[::-1]
        shown = [np.nan if v == -9999 else v for v in vals]
        for j, p in enumerate(bias_params):
            model['bias_param_dict'][p].append(vals[j])
        model['bias_param_dict']['ID'].append(k)
        model['bias_ID_dict'][k] = {p: shown[j] for j, p in enumerate(bias_params)}
        model['E_bias'][k] = e_bias[k]
    return ids, model


def _validate_by_dz_mask_device(grids, y_d, x_d, t_d):
    """validate_by_dz_mask (grid_functions.py:346-374) on device arrays: the dz mask (2-D, or mask_3d) interpolated to the
    points must exceed 0.5.  Returns a bool tensor (True: keep) or None when the mask is all True (every point passes: the
    interpolation weights of a point sum to 1).  The interpolated value is summed in ascending column order like the CSR
    product of the reference."""
    dz = grids['dz']
    m3 = getattr(dz, 'mask_3d', None)
    if m3 is not None:
        mask, pts, bds, delta = np.asarray(getattr(m3, 'z', m3)), [y_d, x_d, t_d], dz.bds, dz.delta
    else:
        mask = np.asarray(dz.mask)
        if mask.ndim == 2:
            pts, bds, delta = [y_d, x_d], dz.bds[0:2], dz.delta[0:2]
        else:
            pts, bds, delta = [y_d, x_d, t_d], dz.bds, dz.delta
    if mask.all():
        return None
    from .fd_grid import fd_grid
    tmp = fd_grid(bds, delta)
    _, cols, vals, _, _ = _cabi.interp_coo(tmp.as_struct(), pts)
    mask_d = _cabi.to_dev(mask.ravel().astype(np.float64), torch.float64)
    cols, order = torch.sort(cols.long(), dim=1)
    vals = torch.gather(vals, 1, order)
    acc = torch.zeros(cols.shape[0], dtype=torch.float64, device=vals.device)
    for k in range(cols.shape[1]):
        acc = acc + vals[:, k] * mask_d[cols[:, k]]
    acc = torch.where(torch.isfinite(acc), acc, torch.zeros_like(acc))
    return acc > 0.5


def _shard_plan(grids, cell_ids, dist):
    """Ownership of the row-sharded solve: rank k owns the preconditioner blocks (dz nodes, row-major)
    [k n_b / P, (k + 1) n_b / P), the points whose dz cell has its lower-left node among them, and additionally holds
    halo copies of the points of the cell row below (every point that touches one of its columns).  ``cell_ids``: the
    dz-cell id (= id of the cell's lower-left node) of every point in the internal (sorted) order."""
    ny, nx, _ = (int(v) for v in grids['dz'].shape)
    n_b = ny * nx
    P, k = dist.world, dist.rank
    bounds = [(r * n_b) // P for r in range(P + 1)]
    edges = torch.tensor(bounds + [max(b - nx - 1, 0) for b in bounds[:-1]], dtype=cell_ids.dtype, device=cell_ids.device)
    pos = torch.searchsorted(cell_ids, edges).cpu().numpy()
    starts, halo = pos[:P + 1], pos[P + 1:]
    return dict(block_lo=bounds[k], block_hi=bounds[k + 1], lo=int(starts[k]), hi=int(starts[k + 1]), lo_e=int(halo[k]),
                counts=[int(starts[r + 1] - starts[r]) for r in range(P)], n_blocks=n_b)


def smooth_fit(**kwargs):
    """Fit z0(x, y) + dz(x, y, t) to point data with smoothness constraints and robust editing.

    Same call as the reference (smooth_fit.py:354-717).  ``data`` is any object with the
    pointCollection.data interface (``lssurf_b200.point_data.data`` is provided); its per-point fields may
    be NumPy arrays (uploaded once) or CUDA tensors (used in place).  Features outside the hot-path scope
    (error propagation, mask files, priors, sensor / jitter grids, averaging
    scales ...) raise NotImplementedError instead of being silently ignored.
    """
    args = dict(DEFAULTS)
    args.update(kwargs)
    for field in REQUIRED:
        if field not in kwargs:
            raise ValueError("%s must be defined", field)
    for key in _UNSUPPORTED:
        if args[key]:
            raise NotImplementedError(f'smooth_fit keyword {key!r} is outside the GPU hot-path scope of lssurf_b200')
    if args['gather_outputs'] not in ('all', 'root'):
        raise ValueError("gather_outputs must be 'all' or 'root'")
    dist = args['dist']
    verbose = args['VERBOSE'] and (dist is None or dist.rank == 0)
    args['VERBOSE'] = verbose
    data_in = args['data']
    if 'sensor' not in data_in.fields:
        data_in.assign({'sensor': np.zeros(data_in.shape)})
    timing = dict()
    m, E, R, RMS = {}, {}, {}, {}
    tic = time()
    stages = _Stages(args['time_stages'], timing.setdefault('stages', {}))
    grids, bds = setup_grids(args)
    stages.mark('setup_grids')

    # ---- host -> device: every per-point field the fit needs is uploaded once; validity comes from the
    # interpolation kernel itself (fd_grid.validate_pts semantics, smooth_fit.py:406, 423-425)
    f64 = torch.float64
    if not all(f in data_in.fields for f in ('x', 'y', 'time')):
        raise ValueError('data must have x, y and time fields')
    y_d, x_d, t_d, z_d, s_d = (_field_dev(data_in, f, dist) for f in ('y', 'x', 'time', 'z', 'sigma'))
    extra_d = {}
    if args['bias_params'] is not None:
        extra_d = {p: _field_dev(data_in, p, dist) for p in list(args['bias_params']) + ['sigma_corr']}
    tse_d = _field_dev(data_in, 'three_sigma_edit', dist) if 'three_sigma_edit' in data_in.fields else None
    cov_d = {}
    if args['grid_bias_model_args'] is not None:       # the data parameters the gridded biases multiply
        cov_d = {bm['param']: _field_dev(data_in, bm['param'], dist) for bm in args['grid_bias_model_args']}
    stages.mark('upload')
    op_z0 = lin_op(grids['z0'], name='interp_z')
    op_dz = lin_op(grids['dz'], name='interp_dz')
    rz = _cabi.interp_coo(grids['z0'].as_struct(), [y_d, x_d])
    rd = _cabi.interp_coo(grids['dz'].as_struct(), [y_d, x_d, t_d])
    finite = torch.isfinite(z_d)
    in_mask = _validate_by_dz_mask_device(grids, y_d, x_d, t_d)        # points under a masked-out part of the dz grid are dropped
    if rz[4] or rd[4] or in_mask is not None or not bool(finite.all().item()):
        keep = finite & rz[3].bool() & rd[3].bool()
        if in_mask is not None:
            keep = keep & in_mask
        valid_data = keep.cpu().numpy()
        if not np.any(valid_data):
            if verbose:
                print("smooth_fit: no valid data")
            return _empty_result(m, E, None, grids, valid_data, timing, args)
        data = data_in.copy_subset(valid_data)
        y_d, x_d, t_d, z_d, s_d = (a[keep].contiguous() for a in (y_d, x_d, t_d, z_d, s_d))
        extra_d = {k: v[keep].contiguous() for k, v in extra_d.items()}
        cov_d = {k: v[keep].contiguous() for k, v in cov_d.items()}
        tse_d = None if tse_d is None else tse_d[keep].contiguous()
        rz = rd = None
    else:
        valid_data = None                      # every point is valid: the mask is written once, from the final edit mask
        data = data_in.view() if hasattr(data_in, 'view') else data_in.copy_subset(valid_data)
    n_all = int(z_d.numel())
    if n_all == 0:
        print("\tsmooth_fit.py: after masking, no data found")
        return _empty_result(m, E, data, grids, np.ones(0, dtype=bool) if valid_data is None else valid_data, timing, args)

    stages.mark('interp+valid')
    # sigma_extra groups (smooth_fit.py:443-453)
    if args['sigma_extra_keys'] is not None:
        args['sigma_extra_masks'] = {}
        for key, field_vals in args['sigma_extra_keys'].items():
            sel = np.zeros(n_all, dtype=bool)
            for field, vals in field_vals.items():
                sel |= np.isin(getattr(data, field), vals)
            args['sigma_extra_masks'][key] = sel
    elif args['sigma_extra_masks'] is not None:
        args['sigma_extra_masks'] = {k: (np.asarray(v) if valid_data is None else np.asarray(v)[valid_data == 1])
                                     for k, v in args['sigma_extra_masks'].items()}

    # bias IDs are global (unique parameter combinations over ALL points), assigned before sharding
    bias_model = {}
    bias_id_d = None
    if args['bias_params'] is not None:
        if args['bias_filter'] is not None:
            # the filter returns a container whose parameter fields define the bias groups (bias_functions.py:53-57)
            data_filt = args['bias_filter'](data)
            extra_d.update({p: _field_dev(data_filt, p) for p in args['bias_params']})
        bias_id_d, bias_model = _assign_bias_ID_device(extra_d, args['bias_params'], extra_d['sigma_corr'])
        bias_id_caller = bias_id_d                  # in the caller's order; copied to the host once the row order is known
        if args['bias_nsigma_edit']:
            bias_model['bias_param_dict']['edited'] = np.zeros_like(bias_model['bias_param_dict']['ID'], dtype=bool)

    stages.mark('bias_ids')
    # ---- internal spatial ordering: rows of G are built for points sorted by (dz cell, z0 cell, epoch) so that the
    # gathers of the CSR / ELL products hit neighbouring grid nodes; every per-point output is returned in the
    # caller's order.  Skipped where the caller sees operator rows (return_fit_objects) or per-point host logic
    # addresses rows (bias editing, `editable`).
    order = None
    shard = None
    if (args['sort_points'] and os.environ.get('LSS_SORT_POINTS', '1') != '0' and not args['return_fit_objects'] and args['bias_nsigma_edit'] is None
            and args['bias_edit_vals'] is None and 'editable' not in data.fields and args['sigma_extra_masks'] is None and args['DEM_tol'] is None
            and args['lagrangian_coords'] is None
            and args['data_slope_sensors'] is None and args['sensor_grid_bias_params'] is None
            and args['sigma_extra_bin_spacing'] is None and n_all > 1):
        order, keys = _cabi.spatial_order(grids['z0'].as_struct(), grids['dz'].as_struct(), [y_d, x_d, t_d], return_keys=True)
        if (dist is not None and dist.comm is not None and not os.environ.get('LSS_REPLICATED_SOLVE') and args['grid_bias_model_args'] is None
                and (dist.world > 1 or os.environ.get('LSS_SHARD_SINGLE')) and len(grids['dz'].shape) == 3 and len(grids['z0'].shape) == 2):
            # row-sharded solve: ownership follows the dz cells, which is the leading part of the sort key
            div = int(grids['z0'].shape[0]) * int(grids['z0'].shape[1]) * int(grids['dz'].shape[2])
            shard = _shard_plan(grids, torch.div(keys[order], div, rounding_mode='floor'), dist)
        del keys
        y_d, x_d, t_d, z_d, s_d = (a[order].contiguous() for a in (y_d, x_d, t_d, z_d, s_d))
        tse_d = None if tse_d is None else tse_d[order].contiguous()
        bias_id_d = None if bias_id_d is None else bias_id_d[order].contiguous()
        cov_d = {k: v[order].contiguous() for k, v in cov_d.items()}
        rz = rd = None
        stages.mark('spatial_sort')
    # does this rank hand per-point outputs to its caller?  (with gather_outputs='root' only rank 0 pays for the copies)
    host_out = dist is None or dist.world == 1 or args['gather_outputs'] == 'all' or dist.rank == 0
    if bias_id_d is not None and (host_out or order is None):        # without the internal order the edit loop reads data.bias_ID
        data.assign({'bias_ID': _cabi.to_host(bias_id_caller.to(f64))})      # float on the device: no 10^7-element host cast
    bias_id_caller = None
    # ---- this rank's block of points
    if shard is not None:
        lo, hi = shard['lo_e'], shard['hi']          # rows assembled here: halo copies [lo_e, lo) + own points [lo, hi)
        shard['own_rows'] = (shard['lo'] - lo, hi - lo)
        args['_shard'] = shard
    else:
        lo, hi = (0, n_all) if dist is None else dist.shard(n_all)
        args.pop('_shard', None)
    whole = (lo, hi) == (0, n_all)
    if order is not None:
        # rows are in the internal order: the loop only needs the row count from the point container
        local = point_data().from_dict({'row': np.empty(hi - lo, dtype=np.uint8)})
    else:
        local = data if whole else data.copy_subset(slice(lo, hi))
    if rz is not None and whole:
        # all points valid: the validity pass already produced the interpolation triplets
        for op, res, g in ((op_z0, rz, grids['z0']), (op_dz, rd, grids['dz'])):
            op._set_device(res[0], res[1], res[2], blocks=[(res[1], res[2])] if res[1].dim() == 2 else None)
            op.N_eq = n_all
            op._ind0 = [('arange', n_all)]
            op.TOC['rows'] = {op.name: range(n_all)}
            op.TOC['cols'] = {g.name: np.arange(g.col_0, g.col_0 + g.N_nodes)}
            op._update_shape()
    else:
        rz = rd = None
        op_z0.interp_mtx([y_d[lo:hi].contiguous(), x_d[lo:hi].contiguous()])
        op_dz.interp_mtx([y_d[lo:hi].contiguous(), x_d[lo:hi].contiguous(), t_d[lo:hi].contiguous()])
    G_data = op_z0
    G_data.add(op_dz)
    del rz, rd
    if args['lagrangian_coords'] is not None:
        # height offsets on a grid in the coordinates that move with the ice; points without such coordinates (NaN / outside)
        # get no entry (smooth_fit.py:463-465, interp_mtx with bounds_error=False).  Rows stay in the caller's order.
        lag_pts = [_field_dev(data, f)[lo:hi].contiguous() for f in args['lagrangian_coords']]
        G_data.add(lin_op(grids['lagrangian_dz'], name='interp_lagrangian_dz').interp_mtx(lag_pts, bounds_error=False))
    z_loc, s_loc = z_d[lo:hi].contiguous(), s_d[lo:hi].contiguous()

    stages.mark('G_data')
    constraint_op_list = []
    if verbose:
        print(f"smooth_fit: E_RMS={args['E_RMS']}")
    scaling_masks = None
    if args['constraint_scaling_maps'] is not None:
        # per-node multipliers of the expected constraint values (smooth_fit.py:467-478): maps that can interpolate themselves
        # (the reference's pointCollection grids) or arrays already on the (y, x) nodes of the z0 / dz grid; non-finite -> 1
        scaling_masks = {}
        for key, this_map in args['constraint_scaling_maps'].items():
            g = grids['z0'] if 'z0' in key else grids['dz']
            vals = this_map.interp(g.ctrs[1], g.ctrs[0], gridded=True) if hasattr(this_map, 'interp') else this_map
            vals = np.array(vals, dtype=float)
            if vals.shape != tuple(g.shape[0:2]):
                raise ValueError(f'constraint_scaling_maps[{key!r}] must sit on the (y, x) nodes of the {g.name} grid')
            vals[~np.isfinite(vals)] = 1.
            scaling_masks[key] = vals
    setup_smoothness_constraints(grids, constraint_op_list, args['E_RMS'], args['mask_scale'], scaling_masks=scaling_masks)
    # gridded biases that multiply a data parameter (setup_grid_bias.py:4-43; the live reference driver has no call site for
    # it, smooth_fit.py:487-489 -- the keyword is the one of its predecessor, deprecated/smooth_xytb_fit.py:465-467)
    if args['grid_bias_model_args'] is not None:
        for bm_args in args['grid_bias_model_args']:
            setup_grid_bias(data, G_data, constraint_op_list, grids, _pts=[y_d[lo:hi].contiguous(), x_d[lo:hi].contiguous()],
                            _scale=cov_d[bm_args['param']][lo:hi].contiguous(), **bm_args)
    stages.mark('c:smoothness_ops')
    if args['bias_params'] is not None:
        n_bias = len(bias_model['bias_ID_dict'])
        col_0 = G_data.col_N
        col_N = col_0 + n_bias
        ids_loc = bias_id_d[lo:hi]
        G_bias = lin_op(name='data_bias', col_0=col_0, col_N=col_N).data_bias(
            torch.arange(hi - lo, dtype=torch.int32, device=ids_loc.device), col=(ids_loc + int(col_0)).contiguous(),
            _toc=(range(hi - lo), np.arange(col_0, col_N, dtype=int)), _regular=True)
        ii = np.arange(n_bias, dtype=int)
        Gc_bias = lin_op(name='constraint_data_bias', col_0=col_0, col_N=col_N).data_bias(ii, col=col_0 + ii)
        for key in bias_model['bias_ID_dict']:
            bias_model['bias_ID_dict'][key]['col'] = col_0 + key
        Gc_bias.expected = np.array([bias_model['E_bias'][k] for k in ii])
        if np.any(Gc_bias.expected == 0):
            raise ValueError('found an zero value in the expected biases')
        constraint_op_list.append(Gc_bias)
        G_data.add(G_bias)
        if args['bias_edit_vals'] is not None:
            want = np.c_[[args['bias_edit_vals'][k] for k in args['bias_edit_vals']]].T.tolist()
            have = np.c_[[bias_model['bias_param_dict'][k] for k in args['bias_edit_vals']]].T.tolist()
            for row in want:
                bias_model['bias_param_dict']['edited'][have.index(row)] = True
            bad = [bias_model['bias_param_dict']['ID'][i] for i in np.flatnonzero(bias_model['bias_param_dict']['edited'])]
            if tse_d is None:
                tse_d = torch.ones(n_all, dtype=f64, device=z_d.device)
            tse_d[torch.isin(bias_id_d, _cabi.to_dev(np.asarray(bad, dtype=np.int32)))] = 0.
    row_range = None if whole else (lo, hi)
    # per-sensor planar slope biases (smooth_fit.py:515-525)
    if args['data_slope_sensors'] is not None and len(args['data_slope_sensors']) > 0:
        if args['bias_params'] is None:
            # the reference fails at the very end here (parse_biases adds lists to None, bias_functions.py:131)
            raise ValueError('data_slope_sensors needs bias_params (parse_biases reports the slopes with the biases)')
        sensors = np.asarray(args['data_slope_sensors'])
        sensor_of = data.sensor
        args['data_slope_sensors'] = sensors[np.isin(sensors, _cabi.to_host(sensor_of) if isinstance(sensor_of, torch.Tensor) else np.asarray(sensor_of))]
        if len(args['data_slope_sensors']) > 0:
            bias_model['E_slope_bias'] = args['E_slope_bias']
            G_slope_bias, Gc_slope_bias, bias_model = data_slope_bias(
                data, bias_model, sensors=args['data_slope_sensors'], col_0=G_data.col_N, E_rms_bias=args['E_slope_bias'],
                _row_range=row_range)
            if G_slope_bias is not None:
                G_data.add(G_slope_bias)
                constraint_op_list.append(Gc_slope_bias)
    # gridded per-sensor biases (smooth_fit.py:527-572); the jitter / stripe variants need a DEM file (GDAL): out of scope
    if args['sensor_grid_bias_params'] is not None:
        for params in args['sensor_grid_bias_params']:
            if any('jitter' in k or 'stripe' in k for k in params):
                raise NotImplementedError('jitter_* / stripe_* sensor grid biases need DEM files (GDAL): out of scope')
            if len(params) and 'spacing' in params:
                setup_sensor_grid_bias(data, grids, G_data, constraint_op_list, _row_range=row_range, **params)
    # prior constraints from an existing dz model (smooth_fit.py:575-576; in-memory data sets only)
    if args['prior_args'] is not None:
        constraint_op_list += match_prior_dz(grids, **args['prior_args'])
    no_priors = all(op.prior is None for op in constraint_op_list)      # the default: every constraint pulls towards zero
    for op in constraint_op_list:
        if op.prior is None:
            op.prior = np.zeros_like(op.expected)

    stages.mark('c:bias_ops')
    Gc = lin_op(None, name='constraints').vstack(constraint_op_list)
    stages.mark('c:vstack')
    names = [op.name for op in constraint_op_list]
    if len(set(names)) == len(names) and sum(int(op.N_eq) for op in constraint_op_list) == int(Gc.N_eq) \
            and all(np.size(op.expected) == op.N_eq for op in constraint_op_list):
        Ec = np.concatenate([np.asarray(op.expected, dtype=float).ravel() for op in constraint_op_list])    # rows are stacked in list order
    else:
        Ec = np.zeros(Gc.N_eq)
        for op in constraint_op_list:
            Ec[rows_of(Gc.TOC['rows'][op.name])] = op.expected
    if bool((s_loc == 0).any().item()):
        raise ValueError('zero value found in data sigma')
    if Ec.size and Ec.min() <= 0 and np.any(Ec == 0):
        raise ValueError('zero value found in constraint sigma')
    # TCinv and rhs stay on the device for the data rows (iterate_fit accepts tensors or NumPy)
    Ec_d = _cabi.to_dev(Ec, f64)
    TCinv = DiagonalOp(torch.cat((1. / s_loc, 1. / Ec_d)))
    if no_priors:
        rhs = torch.cat((z_loc, torch.zeros(int(Gc.N_eq), dtype=f64, device=z_loc.device)))
    else:
        rhs = torch.cat((z_loc, _cabi.to_dev(np.concatenate([op.prior for op in constraint_op_list]), f64)))
    stages.mark('c:Ec_rhs')
    if int(G_data.col_N) != int(Gc.col_N):
        # the reference stacks G_data on Gc at this point (smooth_fit.py:621): constraints that do not reach the last column
        # of G_data (e.g. no dz constraint at all) make scipy's vstack fail with this message
        raise ValueError('incompatible dimensions for axis 1')
    Ip_c = build_reference_epoch_matrix(G_data, Gc, grids, args['reference_epoch'])
    stages.mark('c:Ip_c')
    timing['setup'] = time() - tic

    if tse_d is not None:
        in_TSE = (tse_d[lo:hi] > 0.01)
    else:
        in_TSE = torch.ones(hi - lo, dtype=torch.bool, device=z_d.device)
    if verbose:
        print("initial: %d:" % (n_all - 1), flush=True)
    if args['return_fit_objects']:
        return {'data': local, 'G_data': G_data, 'Gc': Gc, 'grids': grids, 'Ed': s_loc.cpu().numpy(), 'Ec': Ec}

    averaging_ops = {}
    if args['max_iterations'] > 0:
        tic_iteration = time()
        m0, se_d, tse_out_d, r_d, _ = iterate_fit_device(local, None, rhs, TCinv, G_data, Gc, in_TSE, Ip_c, timing, args,
                                                        grids, bias_model=bias_model,
                                                        sigma_extra_masks=None if args['sigma_extra_masks'] is None else
                                                        {k: v[lo:hi] for k, v in args['sigma_extra_masks'].items()})
        timing['iteration'] = time() - tic_iteration
        stages.mark('iterate_fit')
        # model estimate of every data point, G_data . m0 (smooth_fit.py:662); the CSR is already cached
        G_rows = G_data._csr_device(col_N=m0.size, row_N=hi - lo)
        z_own, s_own, ids_own = z_loc, s_loc, (None if bias_id_d is None else bias_id_d[lo:hi])
        if shard is not None:            # results are reported for the own rows; the halo copies belong to other ranks
            o0, o1 = shard['own_rows']
            if (o0, o1) != (0, hi - lo):
                G_rows = G_rows.row_slice(o0, o1)
                z_own, s_own = z_loc[o0:o1].contiguous(), s_loc[o0:o1].contiguous()
                ids_own = None if ids_own is None else ids_own[o0:o1].contiguous()
        z_est_d = G_rows.dot(_cabi.to_dev(m0, f64))
        multi = dist is not None and dist.world > 1

        def to_caller(t):
            """device rows of this rank -> NumPy over all points in the caller's order"""
            if shard is not None:
                t = dist.gather_var_device(t, shard['counts'])
            elif multi:
                t = dist.gather_rows_device(t, n_all)
            if not host_out:                      # took part in the gather; the copy to the host is rank 0's
                return None
            if order is not None:
                out = torch.empty_like(t)
                out[order] = t                    # undo the internal spatial ordering on the device
                t = out
            return _cabi.to_host(t)

        in_TSE_all = to_caller(tse_out_d)
        if host_out:
            in_TSE_all = in_TSE_all.view(np.bool_) if in_TSE_all.dtype == np.uint8 else in_TSE_all.astype(bool)    # 0 / 1 bytes
        sigma_extra_all = to_caller(se_d)
        z_est_all = to_caller(z_est_d)
        if args.get('history') is not None and order is not None and (not multi or shard is not None):
            for h in args['history']:             # per-iteration masks were recorded in the internal order
                mk = torch.empty(n_all, dtype=torch.uint8, device=z_d.device)
                mk[order] = _cabi.to_dev(np.unpackbits(h['mask_out'])[:n_all])
                h['mask_out'] = np.packbits(_cabi.to_host(mk))
        if host_out:
            if valid_data is None:
                valid_data = in_TSE_all.copy()
            else:
                valid_data[valid_data] = in_TSE_all
            data.assign({'three_sigma_edit': in_TSE_all})
            data.assign({'sigma_extra': sigma_extra_all})
            data.assign({'z_est': np.reshape(z_est_all, data.shape)})
        dev_fields = {'z': z_own, 'sigma': s_own, 'z_est': z_est_d, 'three_sigma_edit': tse_out_d, 'sigma_extra': se_d,
                      'bias_ID': ids_own, 'G_rows': G_rows, 'Ec': Ec_d}
        if args['mask_update_function'] is not None:
            # the caller's hook sees the parsed model (no averaging grids yet) and may change the grid masks / cell areas
            # before the averaging operators are built (smooth_fit.py:664-666)
            parse_model(m, m0, dev_fields, R, RMS, G_data, {}, Gc, Ec, grids, bias_model, args,
                        dist=dist if (dist is not None and dist.world > 1) else None)
            args['mask_update_function'](grids, m, args)
        averaging_ops = setup_averaging_ops(grids['dz'], grids['dz'].col_N, args, grids['dz'].cell_area)
        averaging_ops.update(setup_z0_avg(grids, grids['dz'].col_N, args))
        averaging_ops.update(setup_avg_mask_ops(grids['dz'], G_data.col_N, args['avg_masks'], args['dzdt_lags']))
        stages.mark('outputs_d2h')
        parse_model(m, m0, dev_fields, R, RMS, G_data, averaging_ops, Gc, Ec, grids, bias_model, args,
                    dist=dist if (dist is not None and dist.world > 1) else None, stages=stages)
        stages.mark('parse_model')
        if args['compute_E']:
            # error propagation (smooth_fit.py:689-712): data errors inflated by sigma_extra, edited rows out
            if dist is not None and dist.world > 1:
                raise NotImplementedError('compute_E runs on one GPU')
            from .propagate_errors import covariance_dense, calc_and_parse_errors
            w_e = torch.cat((tse_out_d.to(f64) / (s_loc * s_loc + se_d * se_d), 1. / Ec_d ** 2))
            fixed = np.setdiff1d(np.arange(m0.size), np.asarray(Ip_c.free_cols)) if not sp.issparse(Ip_c) else \
                np.setdiff1d(np.arange(m0.size), np.unique(Ip_c.tocoo().row))
            cov = covariance_dense(G_rows, Gc._csr_device(col_N=m0.size, row_N=int(Gc.N_eq)), w_e, fixed)
            calc_and_parse_errors(E, cov, grids, G_data, Gc, averaging_ops, bias_model, args['bias_params'])
            del cov
            stages.mark('compute_E')
    else:
        averaging_ops = setup_averaging_ops(grids['dz'], grids['dz'].col_N, args, grids['dz'].cell_area)
        averaging_ops.update(setup_z0_avg(grids, grids['dz'].col_N, args))
        averaging_ops.update(setup_avg_mask_ops(grids['dz'], G_data.col_N, args['avg_masks'], args['dzdt_lags']))
        if valid_data is None:              # no iterations: the validity mask as it stands (every point valid)
            valid_data = np.ones(n_all, dtype=bool)
        if args['compute_E']:
            # the reference propagates errors without a fit as well (smooth_fit.py:689-712): the initial edit mask and the
            # sigma_extra the data carry; without that field it estimates one from data.z_est, which no fit has produced
            if 'sigma_extra' not in data.fields:
                raise AttributeError("'data' object has no attribute 'z_est'")
            if dist is not None and dist.world > 1:
                raise NotImplementedError('compute_E runs on one GPU')
            from .propagate_errors import covariance_dense, calc_and_parse_errors
            se0 = _field_dev(data, 'sigma_extra')
            if order is not None:
                se0 = se0[order].contiguous()
            n_cols = int(G_data.col_N)
            w_e = torch.cat((in_TSE.to(f64) / (s_loc * s_loc + se0 * se0), 1. / Ec_d ** 2))
            fixed = np.setdiff1d(np.arange(n_cols), np.asarray(Ip_c.free_cols)) if not sp.issparse(Ip_c) else \
                np.setdiff1d(np.arange(n_cols), np.unique(Ip_c.tocoo().row))
            cov = covariance_dense(G_data._csr_device(col_N=n_cols, row_N=hi - lo), Gc._csr_device(col_N=n_cols, row_N=int(Gc.N_eq)), w_e, fixed)
            calc_and_parse_errors(E, cov, grids, G_data, Gc, averaging_ops, bias_model, args['bias_params'])
            del cov
            stages.mark('compute_E')
    stages.mark('epilogue')
    return {'m': m, 'E': E, 'data': data, 'grids': grids, 'valid_data': valid_data, 'TOC': Gc.TOC, 'R': R, 'RMS': RMS,
            'timing': timing, 'E_RMS': args['E_RMS'], 'dzdt_lags': args['dzdt_lags']}
