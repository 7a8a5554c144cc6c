"""Smoothness constraints and the reference-epoch column selection
(API mirror of LSsurf/constraint_functions.py:23-151)."""
import numpy as np
import scipy.sparse as sp

from .lin_op import lin_op


def _scaled(op, sigma, mask_scale, scaling_masks, keys):
    op.expected = sigma * op.mask_for_ind0(mask_scale)
    for key in keys:
        if key in scaling_masks:
            op.expected *= op.mask_for_ind0(mask=scaling_masks[key])
            break
    return op


def setup_smoothness_constraints(grids, constraint_op_list, E_RMS, mask_scale, scaling_masks=None):
    """Append the constraint operators selected by the keys of ``E_RMS`` with their per-row expected
    sigma = E_RMS[key] / sqrt(cell area or volume) x mask scale (constraint_functions.py:40-105).

    Stacking order: grad2_z0, grad_z0, mag_z0, grad2_dzdt, grad_dzdt, d2z_dt2 (+ dz_zero).  As in the
    reference the z0 area factor only exists when 'd2z0_dx2' is given (:41, :50).
    """
    if scaling_masks is None:
        scaling_masks = {}
    z0, dz = grids['z0'], grids['dz']
    if 'd2z0_dx2' in E_RMS:
        root_area_z0 = np.sqrt(np.prod(z0.delta))
        constraint_op_list += [_scaled(lin_op(z0, name='grad2_z0').grad2(DOF='z0'), E_RMS['d2z0_dx2'] / root_area_z0,
                                       mask_scale, scaling_masks, ['d2z0_dx2'])]
    if 'dz0_dx' in E_RMS:
        constraint_op_list += [_scaled(lin_op(z0, name='grad_z0').grad(DOF='z0'), E_RMS['dz0_dx'] / root_area_z0,
                                       mask_scale, scaling_masks, ['dz0_dx', 'd2z0_dx2'])]
    if 'z0' in E_RMS and E_RMS['z0'] is not None:
        mag = lin_op(z0, name='mag_z0').one(DOF='z0')
        mag.expected = E_RMS['z0'] / root_area_z0 * np.ones(mag.N_eq)
        constraint_op_list += [mag]
    root_vol_dz = np.sqrt(np.prod(dz.delta))
    if E_RMS.get('d3z_dx2dt') is not None:
        constraint_op_list += [_scaled(lin_op(dz, name='grad2_dzdt').grad2_dzdt(DOF='z', t_lag=1),
                                       E_RMS['d3z_dx2dt'] / root_vol_dz, mask_scale, scaling_masks, ['d3z_dx2dt'])]
    if E_RMS.get('d2z_dxdt') is not None:
        constraint_op_list += [_scaled(lin_op(dz, name='grad_dzdt').grad_dzdt(DOF='z', t_lag=1),
                                       E_RMS['d2z_dxdt'] / root_vol_dz, mask_scale, scaling_masks,
                                       ['d2z_dx2dt', 'd3z_dx2dt'])]
    if E_RMS.get('d2z_dt2') is not None:
        op = lin_op(dz, name='d2z_dt2').d2z_dt2(DOF='z')
        op.expected = np.zeros(op.N_eq) + E_RMS['d2z_dt2'] / root_vol_dz
        if 'd2z_dt2' in scaling_masks:
            op.expected *= op.mask_for_ind0(mask=scaling_masks['d2z_dt2'])
        constraint_op_list += [op]
    if 'dz' in scaling_masks:
        probe = lin_op(dz, name='dz_zero').one(DOF='dz')
        e_dz = probe.mask_for_ind0(mask=scaling_masks['dz'])
        ind = np.flatnonzero((e_dz > 0) & (np.isfinite(e_dz)))
        mag_dz = lin_op(dz, name='dz_zero').one(DOF='dz', which_nodes=ind + dz.col_0)
        mag_dz.expected = e_dz[ind]
        constraint_op_list += [mag_dz]
    if E_RMS.get('lagrangian_dz') is not None or E_RMS.get('lagrangian_dzdx') is not None:
        lag = grids['lagrangian_dz']
        root_area_lag = np.sqrt(lag.delta[0] * lag.delta[1])
        if E_RMS.get('lagrangian_dz') is not None:             # magnitude of the lagrangian offsets (:95-99)
            op = lin_op(lag, name='lagrangian_rms').one(DOF='lagrangian_dz')
            op.expected = np.zeros(op.N_eq) + E_RMS['lagrangian_dz'] / root_area_lag
            constraint_op_list += [op]
        if E_RMS.get('lagrangian_dzdx') is not None:           # and of their gradient (:101-105)
            op = lin_op(lag, name='lagrangian_rms_grad').grad(DOF='lagrangian_dz')
            op.expected = np.zeros(op.N_eq) + E_RMS['lagrangian_dzdx'] / root_area_lag
            constraint_op_list += [op]
    for op in constraint_op_list:
        if np.any(op.expected == 0):
            raise ValueError(f'found zero value in the expected values for {op.name}')


def reference_epoch_columns(grids, reference_epoch):
    """Global columns of dz[:, :, reference_epoch] (constraint_functions.py:128-131)."""
    dz = grids['dz']
    iy, ix = np.meshgrid(np.arange(dz.shape[0]), np.arange(dz.shape[1]), indexing='ij')
    return dz.global_ind([iy.ravel(), ix.ravel(), reference_epoch + np.zeros(iy.size, dtype=int)])


def build_reference_epoch_matrix(G_data, Gc, grids, reference_epoch, dz_mask=None):
    """Column-selection matrix Ip_c (col_N x n_free) that drops the reference-epoch columns
    (constraint_functions.py:112-151).  The solver never multiplies by it: the dropped columns are
    held at zero inside the PCG (they are recovered here from the matrix' empty rows)."""
    if dz_mask is not None:
        raise NotImplementedError('dz_mask column rejection is outside the hot-path scope')
    is_kept = np.ones(G_data.col_N, dtype=bool)
    is_kept[reference_epoch_columns(grids, reference_epoch)] = False
    keep = np.flatnonzero(is_kept)                      # ascending, as setdiff1d returns it
    # one entry per column: the CSC arrays can be written down directly
    return sp.csc_matrix((np.ones_like(keep), keep, np.arange(keep.size + 1)), shape=(Gc.col_N, keep.size))
