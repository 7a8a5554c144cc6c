"""Gridded bias of one sensor's data (API mirror of LSsurf/setup_sensor_grid_bias.py:6-68).

The interpolation rows come from the same kernel as every other interpolation operator (``lss_interp_coo``); the
only extra step is the row remap ``sensor_rows[r]`` (setup_sensor_grid_bias.py:42), a device gather here.
"""
import re

import numpy as np
import torch

from . import _cabi
from .fd_grid import fd_grid
from .lin_op import lin_op
from .point_data import grid_data


def _host(v):
    return _cabi.to_host(v) if isinstance(v, torch.Tensor) else np.asarray(v)


def setup_sensor_grid_bias(data, grids, G_data, constraint_op_list, sensor=None, spacing=None, expected_rms=None,
                           expected_val=0, expected_rms_grad2=None, expected_rms_grad=None, filename=None, dims=2,
                           _row_range=None, **params):
    """Append a 2-D grid of bias columns sampled (bilinearly) by the points of ``sensor``.

    Same arguments and side effects as the reference: the new grid goes into ``grids['sensor_<sensor>_bias']`` with
    its columns after G_data's, curvature rows (``expected_rms_grad2 / sqrt(cell area)``) and magnitude rows
    (``expected_rms``, prior ``expected_val``) are appended to ``constraint_op_list``, the interpolation rows are
    added to ``G_data``.  Returns None (also when the sensor has no points: setup_sensor_grid_bias.py:35-37).
    ``_row_range=(lo, hi)`` (not in the reference) keeps only this rank's block of data rows, renumbered from 0.
    """
    sensor_rows = np.flatnonzero(_host(data.sensor) == sensor)
    x, y = _host(data.x)[sensor_rows], _host(data.y)[sensor_rows]
    name = f'sensor_{sensor}_bias'
    try:
        xr = [np.floor(np.min(x) / spacing) * spacing, np.ceil(np.max(x / spacing)) * spacing]
        yr = [np.floor(np.min(y) / spacing) * spacing, np.ceil(np.max(y / spacing)) * spacing]
        grid = fd_grid([yr, xr], spacing * np.ones(2), name=name)
    except (IndexError, ValueError):
        return                                  # grid contains no nodes
    grids[name] = grid
    grid.col_0 = G_data.col_N
    grid.col_N = grid.col_0 + grid.N_nodes
    op = lin_op(grid=grid, name=name).interp_mtx([y, x])
    rows_d = _cabi.to_dev(sensor_rows, torch.int32)
    r, c, v = op._device('r'), op._device('c'), op._device('v')
    r = rows_d[r.reshape(-1).long()].reshape(r.shape)
    if _row_range is not None:
        lo, hi = _row_range
        keep = (r[:, 0] >= lo) & (r[:, 0] < hi)
        r, c, v = (r[keep] - lo).contiguous(), c[keep].contiguous(), v[keep].contiguous()
    op._set_device(r.contiguous(), c, v)
    if expected_rms_grad2 is not None:
        grad2_b = lin_op(grid, name='grad2_' + name).grad2(DOF=name)
        grad2_b.expected = expected_rms_grad2 / np.sqrt(np.prod(grid.delta)) + np.zeros(grad2_b.N_eq)
        constraint_op_list.append(grad2_b)
    if expected_rms is not None:
        mag_b = lin_op(grid, name='mag_' + name).data_bias(ind=np.arange(grid.N_nodes), col=np.arange(grid.col_0, grid.col_N))
        mag_b.expected = expected_rms + np.zeros(mag_b.N_eq)
        mag_b.prior = expected_val + np.zeros(mag_b.N_eq)
        constraint_op_list.append(mag_b)
    G_data.add(op)


def parse_sensor_bias_grids(m0, G_data, grids):
    """{grid name: gridded bias values} for every 'sensor_*_bias' column block (setup_sensor_grid_bias.py:58-68)."""
    sensor_bias_re = re.compile('sensor_.*_bias')
    m = {}
    for key, cols in G_data.TOC['cols'].items():
        if sensor_bias_re.search(key) is None:
            continue
        m[key] = grid_data().from_dict({'x': grids[key].ctrs[1], 'y': grids[key].ctrs[0],
                                        'z': np.reshape(m0[cols], grids[key].shape)})
    return m
