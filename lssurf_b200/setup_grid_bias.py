"""Gridded bias proportional to a data parameter (API mirror of LSsurf/setup_grid_bias.py:4-43)."""
import numpy as np
import torch

from . import _cabi
from .lin_op import lin_op


def setup_grid_bias(data, G_data, constraint_op_list, grids, name=None, param=None, grid=None, expected_rms=None,
                    expected_value=None, expected_rms_grad=None, _pts=None, _scale=None):
    """Columns = nodes of ``grid`` appended after G_data's; entries = interpolation weights x data.<param>.

    Replicates the reference including its operator-precedence quirk: the curvature rows get
    ``expected = expected_rms_grad`` (the division by sqrt(cell area) applies to a zero array,
    setup_grid_bias.py:33).  ``_pts`` / ``_scale`` (not in the reference): device arrays [y, x] and the parameter values of
    the rows G_data holds, when smooth_fit has reordered or sharded the points (``data`` is then not consulted).
    """
    if name is None:
        name = param + '_bias'
    grids[name] = grid
    grid.col_0 = G_data.col_N
    grid.col_N = grid.col_0 + grid.N_nodes
    op = lin_op(grid=grid, name=name).interp_mtx(data.coords()[0:2] if _pts is None else _pts)
    if _scale is None:
        field = getattr(data, param)                 # NumPy array, or a tensor when the caller keeps its fields on the device
        scale = _cabi.to_dev(field if isinstance(field, torch.Tensor) else np.asarray(field, dtype=float), torch.float64)
    else:
        scale = _scale
    r, c, v = op._device('r'), op._device('c'), op._device('v')
    vs = (v * scale[r.long()]).contiguous()
    op._set_device(r, c, vs, blocks=[(c, vs)] if op._blocks and c.dim() == 2 else None)      # still row-regular: no sort for toCSR
    G_data.add(op)
    if expected_rms_grad is not None:
        curv = lin_op(grid, name='grad2_' + name).grad2(DOF=param)
        curv.expected = expected_rms_grad + np.zeros(curv.N_eq) / np.sqrt(np.prod(grid.delta))
        constraint_op_list.append(curv)
    if expected_rms is not None:
        mag = lin_op(grid, name='mag_' + name).data_bias(ind=np.arange(grid.N_nodes), col=np.arange(grid.col_0, grid.col_N))
        mag.expected = expected_rms + np.zeros(mag.N_eq)
        mag.expected_val = (0 if expected_value is None else expected_value) + np.zeros(mag.N_eq)
        constraint_op_list.append(mag)
