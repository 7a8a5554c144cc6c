"""Per-sensor planar (x, y) slope biases (API mirror of LSsurf/data_slope_bias.py:11-65)."""
import numpy as np
import torch

from . import _cabi
from .lin_op import lin_op


def _host(v):
    return _cabi.to_host(v) if isinstance(v, torch.Tensor) else np.asarray(v)


def data_slope_bias(data, bias_model, col_0=0, sensors=[], op_name='data_slope', E_rms_bias=1.e-5, _row_range=None):
    """Two columns per sensor holding (x - mean x) / 1000 and (y - mean y) / 1000 of that sensor's points.

    Returns ``G_bias, Gc_bias, bias_model`` exactly as the reference: G_bias carries the data-row triplets (its
    N_eq / shape are left untouched, data_slope_bias.py:54-57), Gc_bias one identity row per new column with
    ``expected = E_rms_bias * 1000`` and zero prior; ``bias_model['slope_bias_dict'][sensor]`` = the two columns.
    The means are taken over ALL points of the sensor on the host with NumPy (bit-identical values); the triplets
    are formed on the device.  ``_row_range=(lo, hi)`` (not in the reference) keeps this rank's block of data rows.
    """
    sensors = np.asarray(sensors)
    sensor_of = _host(data.sensor)
    these_sensors = sensors[np.isin(sensors, sensor_of)]
    if len(these_sensors) == 0:
        return None, None, bias_model
    if 'slope_bias_dict' not in bias_model:
        bias_model['slope_bias_dict'] = {}
    col_N = col_0 + 2 * len(these_sensors)
    rescale = 1000
    rr, cc, vv = [], [], []
    lo, hi = (0, sensor_of.size) if _row_range is None else _row_range
    for d_col_1, sensor in enumerate(these_sensors):
        rows = np.flatnonzero(sensor_of == sensor)
        bias_model['slope_bias_dict'][sensor] = col_0 + d_col_1 * 2 + np.array([0, 1])
        mine = rows[(rows >= lo) & (rows < hi)]
        rows_d = _cabi.to_dev(mine - lo, torch.int32)
        for d_col_2, var in enumerate(['x', 'y']):
            full = _host(getattr(data, var))
            mean = np.nanmean(full[rows])
            delta = _cabi.to_dev(full[mine], torch.float64)
            rr.append(rows_d)
            cc.append(torch.full_like(rows_d, int(col_0 + d_col_1 * 2 + d_col_2)))
            vv.append(_cabi.true_div(delta - float(mean), float(rescale)))
    G_bias = lin_op(col_0=col_0, col_N=col_N, name=op_name)
    G_bias._set_device(torch.cat(rr).contiguous(), torch.cat(cc).contiguous(), torch.cat(vv).contiguous())
    ii = np.arange(col_0, col_N, dtype=int)
    Gc_bias = lin_op(name='constraint_' + op_name, col_0=col_0, col_N=col_N).data_bias(ii - ii[0], col=ii)
    Gc_bias.expected = np.zeros(Gc_bias.shape[0]) + E_rms_bias * rescale
    Gc_bias.prior = np.zeros(Gc_bias.shape[0])
    return G_bias, Gc_bias, bias_model
