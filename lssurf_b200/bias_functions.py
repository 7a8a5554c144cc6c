"""Per-group data biases (API mirror of LSsurf/bias_functions.py:25-164)."""
import numpy as np

from .lin_op import lin_op
from .unique_by_rows import unique_by_rows


def assign_bias_ID(data, bias_params=None, bias_name='bias_ID', key_name=None, bias_filter=None, bias_model=None):
    """One bias ID per unique combination of the ``bias_params`` fields (NaN counts as one value),
    with expected magnitude = nanmedian(sigma_corr) of the group (bias_functions.py:25-82)."""
    if bias_model is None:
        bias_model = {'E_bias': dict(), 'bias_param_dict': dict(), 'bias_ID_dict': dict()}
    first = len(bias_model['bias_ID_dict'].keys())
    if bias_params is None:
        bias_model['bias_ID_dict'][first + 1] = key_name
        bias_model['E_bias'][first + 1] = np.nanmedian(data.sigma_corr)
        data.assign({bias_name: first + 1})
        return data, bias_model
    src = bias_filter(data) if bias_filter is not None else data
    NDV = -9999
    table = np.nan_to_num(np.column_stack([getattr(src, bp) for bp in bias_params]), nan=NDV)
    uniq, members = unique_by_rows(table, return_dict=True)
    bias_model['bias_param_dict'].update({p: list() for p in bias_params})
    bias_model['bias_param_dict'].update({'ID': list()})
    bias_ID = np.zeros(data.size)
    for k, vals in enumerate(uniq):
        idx = members[tuple(vals)]
        shown = vals.copy().astype(float)
        shown[shown == NDV] = np.nan
        for j, p in enumerate(bias_params):
            bias_model['bias_param_dict'][p].append(vals[j])
        bias_model['bias_param_dict']['ID'].append(first + k)
        bias_ID[idx] = first + k
        bias_model['bias_ID_dict'][first + k] = {p: shown[j] for j, p in enumerate(bias_params)}
        bias_model['E_bias'][first + k] = np.nanmedian(data.sigma_corr[idx])
    data.assign({bias_name: bias_ID})
    return data, bias_model


def setup_bias_fit(data, bias_model, G_data, constraint_op_list, bias_param_name='data_bias', op_name='data_bias',
                   n_bias=None):
    """Add one column per bias to G_data (a 1 on each of its data rows) and one identity constraint row
    per bias with sigma = E_bias (bias_functions.py:84-115).  ``n_bias`` (new, optional) is the global
    number of biases when ``data`` is one rank's shard of the points; default max(ID) + 1 as in the reference."""
    col = getattr(data, bias_param_name).astype(int)
    if n_bias is None:
        n_bias = np.max(col) + 1
    col_0 = G_data.col_N
    col_N = G_data.col_N + n_bias
    G_bias = lin_op(name=op_name, col_0=col_0, col_N=col_N).data_bias(np.arange(data.size), col=col_0 + col)
    ii = np.arange(n_bias, dtype=int)
    Gc_bias = lin_op(name='constraint_' + op_name, col_0=col_0, col_N=col_N).data_bias(ii, col=col_0 + ii)
    for key in bias_model['bias_ID_dict']:
        bias_model['bias_ID_dict'][key]['col'] = col_0 + key
    Gc_bias.expected = np.array([bias_model['E_bias'][k] for k in ii])
    if np.any(Gc_bias.expected == 0):
        raise ValueError('found an zero value in the expected biases')
    constraint_op_list.append(Gc_bias)
    G_data.add(G_bias)


def parse_biases(m, bias_model, bias_params, data=None):
    """Bias values (and per-bias residual statistics) from the model vector (bias_functions.py:118-164)."""
    slope_bias_dict = {}
    b_dict = {p: list() for p in bias_params + ['val', 'ID', 'expected']}
    if data is not None:
        for k in ('rms_data_raw', 'rms_data_edited', 'N_raw', 'N_edited'):
            b_dict[k] = []
        r = data.z - data.z_est
    for item in bias_model['bias_ID_dict']:
        b_dict['val'].append(m[bias_model['bias_ID_dict'][item]['col']])
        b_dict['ID'].append(item)
        b_dict['expected'].append(bias_model['E_bias'][item])
        for p in bias_params:
            b_dict[p].append(bias_model['bias_ID_dict'][item][p])
        if data is not None:
            sel = data.bias_ID == item
            r1 = r[sel]
            b_dict['rms_data_raw'].append(np.nan if len(r1) == 0 else np.nanstd(r1))
            b_dict['N_raw'].append(len(r1))
            r2 = r[sel & (data.three_sigma_edit == 1)]
            b_dict['N_edited'].append(len(r2))
            b_dict['rms_data_edited'].append(np.nan if len(r2) == 0 else np.nanstd(r2))
    if 'slope_bias_dict' in bias_model:
        for key, cols in bias_model['slope_bias_dict'].items():
            slope_bias_dict[key] = {'slope_x': m[cols[0]], 'slope_y': m[cols[1]]}
    return b_dict, slope_bias_dict
