"""Unique rows of a 2-D array (host helper; role of LSsurf/unique_by_rows.py:9-52)."""
import numpy as np


def unique_by_rows(x, return_dict=False, return_index=False, return_inverse=False):
    """Lexicographically sorted unique rows of ``x``.

    Optional outputs follow the reference: a dict {row tuple: indices of the rows with that value},
    the index of a representative of each unique row, and the inverse index.
    """
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    uniq, index, inverse = np.unique(x, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).ravel()
    if return_dict:
        order = np.argsort(inverse, kind='stable')
        cuts = np.flatnonzero(np.diff(inverse[order])) + 1
        return uniq, {tuple(uniq[inverse[g[0]]]): g for g in np.split(order, cuts)}
    if return_index and return_inverse:
        return uniq, index, inverse
    if return_index:
        return uniq, index
    if return_inverse:
        return uniq, inverse
    return uniq
