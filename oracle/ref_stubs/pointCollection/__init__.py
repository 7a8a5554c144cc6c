"""Minimal duck-typed stand-in for SmithB/pointCollection, enough for LSsurf.smooth_fit.

TEST INFRASTRUCTURE ONLY: used by oracle/make_golden.py to run the unmodified upstream python
from /root/reference in the build container (pointCollection itself is not installed and there
is no network). It is never imported by the product package.
"""
import copy as _copy

import numpy as np


class data(object):
    def __init__(self, fields=None, **kw):
        self.fields = [] if fields is None else list(fields)
        self.shape = None
        self.size = None

    def _update(self):
        if self.fields:
            a = getattr(self, self.fields[0])
            self.shape = a.shape
            self.size = a.size
        else:
            self.shape, self.size = (0,), 0
        return self

    def from_dict(self, d, **kw):
        for k, v in d.items():
            setattr(self, k, np.asarray(v))
            if k not in self.fields:
                self.fields.append(k)
        return self._update()

    def from_list(self, lst, **kw):
        fields = list(lst[0].fields)
        for f in fields:
            setattr(self, f, np.concatenate([np.asarray(getattr(D, f)).ravel() for D in lst]))
        self.fields = fields
        return self._update()

    def assign(self, d=None, **kw):
        if d is None:
            d = kw
        for k, v in d.items():
            setattr(self, k, v)
            if k not in self.fields:
                self.fields.append(k)
        return self._update()

    def coords(self):
        if 'time' in self.fields:
            return (self.y, self.x, self.time)
        if 't' in self.fields:
            return (self.y, self.x, self.t)
        return (self.y, self.x)

    def copy(self):
        return _copy.deepcopy(self)

    def copy_subset(self, ind, datasets=None, **kw):
        out = data()
        for f in (self.fields if datasets is None else datasets):
            out.assign({f: np.asarray(getattr(self, f))[ind].copy()})
        return out

    def index(self, ind, **kw):
        for f in self.fields:
            setattr(self, f, np.asarray(getattr(self, f))[ind])
        return self._update()

    def __getitem__(self, ind):
        return self.copy_subset(ind)


class _grid_data(object):
    def __init__(self, *a, **kw):
        self.fields = []

    def from_dict(self, d, **kw):
        for k, v in d.items():
            setattr(self, k, v)
            if k not in ('x', 'y', 'time', 't'):
                self.fields.append(k)
        return self

    def assign(self, d=None, **kw):
        if d is None:
            d = kw
        for k, v in d.items():
            setattr(self, k, v)
            if k not in self.fields:
                self.fields.append(k)
        return self


class _grid_ns:
    data = _grid_data


grid = _grid_ns()


def unique_by_rows(x, return_dict=False, **kw):
    """Unique rows (sorted lexicographically) and, optionally, {row tuple: indices (ascending)}."""
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[:, None]
    u, inv = np.unique(x, axis=0, return_inverse=True)
    inv = np.asarray(inv).ravel()
    if not return_dict:
        return u
    order = np.argsort(inv, kind='stable')
    cuts = np.flatnonzero(np.diff(inv[order])) + 1
    groups = np.split(order, cuts)
    return u, {tuple(u[inv[g[0]]]): g for g in groups}
