"""Minimal point / grid containers with the pointCollection interface smooth_fit relies on.

The reference passes ``pointCollection.data`` objects (SmithB/pointCollection, not installed here and
outside the hot path).  smooth_fit only needs duck typing: attribute arrays,
``fields``, ``size``, ``shape``, ``coords()``, ``copy()``, ``copy_subset()``, ``index()``,
``assign()``, ``__getitem__``.  Any object with that interface (including a real pc.data) works.
"""
import copy as _copy

import numpy as np


def _as_field(val):
    """NumPy array for host data; CUDA / CPU torch tensors are kept as they are (device-resident inputs)."""
    if type(val).__module__.startswith('torch'):
        return val
    return np.asarray(val)


def _shape_size(val):
    if type(val).__module__.startswith('torch'):
        return tuple(val.shape), int(val.numel())
    a = np.asarray(val)
    return a.shape, a.size


class data(object):
    def __init__(self, fields=None, **kw):
        self.fields = [] if fields is None else list(fields)
        self.shape = None
        self.size = None

    def __update_size_and_shape__(self):
        if self.fields:
            self.shape, self.size = _shape_size(getattr(self, self.fields[0]))
        else:
            self.shape, self.size = (0,), 0
        return self

    def from_dict(self, fields_dict, **kw):
        for key, val in fields_dict.items():
            setattr(self, key, _as_field(val))
            if key not in self.fields:
                self.fields.append(key)
        return self.__update_size_and_shape__()

    def from_list(self, items, **kw):
        names = list(items[0].fields)
        for f in names:
            setattr(self, f, np.concatenate([np.asarray(getattr(it, f)).ravel() for it in items]))
        self.fields = names
        return self.__update_size_and_shape__()

    def assign(self, fields_dict=None, **kw):
        if fields_dict is None:
            fields_dict = kw
        for key, val in fields_dict.items():
            setattr(self, key, val)
            if key not in self.fields:
                self.fields.append(key)
        return self.__update_size_and_shape__()

    def coords(self):
        if 'time' in self.fields:
            return (self.y, self.x, self.time)
        if 't' in self.fields:
            return (self.y, self.x, self.t)
        return (self.y, self.x)

    def copy(self):
        return _copy.deepcopy(self)

    def copy_subset(self, index, datasets=None, **kw):
        out = data()
        for f in (self.fields if datasets is None else datasets):
            v = _as_field(getattr(self, f))[index]
            out.assign({f: v.clone() if hasattr(v, 'clone') else v.copy()})
        return out

    def view(self):
        """Shallow copy: new container sharing the field arrays."""
        out = data()
        for f in self.fields:
            out.assign({f: getattr(self, f)})
        return out

    def index(self, index, **kw):
        for f in self.fields:
            setattr(self, f, _as_field(getattr(self, f))[index])
        return self.__update_size_and_shape__()

    def __getitem__(self, index):
        return self.copy_subset(index)


class grid_data(object):
    """Gridded output container (stand-in for pointCollection.grid.data)."""

    def __init__(self, *a, **kw):
        self.fields = []
        self.shape = None

    def from_dict(self, fields_dict, **kw):
        for key, val in fields_dict.items():
            setattr(self, key, val)
            if key not in ('x', 'y', 'time', 't') and key not in self.fields:
                self.fields.append(key)
                if self.shape is None and hasattr(val, 'shape'):
                    self.shape = val.shape
        return self

    def assign(self, fields_dict=None, **kw):
        if fields_dict is None:
            fields_dict = kw
        for key, val in fields_dict.items():
            setattr(self, key, val)
            if key not in self.fields:
                self.fields.append(key)
        return self
