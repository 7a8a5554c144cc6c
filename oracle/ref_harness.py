"""Import the UNMODIFIED upstream LSsurf python from /root/reference under stub dependencies.

TEST INFRASTRUCTURE ONLY. Only usable in the build container (``/root/reference`` does not exist
on the GPU box); used by ``oracle/make_golden.py`` to pin the oracle and to generate the fixtures
under ``tests/golden/``.  Nothing in the product package imports this file.

Recipe: the five third-party modules the reference imports but which are
not installed (osgeo, geopandas, requests, pointCollection, sparseqr) are provided by
``oracle/ref_stubs``; the Cython extension ``LSsurf.inv_tr_upper`` (only used by the
``compute_E`` error-propagation path) is replaced by a module holding a dense NumPy stand-in (``inv_tr_upper`` below)
so that nothing has to be copied or compiled from the reference tree.
"""
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get('LSSURF_REFERENCE', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_stubs')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'LSsurf'))


def inv_tr_upper(R, nnz, tol):
    """Stand-in for the reference's Cython extension (LSsurf/inv_tr_upper.pyx:19-94, not compiled here: nothing is built
    from the reference tree): the inverse of the upper-triangular CSR matrix R by back substitution, column by column,
    keeping the diagonal and every entry larger than ``tol`` in magnitude; returns (rows, cols, values, status) with
    status 1 when more than ``nnz`` entries would be needed.  Dense triangular solve, small fixtures only."""
    import numpy as np
    import scipy.linalg as sl
    Rd = np.asarray(R.todense())
    Ri = sl.solve_triangular(Rd, np.eye(Rd.shape[0]), lower=False)
    keep = (np.abs(Ri) > tol) | np.eye(Rd.shape[0], dtype=bool)
    keep &= np.triu(np.ones_like(keep))
    rr, cc = np.nonzero(keep)
    if rr.size >= nnz - 1:
        return rr[:0].astype(np.int32), cc[:0].astype(np.int32), Ri[rr[:0], cc[:0]], 1
    return rr.astype(np.int32), cc.astype(np.int32), Ri[rr, cc], 0


def load():
    """Return the upstream ``LSsurf`` package (module object) with stubs in place."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    if 'LSsurf' in sys.modules and getattr(sys.modules['LSsurf'], '__file__', '').startswith(REF_ROOT):
        return sys.modules['LSsurf']
    for p in (REF_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    placeholder = types.ModuleType('LSsurf.inv_tr_upper')

    placeholder.inv_tr_upper = inv_tr_upper
    sys.modules['LSsurf.inv_tr_upper'] = placeholder
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import LSsurf  # noqa: E402
    return LSsurf


def smooth_fit_module():
    """The *module* LSsurf.smooth_fit (the package attribute of that name is the function)."""
    load()
    return sys.modules['LSsurf.smooth_fit']


def point_data(**fields):
    """Build the stub pointCollection.data the reference expects."""
    load()
    import pointCollection as pc
    return pc.data().from_dict(fields)
