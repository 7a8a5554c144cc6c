"""Python handle on the matrix-free normal-equation PCG solver (``lss_solver_*`` in the C ABI).

Stands in for the single line ``m0 = Ip_c.dot(sparseqr.solve(Ip_r.dot(TCinv.dot(Gcoo)), Ip_r.dot(TCinv.dot(rhs))))``
(LSsurf/smooth_fit.py:142).  The operator A = [G_local; C_local] and its transpose stay in GPU
memory for the whole edit loop; only the row-weight vector changes between outer iterations.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from .lin_op import DeviceCSR


class NormalSolver(object):
    def __init__(self, G, Cn, z, prior, sigma, Ec, fixed_cols, dist=None, ell=True):
        """G: DeviceCSR of this rank's data rows; Cn: DeviceCSR of this rank's constraint rows;
        z / sigma: data values and errors (CUDA fp64); prior / Ec: constraint right-hand side and sigma;
        fixed_cols: columns held at zero (reference epoch)."""
        L = _cabi.lib()
        self.n_data = G.shape[0]
        self.n_cols = G.shape[1]
        self.A = DeviceCSR.vstack([G, Cn]) if Cn.shape[0] > 0 else G
        # the data rows are the leading rows of A: a view, no second copy of G
        self.G = DeviceCSR(self.A.indptr[:self.n_data + 1], self.A.indices, self.A.data, (self.n_data, self.n_cols))
        self.At = self.A.transpose()
        self.rhs = torch.cat([z, prior]).contiguous()
        self.sigma = sigma
        self.w = torch.empty(self.A.shape[0], dtype=torch.float64, device=z.device)
        self.w[self.n_data:] = 1. / (Ec * Ec)                       # constraint rows: fixed weight 1/E_c^2
        self.fixed = _cabi.to_dev(np.asarray(fixed_cols), torch.int32)
        self.dist = dist
        self.handle = C.c_void_p(0)
        comm = dist.comm if dist is not None else None
        _cabi.check(L.lss_solver_create(C.byref(self.handle), self.A.shape[0], self.n_cols, _cabi.ptr(self.A.indptr),
                                        _cabi.ptr(self.A.indices), _cabi.ptr(self.A.data), _cabi.ptr(self.At.indptr),
                                        _cabi.ptr(self.At.indices), _cabi.ptr(self.At.data), _cabi.ptr(self.rhs),
                                        _cabi.ptr(self.fixed), int(self.fixed.numel()), comm, _cabi.stream()))
        # data rows have a fixed, short row length: apply them in column-major (ELL) form
        self.C = Cn
        row_len = int((G.indptr[1:] - G.indptr[:-1]).max().item()) if self.n_data > 0 else 0
        self.ell = bool(ell) and 1 <= row_len <= 16 and not os.environ.get('LSS_NO_ELL')
        if self.ell:
            has_c = Cn.shape[0] > 0
            _cabi.check(L.lss_solver_split_rows(self.handle, self.n_data, row_len, _cabi.ptr(Cn.indptr) if has_c else None,
                                                _cabi.ptr(Cn.indices) if has_c else None, _cabi.ptr(Cn.data) if has_c else None))
        self.x = torch.zeros(self.n_cols, dtype=torch.float64, device=z.device)
        self.last = {}

    def set_rows(self, in_TSE, sigma_extra=None):
        """Select data rows (uint8 mask) and, on the relaxed last pass, inflate their errors."""
        _cabi.data_weights(self.sigma, sigma_extra, in_TSE, self.w[:self.n_data])
        _cabi.check(_cabi.lib().lss_solver_set_weights(self.handle, _cabi.ptr(self.w)))

    def set_profile(self, on):
        _cabi.check(_cabi.lib().lss_solver_set_profile(self.handle, int(bool(on))))

    def solve(self, rtol=1e-12, max_iter=20000, warm_start=True):
        if not warm_start:
            self.x.zero_()
        it, rel = C.c_int32(0), C.c_double(0.0)
        _cabi.check(_cabi.lib().lss_solver_solve(self.handle, _cabi.ptr(self.x), float(rtol), int(max_iter),
                                                 C.byref(it), C.byref(rel)))
        st = (C.c_double * 6)()
        _cabi.check(_cabi.lib().lss_solver_stats(self.handle, st))
        self.last = dict(iterations=it.value, relres=rel.value, ms=st[0], ms_A=st[1], ms_At=st[2], n_apply=int(st[3]),
                         launches=int(st[4]), nnz=int(st[5]), rows=int(self.A.shape[0]))
        return self.x

    def residual(self, x, z, out=None):
        """r = z - G x over this rank's data rows (LSsurf/smooth_fit.py:146)."""
        return self.G.dot(x, mode=_cabi.SPMV_RESIDUAL, s=z, out=out)

    def close(self):
        if self.handle:
            _cabi.load().lss_solver_destroy(self.handle)
            self.handle = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
