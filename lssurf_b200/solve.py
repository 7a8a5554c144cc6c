"""``solve(A, b)``: drop-in for ``sparseqr.solve(A, b)`` -- the one-line seam of the reference
(LSsurf/smooth_fit.py:142; notebooks/smooth_fit_demo_aniso.ipynb cells 6, 13, 16, 18, 20).

Least-squares solution of the full-column-rank sparse system A x = b, where the caller has already scaled rows by the
inverse expected errors (``TCinv.dot(Gcoo)``, ``TCinv.dot(rhs)``).  The solve runs on the GPU as fp64 preconditioned CG on
the normal equations A^T A x = A^T b through the same ``lss_solver_*`` entry points smooth_fit uses (explicit normal
matrix when every row holds <= 24 entries, which is the case for interpolation and stencil rows; two CSR products per
iteration otherwise).  There is no CPU fallback.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from .lin_op import DeviceCSR, lin_op


def _device_csr(A):
    if isinstance(A, DeviceCSR):
        return A
    if isinstance(A, lin_op):
        return A._csr_device()
    M = sp.csr_matrix(A)
    M.sum_duplicates()
    M.sort_indices()
    if M.shape[0] >= 2 ** 31 or M.nnz >= 2 ** 31:
        raise ValueError('matrix too large for the int32 CSR format')
    return DeviceCSR(_cabi.to_dev(M.indptr.astype(np.int32), torch.int32), _cabi.to_dev(M.indices.astype(np.int32), torch.int32),
                     _cabi.to_dev(M.data.astype(np.float64), torch.float64), M.shape)


def solve(A, b, tolerance=None, rtol=1e-12, maxiter=1000000, x0=None, info=None):
    """x = argmin |A x - b|.  A: scipy.sparse matrix (any format), ``lin_op`` or ``DeviceCSR``; b: array of A.shape[0]
    values.  ``tolerance`` is accepted for signature compatibility with sparseqr.solve (a rank tolerance there) and
    ignored; ``rtol`` is the PCG stopping tolerance sqrt(r^T M^-1 r / b^T M^-1 b).  Columns without entries come back as
    0.  Returns a NumPy vector of A.shape[1] values (``info``, if a dict, receives iterations / relres / form)."""
    from . import solver as _solver
    G = _device_csr(A)
    n_rows, n_cols = G.shape
    bd = _cabi.to_dev(np.asarray(_cabi.to_host(b) if isinstance(b, torch.Tensor) else b, dtype=float).ravel(), torch.float64)
    if int(bd.numel()) != n_rows:
        raise ValueError('b must have one value per row of A')
    dev = bd.device
    empty_f = torch.zeros(0, dtype=torch.float64, device=dev)
    Cn = DeviceCSR(torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev), empty_f, (0, n_cols))
    s = _solver.NormalSolver(G, Cn, bd, empty_f, torch.ones(n_rows, dtype=torch.float64, device=dev), empty_f,
                             np.zeros(0, dtype=np.int32))
    try:
        s.set_rows(torch.ones(n_rows, dtype=torch.uint8, device=dev))
        if x0 is not None:
            s.x.copy_(_cabi.to_dev(np.asarray(x0, dtype=float).ravel(), torch.float64))
        x = s.solve(rtol=rtol, max_iter=maxiter, warm_start=x0 is not None)
        last = dict(s.last)
        out = x.cpu().numpy().copy()
    finally:
        s.close()
    if not np.isfinite(last['relres']):
        raise _cabi.LssError('solve: the PCG iteration broke down (residual is not finite)')
    if not last['relres'] <= rtol:
        import warnings
        warnings.warn('solve: PCG stopped at %d iterations with relative residual %.2e > rtol %.1e' % (last['iterations'], last['relres'], rtol),
                      RuntimeWarning)
    if info is not None:
        info.update(iterations=last['iterations'], relres=last['relres'], form=last['form'],
                    converged=bool(last['relres'] <= rtol))
    return out
