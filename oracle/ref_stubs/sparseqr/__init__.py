"""CPU stand-in for PySPQR's `sparseqr.solve` (SuiteSparseQR is not installable here).

TEST INFRASTRUCTURE ONLY. `solve(A, b)` returns the least-squares solution of the full-column-
rank system via SuperLU on the normal equations followed by iterative refinement with the
residual formed through A; small systems are solved by dense lstsq (gold).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

DENSE_MAX = 3000
LAST = {}


def solve(A, b, tolerance=None):
    A = sp.csr_matrix(A)
    b = np.asarray(b, dtype=float).ravel()
    if A.shape[1] <= DENSE_MAX:
        x = np.linalg.lstsq(A.toarray(), b, rcond=None)[0]
        LAST['method'] = 'dense lstsq'
        return x
    At = A.T.tocsr()
    N = (At @ A).tocsc()
    lu = spl.splu(N)
    x = lu.solve(At @ b)
    for _ in range(3):
        x = x + lu.solve(At @ (b - A @ x))
    LAST['method'] = 'splu(AtA)+3 refinements'
    return x


def rz(*a, **kw):
    raise NotImplementedError('sparseqr.rz (compute_E path) is out of scope')
