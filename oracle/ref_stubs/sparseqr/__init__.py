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


def rz(A, b, tolerance=None):
    """Stand-in for PySPQR's ``rz`` (Q-less QR: returns Z = Q^T b, R, the column permutation and the rank) as the
    reference's error propagation calls it (LSsurf/smooth_fit.py:219).  Dense Householder QR without column pivoting
    (LAPACK through NumPy), so the permutation is the identity; only for the small systems of the compute_E fixture."""
    A = sp.csr_matrix(A)
    if A.shape[1] > DENSE_MAX:
        raise NotImplementedError('sparseqr.rz stand-in: dense QR only (%d columns)' % A.shape[1])
    b = np.asarray(b.todense() if sp.issparse(b) else b, dtype=float).reshape(A.shape[0], -1)
    Q, R = np.linalg.qr(A.toarray(), mode='reduced')
    return Q.T @ b, sp.coo_matrix(np.triu(R)), np.arange(A.shape[1]), A.shape[1]
