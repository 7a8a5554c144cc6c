"""Error propagation for the fitted model (``compute_E``; LSsurf/smooth_fit.py:212-274, 689-712, lin_op.py:540-553).

The reference takes the R factor of a sparse QR of the scaled system, inverts it (``inv_tr_upper``, entries below 1e-5
dropped) and reads per-unknown errors as the row norms of R^-1, i.e. sqrt(diag((A^T A)^-1)); operator errors
(dz/dt, block averages) are the row norms of Op R^-1, i.e. sqrt(diag(Op (A^T A)^-1 Op^T)).  A PCG solve has no R factor.

What is built here (design spike): the covariance C = N^-1 of the explicit normal matrix
N = A^T W A (the same device matrix the solver iterates on, data errors inflated by sigma_extra, edited rows out, reference
epoch held at zero), formed DENSE on the GPU and inverted by Cholesky -- exact, for fits of up to ``MAX_DENSE`` free
unknowns.  This is a small-problem path through library calls (torch.linalg -> cuSOLVER); it is not part of the hot path.
Larger fits raise: the design for them is probing -- batched multi-right-hand-side PCG on N with the
two-level preconditioner, right-hand sides = sums of unit vectors on a distance-d colouring of the grid nodes, the error of
each entry bounded by the covariance mass the solves show beyond d / 2.
"""
import numpy as np
import torch

from . import _cabi
from .bias_functions import parse_biases
from .lin_op import DeviceCSR
from .point_data import grid_data
from .solver import NormalMatrix

MAX_DENSE = 20000


def covariance_dense(G_csr, C_csr, w, fixed_cols):
    """C = (A^T diag(w) A)^-1 on the free columns, as a dense fp64 CUDA tensor [n_cols, n_cols] (zero rows / columns at the
    fixed ones).  A = [G; C] as DeviceCSR blocks, w: row weights (CUDA fp64)."""
    A = DeviceCSR.vstack([G_csr, C_csr]) if C_csr.shape[0] > 0 else G_csr
    n = A.shape[1]
    free = np.setdiff1d(np.arange(n), np.asarray(fixed_cols, dtype=np.int64))
    if free.size > MAX_DENSE:
        raise NotImplementedError('compute_E: %d free unknowns exceed the dense error-propagation path (%d); the probing design '
                                  'for large fits is described in DESIGN.md section 9 and not built' % (free.size, MAX_DENSE))
    nm = NormalMatrix(A, fixed_cols)
    try:
        nm.update(w.contiguous())
        indptr, indices, data = nm.export()
    finally:
        nm.close()
    dev = data.device
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (indptr[1:] - indptr[:-1]).long())
    N = torch.zeros((n, n), dtype=torch.float64, device=dev)
    N[rows, indices.long()] = data
    fr = torch.from_numpy(free).to(dev)
    Nf = N[fr][:, fr]
    Nf = 0.5 * (Nf + Nf.T)                       # the two triangles are accumulated in different orders
    L = torch.linalg.cholesky(Nf)
    Cf = torch.cholesky_inverse(L)
    C = torch.zeros_like(N)
    C[fr.unsqueeze(1), fr.unsqueeze(0)] = Cf
    return C


def calc_and_parse_errors(E, cov, grids, G_data, Gc, avg_ops, bias_model, bias_params):
    """Fill E like the reference (smooth_fit.py:249-274): per-node errors of z0 and dz, errors of every averaging / dz-dt
    operator on its destination grid (lin_op.grid_error), errors of the biases."""
    E0 = torch.sqrt(torch.clamp(torch.diagonal(cov), min=0.)).cpu().numpy()
    z0g, dzg = grids['z0'], grids['dz']
    E['sigma_z0'] = grid_data().from_dict({'x': z0g.ctrs[1], 'y': z0g.ctrs[0],
                                           'sigma_z0': np.reshape(E0[Gc.TOC['cols']['z0']], z0g.shape)})
    E['sigma_dz'] = grid_data().from_dict({'x': dzg.ctrs[1], 'y': dzg.ctrs[0], 'time': dzg.ctrs[2],
                                           'sigma_dz': np.reshape(E0[Gc.TOC['cols']['dz']], dzg.shape)})
    n = cov.shape[0]
    for key, op in avg_ops.items():
        grid = op.grid if op.dst_grid is None else op.dst_grid
        ind0 = op.ind0 if op.dst_ind0 is None else op.dst_ind0
        M = op._csr_device(row_N=ind0.size, col_N=n)
        Md = torch.zeros((M.shape[0], n), dtype=torch.float64, device=cov.device)
        rr = torch.repeat_interleave(torch.arange(M.shape[0], device=cov.device), (M.indptr[1:] - M.indptr[:-1]).long())
        Md[rr, M.indices.long()] = M.data
        var = ((Md @ cov) * Md).sum(dim=1)                                   # diag(Op C Op^T)
        out = np.zeros(op.col_N) + np.nan
        out[ind0] = torch.sqrt(torch.clamp(var, min=0.)).cpu().numpy()
        fields = {coord: ctr for coord, ctr in zip(grid.coords, grid.ctrs)}
        fields['sigma_' + key] = out[grid.col_0:grid.col_N].reshape(grid.shape)
        E['sigma_' + key] = grid_data().from_dict(fields)
    if len(bias_model.keys()) > 0:
        E['sigma_bias'], E['sigma_slope_bias'] = parse_biases(E0, bias_model, bias_params)
    return E
