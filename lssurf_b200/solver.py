"""Python handle on the normal-equation PCG solver (``lss_solver_*`` in the C ABI).

Stands in for the single line ``m0 = Ip_c.dot(sparseqr.solve(Ip_r.dot(TCinv.dot(Gcoo)), Ip_r.dot(TCinv.dot(rhs))))``
(LSsurf/smooth_fit.py:142).  The operator stays in GPU memory for the whole edit loop; only the row-weight vector
changes between outer iterations.

Three forms (include/lssurf_b200.h):
  * explicit normal matrix (default): N = A^T W A is formed on the device from the ELL rows of A = [G; C] (it has fewer
    non-zeros than A) and CG multiplies by it with one gather product per iteration; N is updated incrementally when
    the edit mask changes (``LSS_NO_EXPLICIT_N=1`` or ``lss_set_option('explicit_normal', 0)`` selects the next form);
  * block form (default when every data row holds <= 24 entries: interpolation rows with bias, slope and sensor-grid columns):
    G as an ELL copy applied by the fused ``q += G^T (w .* (G p))`` kernel, the constraint rows as a CSR pair
    (C, C^T) -- one pass over G per CG iteration, no transpose of G;
  * generic form: A = [G; C] and an explicit A^T as CSR, two tiled CSR products per iteration (deterministic;
    selected with ``fused=False`` or ``LSS_NO_FUSED=1``).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi, structured
from .lin_op import DeviceCSR


class NormalMatrix(object):
    """N = A^T diag(w) A on the device for a DeviceCSR A with short rows (``lss_normal_*``)."""

    def __init__(self, A, fixed_cols=None):
        L = _cabi.lib()
        self.n_cols = A.shape[1]
        row_len = int((A.indptr[1:] - A.indptr[:-1]).max().item()) if A.shape[0] > 0 else 1
        fixed = _cabi.to_dev(np.asarray(fixed_cols if fixed_cols is not None else [], dtype=np.int32), torch.int32)
        self.handle = C.c_void_p(0)
        nnz = C.c_int64(0)
        _cabi.check(L.lss_normal_create(C.byref(self.handle), A.shape[0], self.n_cols, max(row_len, 1), _cabi.ptr(A.indptr),
                                        _cabi.ptr(A.indices), _cabi.ptr(A.data), _cabi.ptr(fixed), int(fixed.numel()),
                                        C.byref(nnz), _cabi.stream()))
        self.nnz = int(nnz.value)

    def update(self, w):
        _cabi.check(_cabi.lib().lss_normal_update(self.handle, _cabi.ptr(w)))

    def export(self):
        dev = _cabi.device()
        indptr = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
        data = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=dev)
        _cabi.check(_cabi.lib().lss_normal_export(self.handle, _cabi.ptr(indptr), _cabi.ptr(indices), _cabi.ptr(data)))
        return indptr, indices[:self.nnz], data[:self.nnz]

    def dot(self, x):
        y = torch.empty(self.n_cols, dtype=torch.float64, device=x.device)
        _cabi.check(_cabi.lib().lss_normal_apply(self.handle, _cabi.ptr(x), _cabi.ptr(y)))
        return y

    def close(self):
        if self.handle:
            _cabi.load().lss_normal_destroy(self.handle)
            self.handle = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _structure_device():
    """Where the index tables of the solver structure are built: the GPU when there is one (they are handed to the C ABI
    as device arrays); the CPU otherwise, so that the tables themselves can be checked without a device."""
    return _cabi.device() if torch.cuda.is_available() else torch.device('cpu')


def grid_structure(grids, fixed_cols, bias_cols=None, max_block=64):
    """Column structure handed to ``lss_solver_set_structure``: one preconditioner block per dz node (its free epochs
    plus the z0 nodes nearest to it), and the bias column range to eliminate.  ``grids``: the dict of fd_grid objects
    smooth_fit builds (needs 'z0' with 2 and 'dz' with 3 dimensions, same extent); returns None without them.
    ``block_ptr`` / ``block_cols`` are int32 device tensors (built with tensor ops on the device: 4e5 columns)."""
    out = {}
    z0, dz = grids.get('z0'), grids.get('dz')
    if z0 is not None and dz is not None and len(z0.shape) == 2 and len(dz.shape) == 3:
        dev = _structure_device()
        ny, nx, nt = [int(v) for v in dz.shape]
        # nearest dz node of every z0 row / column (tiny host tables), then everything per column on the device
        jy_tab = np.clip(np.floor((z0.ctrs[0] - dz.ctrs[0][0]) / dz.delta[0] + 0.5), 0, ny - 1).astype(np.int64)
        jx_tab = np.clip(np.floor((z0.ctrs[1] - dz.ctrs[1][0]) / dz.delta[1] + 0.5), 0, nx - 1).astype(np.int64)
        jy_d, jx_d = torch.from_numpy(jy_tab).to(dev), torch.from_numpy(jx_tab).to(dev)
        n0, nd = int(z0.N_nodes), int(dz.N_nodes)
        k = torch.arange(n0, device=dev)
        blk0 = jy_d[k // int(z0.shape[1])] * nx + jx_d[k % int(z0.shape[1])]
        kd = torch.arange(nd, device=dev)
        cols = torch.cat([int(z0.col_0) + k, int(dz.col_0) + kd])
        blk = torch.cat([blk0, kd // nt])
        free = torch.ones(int(cols.max().item()) + 1, dtype=torch.bool, device=dev)
        if len(fixed_cols):
            free[torch.from_numpy(np.asarray(fixed_cols, dtype=np.int64)).to(dev)] = False
        sel = free[cols]
        cols, blk = cols[sel], blk[sel]
        order = torch.argsort(blk, stable=True)
        counts = torch.bincount(blk, minlength=ny * nx)
        counts = counts[counts > 0]
        if counts.numel() and int(counts.max().item()) <= max_block:
            ptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=dev)
            ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            out.update(block_ptr=ptr, block_cols=cols[order].to(torch.int32).contiguous(), max_block=int(counts.max().item()))
        out['grids_desc'] = (int(z0.col_0), tuple(int(v) for v in z0.shape), int(dz.col_0), tuple(int(v) for v in dz.shape))
    if bias_cols is not None and len(bias_cols) > 0:
        b = np.asarray(bias_cols, dtype=np.int64)
        if b.max() - b.min() + 1 == b.size:
            out['elim'] = (int(b.min()), int(b.max()) + 1)
    return out or None


def _coarse_1d(nf, c):
    """Bilinear coarsening of a line of nf nodes by the factor c (every c-th node plus the last one): per fine node the
    lower / upper coarse parent and the weight of the upper one (0 where the fine node is a coarse node)."""
    coarse = np.unique(np.r_[np.arange(0, nf, c), nf - 1])
    i = np.arange(nf)
    lo = np.searchsorted(coarse, i, side='right') - 1
    hi = np.minimum(lo + 1, coarse.size - 1)
    span = np.maximum(coarse[hi] - coarse[lo], 1)
    w = np.where(coarse[lo] == i, 0., (i - coarse[lo]) / span)
    return coarse.size, lo.astype(np.int64), hi.astype(np.int64), w.astype(np.float64)


def coarse_structure(grids, fixed_cols, n_cols, max_coarse=None, owned=None):
    """Coarse space of the two-level preconditioner (``lss_solver_set_coarse``): bilinear functions of (y, x) on a grid
    c times coarser than z0 / dz (same factor for both, every epoch of dz kept), c the smallest power of two that
    leaves at most ``max_coarse`` coarse unknowns.  Returns device tensors: per fine column its <= 4 parents and
    weights (``cp_idx`` int32 [n_cols, 4], -1 = none; ``cp_w`` float32) and P^T as CSR by coarse row.
    Fixed columns (reference epoch) and columns outside the two grids have no parents.  With ``owned`` (uint8 per column)
    P^T keeps the entries at those columns only (one rank's share of the restriction)."""
    z0, dz = grids.get('z0'), grids.get('dz')
    if z0 is None or dz is None or len(z0.shape) != 2 or len(dz.shape) != 3:
        return None
    if max_coarse is None:
        max_coarse = min(int(os.environ.get('LSS_COARSE_MAX', '1024')), 1024)      # CS_MAX of the library
    ny0, nx0 = (int(v) for v in z0.shape)
    ny, nx, nt = (int(v) for v in dz.shape)
    c = 2
    while True:
        n_c = _coarse_1d(ny0, c)[0] * _coarse_1d(nx0, c)[0] + _coarse_1d(ny, c)[0] * _coarse_1d(nx, c)[0] * nt
        if n_c <= max_coarse:
            break
        c *= 2
        if c > 4 * max(ny0, nx0, ny, nx):
            return None
    dev = _structure_device()
    idx = torch.full((n_cols, 4), -1, dtype=torch.int32, device=dev)
    wts = torch.zeros((n_cols, 4), dtype=torch.float32, device=dev)

    def fill(col_0, shp, nt_, base):
        cy, ly, hy, wy = _coarse_1d(shp[0], c)
        cx, lx, hx, wx = _coarse_1d(shp[1], c)
        ly, hy, lx, hx = (torch.from_numpy(a).to(dev) for a in (ly, hy, lx, hx))
        wy, wx = torch.from_numpy(wy).to(dev), torch.from_numpy(wx).to(dev)
        k = torch.arange(shp[0] * shp[1] * nt_, device=dev)
        node, t = k // nt_, k % nt_
        iy, ix = node // shp[1], node % shp[1]
        cols = slice(col_0, col_0 + k.numel())
        for s_, (py, px, w_) in enumerate(((ly[iy], lx[ix], (1 - wy[iy]) * (1 - wx[ix])), (ly[iy], hx[ix], (1 - wy[iy]) * wx[ix]),
                                           (hy[iy], lx[ix], wy[iy] * (1 - wx[ix])), (hy[iy], hx[ix], wy[iy] * wx[ix]))):
            ci = base + (py * cx + px) * nt_ + t
            idx[cols, s_] = torch.where(w_ > 0, ci, torch.full_like(ci, -1)).to(torch.int32)
            wts[cols, s_] = w_.to(torch.float32)
        return base + cy * cx * nt_

    base = fill(int(z0.col_0), (ny0, nx0), 1, 0)
    n_c = fill(int(dz.col_0), (ny, nx), nt, base)
    if len(fixed_cols):
        fc = torch.from_numpy(np.asarray(fixed_cols, dtype=np.int64)).to(dev)
        idx[fc] = -1
        wts[fc] = 0
    flat = idx.reshape(-1)
    keep = flat >= 0
    if owned is not None:          # a rank of the row-sharded solver restricts from its own columns only
        keep = keep & owned.to(dev).bool().repeat_interleave(4)
    sel = torch.nonzero(keep).reshape(-1)
    order = torch.argsort(flat[sel], stable=True)
    ent = sel[order]
    counts = torch.bincount(flat[ent].to(torch.int64), minlength=n_c)
    pt_ptr = torch.zeros(n_c + 1, dtype=torch.int32, device=dev)
    pt_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return dict(n_c=int(n_c), factor=c, cp_idx=idx.contiguous(), cp_w=wts.contiguous(), pt_ptr=pt_ptr,
                pt_col=(ent // 4).to(torch.int32).contiguous(), pt_w=wts.reshape(-1)[ent].contiguous())


class NormalSolver(object):
    def __init__(self, G, Cn, z, prior, sigma, Ec, fixed_cols, dist=None, fused=True, structure=None, G_local=None, stages=None,
                 shard=None):
        """G: DeviceCSR of this rank's data rows; Cn: DeviceCSR of this rank's constraint rows;
        z / sigma: data values and errors (CUDA fp64); prior / Ec: constraint right-hand side and sigma;
        fixed_cols: columns held at zero (reference epoch)."""
        L = _cabi.lib()
        self.n_data, self.n_con, self.n_cols = G.shape[0], Cn.shape[0], G.shape[1]
        self.G, self.C = G, Cn
        self.G_local = G if G_local is None else G_local      # rows whose residual this rank owns (replicated solve: its shard)
        self.rhs = torch.cat([z, prior]).contiguous()
        self.sigma = sigma
        self.w = torch.empty(self.n_data + self.n_con, dtype=torch.float64, device=z.device)
        self.w[self.n_data:] = 1. / (Ec * Ec)                       # constraint rows: fixed weight 1/E_c^2
        self.fixed = _cabi.to_dev(np.asarray(fixed_cols), torch.int32)
        self.dist = dist
        self.handle = C.c_void_p(0)
        comm = dist.comm if dist is not None else None
        row_len = int((G.indptr[1:] - G.indptr[:-1]).max().item()) if self.n_data > 0 else 1
        con_len = int((Cn.indptr[1:] - Cn.indptr[:-1]).max().item()) if self.n_con > 0 else 0
        self.fused = bool(fused) and row_len <= 24 and not os.environ.get('LSS_NO_FUSED')
        self.all_ell = False
        self.shard = shard
        if shard is not None:
            # row-sharded N: G holds this rank's own points plus halo copies, Cn every constraint row; the rank owns the
            # preconditioner blocks [block_lo, block_hi) and exchanges p with its peers through the P2P mailbox
            if not (self.fused and con_len <= 24 and structure is not None and 'block_ptr' in structure):
                raise _cabi.LssError('the row-sharded solver needs short rows and a block structure')
            self.all_ell = True
            self.A = DeviceCSR.vstack([G, Cn]) if self.n_con > 0 else G
            e0, e1 = structure.get('elim', (0, 0))
            self._owned = shard['owned_cols']
            _cabi.check(L.lss_solver_create_sharded(
                C.byref(self.handle), self.n_data + self.n_con, self.n_cols, max(row_len, con_len, 1), _cabi.ptr(self.A.indptr),
                _cabi.ptr(self.A.indices), _cabi.ptr(self.A.data), _cabi.ptr(self.rhs), _cabi.ptr(self.fixed), int(self.fixed.numel()),
                _cabi.ptr(self._owned), int(shard['block_lo']), int(shard['block_hi']), int(shard['own_rows'][0]), int(shard['own_rows'][1]),
                int(self.n_data), int(e0), int(e1), dist.comm, _cabi.stream()))
            self.A = None
            if stages is not None:
                stages.mark('it:solver_create(ELL+symbolic N)')
        elif self.fused:
            p = _cabi.ptr
            nul = C.c_void_p(0)
            self.all_ell = con_len <= 24 and not os.environ.get('LSS_CSR_CONSTRAINTS')
            if self.all_ell:
                # constraint stencils are short rows too (<= 9 entries): one ELL block [G; C], one kernel per product
                self.A = DeviceCSR.vstack([G, Cn]) if self.n_con > 0 else G
                if stages is not None:
                    stages.mark('it:solver_vstack')
                _cabi.check(L.lss_solver_create_blocks(
                    C.byref(self.handle), self.n_data + self.n_con, 0, self.n_cols, max(row_len, con_len, 1),
                    p(self.A.indptr), p(self.A.indices), p(self.A.data), nul, nul, nul, nul, nul, nul,
                    p(self.rhs), p(self.fixed), int(self.fixed.numel()), comm, _cabi.stream()))
                self.A = None                      # the solver keeps its own ELL copy
                if stages is not None:
                    stages.mark('it:solver_create(ELL+symbolic N)')
            else:
                self.Ct = Cn.transpose() if self.n_con > 0 else None
                _cabi.check(L.lss_solver_create_blocks(
                    C.byref(self.handle), self.n_data, self.n_con, self.n_cols, max(row_len, 1),
                    p(G.indptr), p(G.indices), p(G.data),
                    p(Cn.indptr) if self.n_con else nul, p(Cn.indices) if self.n_con else nul, p(Cn.data) if self.n_con else nul,
                    p(self.Ct.indptr) if self.n_con else nul, p(self.Ct.indices) if self.n_con else nul,
                    p(self.Ct.data) if self.n_con else nul,
                    p(self.rhs), p(self.fixed), int(self.fixed.numel()), comm, _cabi.stream()))
        else:
            self.A = DeviceCSR.vstack([G, Cn]) if self.n_con > 0 else G
            self.At = self.A.transpose()
            _cabi.check(L.lss_solver_create(C.byref(self.handle), self.A.shape[0], self.n_cols, _cabi.ptr(self.A.indptr),
                                            _cabi.ptr(self.A.indices), _cabi.ptr(self.A.data), _cabi.ptr(self.At.indptr),
                                            _cabi.ptr(self.At.indices), _cabi.ptr(self.At.data), _cabi.ptr(self.rhs),
                                            _cabi.ptr(self.fixed), int(self.fixed.numel()), comm, _cabi.stream()))
        self.x = torch.zeros(self.n_cols, dtype=torch.float64, device=z.device)
        self.last = {}
        info = (C.c_int64 * 8)()
        _cabi.check(L.lss_solver_info(self.handle, info))
        self.form = ('csr_pair', 'fused_ell', 'explicit_normal')[int(info[0])]
        self.nnz_N, self.n_pairs = int(info[1]), int(info[2])
        self.n_blocks = self.n_elim = self.n_coarse = 0
        if structure is not None and self.form == 'explicit_normal' and not os.environ.get('LSS_NO_STRUCTURE'):
            bp, bc = structure.get('block_ptr'), structure.get('block_cols')
            e0, e1 = structure.get('elim', (0, 0))
            if os.environ.get('LSS_NO_ELIM'):
                e0 = e1 = 0
            nb = 0 if bp is None or os.environ.get('LSS_NO_BLOCKS') else int(bp.numel()) - 1
            zero = torch.zeros(1, dtype=torch.int32, device=z.device)
            self._bp = _cabi.to_dev(bp, torch.int32) if nb else zero
            self._bc = _cabi.to_dev(bc, torch.int32) if nb else zero
            rc = L.lss_solver_set_structure(self.handle, nb, _cabi.ptr(self._bp), _cabi.ptr(self._bc),
                                            int(structure.get('max_block', 1)), int(e0), int(e1))
            if rc == -1 and e1 > e0:         # LSS_EINVAL: the candidate block is not diagonal -> keep those columns in the solve
                rc = L.lss_solver_set_structure(self.handle, nb, _cabi.ptr(self._bp), _cabi.ptr(self._bc),
                                                int(structure.get('max_block', 1)), 0, 0)
            _cabi.check(rc)
            _cabi.check(L.lss_solver_info(self.handle, info))
            self.n_blocks, self.n_elim = int(info[4]), int(info[5])
            cs = structure.get('coarse')
            if cs is not None and self.n_blocks > 0:
                self._coarse = cs          # keeps the device tensors alive until the copies inside the library are made
                _cabi.check(L.lss_solver_set_coarse(self.handle, cs['n_c'], _cabi.ptr(cs['cp_idx']), _cabi.ptr(cs['cp_w']),
                                                    _cabi.ptr(cs['pt_ptr']), _cabi.ptr(cs['pt_col']), _cabi.ptr(cs['pt_w']),
                                                    int(cs['pt_col'].numel()), int(structure.get('coarse_refresh', 1))))
                self.n_coarse = cs['n_c']
            self.structured = None
            gd = structure.get('grids_desc')
            # opt-in (LSS_STRUCTURED_N=1): 0.073 ms instead of 0.092 ms per product on the bench workload, but the tables cost
            # ~12 ms per fit to build, and the kernel is still issue-bound (profiles/structured_product_r02.md)
            if gd is not None and os.environ.get('LSS_STRUCTURED_N', '0') == '1' and shard is None:
                self._set_structured(gd, int(e0) if (e1 > e0 and e1 == self.n_cols) else self.n_cols)

    def _set_structured(self, gd, rows):
        """Grid-structured copy of N for the CG product (structured.py): tables from the exported pattern, handed to the
        library, which fills the values after every numeric update.  Skipped when the layout does not apply."""
        L = _cabi.lib()
        dev = self.x.device
        indptr = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(self.nnz_N, 1), dtype=torch.int32, device=dev)
        _cabi.check(L.lss_solver_export_pattern(self.handle, _cabi.ptr(indptr), _cabi.ptr(indices)))
        st = structured.build(indptr, indices, rows, *gd)
        del indptr, indices
        if st is None:
            return
        d = _cabi.lss_structured_t()
        for k in ('nt', 'nn', 'n0', 'n_diag', 'dd_stride', 'node_len', 'rec_len', 'max_pairs', 'n_zd', 'z0_col0', 'dz_col0', 'rows', 'nnz', 'n_vals', 'zd_base', 'zz_base'):
            setattr(d, k, int(st[k]))
        d.rem_n_rows, d.rem_nnz = int(st['rem_rows'].numel()), int(st['rem_src'].numel())
        for k in ('map', 'dd_b', 'zd_ptr', 'zd_k', 'zz_off', 'rem_src', 'rem_rows', 'rem_indptr', 'rem_indices'):
            setattr(d, k, st[k].data_ptr() if st[k].numel() else None)
        _cabi.check(L.lss_solver_set_structured(self.handle, C.byref(d)))
        self.structured = st          # the library borrows the tables: they live as long as the solver
        self.structured_bytes = structured.product_bytes(st)

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
        st = (C.c_double * 8)()
        _cabi.check(_cabi.lib().lss_solver_stats(self.handle, st))
        info = (C.c_int64 * 8)()
        _cabi.check(_cabi.lib().lss_solver_info(self.handle, info))
        self.last = dict(iterations=it.value, relres=rel.value, ms=st[0], ms_A=st[1], ms_At=st[2], ms_step=st[6], n_apply=int(st[3]),
                         launches=int(st[4]), nnz=int(st[5]), rows=self.n_data + self.n_con, fused=self.fused, all_ell=self.all_ell,
                         n_data=self.n_data, n_con=self.n_con, nnz_G=self.G.nnz, nnz_C=self.C.nnz,
                         form=self.form, nnz_N=self.nnz_N, n_cols=self.n_cols, precond_blocks=self.n_blocks, eliminated=self.n_elim,
                         nnz_product=int(info[7]), bad_blocks=int(info[6]), coarse=self.n_coarse,
                         structured_bytes=getattr(self, 'structured_bytes', 0) if getattr(self, 'structured', None) is not None else 0)
        return self.x

    def residual(self, x, z, out=None):
        """r = z - G x over this rank's data rows (LSsurf/smooth_fit.py:146)."""
        return self.G_local.dot(x, mode=_cabi.SPMV_RESIDUAL, s=z, out=out)

    def close(self):
        if self.handle:
            _cabi.load().lss_solver_destroy(self.handle)
            self.handle = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
