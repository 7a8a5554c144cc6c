"""ctypes binding of liblssurf_b200.so (the C ABI declared in include/lssurf_b200.h).

PyTorch is used only as the owner of device memory and streams: every call below takes raw device
pointers (``tensor.data_ptr()``) and the current CUDA stream.  There is no CPU fallback: if the
shared library is missing, or no CUDA device is visible, the compute entry points raise.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblssurf_b200.so')
_lib = None

LSS_MAX_DIMS = 3
SPMV_PLAIN, SPMV_SCALE, SPMV_RESIDUAL = 0, 1, 2


class LssError(RuntimeError):
    pass


class lss_grid_t(C.Structure):
    _fields_ = [('ndim', C.c_int32), ('shape', C.c_int64 * 3), ('stride', C.c_int64 * 3), ('lo', C.c_double * 3),
                ('hi', C.c_double * 3), ('delta', C.c_double * 3), ('col_0', C.c_int64)]


class lss_structured_t(C.Structure):
    """Tables of the grid-structured normal-matrix layout (include/lssurf_b200.h; built by lssurf_b200/structured.py)."""
    _fields_ = [(k, C.c_int32) for k in ('nt', 'nn', 'n0', 'n_diag', 'dd_stride', 'node_len', 'rec_len', 'max_pairs', 'n_zd', 'rem_n_rows')] + \
               [(k, C.c_int64) for k in ('z0_col0', 'dz_col0', 'rows', 'nnz', 'n_vals', 'zd_base', 'zz_base', 'rem_nnz')] + \
               [(k, C.c_void_p) for k in ('map', 'dd_b', 'zd_ptr', 'zd_k', 'zz_off', 'rem_src', 'rem_rows', 'rem_indptr', 'rem_indices')]


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pi64, _pi32, _pf64 = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)

# name -> argtypes; every function returns int except where noted.  This table is also what the CPU
# test suite checks against the header (tests/test_cabi_symbols.py).
SIGNATURES = {
    'lss_version': [],
    'lss_device_count': [],
    'lss_launch_count': [],
    'lss_interp_coo': [C.POINTER(lss_grid_t), C.POINTER(_vp), _i64, _i64, _vp, _vp, _vp, _vp, _pi64, _vp],
    'lss_stencil_count': [C.POINTER(lss_grid_t), _i32, _pi32, _i32, _pi64, _pi64, _pi64],
    'lss_stencil_coo': [C.POINTER(lss_grid_t), _i32, _pi32, _pf64, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    'lss_coo_extent': [_i64, _vp, _vp, _pi64, _pi64, _vp],
    'lss_coo_to_csr': [_i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _pi64, _vp],
    'lss_blocks_to_csr': [_i64, _i64, _i32, _pi32, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _pi64, _pi64, _vp],
    'lss_csr_transpose': [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    'lss_cell_keys': [C.POINTER(lss_grid_t), C.POINTER(lss_grid_t), C.POINTER(_vp), _i64, _vp, _pi32, _vp],
    'lss_argsort_u64': [_i64, _vp, _i32, _vp, _vp],
    'lss_dense_rank_f64': [_vp, _i64, _f64, _vp, _pf64, _i64, _pi64, _vp],
    'lss_group_median': [_vp, _vp, _i64, _i32, _vp, _vp],
    'lss_spmv': [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    'lss_csr_rmatvec': [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    'lss_spmv_plan_create': [C.POINTER(_vp), _i64, _vp, _vp],
    'lss_spmv_plan_destroy': [_vp],
    'lss_spmv_planned': [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    'lss_set_option': [C.c_char_p, _i32],
    'lss_comm_unique_id': [C.c_char_p],
    'lss_comm_create': [C.POINTER(_vp), C.c_char_p, _i32, _i32],
    'lss_comm_destroy': [_vp],
    'lss_comm_allreduce_f64': [_vp, _vp, _i64, _vp],
    'lss_comm_box_alloc': [_vp, _i64, C.c_char_p],
    'lss_comm_box_open': [_vp, C.c_char_p],
    'lss_comm_box_bytes': [_vp, _pi64],
    'lss_solver_create': [C.POINTER(_vp), _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    'lss_solver_destroy': [_vp],
    'lss_solver_create_blocks': [C.POINTER(_vp), _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    'lss_solver_create_sharded': [C.POINTER(_vp), _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _vp, _vp],
    'lss_solver_set_weights': [_vp, _vp],
    'lss_solver_solve': [_vp, _vp, _f64, _i32, _pi32, _pf64],
    'lss_solver_set_profile': [_vp, _i32],
    'lss_solver_stats': [_vp, _pf64],
    'lss_solver_info': [_vp, _pi64],
    'lss_solver_set_structure': [_vp, _i32, _vp, _vp, _i32, _i64, _i64],
    'lss_solver_set_coarse': [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32],
    'lss_solver_export_pattern': [_vp, _vp, _vp],
    'lss_solver_set_structured': [_vp, C.POINTER(lss_structured_t)],
    'lss_normal_create': [C.POINTER(_vp), _i64, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _pi64, _vp],
    'lss_normal_update': [_vp, _vp],
    'lss_normal_export': [_vp, _vp, _vp, _vp],
    'lss_normal_apply': [_vp, _vp, _vp],
    'lss_normal_destroy': [_vp],
    'lss_rde': [_vp, _vp, _vp, _f64, _i64, _vp, _pf64, _pi64, _vp],
    'lss_sigma_extra': [_vp, _vp, _vp, _i64, _vp, _pf64, _pi32, _vp],
    'lss_edit_mask': [_vp, _vp, _vp, _i64, _f64, _vp, _pf64, _vp, _vp],
    'lss_data_weights': [_vp, _vp, _vp, _i64, _vp, _vp],
}


def load():
    """Load the shared library (no CUDA call is made).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LssError(f'{LIB_PATH} not found: build it with `python -m lssurf_b200.csrc.build` '
                       '(there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.lss_launch_count.restype = C.c_int64
    lib.lss_last_error.argtypes = []
    lib.lss_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def lib():
    """The library, after checking that a CUDA device is usable.  No silent CPU path."""
    L = load()
    if not torch.cuda.is_available() or L.lss_device_count() < 1:
        raise LssError('lssurf_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists')
    return L


def check(rc):
    if rc != 0:
        msg = load().lss_last_error().decode(errors='replace')
        raise LssError(f'lssurf_b200 error {rc}: {msg}')


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return C.c_void_p(t.data_ptr())


def device():
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(a, dtype=None):
    """numpy / tensor -> contiguous CUDA tensor (tensor plumbing only)."""
    if isinstance(a, torch.Tensor):
        t = a.to(device())
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device())
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def true_div(t, s):
    """``t / s`` for a Python scalar ``s`` with IEEE division on every element.  torch's CUDA ``tensor / scalar``
    multiplies by the rounded reciprocal instead (last-bit differences from NumPy); a 0-dim device divisor takes the
    exact path, which is what bit-exact parity with the reference's NumPy arithmetic needs."""
    return torch.div(t, torch.tensor(s, dtype=t.dtype, device=t.device))


class _PinnedPool(object):
    """Page-locked host blocks handed out as NumPy arrays and taken back when the array is garbage collected.
    Two blocks are created the first time a size is asked for: a caller that still holds the previous fit's outputs
    while the next fit runs (the usual pattern) then never waits for cudaHostAlloc again."""

    def __init__(self):
        self.free = {}

    def take(self, nbytes):
        import weakref
        key = (nbytes + (1 << 20) - 1) >> 20
        blocks = self.free.setdefault(key, None)
        if blocks is None:
            blocks = self.free[key] = [torch.empty(key << 20, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        buf = blocks.pop() if blocks else torch.empty(key << 20, dtype=torch.uint8, pin_memory=True)
        return buf, (lambda b=buf, bl=blocks: bl.append(b))


_pool = _PinnedPool()


def to_host(t):
    """CUDA tensor -> NumPy array owned by the caller.  Large tensors are copied by DMA into a page-locked block of
    their own (no second host copy); the block goes back to a pool when the array is released."""
    import weakref
    if not t.is_cuda:
        return t.numpy().copy()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20):
        return t.cpu().numpy()
    buf, give_back = _pool.take(nbytes)
    view = buf[:nbytes].view(t.dtype).reshape(t.shape)
    view.copy_(t, non_blocking=False)
    arr = view.numpy()
    weakref.finalize(arr, give_back)
    return arr


def grid_struct(shape, stride, lo, hi, delta, col_0):
    g = lss_grid_t()
    nd = len(shape)
    g.ndim = nd
    for d in range(3):
        g.shape[d] = int(shape[d]) if d < nd else 1
        g.stride[d] = int(stride[d]) if d < nd else 0
        g.lo[d] = float(lo[d]) if d < nd else 0.0
        g.hi[d] = float(hi[d]) if d < nd else 0.0
        g.delta[d] = float(delta[d]) if d < nd else 1.0
    g.col_0 = int(col_0)
    return g


# ---------------------------------------------------------------------------------------------
# thin functional wrappers (device tensors in, device tensors out)
# ---------------------------------------------------------------------------------------------

def interp_coo(g, pts, row_0=0):
    """-> rows[int32 n,2^d], cols[int32 n,2^d], vals[f64 n,2^d], valid[uint8 n], n_invalid"""
    L = lib()
    nd = g.ndim
    n = int(pts[0].numel())
    nc = 1 << nd
    dev = device()
    rows = torch.empty((n, nc), dtype=torch.int32, device=dev)
    cols = torch.empty((n, nc), dtype=torch.int32, device=dev)
    vals = torch.empty((n, nc), dtype=torch.float64, device=dev)
    valid = torch.empty(n, dtype=torch.uint8, device=dev)
    parr = (C.c_void_p * nd)(*[p.data_ptr() for p in pts[:nd]])
    bad = C.c_int64(0)
    check(L.lss_interp_coo(C.byref(g), parr, n, row_0, ptr(rows), ptr(cols), ptr(vals), ptr(valid), C.byref(bad), stream()))
    return rows, cols, vals, valid, bad.value


def stencil_coo(g, offsets, coeffs, row_0=0, valid_only=True):
    L = lib()
    offs = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(g.ndim, -1))
    n_off = offs.shape[1]
    coef = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64).ravel())
    if coef.size != n_off:
        raise ValueError('one coefficient per offset is required')
    n_eq = C.c_int64(0)
    check(L.lss_stencil_count(C.byref(g), n_off, offs.ctypes.data_as(_pi32), int(valid_only), C.byref(n_eq), None, None))
    n_eq = n_eq.value
    dev = device()
    rows = torch.empty((n_eq, n_off), dtype=torch.int32, device=dev)
    cols = torch.empty((n_eq, n_off), dtype=torch.int32, device=dev)
    vals = torch.empty((n_eq, n_off), dtype=torch.float64, device=dev)
    ind0 = torch.empty(n_eq, dtype=torch.int64, device=dev)
    check(L.lss_stencil_coo(C.byref(g), n_off, offs.ctypes.data_as(_pi32), coef.ctypes.data_as(_pf64), row_0,
                            int(valid_only), n_eq, ptr(rows), ptr(cols), ptr(vals), ptr(ind0), stream()))
    return rows, cols, vals, ind0


def coo_extent(rows, vals):
    L = lib()
    mx, cnt = C.c_int64(-1), C.c_int64(0)
    check(L.lss_coo_extent(int(vals.numel()), ptr(rows), ptr(vals), C.byref(mx), C.byref(cnt), stream()))
    return mx.value, cnt.value


def coo_to_csr(rows, cols, vals, n_rows, n_cols):
    """Canonical CSR (device tensors indptr[int32], indices[int32], data[f64])."""
    L = lib()
    nnz = int(vals.numel())
    dev = device()
    indptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
    out = C.c_int64(0)
    check(L.lss_coo_to_csr(nnz, ptr(rows), ptr(cols), ptr(vals), n_rows, n_cols, ptr(indptr), ptr(indices), ptr(data),
                           C.byref(out), stream()))
    k = out.value
    return indptr, indices[:k].clone() if k < nnz else indices[:k], data[:k].clone() if k < nnz else data[:k]


def blocks_to_csr(blocks, n_rows, n_cols):
    """Canonical CSR of row-regular COO blocks [(cols[n_rows, k], vals[n_rows, k]), ...] without a sort
    (``lss_blocks_to_csr``); None when some row is not in ascending column order (the caller sorts instead)."""
    L = lib()
    nb = len(blocks)
    ks = (C.c_int32 * nb)(*[int(c.shape[1]) for c, _ in blocks])
    cols = (_vp * nb)(*[c.data_ptr() for c, _ in blocks])
    vals = (_vp * nb)(*[v.data_ptr() for _, v in blocks])
    cap = int(n_rows) * sum(int(c.shape[1]) for c, _ in blocks)
    dev = device()
    indptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    data = torch.empty(max(cap, 1), dtype=torch.float64, device=dev)
    nnz, bad = C.c_int64(0), C.c_int64(0)
    check(L.lss_blocks_to_csr(int(n_rows), int(n_cols), nb, ks, cols, vals, ptr(indptr), ptr(indices), ptr(data), C.byref(nnz),
                              C.byref(bad), stream()))
    if bad.value:
        return None
    k = nnz.value
    return indptr, (indices[:k].clone() if k < cap else indices), (data[:k].clone() if k < cap else data)


def csr_transpose(indptr, indices, data, n_rows, n_cols):
    L = lib()
    nnz = int(data.numel())
    dev = device()
    t_indptr = torch.empty(n_cols + 1, dtype=torch.int32, device=dev)
    t_indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    t_data = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
    check(L.lss_csr_transpose(n_rows, n_cols, nnz, ptr(indptr), ptr(indices), ptr(data), ptr(t_indptr), ptr(t_indices),
                              ptr(t_data), stream()))
    return t_indptr, t_indices[:nnz], t_data[:nnz]


def spatial_order(fine, coarse, pts, return_keys=False):
    """Permutation (int64 CUDA tensor) that sorts points by (coarse cell, fine cell, time cell)."""
    L = lib()
    n = int(pts[0].numel())
    keys = torch.empty(n, dtype=torch.int64, device=device())
    nd = coarse.ndim
    parr = (C.c_void_p * nd)(*[p.data_ptr() for p in pts[:nd]])
    bits = C.c_int32(0)
    check(L.lss_cell_keys(C.byref(fine), C.byref(coarse), parr, n, ptr(keys), C.byref(bits), stream()))
    order = argsort_u64(keys, bits.value).to(torch.int64)
    if return_keys:
        return order, keys
    return order


def argsort_u64(keys, key_bits):
    L = lib()
    n = int(keys.numel())
    order = torch.empty(n, dtype=torch.int32, device=device())
    check(L.lss_argsort_u64(n, ptr(keys), int(key_bits), ptr(order), stream()))
    return order


def dense_rank(x, nan_value=-9999.0, max_unique=1 << 20):
    """-> (rank int32 tensor, ascending unique values as NumPy)   [np.unique(x, return_inverse=True)]"""
    L = lib()
    n = int(x.numel())
    rank = torch.empty(n, dtype=torch.int32, device=device())
    cap = min(max(n, 1), max_unique)
    uniq = np.empty(cap, dtype=np.float64)
    nu = C.c_int64(0)
    check(L.lss_dense_rank_f64(ptr(x), n, float(nan_value), ptr(rank), uniq.ctypes.data_as(_pf64), cap, C.byref(nu), stream()))
    return rank, uniq[:nu.value].copy()


def group_median(ids, vals, n_groups):
    L = lib()
    out = torch.empty(int(n_groups), dtype=torch.float64, device=device())
    check(L.lss_group_median(ptr(ids), ptr(vals), int(vals.numel()), int(n_groups), ptr(out), stream()))
    return out


def set_option(name, value):
    check(load().lss_set_option(name.encode(), int(value)))


def launch_count():
    return int(load().lss_launch_count())


def spmv(indptr, indices, data, x, n_rows, n_cols, mode=SPMV_PLAIN, s=None, out=None):
    L = lib()
    y = out if out is not None else torch.empty(n_rows, dtype=torch.float64, device=device())
    check(L.lss_spmv(n_rows, n_cols, ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y), mode, ptr(s), stream()))
    return y


def csr_rmatvec(indptr, indices, data, u, n_rows, n_cols):
    """q = A^T u from the CSR rows of A (scatter with atomics)."""
    q = torch.empty(n_cols, dtype=torch.float64, device=device())
    check(lib().lss_csr_rmatvec(n_rows, n_cols, ptr(indptr), ptr(indices), ptr(data), ptr(u), ptr(q), stream()))
    return q


class SpmvPlan(object):
    """Owner of an lss_spmv_plan (structure-dependent set-up of the tiled CSR product)."""

    def __init__(self, indptr, n_rows):
        self.h = C.c_void_p(0)
        check(lib().lss_spmv_plan_create(C.byref(self.h), int(n_rows), ptr(indptr), stream()))

    def __del__(self):
        try:
            if self.h:
                load().lss_spmv_plan_destroy(self.h)
                self.h = C.c_void_p(0)
        except Exception:
            pass


def spmv_planned(plan, indptr, indices, data, x, n_rows, mode=SPMV_PLAIN, s=None, out=None):
    L = lib()
    y = out if out is not None else torch.empty(n_rows, dtype=torch.float64, device=device())
    check(L.lss_spmv_planned(plan.h, n_rows, ptr(indptr), ptr(indices), ptr(data), ptr(x), ptr(y), mode, ptr(s), stream()))
    return y


def rde(r, sigma=None, mask=None, s1=0.0, comm=None):
    L = lib()
    out, n_used = C.c_double(0.0), C.c_int64(0)
    check(L.lss_rde(ptr(r), ptr(sigma), ptr(mask), float(s1), int(r.numel()), comm, C.byref(out), C.byref(n_used), stream()))
    return out.value, n_used.value


def sigma_extra(r, sigma, mask, comm=None):
    L = lib()
    out, nfev = C.c_double(0.0), C.c_int32(0)
    check(L.lss_sigma_extra(ptr(r), ptr(sigma), ptr(mask), int(r.numel()), comm, C.byref(out), C.byref(nfev), stream()))
    return out.value, nfev.value


def edit_mask(r, sigma, sigma_extra_t, n_sigma=3.0, comm=None, out=None):
    L = lib()
    n = int(r.numel())
    m = out if out is not None else torch.empty(n, dtype=torch.uint8, device=device())
    stats = (C.c_double * 3)()
    check(L.lss_edit_mask(ptr(r), ptr(sigma), ptr(sigma_extra_t), n, float(n_sigma), ptr(m), stats, comm, stream()))
    return m, dict(count=int(stats[0]), min_sigma=float(stats[1]), near_threshold=int(stats[2]))


def data_weights(sigma, sigma_extra_t, mask, out):
    L = lib()
    check(L.lss_data_weights(ptr(sigma), ptr(sigma_extra_t), ptr(mask), int(sigma.numel()), ptr(out), stream()))
    return out
