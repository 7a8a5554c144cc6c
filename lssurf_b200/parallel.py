"""Data parallelism over the GPUs of one node: one process per GPU.

The reference is single-process.  Here each rank owns a contiguous block of the spatially sorted
data rows: it assembles its rows of G and does all per-point work (residuals, order statistics, edit masks) on them.
Three ways to run the solve (smooth_fit.iterate_fit_device picks):
  * row-sharded normal matrix (default with more than one rank on the z0 + dz grids): every rank owns a range of
    preconditioner blocks -- its rows of N = A^T W A, built from its own points plus halo copies -- and the ranks exchange
    p.q, their slice of z, the coarse restriction and the bias coupling from inside the CG step kernel through a
    peer-to-peer mailbox (``ensure_box``: one device block per rank, CUDA-IPC handles opened on every peer, plain stores over
    NVLink); no collective inside the iteration;
  * replicated solve (``LSS_REPLICATED_SOLVE=1``): the assembled rows are all-gathered once (NCCL, ~2 GB at 10^7 points) and
    every rank runs the same PCG; the solution is broadcast from rank 0 after every solve so all ranks continue from
    identical bits;
  * all-reduce solve (``LSS_SHARDED_SOLVE=1`` with the replicated mode, or rows too long for the normal-matrix forms): each
    rank keeps its rows plus a 1/P share of the constraint rows and the products are summed with one NCCL all-reduce
    (fp64) per CG iteration, issued from the C++ solver loop through the communicator created here.
torch.distributed is plumbing only: the host transport for the 128-byte NCCL unique id and the IPC handles (any backend,
gloo on CPU works), the all-gathers of masks / per-point outputs and the broadcast of the solution.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi


class DistContext(object):
    """rank / world size + the NCCL communicator handle passed to the C ABI (None on one GPU)."""

    def __init__(self, rank=0, world=1, comm=None):
        self.rank, self.world, self.comm = rank, world, comm

    def shard(self, n):
        """[lo, hi) of this rank's contiguous block of n items."""
        return (self.rank * n) // self.world, ((self.rank + 1) * n) // self.world

    def gather_rows(self, arr):
        """Concatenate per-rank 1-D NumPy arrays in rank order (host transport)."""
        if self.world == 1:
            return arr
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, arr)
        return np.concatenate(parts)

    def gather_rows_device(self, t, n_total):
        """Concatenate the per-rank row blocks of a 1-D CUDA tensor in rank order on every rank (NCCL all-gather
        of equal-size padded blocks; the block sizes follow from shard()).  Falls back to the host transport
        when torch.distributed is not running on NCCL."""
        if self.world == 1:
            return t
        import torch.distributed as dist
        if dist.get_backend() != 'nccl':
            return _cabi.to_dev(self.gather_rows(_cabi.to_host(t)))
        sizes = [((r + 1) * n_total) // self.world - (r * n_total) // self.world for r in range(self.world)]
        m = max(sizes)
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        out = torch.empty(m * self.world, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[i * m:i * m + sizes[i]] for i in range(self.world)])

    def gather_counts(self, n):
        """Every rank's value of the integer n, in rank order (host transport)."""
        if self.world == 1:
            return [int(n)]
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, int(n))
        return [int(v) for v in parts]

    def gather_var_device(self, t, counts):
        """Concatenate 1-D CUDA tensors of different lengths (counts[r] on rank r) in rank order on every rank."""
        if self.world == 1:
            return t
        import torch.distributed as dist
        if dist.get_backend() != 'nccl':        # host transport (CPU tests under gloo)
            return torch.from_numpy(self.gather_rows(t.cpu().numpy())).to(t.device)
        m = max(max(counts), 1)
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        out = torch.empty(m * self.world, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[i * m:i * m + counts[i]] for i in range(self.world)])

    def gather_csr(self, A):
        """Stack the per-rank row blocks of a DeviceCSR in rank order on every rank (NCCL all-gather over NVLink)."""
        from .lin_op import DeviceCSR
        if self.world == 1:
            return A
        rows = self.gather_counts(A.shape[0])
        nnzs = self.gather_counts(A.nnz)
        lens = self.gather_var_device((A.indptr[1:] - A.indptr[:-1]).contiguous(), rows)
        indptr = torch.zeros(sum(rows) + 1, dtype=torch.int32, device=lens.device)
        indptr[1:] = torch.cumsum(lens, 0)
        return DeviceCSR(indptr, self.gather_var_device(A.indices, nnzs), self.gather_var_device(A.data, nnzs), (sum(rows), A.shape[1]))

    def broadcast(self, t, src=0):
        """In-place broadcast of a CUDA tensor from rank src (keeps replicated solves bit-identical across ranks)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.broadcast(t, src=src)
        return t

    def allreduce_sum(self, t):
        """In-place fp64 sum over ranks of a CUDA tensor through the library's NCCL communicator."""
        if self.world == 1 or self.comm is None:
            return t
        t = t.contiguous()
        _cabi.check(_cabi.lib().lss_comm_allreduce_f64(self.comm, _cabi.ptr(t), int(t.numel()), _cabi.stream()))
        return t

    def allreduce_min(self, t):
        """Minimum over ranks of a small fp64 CUDA tensor (sum all-reduce of one slot per rank)."""
        if self.world == 1:
            return t
        slots = torch.zeros((self.world,) + tuple(t.shape), dtype=torch.float64, device=t.device)
        slots[self.rank] = t
        return self.allreduce_sum(slots).amin(dim=0)

    def ensure_box(self, nbytes):
        """Peer-to-peer mailbox of at least nbytes on every rank (``lss_comm_box_*``): allocated once per communicator,
        CUDA IPC handles exchanged through torch.distributed (host transport), mapped into every rank."""
        if self.comm is None:
            raise _cabi.LssError('the peer-to-peer mailbox needs a communicator')
        L = _cabi.lib()
        have = C.c_int64(0)
        _cabi.check(L.lss_comm_box_bytes(self.comm, C.byref(have)))
        if have.value >= nbytes:
            return
        if have.value > 0:
            raise _cabi.LssError(f'mailbox of {have.value} bytes is too small ({nbytes} needed); set LSS_BOX_BYTES before the first fit')
        size = max(int(nbytes), int(os.environ.get('LSS_BOX_BYTES', str(32 << 20))))
        if self.world > 1:
            import torch.distributed as dist
            sizes = [None] * self.world
            dist.all_gather_object(sizes, size)
            size = max(sizes)
        buf = C.create_string_buffer(64)
        _cabi.check(L.lss_comm_box_alloc(self.comm, size, buf))
        handles = None
        if self.world > 1:
            import torch.distributed as dist
            parts = [None] * self.world
            dist.all_gather_object(parts, bytes(buf.raw))
            handles = b''.join(parts)
        _cabi.check(L.lss_comm_box_open(self.comm, handles))
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def gather_ints(self, values):
        """Every rank's list of integers, as a list of lists in rank order (host transport)."""
        if self.world == 1:
            return [list(values)]
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, [int(v) for v in values])
        return parts

    def destroy(self):
        if self.comm is not None:
            _cabi.load().lss_comm_destroy(self.comm)
            self.comm = None


def single_rank_context():
    """A one-rank context WITH a communicator handle: lets the row-sharded solver (mailbox exchange with itself) run
    on a single GPU, which is how its kernels are tested where only one device is available."""
    handle = C.c_void_p(0)
    _cabi.check(_cabi.lib().lss_comm_create(C.byref(handle), None, 0, 1))
    return DistContext(0, 1, handle)


def init_from_env():
    """Create the context from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world == 1:
        return DistContext()
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(backend='nccl' if torch.cuda.is_available() else 'gloo')
    L = _cabi.lib()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _cabi.check(L.lss_comm_unique_id(buf))
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0)
    handle = C.c_void_p(0)
    _cabi.check(L.lss_comm_create(C.byref(handle), box[0], rank, world))
    return DistContext(rank, world, handle)
