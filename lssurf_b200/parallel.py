"""Row-sharded data parallelism over the GPUs of one node: one process per GPU.

The reference is single-process.  Here each rank owns a contiguous block of data
rows of G plus a 1/P share of the constraint rows; the grid unknowns are replicated.  The only
data-path collective is one NCCL all-reduce (fp64 sum) of A^T W (A p) plus the dot-product partials
per CG iteration, issued from the C++ solver loop through the communicator created here
(``lss_comm_create`` -> ncclCommInitRank, NVLink 5 / NVSwitch).  torch.distributed (any backend,
gloo on CPU works) is used ONLY as the host transport for the 128-byte NCCL unique id and for
gathering per-point outputs at the end.
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

    def allreduce_sum(self, t):
        """In-place fp64 sum over ranks of a CUDA tensor through the library's NCCL communicator."""
        if self.world == 1 or self.comm is None:
            return t
        t = t.contiguous()
        _cabi.check(_cabi.lib().lss_comm_allreduce_f64(self.comm, _cabi.ptr(t), int(t.numel()), _cabi.stream()))
        return t

    def destroy(self):
        if self.comm is not None:
            _cabi.load().lss_comm_destroy(self.comm)
            self.comm = None


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
