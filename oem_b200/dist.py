"""Row-sharded multi-process plumbing: one process per GPU, torch.distributed for the rendezvous,
NCCL (over NVLink / NVSwitch) for the one exchange step of the path -- the sum all-reduce of the
packed sufficient statistics [Gram | X'y | column sums | counts].

The C library never links NCCL; it calls back into `Comm.callback` with a device pointer, a count
and the CUDA stream the data is ordered on (include/oem_b200.h: oemb200_allreduce_fn).  An R / C++
host would pass a thin wrapper around ncclAllReduce instead."""
import ctypes

import numpy as np

from .api import ALLREDUCE_FN


def shard_rows(n, rank, world, align=72):
    """Contiguous row block of `rank` among `world`: blocks are multiples of
    `align` rows (2 Gram k-tiles) except the last, like the OpenMP slices of
    src/oem_dense.h:328-358 (floor(n/ncores) rows each, last takes the remainder)."""
    per = (n // world) // align * align
    if per == 0:
        per = n // world
    r0 = rank * per
    r1 = n if rank == world - 1 else r0 + per
    return r0, r1


def shard_csc_rows(x, rank, world, align=72):
    """Row block of `rank` of a sparse design as its own dgCMatrix (scipy CSC with sorted indices), the input
    oem_fit_sparse expects on every rank of a row-sharded run; same blocks as shard_rows."""
    import scipy.sparse as sps
    r0, r1 = shard_rows(x.shape[0], rank, world, align)
    blk = sps.csc_matrix(sps.csr_matrix(x)[r0:r1])
    blk.sum_duplicates()
    blk.sort_indices()
    return blk, r0, r1


class Comm:
    """All-reduce callback over torch.distributed (NCCL for CUDA buffers, gloo for host buffers)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.calls = 0
        self.doubles = 0
        self.callback = ALLREDUCE_FN(self._allreduce)

    def _allreduce(self, buf, count, stream, ctx):
        try:
            import torch
            self.calls += 1
            self.doubles += int(count)
            # 0 / 1 / 2 are the CUDA handles of the null, legacy-default and per-thread default streams
            sid = int(stream) if stream else 0
            ext = torch.cuda.default_stream() if sid in (0, 1) else torch.cuda.ExternalStream(sid)
            with torch.cuda.stream(ext):
                t = _wrap_device_f64(buf, int(count))
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            # NCCL work is enqueued on torch's communication stream; make the library stream wait for it
            ext.synchronize()
            return 0
        except Exception as e:      # never raise through the C ABI
            import sys
            print(f"[oem_b200.dist] all-reduce failed: {e!r}", file=sys.stderr)
            return 1

    def allreduce_host(self, arr):
        """Sum all-reduce of a host numpy array (gloo); used by the CPU tests of the sharding logic."""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.numpy()


class _CudaArrayView:
    """__cuda_array_interface__ shim so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _wrap_device_f64(ptr, count):
    import torch
    return torch.as_tensor(_CudaArrayView(ptr, count), device="cuda")
