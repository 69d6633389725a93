"""Row-sharded multi-process plumbing: one process per GPU, torch.distributed for the rendezvous,
NVLink / NVSwitch for the one exchange step of the path -- the sum all-reduce of the packed sufficient
statistics [Gram | X'y | column sums | counts], and of the (p+1)-vector gradient once per IRLS iteration
of the logistic entry.

Two ways to give the library its all-reduce (include/oem_b200.h):
  * `LibComm` (default for GPU runs): the library owns the communicator (oemb200_comm_create) and issues the
    collectives itself on its own stream -- ncclAllReduce from the libnccl already loaded in the process for the
    big bundle, a one-shot NVLink peer-memory kernel for small vectors.  torch.distributed only carries the
    128-byte NCCL unique id at start-up.
  * `Comm`: a host callback (`oemb200_allreduce_fn`) bound to torch.distributed.all_reduce; kept as the fallback
    and for the CPU (gloo) tests of the sharding logic."""
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


class LibComm:
    """In-library communicator (oemb200_comm_*): no Python in the all-reduce path.  Collective constructor: every rank
    of `group` must create it at the same time (ncclCommInitRank + the IPC exchange of the peer mailboxes)."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        from . import api
        L = api.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.handle = None
        self._L = L
        if self.world <= 1:
            return
        uid = np.zeros(128, dtype=np.uint8)
        on_gpu = dist.get_backend(group) == "nccl"
        # every rank asks for an id (rank 0's is the one that is used): this proves libnccl can be loaded HERE before any
        # rank enters a collective it could hang in; the ranks then agree on the outcome
        err = None
        try:
            api._check(L.oemb200_comm_unique_id(uid.ctypes.data))
        except api.OemB200Error as e:
            err = e
        ok = torch.tensor([0.0 if err else 1.0], dtype=torch.float64)
        if on_gpu:
            ok = ok.cuda(self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) < 0.5:
            raise RuntimeError(f"in-library communicator unavailable on at least one rank ({err or 'another rank failed'}); "
                               "use oem_b200.dist.Comm (host-callback all-reduce) instead")
        t = torch.from_numpy(uid)
        if on_gpu:
            t = t.cuda(self.device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = t.cpu().numpy().copy()
        h = ctypes.c_void_p()
        api._check(L.oemb200_comm_create(uid.ctypes.data, self.rank, self.world, self.device, ctypes.byref(h)))
        self.handle = h.value

    @property
    def p2p(self):
        return bool(self.handle) and bool(self._L.oemb200_comm_p2p_enabled(self.handle))

    def all_reduce(self, tensor, stream=None, timed=False):
        """In-place sum of a float64 CUDA tensor through the library's communicator; returns microseconds if timed."""
        from . import api
        us = ctypes.c_double(0.0)
        api._check(self._L.oemb200_comm_allreduce(self.handle, tensor.data_ptr(), tensor.numel(),
                                                  int(stream) if stream else None, ctypes.byref(us) if timed else None))
        return us.value if timed else None

    def close(self):
        if self.handle:
            self._L.oemb200_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArrayView:
    """__cuda_array_interface__ shim so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _wrap_device_f64(ptr, count):
    import torch
    return torch.as_tensor(_CudaArrayView(ptr, count), device="cuda")
