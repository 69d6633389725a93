"""torchrun worker: row-sharded fits over NCCL (one process per GPU) checked against the single-process
CPU oracle on the full data.  Launched by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import args_xy, assert_same_fit, binomial_problem, gaussian_problem, sparse_problem   # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import oem_b200
    from oem_b200.dist import Comm, LibComm, shard_csc_rows, shard_rows
    from oracle import oracle as orc
    # default: the in-library communicator (ncclAllReduce / one-shot NVLink peer kernel issued by the library itself);
    # OEMB200_TEST_COMM=callback exercises the host-callback fallback (torch.distributed.all_reduce)
    use_cb = os.environ.get("OEMB200_TEST_COMM") == "callback"
    comm = Comm() if use_cb else LibComm()
    if not use_cb:
        # the communicator on its own: small (peer-memory path when available) and large (NCCL) sums, in place
        for cnt in (1, 7, 1001, 8192, 8193, 1 << 20):
            t = torch.full((cnt,), float(rank + 1), dtype=torch.float64, device="cuda") + torch.arange(cnt, dtype=torch.float64, device="cuda")
            comm.all_reduce(t)
            torch.cuda.synchronize()
            want = world * (world + 1) / 2 + world * torch.arange(cnt, dtype=torch.float64, device="cuda")
            assert torch.equal(t, want), (cnt, t[:4], want[:4])

    def shard(a, X, y):
        r0, r1 = shard_rows(X.shape[0], rank, world)
        a = list(a)
        a[0], a[1] = np.asfortranarray(X[r0:r1]), y[r0:r1]
        return a, r0, r1

    # big.oem
    X, y = gaussian_problem(105, 9000, 70, mean_x=0.2)
    a = args_xy(X, y, "gaussian", ["lasso", "scad", "mcp"], gamma=[3.0, 3.7, 3.0], nlambda=30)
    ref = orc.oem_fit_big(*a) if rank == 0 else None
    sa, _, _ = shard(a, X, y)
    got = oem_b200.oem_fit_big(*sa, comm=comm)
    if rank == 0:
        assert_same_fit(got, ref)
    # oem (centred / scaled: two all-reduces)
    a = args_xy(X, y, "gaussian", ["lasso", "mcp"], nlambda=30, opts=dict(tol=1e-10))
    ref = orc.oem_fit_dense(*a) if rank == 0 else None
    sa, _, _ = shard(a, X, y)
    got = oem_b200.oem_fit_dense(*sa, comm=comm)
    if rank == 0:
        assert_same_fit(got, ref)
    # logistic: two-sweep route (p = 30) and slab route (p = 200), one (p+1)-vector all-reduce per data pass
    for nb, pb in ((6000, 30), (6001, 200)):
        Xb, yb = binomial_problem(104, nb, pb)
        a = args_xy(Xb, yb, "binomial", ["lasso"], nlambda=10, lmin_ratio=1e-2)
        ref = orc.oem_fit_logistic_dense(*a) if rank == 0 else None
        sa, _, _ = shard(a, Xb, yb)
        got = oem_b200.oem_fit_logistic_dense(*sa, comm=comm)
        assert got["stats"]["allreduce_calls"] >= got["stats"]["data_passes"] + 1
        if rank == 0:
            assert_same_fit(got, ref)
    # xval
    rng = np.random.default_rng(3)
    foldid = 1 + rng.permutation(9000) % 5
    a = args_xy(X, y, "gaussian", ["lasso", "mcp"], nlambda=20)
    xa = a[:17] + [5, foldid, False, "mse", a[18]]
    ref = orc.oem_xval_dense(*xa) if rank == 0 else None
    r0, r1 = shard_rows(9000, rank, world)
    sx = list(xa)
    sx[0], sx[1], sx[18] = np.asfortranarray(X[r0:r1]), y[r0:r1], foldid[r0:r1]
    got = oem_b200.oem_xval_dense(*sx, comm=comm)
    if rank == 0:
        assert_same_fit(got, ref)
        for pp in range(2):
            assert np.allclose(got["cvm"][pp], ref["cvm"][pp], rtol=1e-9)
            assert np.allclose(got["cvsd"][pp], ref["cvsd"][pp], rtol=1e-8)
    # sparse (dgCMatrix row blocks; the loss partials are a second all-reduce)
    Xs, ysp = sparse_problem(106, 8000, 50, density=0.06, shift_y=1.0)
    a = args_xy(Xs, ysp, "gaussian", ["lasso", "mcp"], nlambda=15, compute_loss=True)
    ref = orc.oem_fit_sparse(*a) if rank == 0 else None
    blk, r0, r1 = shard_csc_rows(Xs, rank, world)
    sa = list(a)
    sa[0], sa[1] = blk, ysp[r0:r1]
    got_s = oem_b200.oem_fit_sparse(*sa, comm=comm)
    if rank == 0:
        assert_same_fit(got_s, ref)
        for lg, lr in zip(got_s["loss"], ref["loss"]):
            assert np.allclose(lg[:len(lr)], lr, rtol=1e-10, atol=0)
    # every rank holds the same replicated result
    chk = torch.tensor([float(np.sum(got["beta"][0]))], dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item()
    if rank == 0:
        how = f"callback calls={comm.calls}" if use_cb else f"libcomm p2p={comm.p2p}"
        print(f"DIST_OK world={world} {how}")
    if not use_cb:
        comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
