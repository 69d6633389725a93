"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous row shards, one sum all-reduce of
the packed sufficient-statistics bundle [Gram | column sums | X'y | sum x^2 | sum y | n], then the
replicated path -- the same sequence oemb200_fit_big runs per rank with NCCL.  The partial
bundles are computed by the oracle here (no GPU); the reduced bundle must reproduce the single-process fit."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import args_xy, gaussian_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oem_b200.dist import Comm, shard_rows
    from oracle import oracle as orc
    X, y = gaussian_problem(77, 5000, 12)
    n, p = X.shape
    r0, r1 = shard_rows(n, rank, world)
    Xs, ys = X[r0:r1], y[r0:r1]
    bundle = np.concatenate([(Xs.T @ Xs).ravel(order="F"), Xs.sum(0), Xs.T @ ys, (Xs ** 2).sum(0),
                             [ys.sum(), (ys ** 2).sum(), float(r1 - r0)]])
    comm = Comm()
    red = comm.allreduce_host(bundle.copy())
    G = red[:p * p].reshape(p, p, order="F")
    colsum, xty, sq = red[p * p:p * p + p], red[p * p + p:p * p + 2 * p], red[p * p + 2 * p:p * p + 3 * p]
    ysum, ntot = red[-3], red[-1]
    # assemble like oem_big and run the path through oem_xtx semantics on the augmented Gram
    w = 1.0 / np.sqrt(sq / (ntot - 1.0))
    XX = np.zeros((p + 1, p + 1))
    XX[1:, 1:] = w[:, None] * G * w[None, :]
    XX[0, 1:] = XX[1:, 0] = colsum * w
    XX[0, 0] = ntot
    XX /= ntot
    XY = np.concatenate([[ysum], xty * w]) / ntot
    q.put((rank, r0, r1, XX, XY, comm.world))
    dist.destroy_process_group()


def test_row_sharded_bundle_allreduce_reproduces_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (_, a0, a1, XX0, XY0, w0), (_, b0, b1, XX1, XY1, _) = outs
    assert w0 == 2 and a0 == 0 and a1 == b0 and b1 == 5000 and a1 % 72 == 0
    assert np.array_equal(XX0, XX1) and np.array_equal(XY0, XY1)          # replicated result
    X, y = gaussian_problem(77, 5000, 12)
    n, p = X.shape
    w = 1.0 / np.sqrt((X ** 2).sum(0) / (n - 1.0))
    XXf = np.zeros((p + 1, p + 1))
    XXf[1:, 1:] = w[:, None] * (X.T @ X) * w[None, :]
    XXf[0, 1:] = XXf[1:, 0] = X.sum(0) * w
    XXf[0, 0] = n
    XXf /= n
    assert np.allclose(XX0, XXf, rtol=1e-13, atol=1e-15)
    assert np.allclose(XY0, np.concatenate([[y.sum()], (X.T @ y) * w]) / n, rtol=1e-12)


def test_shard_rows_cover_and_align():
    from oem_b200.dist import shard_rows
    for n, world in [(100_000_000, 8), (12_500_000, 1), (5001, 2), (1000, 4), (71, 2)]:
        blocks = [shard_rows(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a1 > a0
        if n // world >= 72:
            assert all(b[0] % 72 == 0 for b in blocks)


def test_shard_csc_rows_partition_the_statistics():
    """Sparse row blocks: the per-rank dgCMatrix pieces tile the matrix and their sufficient statistics add up to the
    whole -- what the single all-reduce of oem_fit_sparse's bundle relies on."""
    import scipy.sparse as sps
    from cases import sparse_problem
    from oem_b200.dist import shard_csc_rows
    X, y = sparse_problem(8, 1003, 11, density=0.2)
    G, xty, nnz, rows = np.zeros((11, 11)), np.zeros(11), 0, 0
    for rank in range(3):
        blk, r0, r1 = shard_csc_rows(X, rank, 3)
        assert sps.isspmatrix_csc(blk) and blk.has_sorted_indices and blk.shape == (r1 - r0, 11)
        assert np.array_equal(blk.toarray(), X.toarray()[r0:r1])
        G += (blk.T @ blk).toarray()
        xty += blk.T @ y[r0:r1]
        nnz += blk.nnz
        rows += r1 - r0
    assert rows == 1003 and nnz == X.nnz
    assert np.allclose(G, (X.T @ X).toarray(), rtol=1e-13, atol=1e-13) and np.allclose(xty, X.T @ y, rtol=1e-13)
