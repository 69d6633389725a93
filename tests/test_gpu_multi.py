"""-m gpu: the row-sharded N>1 path over NCCL, two processes on two GPUs of one box (skipped on a 1-GPU box;
the host-side sharding logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["libcomm", "callback"])
def test_two_rank_nccl_parity(lib, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731" if transport == "libcomm" else "29732", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    env = dict(os.environ, OEMB200_TEST_COMM=transport)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_OK world=2" in out.stdout


def test_libcomm_single_gpu_world_of_one(lib):
    # world = 1: the communicator degenerates to a no-op and the fit equals the plain single-process call
    import torch
    import torch.distributed as dist
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import args_xy, gaussian_problem
    from oem_b200.dist import LibComm
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29741", rank=0, world_size=1)
    try:
        comm = LibComm()
        X, y = gaussian_problem(5, 3000, 20)
        a = args_xy(X, y, "gaussian", ["lasso"], nlambda=10)
        g1, g2 = lib.oem_fit_big(*a, comm=comm), lib.oem_fit_big(*a)
        assert np.array_equal(g1["beta"][0], g2["beta"][0]) and g1["stats"]["allreduce_calls"] == 0
    finally:
        if own:
            dist.destroy_process_group()
