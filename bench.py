#!/usr/bin/env python
"""bench.py -- headline benchmark of the OEM hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[4], "big.oem-scale tall data n=1e8 p=1000 FP64 ... at 1/2/4/8 B200"):
one STEP = one full big.oem fit -- the one-pass Gram / X'y / column-sum build over the rank's row
shard (n = 1.25e7 x p = 1000 FP64 = 100 GB per GPU), the NCCL all-reduce of the packed sufficient
statistics, the on-device Lanczos top eigenvalue and the warm-started 100-lambda path for lasso +
SCAD + MCP batched in one call.  Weak scaling: every rank holds 1.25e7 rows, so --gpus 8 is exactly
configs[4] (n = 1e8).  metric = full lambda-path fit time (s), lower is better.

  value      device-resident inputs (X, y already in HBM when the timed region starts)
  e2e        the same fit through the C ABI with HOST buffers (pinned); the library streams row chunks
             host->device inside the timed region and returns beta on the host
  roofline   the Gram kernel (FP64 DMMA SYRK): algorithmic n*p*(p+1) flops / CUDA-event kernel time,
             against the FP64 tensor peak measured in-run with cuBLAS DGEMM (MEASURED_PEAKS.json has
             no FP64 entry)
  cpu_baseline / --impl reference   the CPU oracle (restated reference algorithm, numpy/OpenBLAS +
             plain C; the reference itself cannot be built here: no R / Rcpp / Eigen) on the box's host
             cores, on a bounded row sample, data passes extrapolated linearly in n
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 1000
ROWS_PER_GPU = 12_500_000
PENALTIES = ["lasso", "scad", "mcp"]
GAMMAS = [3.0, 3.7, 3.0]
NLAMBDA = 100
OPTS = dict(maxit=500, tol=1e-7)


def fit_args(X, y):
    p = P
    return [X, y, "gaussian", PENALTIES, [], [], [], [], [], NLAMBDA, 1e-4, 1.0, GAMMAS, 0.5, np.ones(p), True, True,
            False, dict(OPTS)]


def workload_name(n_gpus, rows):
    return (f"configs[4] big.oem n={rows * n_gpus:.3g} x p={P} FP64 ({rows:.3g} rows = {rows * P * 8 / 1e9:.0f} GB "
            f"per GPU, weak scaling), lasso+scad+mcp, 100 lambdas, standardize, intercept, tol 1e-7")


# ----------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, w in zip(sm, power) if w > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_fit_time_one_thread(rows_full, sample_rows, seed=1234):
    """The same port with BLAS limited to ONE thread (the reference's default is ncores = 1, R/oem.R:191): one untimed and
    one timed fit on a quarter of the all-cores sample, extrapolated like cpu_fit_time.  None if threadpoolctl is missing."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        return None
    with threadpool_limits(limits=1):
        sec, info = cpu_fit_time(rows_full, max(20_000, sample_rows // 4), 1, 1, seed)
    return {"value": sec, "unit": "s", "cores": 1, "sample_rows": info["sample_rows"]}


def cpu_fit_time(rows_full, sample_rows, steps, warmup, seed=1234):
    """Times oracle.oem_fit_big (numpy/OpenBLAS X'X + column sweeps, plain-C OEM iterations) on
    `sample_rows` rows with all host threads.  The O(n) data passes are timed again on their own and
    extrapolated linearly to rows_full; the O(p^2) path phase (total - data passes) is not scaled.
    Returns (seconds_full, details)."""
    from oracle import oracle as orc
    orc.build()
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((sample_rows, P)))
    b = np.zeros(P)
    b[:25] = rng.uniform(-0.5, 0.5, 25)
    y = X @ b + rng.standard_normal(sample_rows)
    cores = len(os.sched_getaffinity(0))
    est, parts = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.oem_fit_big(*fit_args(X, y))
        total = time.perf_counter() - t0
        t0 = time.perf_counter()
        _ = (X ** 2).sum(axis=0); _ = X.T @ y; _ = X.sum(axis=0); _ = X.T @ X      # the passes over X of oem_fit_big
        t_data = time.perf_counter() - t0
        t_path = max(total - t_data, 0.0)
        if it >= warmup:
            est.append(t_data * (rows_full / sample_rows) + t_path)
            parts.append((t_data, t_path))
    return float(np.mean(est)), {"cores": cores, "sample_rows": sample_rows, "kind": "port",
                                 "t_data_sample_s": float(np.mean([p[0] for p in parts])),
                                 "t_path_s": float(np.mean([p[1] for p in parts]))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.rows
    sec, info = cpu_fit_time(rows * args.gpus, args.cpu_sample_rows, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "full lambda-path fit time", "value": sec, "unit": "s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus, rows)},
            "cpu_baseline": {"value": sec, "unit": "s", "cores": info["cores"], "kind": "port",
                             "sample": f"oracle.oem_fit_big on {info['sample_rows']} x {P} rows per step; data passes "
                                       f"scaled x{rows * args.gpus / info['sample_rows']:.0f} to n={rows * args.gpus:.3g}, "
                                       "path phase unscaled"},
            "e2e": {"value": sec, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gen_shard(torch, rows, seed, device):
    """X ~ N(0,1) rows x P column-major on the device, y = X b + N(0,1)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = rows + (rows & 1)
    Xt = torch.empty((P, ld), dtype=torch.float64, device=device)        # row j = column j of X
    gb = torch.Generator(device="cpu"); gb.manual_seed(105)
    b = torch.zeros(P, dtype=torch.float64)
    b[:25] = torch.rand(25, generator=gb, dtype=torch.float64) - 0.5
    b = b.to(device)
    y = torch.randn(rows, generator=g, dtype=torch.float64, device=device)
    step = 25
    for j in range(0, P, step):
        blk = Xt[j:j + step]
        blk.normal_(generator=g)
        y += blk[:, :rows].t() @ b[j:j + step]
    if ld != rows:
        Xt[:, rows:] = 0
    return Xt.t()[:rows], y            # view with stride (1, ld)


def measure_fp64_peak(torch, device):
    N = 8192
    a = torch.randn(N, N, dtype=torch.float64, device=device)
    b = torch.randn(N, N, dtype=torch.float64, device=device)
    for _ in range(2):
        torch.matmul(a, b)
    best = 1e9
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / 1e3)
    del a, b
    return 2.0 * N ** 3 / best / 1e12


def run_ours(args):
    import torch
    import oem_b200
    from oem_b200 import api
    from oem_b200.dist import Comm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        comm = Comm()
    api.load()
    rows = args.rows
    fp64_peak = measure_fp64_peak(torch, device)
    X, y = gen_shard(torch, rows, 105_000 + rank, device)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    opts = api.make_opts(dict(OPTS, stream=stream.cuda_stream), comm)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """barrier + sync on both sides, CUDA events on the launching stream, max over ranks."""
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        outs, walls = [], []
        for _ in range(steps):
            t0 = time.perf_counter()
            outs.append(fn())
            walls.append(round((time.perf_counter() - t0) * 1e3, 2))
        e.record(stream)
        timed.last_walls = walls
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()) / steps, outs

    a_dev = fit_args(X, y)
    a_dev[-1] = opts

    def step_dev():
        return oem_b200.oem_fit_big(*a_dev)

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, outs = timed(step_dev, args.steps)
    walls_dev = timed.last_walls
    clocks = sampler.stop()
    st = outs[-1]["stats"]
    gram_ms = float(np.mean([o["stats"]["ms_gram"] / max(1, o["stats"]["gram_launches"]) for o in outs]))
    gram_flops = st["gram_flops"] / max(1, st["gram_launches"])
    achieved = gram_flops / (gram_ms / 1e3) / 1e12

    traffic, traffic_note = None, None
    try:       # per-launch DRAM bytes of the Gram kernel from the committed ncu capture of this exact launch shape
        tj = json.load(open(os.path.join(ROOT, "profiles", "gram_traffic.json")))
        if tj["rows"] == rows and tj["p"] == P:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_note = "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu capture in profiles/gram_traffic.json"
    except Exception:
        pass
    sm_count = torch.cuda.get_device_properties(device).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 0.0
    dmma_limit = 4 * 512 / 16 * sm_count * sm_mhz * 1e6 / 1e12
    line = {"metric": "full lambda-path fit time", "value": ms_step / 1e3, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world, rows), "l2": "inputs (100 GB/GPU) exceed L2; no flush needed",
                       "phases_ms": {k: st[k] for k in ("ms_colstats", "ms_gram", "ms_gram_reduce", "ms_allreduce",
                                                        "ms_assemble", "ms_path", "ms_total")},
                       "oem_iterations": st["total_oem_iters"], "lanczos_steps": st["lanczos_steps"],
                       "host_wall_ms_per_step": walls_dev, "lib_ms_total_per_step": [round(o["stats"]["ms_total"], 2) for o in outs]},
            "clocks": clocks, "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "roofline": {"bound": "tensor", "kernel": "gram_syrk_kernel (FP64 DMMA.8x8x4, TMA-staged)",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_note,
                         "algorithmic_bytes_per_launch": 8.0 * rows * P,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_launch": gram_flops, "ms_per_launch": gram_ms,
                         # second yardstick: one DMMA.8x8x4 (512 flop) per 16 cycles per SM sub-partition
                         # (tools/micro/fp64_lat.cu) x 4 sub-partitions x SMs x the sampled SM clock
                         "dmma_issue_limit_tflops": dmma_limit,
                         "frac_of_dmma_issue_limit": achieved / dmma_limit if dmma_limit else None}}

    # ---- e2e: HOST buffers through the same C-ABI call ----
    if not args.no_e2e:
        e2e = run_e2e(torch, api, oem_b200, X, y, rows, opts, timed, args, st)
        line["e2e"] = e2e
    # ---- CPU baseline (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu:
        del X, y
        torch.cuda.empty_cache()
        sec, info = cpu_fit_time(rows, args.cpu_sample_rows, 1, 1)
        line["cpu_baseline"] = {"value": sec, "unit": "s", "cores": info["cores"], "kind": "port",
                                "sample": f"oracle.oem_fit_big on {info['sample_rows']} x {P} rows; data passes scaled "
                                          f"x{rows / info['sample_rows']:.0f} to n={rows:.3g}, path phase unscaled"}
        try:
            one = cpu_fit_time_one_thread(rows, args.cpu_sample_rows)
        except Exception as e:                              # a reported extra, never a reason to lose the line
            one = {"error": repr(e)}
        if one:
            line["cpu_baseline"]["one_thread"] = one        # the reference's default ncores = 1
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_e2e(torch, api, oem_b200, X, y, rows, opts, timed, args, st_dev):
    """Same fit, inputs in pinned HOST memory; H2D of the row chunks and D2H of beta inside the timed region."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
    budget = int(avail * 0.8 / world)
    rows_h = min(rows, budget // (P * 8) // 72 * 72)
    if rows_h < rows and not args.allow_partial_e2e:
        return {"value": None, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": f"host RAM cannot hold the {rows * P * 8 / 1e9:.0f} GB shard per rank ({avail / 1e9:.0f} GB available "
                        f"for {world} ranks); no end-to-end number"}
    Xh = torch.empty((P, rows_h), dtype=torch.float64)
    cudart = torch.cuda.cudart()
    nbytes = Xh.numel() * 8
    rc = cudart.cudaHostRegister(Xh.data_ptr(), nbytes, 0)
    pinned = (int(rc) == 0)
    if not pinned:
        # a box whose pinned-memory limit cannot take every rank's shard (seen: 2 x 100 GB on a 251 GB host).  The failed
        # call leaves its code in the runtime's last-error slot, where torch's next launch check would find it: drain it
        # with a throw-away launch, then run this rank's leg from pageable memory (slower H2D, still a valid e2e number)
        print(f"[bench] rank {os.environ.get('RANK', '0')}: cudaHostRegister({nbytes / 1e9:.0f} GB) failed with code {int(rc)}; "
              "e2e leg uses pageable host memory on this rank", file=sys.stderr, flush=True)
        try:
            torch.zeros(1, device=X.device).add_(1)
            torch.cuda.synchronize()
        except RuntimeError:
            pass
    all_pinned = pinned
    if world > 1:
        flag = torch.tensor([1.0 if pinned else 0.0], dtype=torch.float64, device=X.device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        all_pinned = bool(flag.item() > 0.5)
    step = 50
    for j in range(0, P, step):
        Xh[j:j + step].copy_(X.t()[j:j + step, :rows_h], non_blocking=False)
    yh = y[:rows_h].cpu().numpy()
    Xh_np = Xh.numpy().T                       # rows_h x P, column-major view
    a = fit_args(Xh_np, yh)
    a[-1] = opts
    opts.gigs = args.gigs

    def step_host():
        return oem_b200.oem_fit_big(*a)

    for _ in range(max(1, args.warmup - 1)):
        step_host()
    ms, outs = timed(step_host, args.steps)
    s = outs[-1]["stats"]
    if pinned:
        cudart.cudaHostUnregister(Xh.data_ptr())
    out = {"value": ms / 1e3, "unit": "s", "h2d_bytes_per_step": int(s["h2d_bytes"]),
           "d2h_bytes_per_step": int(s["d2h_bytes"]), "host_memory": "pinned (cudaHostRegister)" if all_pinned else "pageable on at least one rank (cudaHostRegister refused)",
           "rows_per_gpu": rows_h, "stream_chunk_gb": args.gigs}
    if rows_h < rows:
        out["note"] = f"host buffer holds {rows_h} of {rows} rows per rank"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU (default: the 100 GB shard of configs[4])")
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--gigs", type=float, default=2.0, help="host->device streaming chunk (GB) for the e2e leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--allow-partial-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
