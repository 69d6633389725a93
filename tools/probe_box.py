"""One-off probe of the GPU box: host cores/RAM, FP64 GEMM/SYRK peak (cuBLAS via torch), H2D bandwidth.
Writes gpurun_out/probe.json.  Not part of the product path."""
import json, os, time, subprocess
import torch

out = {}
out["nproc"] = os.cpu_count()
try:
    out["affinity"] = len(os.sched_getaffinity(0))
except Exception:
    pass
with open("/proc/meminfo") as f:
    mi = f.read().splitlines()
out["meminfo"] = mi[:3]
try:
    out["cgroup_mem_max"] = open("/sys/fs/cgroup/memory.max").read().strip()
    out["cgroup_cpu_max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
except Exception as e:
    out["cgroup_err"] = str(e)
out["cpu_model"] = subprocess.run("grep -m1 'model name' /proc/cpuinfo", shell=True, capture_output=True, text=True).stdout.strip()
out["nvidia_smi"] = subprocess.run("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv", shell=True, capture_output=True, text=True).stdout
dev = torch.device("cuda:0")
free, total = torch.cuda.mem_get_info()
out["mem_free_total"] = [free, total]

def ev_time(fn, iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 1e3 / iters

# FP64 GEMM peak, square
res = {}
for N in (4096, 8192):
    a = torch.randn(N, N, dtype=torch.float64, device=dev); b = torch.randn(N, N, dtype=torch.float64, device=dev)
    for _ in range(3): torch.matmul(a, b)
    best = min(ev_time(lambda: torch.matmul(a, b), 3) for _ in range(5))
    res[f"dgemm_{N}"] = 2.0 * N ** 3 / best / 1e12
    t0 = time.time(); it = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b); it += 1
        if it % 8 == 0: torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    res[f"dgemm_{N}_sustained"] = 2.0 * N ** 3 * it / (s.elapsed_time(e) / 1e3) / 1e12
    del a, b
# tall-skinny X'X like ours: (1000 x K) @ (K x 1000)
K = 2_000_000
x = torch.randn(K, 1000, dtype=torch.float64, device=dev)   # row-major K x p  == col-major p x K
for _ in range(2): torch.matmul(x.t(), x)
best = min(ev_time(lambda: torch.matmul(x.t(), x), 2) for _ in range(4))
res["dgemm_xtx_p1000_full_flops_TF"] = 2.0 * K * 1000 * 1000 / best / 1e12
xc = x.t().contiguous()   # p x K row-major -> X col-major n x p analogue: X = xc.t()
X = xc.t()
for _ in range(2): torch.matmul(X.t(), X)
best = min(ev_time(lambda: torch.matmul(X.t(), X), 2) for _ in range(4))
res["dgemm_xtx_colmajor_p1000_TF"] = 2.0 * K * 1000 * 1000 / best / 1e12
del x, xc, X
out["fp64"] = res
# H2D bandwidth pinned / pageable, 2 GiB
nb = 2 << 30
hp = torch.empty(nb, dtype=torch.uint8).pin_memory()
hn = torch.empty(nb, dtype=torch.uint8); hn.fill_(1)
d = torch.empty(nb, dtype=torch.uint8, device=dev)
for name, h in (("pinned", hp), ("pageable", hn)):
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.time(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
    out[f"h2d_{name}_GBs"] = nb / dt / 1e9
t = time.time(); big = torch.empty(8 << 30, dtype=torch.uint8).pin_memory(); out["pin_8GiB_s"] = time.time() - t
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
