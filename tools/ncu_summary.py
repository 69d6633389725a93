"""Key figures of one .ncu-rep (first kernel in the report) as a markdown table row set: duration, pipe / memory
utilisation, DRAM traffic, stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof_gram.ncu-rep [...]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (DFMA/DMMA) active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"### {rep}: no kernels\n")
            continue
        hdr, units, val = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, v, u in zip(hdr, val, units)}
        print(f"### {rep.split('/')[-1]} — `{d.get('Kernel Name', ('?', ''))[0][:100]}`\n")
        print("| metric | value |\n|---|---|")
        for k, name in KEYS:
            if k in d:
                print(f"| {name} (`{k}`) | {d[k][0]} {d[k][1]} |")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(d[h][0].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]) + " |\n")


if __name__ == "__main__":
    main()
