"""Condense an .ncu-rep (ncu --set full) into the small per-kernel CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.csv [--first-per-kernel]
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    first = "--first-per-kernel" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    seen, cols = set(), []
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        if first and name in seen:
            continue
        seen.add(name)
        cols.append((name, r))
    wanted = [m for m in METRICS if m in hdr] + [h for h in hdr if STALLS.match(h)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [c[0] for c in cols])
        for m in wanted:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [c[1][i] for c in cols])
    print(f"{out}: {len(cols)} kernels, {len(wanted)} metrics")


if __name__ == "__main__":
    main()
