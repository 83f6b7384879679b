"""Read an .ncu-rep: per-kernel headline metrics, stall reasons, and the hottest SASS regions."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
for k, r in enumerate(rows[2:]):
    print(f"=== kernel {k}: {r[hdr.index('Kernel Name')][:50]}")
    for w in want:
        idx = [i for i, h in enumerate(hdr) if h == w]
        if idx:
            print(f"  {w:75s} {r[idx[0]]} {rows[1][idx[0]]}")
    items = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    tot = sum(v for _, v in items) or 1
    print("  stalls: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.1f}%" for h, v in sorted(items, key=lambda t: -t[1])[:9]))
