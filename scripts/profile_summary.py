#!/usr/bin/env python3
"""Text summary of an `ncu --set full` report for the committed profiles/ directory.

    python scripts/profile_summary.py <report.ncu-rep> [<lib.so> <mangled-kernel-substring>]
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, report {rep}")
    for n, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            continue
        print(f"\n## launch {n}: {r[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:70s} {r[i]:>16s} {units[i]}")
        print("# warps stalled per issue-active cycle, by reason (> 0.05)")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                try:
                    val = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if val > 0.05:
                    name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                    print(f"  {name:28s} {val:8.3f}")
        break_after_first = len(sys.argv) <= 4
        if break_after_first:
            break
    if len(sys.argv) >= 4:
        print("\n# ---- executed instructions attributed to source lines (scripts/ncu_lines.py) ----")
        sys.stdout.flush()
        subprocess.run([sys.executable, __file__.replace("profile_summary.py", "ncu_lines.py"), sys.argv[2], sys.argv[3], rep, "32"])


if __name__ == "__main__":
    main()
