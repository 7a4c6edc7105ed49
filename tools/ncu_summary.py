#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py gpurun_out/x.ncu-rep [extra metric substrings]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    for r in rows[2:]:
        print("---- " + r[H.index("Kernel Name")][:150])
        for w in WANT:
            if w in H:
                print("%-75s %s %s" % (w, r[H.index(w)], rows[1][H.index(w)]))
        stalls = [(float(r[i].replace(",", "") or 0), h) for i, h in enumerate(H)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("  stall %-60s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        for e in extra:
            for i, h in enumerate(H):
                if e in h:
                    print("%-75s %s" % (h, r[i]))


if __name__ == "__main__":
    main()
