#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per profiled launch.
Usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv
import subprocess
import sys

WANT = [
    ("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"), ("time_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
]


def to_mb(val, unit):
    v = float(val)
    u = unit.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([n for n, _ in WANT])
        for r in rows[2:]:
            line = []
            for n, k in WANT:
                if k not in idx:
                    line.append("")
                    continue
                v, u = r[idx[k]], units[idx[k]]
                if n == "kernel":
                    v = v.split("(")[0].replace("void ", "")
                elif n.startswith("dram_") and n.endswith("MB"):
                    v = "%.1f" % to_mb(v, u)
                elif n == "time_us":
                    v = "%.1f" % (float(v) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1))
                else:
                    try:
                        v = "%.3g" % float(v)
                    except ValueError:
                        pass
                line.append(v)
            w.writerow(line)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
