#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into profiles/: selected raw metrics per
launch and the source-line hot spots.  Usage: tools/ncu_summary.py <rep> <out_prefix>"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out + "_raw.csv", "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if len(r) > 8 and r[2] == "-":
            try:
                agg[(cur, int(r[0]))] = (int(r[4] or 0), int(r[7] or 0), int(r[8] or 0), r[1].strip()[:110])
            except ValueError:
                pass
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    with open(out + "_hotspots.txt", "w") as f:
        f.write(f"first profiled launch; total stall samples {ts}, warp instructions {ti}\n")
        f.write("file:line  warp-inst (share)  samples (share)  active lanes per inst | source\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            f.write(f"{k[0]}:{k[1]:<4d} inst={v[1]:8d} ({100 * v[1] / ti:4.1f}%) samp={v[0]:5d} "
                    f"({100 * v[0] / ts:4.1f}%) lanes={v[2] / max(v[1], 1):4.1f} | {v[3]}\n")
    print("wrote", out + "_raw.csv", out + "_hotspots.txt")


if __name__ == "__main__":
    main()
