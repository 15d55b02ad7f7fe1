"""Summarise an .ncu-rep (read with `ncu -i` on the CPU box): headline metrics + top SASS stall sites.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def sass(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    data = []
    for i, r in enumerate(rows[2:]):
        try:
            data.append((i, float(r[ci] or 0), float(r[cw] or 0), r[cs].strip()))
        except (ValueError, IndexError):
            pass
    return data


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    m = raw(rep)
    for k in KEYS:
        if k in m:
            print(f"{k:90s} {m[k][0]:>16s} {m[k][1]}")
    d = sass(rep)
    tot, tots = sum(x[1] for x in d), sum(x[2] for x in d)
    print(f"-- {len(d)} SASS instructions, {tot:.0f} executed, {tots:.0f} stall samples; top by stall samples:")
    for x in sorted(d, key=lambda x: -x[2])[:n]:
        print(f"#{x[0]:5d} {x[1] / tot * 100:5.1f}% inst {x[2] / tots * 100:5.1f}% stall | {x[3][:100]}")


if __name__ == "__main__":
    main()
