"""Condense an .ncu-rep (ncu --set full [--import-source on]) into a small text summary that can travel back from the
GPU box: headline metrics per launch, instruction mix and the hottest SASS lines of the first launch.
    python scripts/ncu_summarize.py report.ncu-rep > summary.txt"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        for w in WANT:
            if w in ix:
                print(f"{w} = {r[ix[w]][:160]}")
        print()
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--kernel-id", ":::1"]))))
    h = None
    for i, r in enumerate(src):
        if "Source" in r and "Instructions Executed" in r:
            h, body = {k: j for j, k in enumerate(r)}, src[i + 1:]
            break
    if h is None:
        return
    ops, hot, total = collections.Counter(), [], 0
    for r in body:
        try:
            n = int(r[h["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        if n == 0:
            continue
        s = r[h["Source"]].strip()
        tok = s.split()
        op = (tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]).split(".")[0]
        ops[op] += n
        total += n
        stall = r[h["Warp Stall Sampling (All Samples)"]] if "Warp Stall Sampling (All Samples)" in h else ""
        hot.append((int(stall or 0), n, s[:110]))
    print(f"SASS lines executed: {len(hot)}   warp instructions (source page): {total}")
    for op, n in ops.most_common(24):
        print(f"  {op:10s} {100.0 * n / total:5.1f} %")
    print("hottest lines by stall samples:")
    for st, n, s in sorted(hot, reverse=True)[:25]:
        print(f"  {st:7d} {n:10d}  {s}")


if __name__ == "__main__":
    main()
