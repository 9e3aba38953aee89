#!/usr/bin/env python
"""Print the headline metrics and the hottest SASS lines of an .ncu-rep (reads it with `ncu -i`)."""
import csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]

WANT = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]

def source(rep, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr = {h: i for i, h in enumerate(r)}
            if data: break
            continue
        if hdr and len(r) > 10:
            try:
                data.append((r[hdr["Source"]], int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]),
                             r[hdr["Avg. Threads Executed"]], r[hdr["L1 Wavefronts Shared"]]))
            except ValueError:
                pass
    ts = sum(d[1] for d in data) or 1
    ti = sum(d[2] for d in data) or 1
    print("total warp instructions", ti, "samples", ts)
    for i, d in enumerate(data):
        if d[2] > ti * 0.006 or d[1] > ts * 0.008:
            print(f"{i:5d} {d[0][:64]:64s} smp {100*d[1]/ts:5.1f}% inst {100*d[2]/ti:5.2f}% thr {d[3]:>5s} wf {d[4]}")
    return data

if __name__ == "__main__":
    rep = sys.argv[1]
    for k in raw(rep):
        for w in WANT:
            if w in k: print(w, k[w])
        print()
    if len(sys.argv) > 2: source(rep)
