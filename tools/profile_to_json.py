#!/usr/bin/env python
"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`) into a small tracked JSON under profiles/.

usage: tools/profile_to_json.py OUT.json REP [REP ...]      (later reports override earlier ones per kernel name)
Also rewrites profiles/ncu_traffic.json (dram bytes per launch of the bench's kernels; bench.py reads it for
roofline.traffic)."""
import csv, io, json, os, subprocess, sys

KEEP = {
    "gpu__time_duration.sum": "time_ms",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "smsp__inst_executed_op_shared_atom.sum": "shared_atomic_instructions",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum": "shared_atomic_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "sm__cycles_elapsed.max": "sm_cycles",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem", "launch__shared_mem_per_block_static": "static_smem",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.sum": "lsu_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
}
STALLS = ["short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "barrier", "branch_resolving", "not_selected",
          "mio_throttle", "lg_throttle", "dispatch_stall", "no_instruction"]


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"kernel": d["Kernel Name"], "source_report": os.path.basename(rep)}
        for m, name in KEEP.items():
            if m in d and d[m] != "":
                v = float(d[m].replace(",", ""))
                unit = u.get(m, "")
                if name == "time_ms":
                    v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
                if name in ("dram_read", "dram_write"):
                    v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                    name_out = name + "_bytes"
                    k[name_out] = v
                    continue
                k[name] = v
        k["stall_per_issue"] = {s: float(d[f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"])
                                for s in STALLS if d.get(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")}
        if "sm_cycles" in k and "shared_wavefronts" in k:
            k["lsu_shared_wavefronts_per_sm_cycle"] = k["shared_wavefronts"] / 148.0 / k["sm_cycles"]
        res.append(k)
    return res


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    kernels = {}
    for rep in reps:
        for k in read(rep):
            kernels[k["kernel"]] = k
    doc = {"how": "ncu --set full --clock-control none --import-source on (one launch per kernel, after the command ran clean without "
                  "ncu); per-launch numbers, cold caches, serialised: use shares, not absolutes",
           "kernels": list(kernels.values())}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    traffic = {"_source": f"{os.path.relpath(out, root)} (ncu --set full, C3 query, one launch each)"}
    def tot(name):
        for k in kernels.values():
            if name in k["kernel"]:
                return k
        return None
    pairs = {"k_sieve": ["k_sieve"], "k_bitmap_moments2+k_scan_columns": ["k_bitmap_moments2w", "k_scan_columns"],
             "k_bitmap_scan<compare>": ["k_bitmap_scan<2, 0, 1>"], "k_bitmap_scan<reduce>+k_scan_columns": ["k_bitmap_scan<2, 0, 0>", "k_scan_columns"],
             "k_bitmap_moments_tab+k_scan_columns": ["k_bitmap_moments_tabw<5, 1>", "k_scan_columns"],
             "k_compact": ["k_compact"], "k_compact_lut": ["k_compact_lut"],
             "k_power_tiles+k_scan_columns+k_power_walk": ["k_power_tiles<2, 0, 16>", "k_scan_columns", "k_power_walk<2, 0, 1, 16>"]}
    for label, names in pairs.items():
        ks = [tot(n) for n in names]
        if any(k is None for k in ks):
            continue
        rd = sum(k.get("dram_read_bytes", 0) for k in ks)
        wr = sum(k.get("dram_write_bytes", 0) for k in ks)
        traffic[label] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr)}
    sv = tot("k_sieve")
    if sv and sv.get("shared_atomic_instructions"):
        traffic["k_sieve_counters"] = {
            "shared_atomic_warp_instructions": int(sv["shared_atomic_instructions"]),
            "shared_atomic_wavefronts": int(sv.get("shared_atomic_wavefronts", 0)),
            "wavefronts_per_atomic": round(sv.get("shared_atomic_wavefronts", 0) / sv["shared_atomic_instructions"], 3),
            "shared_wavefronts_total": int(sv.get("shared_wavefronts", 0)),
            "bank_conflict_wavefronts": int(sv.get("shared_bank_conflicts", 0)),
            "bank_conflict_share_of_shared_wavefronts": round(sv.get("shared_bank_conflicts", 0) / max(sv.get("shared_wavefronts", 1), 1), 3),
            "l1tex_throughput_pct": sv.get("l1tex_throughput_pct"), "issue_active_pct": sv.get("issue_active_pct"),
            "warp_instructions": int(sv.get("warp_instructions", 0)),
            "source": os.path.relpath(out, root) + " (one C3 launch, ncu --set full)"}
    tp = os.path.join(root, "profiles", "ncu_traffic.json")
    old = {}
    if os.path.exists(tp):
        old = json.load(open(tp))
    for k, v in old.items():
        traffic.setdefault(k, v)
    json.dump(traffic, open(tp, "w"), indent=1)
    print("wrote", out, "and", tp)


if __name__ == "__main__":
    main()
