#!/usr/bin/env python
"""Turn an `ncu --set full` report into the two small files kept under profiles/: per-kernel DRAM traffic per launch
(JSON, read by bench.py for `roofline.traffic`) and a one-line-per-metric text summary.
Usage: python tools/ncu_summarise.py profiles/rXX report1.ncu-rep [report2.ncu-rep ...] [--source-hash HASH]
       (writes rXX_ncu_dram_traffic.json, rXX_ncu_metrics.txt; the JSON carries `_meta` = report names + the hash of the CUDA sources
        the capture was taken from -- pass the hash printed by the capture run, default: the current tree -- so that bench.py can tell a
        stale capture from a current one)"""
import csv
import json
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]


def main():
    args = [a for a in sys.argv[1:]]
    src_hash = None
    if "--source-hash" in args:
        i = args.index("--source-hash")
        src_hash = args[i + 1]
        del args[i:i + 2]
    prefix, reps = args[0], args[1:]
    if src_hash is None:
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        src_hash = bench.kernel_source_hash()
    traffic, lines, seen = {"_meta": {"report": [r.split("/")[-1] for r in reps], "source_hash": src_hash}}, [], set()
    for rep in reps:
        summarise(rep, traffic, lines, seen)
    with open(prefix + "_ncu_dram_traffic.json", "w") as f:
        json.dump(traffic, f, indent=1)
    with open(prefix + "_ncu_metrics.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("kernels:", ", ".join(seen))


def summarise(rep, traffic, lines, seen):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").strip()
        dur = float(r[idx["gpu__time_duration.sum"]])
        durn = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[idx["gpu__time_duration.sum"]], 1.0)
        rec = {"dram_read": float(r[idx["dram__bytes_read.sum"]]), "read_unit": units[idx["dram__bytes_read.sum"]],
               "dram_write": float(r[idx["dram__bytes_write.sum"]]), "write_unit": units[idx["dram__bytes_write.sum"]],
               "duration_us": durn}
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rec["dram_bytes"] = rec["dram_read"] * mult.get(rec["read_unit"], 1e6) + rec["dram_write"] * mult.get(rec["write_unit"], 1e6)
        traffic.setdefault(name, []).append(rec)
        if name in seen:
            continue
        seen.add(name)
        lines.append(name + "  grid " + r[idx["Grid Size"]] + " block " + r[idx["Block Size"]])
        for m in METRICS:
            if m in idx:
                lines.append("    %-80s %s %s" % (m, r[idx[m]], units[idx[m]]))


if __name__ == "__main__":
    main()
