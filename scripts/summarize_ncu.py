#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep / launch-list CSVs into the small text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
STALLS = "smsp__average_warps_issue_stalled_"


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            f.write(f"\n## {d['Kernel Name'][1]}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:70s} {d[k][1]:>20s} {d[k][0]}\n")
            f.write("-- warp stall reasons (warps stalled per issue-active cycle) --\n")
            st = [(float(v[1]), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                  if k.startswith(STALLS) and v[1] not in ("", "n/a")]
            for v, k in sorted(st, reverse=True):
                if v > 0.001:
                    f.write(f"  {k:30s} {v:8.3f}\n")


def launches(csv_path, out):
    rows = list(csv.reader(open(csv_path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                u = d["Metric Unit"]
                ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v if u.startswith("m") else v * 1e3
                name = d["Kernel Name"].split("(")[0]
                agg[name][0] += 1
                agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel launch list summary of {csv_path} (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: compare shares)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:36s} launches {v[0]:4d}  total {v[1]:12.3f} ms  share {100 * v[1] / tot:6.2f} %\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3])
