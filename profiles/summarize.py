#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the small per-launch summary that is
committed under profiles/.  Usage: python profiles/summarize.py gpurun_out/x.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__cluster_dim_x",
    "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tma.sum", "sm__inst_executed_pipe_tmem.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
        "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(rep, out, unique=False):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    seen = set()
    for r in rows[2:]:
        if unique:
            if r[hdr.index("Kernel Name")] in seen:
                continue
            seen.add(r[hdr.index("Kernel Name")])
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if u in UNIT and ("bytes" in k or "time" in k):
                    v *= UNIT[u]
                    u = "byte" if "bytes" in k else "s"
                d[k] = {"value": v, "unit": u}
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            d["traffic_bytes_per_launch"] = (d["dram__bytes_read.sum"]["value"] +
                                             d["dram__bytes_write.sum"]["value"])
        res.append(d)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"), d.get("traffic_bytes_per_launch"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], unique="--unique" in sys.argv)
