#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed) into the few numbers the roofline
discussion needs, and a launch list CSV (`--metrics gpu__time_duration.sum`) into per-kernel totals and shares.

  python scripts/summarize_ncu.py rep gpurun_out/leaf.ncu-rep [out.json]
  python scripts/summarize_ncu.py launches gpurun_out/launches.csv [out.json]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__t_bytes.sum", "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
]


def rep(path):
    # a .ncu-rep, or the CSV that `ncu -i x.ncu-rep --page raw --csv` printed on the GPU box (the reports themselves are too big to bring back)
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d.get("Kernel Name", "")[:90], "id": d.get("ID")}
        for k in hdr:
            base = k
            if base in KEYS or any(base.startswith(p) for p in ("smsp__pcsamp_warps_issue_stalled", "smsp__average_warp")):
                try:
                    e[base] = float(d[k].replace(",", ""))
                except Exception:
                    e[base] = d[k]
        e["units"] = {k: rows[1][hdr.index(k)] for k in hdr if k in e and k in KEYS}
        if "dram__bytes_read.sum" in e and "dram__bytes_write.sum" in e:       # ncu picks a unit per column: normalise to bytes
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            rd = e["dram__bytes_read.sum"] * scale.get(e["units"].get("dram__bytes_read.sum"), 1.0)
            wr = e["dram__bytes_write.sum"] * scale.get(e["units"].get("dram__bytes_write.sum"), 1.0)
            e["dram_bytes_total"] = rd + wr
            e["dram_bytes_read"], e["dram_bytes_write"] = rd, wr
        res.append(e)
    return res


def launches(path):
    txt = open(path).read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v_ms
    tot = sum(a[1] for a in agg.values())
    return {"total_ms": tot, "kernels": sorted(({"kernel": k, "launches": a[0], "ms": a[1], "share": a[1] / tot} for k, a in agg.items()),
                                               key=lambda x: -x["ms"])}


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    res = rep(path) if mode == "rep" else launches(path)
    s = json.dumps(res, indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(s)
    print(s[:6000])
