#!/usr/bin/env python
"""Turn Nsight Compute output into the small text / JSON summaries committed here.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
      per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list
  python profiles/summarize_ncu.py report gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full.txt
      one block per profiled launch of an `ncu --set full` report: duration, DRAM bytes and
      throughput, IPC, occupancy, top stall reasons (needs the `ncu` CLI to read the report)
"""

import collections
import csv
import io
import json
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    ours = 0.0
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        try:
            ms = float(r[mv].replace(",", "")) / 1e6
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] = max(a[2], ms)
        total += ms
        if name.startswith("k_"):
            ours += ms
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total:.3f} ms of kernel time "
          f"({ours:.3f} ms in soap_b200 kernels; the rest is torch's synthetic data generation)")
    print(f"{'kernel':60s} {'n':>5s} {'total ms':>10s} {'max ms':>9s} {'share of ours':>14s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if not k.startswith("k_"):
            continue
        print(f"{k[:60]:60s} {a[0]:5d} {a[1]:10.3f} {a[2]:9.3f} {100 * a[1] / ours:13.1f}%")


def _ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def report(path):
    det = _ncu(path, "details")
    h = det[0]
    ki, mi, vi, ui, idi = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    want = ["Duration", "DRAM Throughput", "Memory Throughput", "L2 Hit Rate", "Executed Ipc Active", "Issue Slots Busy",
            "Achieved Occupancy", "Theoretical Occupancy", "Registers Per Thread", "Executed Instructions",
            "Avg. Active Threads Per Warp", "Grid Size", "Block Size"]
    per = collections.OrderedDict()
    for r in det[1:]:
        d = per.setdefault(r[idi], {"kernel": r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")})
        if r[mi] in want:
            d.setdefault(r[mi], f"{r[vi]} {r[ui]}".strip())
    raw = _ncu(path, "raw")
    rh = raw[0]
    cols = [i for i, c in enumerate(rh) if "pcsamp_warps_issue_stalled" in c and "not_issued" not in c]
    rid = rh.index("ID")
    dr = rh.index("dram__bytes_read.sum") if "dram__bytes_read.sum" in rh else None
    dw = rh.index("dram__bytes_write.sum") if "dram__bytes_write.sum" in rh else None
    units = raw[1]
    traffic = {}
    for r in raw[2:]:
        d = per.get(r[rid])
        if d is None:
            continue
        st = sorted(((float(r[i].replace(",", "") or 0), rh[i].replace("smsp__pcsamp_warps_issue_stalled_", ""))
                     for i in cols), reverse=True)
        tot = sum(x for x, _ in st) or 1.0
        d["stalls"] = ", ".join(f"{n} {100 * x / tot:.0f}%" for x, n in st[:5])
        if dr is not None:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r[dr].replace(",", "")) * scale.get(units[dr], 1.0)
            wr = float(r[dw].replace(",", "")) * scale.get(units[dw], 1.0)
            d["dram_bytes"] = int(rd + wr)
            traffic.setdefault(d["kernel"], []).append(int(rd + wr))
    for i, d in per.items():
        print(f"[{i}] {d['kernel']}")
        for k in want + ["dram_bytes", "stalls"]:
            if k in d:
                print(f"      {k:32s} {d[k]}")
    print("# dram traffic per launch (bytes):", json.dumps(traffic))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
