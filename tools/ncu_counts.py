#!/usr/bin/env python
"""FP64 work and DRAM traffic of ONE profiled launch, from an `ncu --set full --import-source on` report:
flops = 2 x DFMA + DADD + DMUL thread instructions (predicated on) + 512 x DMMA.8x8x4 warp instructions, counted on the SASS page;
dram bytes = dram__bytes_read.sum + dram__bytes_write.sum.   usage: ncu_counts.py report.ncu-rep [units]   (units: work items the launch
processed, e.g. instance x interior-point iterations; the per-unit figures are printed when given)"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
get = lambda name: next((float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(u, 1.0)
                         for h, u, v in zip(hdr, rows[1], vals) if h == name), None)
out = {"kernel": next(v for h, v in zip(hdr, vals) if h == "Kernel Name"), "duration_s": get("gpu__time_duration.sum"),
       "dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
       "issue_active_pct": get("smsp__issue_active.avg.pct"), "registers": get("launch__registers_per_thread"),
       "warp_instructions": get("smsp__inst_executed.sum")}
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h = rows[1]
isrc, iw, it = h.index("Source"), h.index("Instructions Executed"), h.index("Predicated-On Thread Instructions Executed")
ops = {}
for r in rows[2:]:
    if len(r) <= it:
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    if op in ("DFMA", "DADD", "DMUL", "DMMA", "DSETP", "MUFU"):
        w, t = ops.get(op, (0, 0))
        ops[op] = (w + int(r[iw] or 0), t + int(r[it] or 0))
flops = 2 * ops.get("DFMA", (0, 0))[1] + ops.get("DADD", (0, 0))[1] + ops.get("DMUL", (0, 0))[1] + 512 * ops.get("DMMA", (0, 0))[0]
out.update({"fp64_flop": flops, "dfma_thread": ops.get("DFMA", (0, 0))[1], "dadd_thread": ops.get("DADD", (0, 0))[1], "dmul_thread": ops.get("DMUL", (0, 0))[1],
            "dmma_warp": ops.get("DMMA", (0, 0))[0], "fp64_tflops_in_launch": flops / out["duration_s"] / 1e12})
if units:
    out.update({"units": units, "flop_per_unit": flops / units, "dram_bytes_per_unit": out["dram_bytes"] / units})
print(json.dumps(out, indent=1))
