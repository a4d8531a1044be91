#!/usr/bin/env python
"""profiles/r2_gae_kernels.txt + gae_dram_bytes.json (+ the GAE entries of traffic.json) from the two ncu captures
tools/capture_profiles.sh takes (gpurun_out/r2f/gae_T{256,1024}.ncu-rep)."""
import importlib.util
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
spec = importlib.util.spec_from_file_location("sp", ROOT / "tools/summarize_profiles.py")
sp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sp)
D = ROOT / (sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2f")
M = 49152
gd = {}
with open(ROOT / "profiles/r2_gae_kernels.txt", "w") as out:
    out.write("# ncu --set full --clock-control none -k regex:cat_gae|cat_adv -s 6 -c 2 : python tools/prof_gae.py --T {256,1024} "
              "(isolated launches, L2 read-flushed before each; final round-2 kernels: memset-free statistics, programmatic dependent launch)\n")
    for T in (256, 1024):
        hdr, units, rows = sp.raw_rows(str(D / f"gae_T{T}.ncu-rep"))
        ik = hdr.index("Kernel Name")
        for r in rows:
            name = r[ik]
            out.write(f"## T={T} M={M}: {name[:90]}\n")
            for m in sp.METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.write(f"{m:80s} {units[i]:12s} {r[i]}\n")
            rd = sp.to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = sp.to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            i = hdr.index("gpu__time_duration.sum")
            us = float(r[i]) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[i], 1)
            alg = 17 if "gae" in name else 8
            out.write(f"dram traffic per launch: {(rd + wr) / 1e6:.1f} MB = {(rd + wr) / (T * M):.2f} B per sample (algorithmic {alg}); "
                      f"{(rd + wr) / us / 1e3:.0f} GB/s of real DRAM traffic under ncu\n")
            gd.setdefault(f"T{T}", {})["gae" if "gae" in name else "normalize"] = rd + wr
json.dump(gd, open(ROOT / "profiles/gae_dram_bytes.json", "w"), indent=1)
t = json.load(open(ROOT / "profiles/traffic.json"))
t["cat_gae-T256xM49152"] = gd["T256"]["gae"]
t["cat_adv_normalize-T256xM49152"] = gd["T256"]["normalize"]
t["cat_gae-T1024xM49152"] = gd["T1024"]["gae"]
json.dump(t, open(ROOT / "profiles/traffic.json", "w"), indent=1)
print(gd)
