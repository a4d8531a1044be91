#!/usr/bin/env python
"""Top source lines of cat_world_kernel by stall samples, with the stall reasons (from an ncu --set full --import-source on capture).
usage: prof_lines.py <report.ncu-rep> <lib.so> [n_lines]"""
import collections, csv, re, subprocess, sys, tempfile
from pathlib import Path
rep, lib = sys.argv[1], sys.argv[2]
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = Path(__file__).resolve().parents[1]
tmp = Path(tempfile.mkdtemp())
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {Path(lib).resolve()} >/dev/null && nvdisasm -g -c *.cubin > dis.txt", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv 2>/dev/null > {tmp}/src.csv", shell=True, check=True)
rows = list(csv.reader(open(tmp / "src.csv")))
kname = rows[0][1] if len(rows[0]) > 1 else ""
hdr = rows[1]
spec = "ILi3ELi90" if re.search(r"<\(?int\)?3, ?\(?int\)?90>|<3, ?90>", " ".join(rows[0])) else "ILi0ELi0"
cur_func = cur_line = None
addr2line = {}
for l in open(tmp / "dis.txt"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined" not in m.group(3):
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur_func and "cat_world_kernel" + spec in cur_func:
        addr2line[int(m.group(1), 16)] = (cur_line, m.group(2))
ia, isamp, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
src = (ROOT / "as_cops_and_thieves_b200/csrc/world_kernel.cuh").read_text().split("\n")
agg = collections.defaultdict(lambda: collections.Counter())
base = None
tot = 0
for r in rows[2:]:
    if not r or r[0] == "Kernel Name":
        break
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    ln, _ = addr2line.get(a - base, (("?", 0), ""))
    c = agg[ln]
    c["samples"] += int(r[isamp]); c["inst"] += int(r[ii]); tot += int(r[isamp])
    for i, h in stall_cols:
        try:
            c[h] += int(r[i])
        except ValueError:
            pass
print(f"kernel {spec}, total samples {tot}")
for ln, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:nl]:
    top = sorted(((v, h[6:]) for h, v in c.items() if h.startswith("stall_") and v), reverse=True)[:3]
    txt = src[ln[1] - 1].strip()[:70] if ln[0].startswith("world_kernel") and ln[1] > 0 else ln[0]
    print(f"{ln[1]:5d} samp {c['samples'] / tot * 100:5.1f}% inst {c['inst']:10d} | {', '.join(f'{h} {v}' for v, h in top):45s} | {txt}")
