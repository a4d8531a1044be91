#!/usr/bin/env python
"""Static SASS instruction count of cat_world_kernel per source function (needs -lineinfo)."""
import collections, re, subprocess, sys, tempfile
from pathlib import Path
lib = Path(sys.argv[1]).resolve()
ROOT = Path(__file__).resolve().parents[1]
tmp = Path(tempfile.mkdtemp())
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {lib} >/dev/null && nvdisasm -g -c *.cubin > dis.txt", shell=True, check=True)
src = (ROOT / "as_cops_and_thieves_b200/csrc/world_kernel.cuh").read_text().split("\n")
funcs = []
for i, l in enumerate(src):
    m = re.match(r"^(?:__device__|__global__|static).*?\b(\w+)\s*\(", l)
    if m and not l.strip().startswith("//"):
        funcs.append((i + 1, m.group(1)))
def owner(f, l):
    if not f.startswith("world_kernel"):
        return "lib:" + f
    name = "?"
    for s, n in funcs:
        if l >= s:
            name = n
    return name
cnt = collections.Counter()
cur_func = cur_line = None
for l in open(tmp / "dis.txt"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", l) and cur_func and ("cat_world_kernel" in cur_func or "hull_closest" in cur_func or "los_blocked" in cur_func or "raster" in cur_func):
        cnt[(("K" if "cat_world_kernel" in cur_func else "F"), owner(*cur_line) if cur_line else "?")] += 1
tot = sum(cnt.values())
print("total", tot, "instr =", tot * 16 / 1024, "KB")
for k, v in cnt.most_common(30):
    print(f"{v:6d}  {k}")
