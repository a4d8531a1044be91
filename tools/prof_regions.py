#!/usr/bin/env python
"""Attribute an ncu --set full capture of cat_world_kernel to source regions / lines.
usage: prof_regions.py <report.ncu-rep> <lib.so> <n_worlds> [--lines N]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

rep, lib, nworlds = sys.argv[1], sys.argv[2], int(sys.argv[3])
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
ROOT = Path(__file__).resolve().parents[1]
tmp = Path(tempfile.mkdtemp())
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {Path(lib).resolve()} >/dev/null && nvdisasm -g -c *.cubin > dis.txt", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv 2>/dev/null > {tmp}/src.csv", shell=True, check=True)
raw = subprocess.run(f"ncu -i {rep} --page raw --csv 2>/dev/null", shell=True, capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
for key in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
    if key in rr[0]:
        i = rr[0].index(key)
        print(f"{key:70s} {rr[1][i]:10s} {rr[2][i]}")
# the instantiation the report's kernel is: cat_world_kernel<3, 90, true, false> -> ILi3ELi90ELb1ELb0E
kname = rr[2][rr[0].index("Kernel Name")] if "Kernel Name" in rr[0] else ""
mt = re.search(r"cat_world_kernel<(\d+), (\d+), (true|false|0|1), (true|false|0|1)>", kname)
INST = (f"ILi{mt.group(1)}ELi{mt.group(2)}ELb{int(mt.group(3) in ('true', '1'))}ELb{int(mt.group(4) in ('true', '1'))}E"
        if mt else "cat_world_kernel")
print("instantiation:", kname, INST)
cur_func = cur_line = None
addr2line = {}
for l in open(tmp / "dis.txt"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_func = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur_func and "cat_world_kernel" in cur_func and INST in cur_func:
        addr2line[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(tmp / "src.csv")))
for _i in range(2, len(rows)):          # several launches in one report: keep the first
    if rows[_i] and rows[_i][0] == "Kernel Name":
        rows = rows[:_i]
        break
hdr = rows[1]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
src = (ROOT / "as_cops_and_thieves_b200/csrc/world_kernel.cuh").read_text().split("\n")


def find(pat, start=0):
    if start is None:
        return None
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    return None


ob = find("__device__ __forceinline__ void observe_world")
marks = [("hull_closest", find("__device__ __noinline__ float3 hull_closest_impl")),
         ("thin_bb_hit", find("__device__ __forceinline__ bool thin_bb_hit")),
         ("fast_sqrt", find("__device__ __forceinline__ float fast_sqrt")),
         ("ray_edge", find("__device__ __forceinline__ float ray_edge")),
         ("wall_hit_normal", find("__device__ __forceinline__ float2 wall_hit_normal")),
         ("ray_circle", find("__device__ __forceinline__ float ray_circle")),
         ("make_ray/los", find("__device__ __forceinline__ Ray make_ray")),
         ("grid_cell", find("__device__ __forceinline__ int grid_cell")),
         ("raster_batch", find("__device__ __noinline__ void raster_batch")), ("raster:pairs", find("for (int p0 = 0; p0 < total; p0 += 32)")), ("observe:near", ob),
         ("rasterise_agent", find("__device__ __noinline__ void rasterise_agent")),
         ("stage_ray_slots", find("__device__ __forceinline__ void stage_ray_slots")),
         ("sweep_lists", find("__device__ __noinline__ void sweep_lists")),
         ("sweep:walk", find("while (__any_sync(0xFFFFFFFFu, cur >= 0))")),
         ("sweep:finish", find("if (__any_sync(0xFFFFFFFFu, fin))")),
         ("write_flat", find("__device__ __noinline__ void write_flat_layouts")),
         ("observe:agent-uniform", find("// ---- warp-uniform, per agent", ob)), ("observe:clear", find("// ---- (1) clear the depth buffer", ob)), ("observe:candidates", find("// ---- (2) lanes = edges", ob)), ("observe:epilogue3", find("// ---- (3) lanes = rays", ob)),
         ("observe:rayinit", find("for (int sub = 0; sub < nsub; ++sub)", ob)),
         ("observe:gridsetup", find("// ---- uniform-grid walk set-up", ob)),
         ("observe:traverse", find("// ---- converged traversal", ob)),
         ("observe:epilogue", find("if (valid) {", find("// ---- converged traversal", ob))),
         ("write_obs", find("__device__ __forceinline__ void write_observation")),
         ("reward", find("__device__ __forceinline__ float agent_reward")),
         ("physics", find("__device__ __forceinline__ void physics_world")),
         ("reset", find("__device__ __forceinline__ void reset_world")),
         ("kernel-main", find("cat_world_kernel(const __grid_constant__")),
         ("after", find("// ------------------------------------------------------------------ state pack"))]
marks = sorted([(n, l) for n, l in marks if l], key=lambda x: x[1])


def region(f, l):
    if not f.startswith("world_kernel"):
        return "lib:" + f
    r = "pre"
    for n, s in marks:
        if l >= s:
            r = n
    return r


lagg = collections.defaultdict(lambda: [0, 0, 0])
hot = collections.Counter()       # distinct instructions executed in >= 10 % of the world-steps, per region
agg = collections.defaultdict(lambda: [0, 0, 0])
llagg = collections.defaultdict(lambda: [0, 0, 0])
hot = collections.Counter()       # distinct instructions executed in >= 10 % of the world-steps, per region
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for r in rows[2:]:
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    ln = addr2line.get(a - base) or ("?", 0)
    v = [int(r[ii]), int(r[it]), int(r[isamp])]
    if v[0] >= 0.1 * nworlds:
        hot[region(*ln)] += 1
    for k in range(3):
        agg[region(*ln)][k] += v[k]
        lagg[ln][k] += v[k]
        tot[k] += v[k]
print(f"total warp-inst/world {tot[0] / nworlds:.0f}  thread-inst/world {tot[1] / nworlds:.0f}  eff {tot[1] / tot[0]:.1f}")
print(f"hot static footprint (instructions executed in >= 10 % of world-steps): {sum(hot.values())} instr = {sum(hot.values()) * 16 / 1024:.1f} KB")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:26s} warp-inst {v[0] / nworlds:8.0f}/world ({v[0] / tot[0] * 100:5.1f}%)  thread-inst/world {v[1] / nworlds:9.0f}  eff {v[1] / max(v[0], 1):5.1f}  samples {v[2] / max(tot[2], 1) * 100:5.1f}%  hot-instr {hot[k]:4d}")
for (f, l), v in sorted(lagg.items(), key=lambda kv: -kv[1][0])[:nlines]:
    txt = src[l - 1].strip()[:100] if f.startswith("world_kernel") and l > 0 else ""
    print(f"{f}:{l:5d} inst {v[0] / tot[0] * 100:5.1f}% eff {v[1] / max(v[0], 1):5.1f} samp {v[2] / max(tot[2], 1) * 100:5.1f}% | {txt}")
