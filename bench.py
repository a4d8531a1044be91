#!/usr/bin/env python
"""Benchmark of the batched cops-and-thieves environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--worlds n]

A "step" is one lockstep transition of every world on a rank: ONE launch of `cat_world_kernel` (termination test,
action impulses, 90-ray sensor sweep per agent, float16 observation chain, rewards, rigid-body step, auto-reset).
Default workload = the configuration the metric is quoted on (BASELINE.json north_star / configs[3] with the real
large-segment map): agh-map, 16384 worlds per GPU.  Weak scaling: the worlds per GPU are fixed, ranks shard the
global world range, there is no data-path collective.

Prints ONE JSON line (rank 0).  `value` = agent-steps/s over all ranks with everything resident in HBM; `e2e` = the
same metric through the host-buffer API (pinned actions in, the step's records back in pinned host memory, copies
inside the timed region).  Every named workload gets the same treatment under `other_workloads` (own roofline, e2e).
Cold L2: one step's working set is a few MB, far smaller than the 126 MB L2, so the timed steps rotate over enough
independent world sets that their buffers exceed L2 ("inputs larger than L2"); a block of K steps is timed back to
back between one CUDA-event pair, and the block is repeated until the timed span is at least 50 ms.
`--workload mappo-agh-map` times BASELINE config 5 instead (MAPPO self-play training on agh-map).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALG_BYTES_PER_AGENT_STEP = 336          # native dtypes, step + raycast obs
FALLBACK_HBM_GBS = 6650.0               # /opt/skills/guides/B200_PROFILING.md fallback
MIN_SPAN_MS = 50.0                      # a timed span shorter than this is dominated by launch / event noise
WORKLOADS = {                           # name -> (map, free-space spawns, worlds per GPU)
    "agh-map-16384": ("agh-map", True, 16384),          # the metric's configuration (default): 496 hull edges
    "squarinth-4096": ("squarinth", False, 4096),       # BASELINE.json configs[1] (the parity configuration)
    "lbirinth-4096": ("lbirinth", False, 4096),         # the fifth named map
    "labyrinth-8192": ("labyrinth", True, 8192),        # configs[2] per-GPU share
    "grandbyrinth-16384": ("grandbyrinth", False, 16384),  # configs[3] as named (16 edges)
}
DEFAULT_WORKLOAD = "agh-map-16384"
TRAIN_WORKLOAD = "mappo-agh-map"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + [TRAIN_WORKLOAD])
    ap.add_argument("--worlds", type=int, default=None, help="worlds per GPU (overrides the workload's)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other workloads / GAE / cpu baseline")
    return ap.parse_args()


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def build_cmap(name, free):
    from as_cops_and_thieves_b200.maps import compile_map, free_space_regions, load_named_map
    m = load_named_map(name)
    return compile_map(m, name=name, spawn_override=free_space_regions(m) if free else None)


def profile_json(name):
    p = ROOT / "profiles" / name
    try:
        return json.load(open(p)) if p.exists() else {}
    except Exception:
        return {}


# ------------------------------------------------------------------ CPU arm (host cores)
def time_cpu_port(map_name, free, n_worlds, seconds=12.0, min_steps=3):
    """The reference's algorithm on the host cores: the fp64 oracle port with OpenMP over worlds."""
    import numpy as np
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    from oracle.cat_oracle import Oracle, num_threads
    cmap = build_cmap(map_name, free)
    orc = Oracle(cmap, seed=0)
    st = orc.new_state(n_worlds)
    out = orc.new_out(n_worlds)
    orc.reset(st, out=out)
    rng = np.random.default_rng(1)
    acts = [rng.integers(0, 4, (n_worlds, orc.A)).astype(np.int32) for _ in range(8)]
    orc.step(st, acts[0], out)  # warm-up
    t0 = time.perf_counter()
    steps = 0
    while steps < min_steps or time.perf_counter() - t0 < seconds:
        orc.step(st, acts[steps % 8], out)
        steps += 1
    dt = time.perf_counter() - t0
    return dict(value=n_worlds * orc.A * steps / dt, steps=steps, seconds=dt, cores=num_threads(),
                worlds=n_worlds, A=orc.A)


def time_pymunk_reference(map_name, free, seconds):
    """The REAL Pymunk path, one process per host core — None when pymunk is not importable
    (also tried with baseline/_ref on sys.path)."""
    from oracle import pymunk_ref
    pm, why = pymunk_ref.probe()
    if pm is None or getattr(pm, "IS_STAND_IN", False):
        return None, why
    return pymunk_ref.time_reference(map_name, free, seconds=seconds), why


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (libgomp reads this at load)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    workload = DEFAULT_WORKLOAD if args.workload == TRAIN_WORKLOAD else args.workload
    map_name, free, _ = WORKLOADS[workload]
    K, W = max(1, args.steps), max(0, args.warmup)
    base = {"impl": "reference", "metric": "agent-steps/s (physics+raycast obs)", "unit": "agent-steps/s",
            "n_gpus": args.gpus, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}
    ref, why = time_pymunk_reference(map_name, free, seconds=10.0)
    if ref is not None:
        # the unmodified engine: one world per process, one process per core, ~10 s each
        value = ref["value"]
        line = dict(base, value=value, steps=K, warmup=W, ms_per_step=None,
                    config={"workload": workload, "map": map_name, "worlds_per_step": ref["cores"], "how": ref["how"]},
                    cpu_baseline={"value": value, "unit": "agent-steps/s", "cores": ref["cores"], "kind": "reference",
                                  "sample": f"{ref['cores']} processes x 1 {map_name} world x {ref['seconds']:.1f} s, {ref['how']}"},
                    e2e={"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    import numpy as np
    from oracle.cat_oracle import Oracle, num_threads
    sample_worlds = 512
    cmap = build_cmap(map_name, free)
    orc = Oracle(cmap, seed=0)
    st = orc.new_state(sample_worlds)
    out = orc.new_out(sample_worlds)
    orc.reset(st, out=out)
    rng = np.random.default_rng(1)
    acts = [rng.integers(0, 4, (sample_worlds, orc.A)).astype(np.int32) for _ in range(8)]
    K, W = min(K, 400), min(W, 20)       # bounded: the whole run ends within a few minutes
    for i in range(W):
        orc.step(st, acts[i % 8], out)
    t0 = time.perf_counter()
    for i in range(K):
        orc.step(st, acts[i % 8], out)
    dt = time.perf_counter() - t0
    value = sample_worlds * orc.A * K / dt
    sample = f"{sample_worlds} {map_name} worlds x {K} steps, OpenMP over worlds (oracle port of the Pymunk path)"
    line = dict(base, value=value, steps=K, warmup=W, ms_per_step=dt / K * 1e3,
                config={"workload": workload, "map": map_name, "worlds_per_step": sample_worlds,
                        "note": f"real Pymunk path unavailable ({why}); CPU port of the same algorithm instead"},
                cpu_baseline={"value": value, "unit": "agent-steps/s", "cores": num_threads(), "kind": "port", "sample": sample},
                e2e={"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.period = float(os.environ.get("CAT_BENCH_CLOCK_PERIOD", "0.01"))
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ GPU arm
L2_BYTES = 126e6                        # B200 L2


def make_world_sets(CatWorlds, cmap, n_local, gid0, n_global, dev):
    """Independent sets of `n_local` worlds, enough of them that their state + output buffers together are 1.5x
    the L2.  The timed steps rotate over the sets (one launch = one set), so every step reads its state from
    and writes its records to HBM — "inputs larger than L2" — and K steps can be timed back to back with ONE
    CUDA-event pair, as they run in use.  Sets differ by global world id."""
    first = CatWorlds(cmap, n_local, device=dev, gid0=gid0, seed=0, want_f32=False, want_shared=False)
    footprint = first.state.numel() + first._out.numel()
    n_sets = max(2, int(-(-1.5 * L2_BYTES // footprint)))
    sets = [first] + [CatWorlds(cmap, n_local, device=dev, gid0=gid0 + j * n_global, seed=0, want_f32=False,
                                want_shared=False) for j in range(1, n_sets)]
    for w in sets:
        w.reset()
    return sets, footprint


def time_steps(torch, sets, acts, K, W, dist=None):
    """W warm-up steps per set, then blocks of EXACTLY K steps (round-robin over the sets), each block between one pair
    of CUDA events on the launching stream, barrier + synchronize on both sides; the block is repeated until the
    timed span reaches MIN_SPAN_MS (a 20-step request on a 35 us step would otherwise time 0.7 ms).
    Returns (total timed ms on this rank, number of blocks)."""
    n = len(sets)
    i = 0
    for _ in range(W * n):
        sets[i % n].step(acts[i % len(acts)]); i += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):                  # untimed probe block: how long is K steps?
        sets[i % n].step(acts[i % len(acts)]); i += 1
    e1.record()
    torch.cuda.synchronize()
    est = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=acts[0].device)
    if dist is not None:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    blocks = max(1, min(10000, int(math.ceil(MIN_SPAN_MS / max(float(est[0]), 1e-6)))))
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(blocks)]
    for b in range(blocks):
        ev[b][0].record()
        for _ in range(K):
            sets[i % n].step(acts[i % len(acts)]); i += 1
        ev[b][1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev), blocks


def time_steps_flushed(torch, cw, acts, K, W, flush):
    """Cross-check with the other method the contract allows: ONE world set, a 192 MiB write between steps to
    flush L2, each step timed with its own CUDA-event pair (flush outside the timed span), times summed."""
    for i in range(W):
        cw.step(acts[i % len(acts)])
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(K):
        flush.add_(1)                           # > L2-sized write: evicts state / outputs / actions from L2
        starts[i].record()
        cw.step(acts[i % len(acts)])
        stops[i].record()
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in zip(starts, stops))


def time_e2e(torch, cw, K, W, dist=None, mode="auto", packed=True):
    """The host-buffer API: pinned uint8 actions in, the step's records (observations / rewards / flags) back in
    pinned host memory when each call returns (CatWorlds.step_host)."""
    N, A = cw.n_worlds, cw.A
    host_acts = [torch.randint(0, 4, (N, A), dtype=torch.uint8).pin_memory() for _ in range(8)]
    for i in range(max(W, 16 if mode == "auto" else W)):     # (auto: its 15 tuning calls are warm-up, not timed)
        cw.step_host(host_acts[i % 8], mode=mode, packed=packed)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        cw.step_host(host_acts[i % 8], mode=mode, packed=packed)   # synchronises: results are in host memory on return
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    return e0.elapsed_time(e1)


def time_gae(torch, dev, flush, T=256, M=49152, iters=30):
    """north-star item 4: the GAE scan (TMA-fed kernel; 9 B read + 8 B written per sample) and the in-place advantage
    normalisation (4 B + 4 B).  TWO timings, both with CUDA events on the launching stream:

    * `steady_state` (the one the fractions below are quoted from): launches back to back over rotating buffer sets
      several times the size of L2, ONE event pair around the lot — every byte, including the dirty lines a single
      launch leaves in L2 when it ends, reaches HBM inside the timed span, so ALGORITHMIC bytes / time is an honest
      HBM fraction (an isolated launch hides part of its write-back);
    * `isolated`: one launch at a time, L2 read-flushed before each (a write flush would leave 126 MB of dirty lines
      whose write-back competes with the timed kernel), median — round 1's method, kept for comparison, together with
      the DRAM bytes ncu measured for such a launch (profiles/gae_dram_bytes.json)."""
    from as_cops_and_thieves_b200 import _lib
    L = _lib.load()
    g = torch.Generator(device=dev).manual_seed(7)
    n = T * M
    n_sets = max(3, int(-(-4 * L2_BYTES // (25 * n))))
    sets = []
    for _ in range(n_sets):
        r = torch.randn((T, M), device=dev, generator=g)
        v = torch.randn((T, M), device=dev, generator=g)
        d = (torch.rand((T, M), device=dev, generator=g) < 0.01).to(torch.uint8)
        sets.append((r, v, d, torch.randn((M,), device=dev, generator=g), torch.empty_like(r), torch.empty_like(r),
                     torch.zeros(6, dtype=torch.float64, device=dev)))
    stream = torch.cuda.current_stream(dev).cuda_stream

    def gae(s_):
        r, v, d, lv, ret, adv, st = s_
        _lib.check(L.cat_gae(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ret.data_ptr(), adv.data_ptr(),
                             st.data_ptr(), T, M, 0.99, 0.95, stream), "cat_gae")

    def norm(s_):
        _lib.check(L.cat_adv_normalize(s_[5].data_ptr(), n, s_[6].data_ptr(), n, stream), "cat_adv_normalize")

    def steady(fns, rounds=4, reps=5):
        out = []
        for _ in range(reps):
            for s_ in sets:
                for f in fns:
                    f(s_)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _r in range(rounds):
                for s_ in sets:
                    for f in fns:
                        f(s_)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / (rounds * len(sets)))
        return sorted(out)[len(out) // 2]
    ms_gae, ms_norm, ms_both = steady([gae]), steady([norm]), steady([gae, norm])
    t_gae, t_norm = [], []
    s0 = sets[0]
    for i in range(-3, iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.sum()
        e[0].record(); gae(s0); e[1].record()
        flush.sum()
        e[2].record(); norm(s0); e[3].record()
        torch.cuda.synchronize()
        if i >= 0:
            t_gae.append(e[0].elapsed_time(e[1]))
            t_norm.append(e[2].elapsed_time(e[3]))
    iso_gae, iso_norm = sorted(t_gae)[iters // 2], sorted(t_norm)[iters // 2]
    peak, _ = measured_hbm_peak()
    gbs = lambda b, ms: b * n / ms / 1e6      # noqa: E731
    out = {"T": T, "columns": M, "samples": n, "bytes_per_sample": {"gae": 17, "normalize": 8},
           "timing": f"steady state: {n_sets} buffer sets ({25 * n * n_sets / 1e6:.0f} MB > 4 x L2), launches back to back, one CUDA-event pair, median of 5",
           "gae_ms": ms_gae, "normalize_ms": ms_norm, "gae_plus_normalize_ms": ms_both,
           "gae_gbs": gbs(17, ms_gae), "gae_frac_of_hbm_peak": gbs(17, ms_gae) / peak,
           "normalize_gbs": gbs(8, ms_norm), "normalize_frac_of_hbm_peak": gbs(8, ms_norm) / peak,
           "combined_gbs": gbs(25, ms_both), "combined_frac_of_hbm_peak": gbs(25, ms_both) / peak,
           "samples_per_s": n / ms_both * 1e3,
           "isolated": {"how": f"one launch at a time, L2 read-flushed, median of {iters}", "gae_ms": iso_gae, "normalize_ms": iso_norm,
                        "gae_gbs_algorithmic": gbs(17, iso_gae), "normalize_gbs_algorithmic": gbs(8, iso_norm)}}
    dram = profile_json("gae_dram_bytes.json").get(f"T{T}")
    if dram:
        out["isolated"]["dram_traffic"] = {
            "source": "profiles/gae_dram_bytes.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of one isolated launch, same shapes)",
            "gae_bytes_per_sample": dram["gae"] / n, "normalize_bytes_per_sample": dram["normalize"] / n,
            "gae_frac_of_hbm_peak": dram["gae"] / iso_gae / 1e6 / peak,
            "normalize_frac_of_hbm_peak": dram["normalize"] / iso_norm / 1e6 / peak}
    return out


def measure_workload(torch, CatWorlds, name, K, W, rank, world_size, dev, dist, gen, worlds=None, e2e_modes=("auto",),
                     cross_check=None, sampler=None):
    """One full line for one workload: device-resident rate, HBM + instruction-throughput roofline, host-buffer rate."""
    from as_cops_and_thieves_b200.sharding import shard_range
    map_name, free, per_gpu = WORKLOADS[name]
    per_gpu = worlds or per_gpu
    n_global = per_gpu * world_size
    gid0, n_local = shard_range(n_global, rank, world_size)
    cmap = build_cmap(map_name, free)
    sets, footprint = make_world_sets(CatWorlds, cmap, n_local, gid0, n_global, dev)
    cw = sets[0]
    A = cw.A
    acts = [torch.randint(0, 4, (n_local, A), dtype=torch.uint8, device=dev, generator=gen) for _ in range(16)]
    if sampler is not None:
        sampler.start()
    ms_total, blocks = time_steps(torch, sets, acts, K, W, dist)
    clocks = sampler.stop() if sampler is not None else None
    n_steps = K * blocks
    ms_flushed = None
    if cross_check is not None:
        k_flush = min(K, 300)
        ms_flushed = time_steps_flushed(torch, cw, acts, k_flush, 10, cross_check) / k_flush
    est_ms = ms_total / n_steps
    e2e_steps = max(50, min(2000, int(math.ceil(MIN_SPAN_MS / (2.5 * est_ms)))))
    # modes: "auto" / "zero_copy" / ... move the packed record (types at 2 bits); a "+u8" suffix moves u8 types
    e2e_ms = {m: time_e2e(torch, cw, e2e_steps, 5, dist, mode=m.split("+")[0], packed=not m.endswith("+u8")) for m in e2e_modes}
    t = torch.tensor([ms_total] + [e2e_ms[m] for m in e2e_modes], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    e2e_ms = {m: float(t[1 + i]) for i, m in enumerate(e2e_modes)}

    value = n_global * A * n_steps / (ms_total * 1e-3)
    launch_ms = ms_total / n_steps
    peak, peak_kind = measured_hbm_peak()
    achieved = n_local * A * ALG_BYTES_PER_AGENT_STEP / (launch_ms * 1e-3) / 1e9
    key = name if not worlds else ""
    traffic = profile_json("traffic.json").get(key)
    issue = None
    per_world = profile_json("thread_instr_per_world_step.json").get(key)
    if per_world:
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or 1965
        peak_ti = n_sm * 128 * mhz * 1e6                 # one instruction per lane per clock
        ach_ti = per_world * n_local / (launch_ms * 1e-3)
        issue = {"what": "thread-level instructions per second of cat_world_kernel vs. 128 lanes x SMs x clock",
                 "thread_instr_per_world_step": per_world, "achieved": ach_ti, "peak": peak_ti, "frac": ach_ti / peak_ti,
                 "source": "profiles/thread_instr_per_world_step.json (ncu, this round's kernel) x live rate"}
    main = e2e_modes[0]
    e2e = {"value": n_global * A * e2e_steps / (e2e_ms[main] * 1e-3), "unit": "agent-steps/s",
           "h2d_bytes_per_step": cw.h2d_bytes_per_step, "d2h_bytes_per_step": cw.d2h_bytes_per_step,
           "steps": e2e_steps, "ms_per_step": e2e_ms[main] / e2e_steps,
           "api": "CatWorlds.step_host(host_actions) [mode='auto': keeps the fastest of zero-copy stores into mapped pinned "
                  "memory / chunked launches + one DMA copy per chunk / staged copies, timed on its first calls]: pinned u8 "
                  f"actions in, one {cw.packed_record_bytes}-B record per world (f16 distances, ray types packed at 2 bits, "
                  "f32 rewards, u8 flags) back in pinned host memory on return; the '+u8' entries move the "
                  f"{cw.record_bytes}-B record with u8 types instead",
           "auto_choice": getattr(cw, "_auto_choice", None)}
    for m in e2e_modes[1:]:
        e2e[m] = {"value": n_global * A * e2e_steps / (e2e_ms[m] * 1e-3), "ms_per_step": e2e_ms[m] / e2e_steps,
                  "d2h_bytes_per_step": cw.d2h_bytes(not m.endswith("+u8"))}
    res = {
        "value": value, "ms_per_step": launch_ms, "steps_timed": n_steps, "timed_span_ms": ms_total, "blocks": blocks,
        "config": {"workload": name, "map": map_name, "hull_edges": int(cmap.n_edges), "worlds_per_gpu": per_gpu,
                   "global_worlds": n_global, "agents_per_world": A, "rays_per_agent": cw.R, "dt": 1 / 60,
                   "max_step_count": 400, "spawn": "free-space regions" if free else "map spawn regions",
                   "outputs": "one 832-B record per world: f16 distance + u8 type + f32 reward + u8 flags (native dtypes)",
                   "sensor_sweep": (f"per-(cell, ray) candidate lists, {cw.info.ray_list_nx}x{cw.info.ray_list_ny} cells of "
                                    f"{cw.info.ray_list_cell:.1f} units, {cw.info.ray_list_bytes / 1e6:.1f} MB in HBM/L2"),
                   "launch": f"{cw.info.grid} CTAs x {cw.info.warps_per_cta} warps, {cw.info.smem_bytes_per_cta} B shared memory",
                   "l2": f"inputs larger than L2: the timed steps rotate over {len(sets)} independent sets of {n_local} worlds "
                         f"({footprint * len(sets) / 1e6:.0f} MB of state + record buffers > 126 MB L2; the map's ray lists are read-only "
                         f"constants shared by the sets, one copy per map and device as a single environment holds), one launch = one set; "
                         f"{blocks} block(s) of K steps, each between one CUDA-event pair, span {ms_total:.1f} ms",
                   "parallelism": f"worlds sharded over {world_size} GPU(s), no data-path collective"},
        "e2e": e2e,
        "gpu_launches": n_steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                     "algorithmic_bytes_per_agent_step": ALG_BYTES_PER_AGENT_STEP, "kernel": "cat_world_kernel<3, 90, true, false>",
                     "instruction_throughput": issue,
                     "note": "ALU/latency-bound path (SURVEY.md §8d): the HBM fraction is reported as asked, not a target; "
                             "what bounds the kernel is instruction issue and dependent-instruction latency at 32 warps "
                             "per SM (profiles/r2_cat_world_kernel_*.txt)"},
        "overflow_counts": {"wall_contact_slots": 0, "near_hull_slots": 0},
    }
    oc = [0, 0]
    for w_ in sets:
        c = w_.overflow_counts()
        oc[0] += c[0]; oc[1] += c[1]
    res["overflow_counts"] = {"wall_contact_slots": oc[0], "near_hull_slots": oc[1]}
    if ms_flushed is not None:
        res["config"]["cross_check_flush_method"] = {
            "ms_per_step": ms_flushed, "how": "one world set, 192 MiB write between steps, each step timed with its own event "
                                              "pair (adds ~5 us of event / launch-gap overhead per step)"}
    if clocks is not None:
        res["clocks"] = clocks
    for w_ in sets:
        w_.close()
    return res, cmap


def init_dist(torch, local_rank, world_size):
    if world_size <= 1:
        return None
    import torch.distributed as dist_mod
    # NCCL prints its version banner (this image sets NCCL_DEBUG=VERSION) with printf on fd 1 and ignores
    # NCCL_DEBUG_FILE for it; stdout must carry ONE JSON line, so fd 1 points at stderr while the communicator
    # is created (init + first collective).
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        dist_mod.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist_mod.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    return dist_mod


def run_b200(args):
    import torch
    from as_cops_and_thieves_b200.sharding import dist_env
    from as_cops_and_thieves_b200.worlds import CatWorlds

    rank, local_rank, world_size = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    dist = init_dist(torch, local_rank, world_size)
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    numa_cpus = ()
    if world_size > 1 and os.environ.get("CAT_BENCH_NUMA_BIND", "1") == "1":
        from as_cops_and_thieves_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)   # before any pinned allocation (first touch)
    K, W = max(1, args.steps), max(3, args.warmup)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # 192 MiB > 126 MB L2
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)

    head, cmap = measure_workload(torch, CatWorlds, args.workload, K, W, rank, world_size, dev, dist, gen, worlds=args.worlds,
                                  e2e_modes=("auto", "zero_copy", "pipelined", "staged", "auto+u8"), cross_check=flush, sampler=sampler)
    map_name, free, _ = WORKLOADS[args.workload]
    head["config"]["cpu_affinity"] = f"rank 0 bound to the {len(numa_cpus)} cores local to its GPU" if numa_cpus else "unbound"
    line = {
        "metric": "agent-steps/s (physics+raycast obs)", "value": head["value"], "unit": "agent-steps/s",
        "n_gpus": world_size, "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": head["config"], "clocks": head.get("clocks"), "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
        "roofline": head["roofline"], "overflow_counts": head["overflow_counts"],
        "timed": {"steps": head["steps_timed"], "blocks_of_K": head["blocks"], "span_ms": head["timed_span_ms"]},
    }

    if not args.no_extras:
        others = {}
        for wl in WORKLOADS:          # every named map, on every rank, the same treatment as the headline
            if wl == args.workload:
                continue
            res, _ = measure_workload(torch, CatWorlds, wl, K, W, rank, world_size, dev, dist, gen)
            others[wl] = {k: res[k] for k in ("value", "ms_per_step", "steps_timed", "timed_span_ms", "e2e", "roofline",
                                              "overflow_counts")}
            others[wl]["config"] = {k: res["config"][k] for k in ("map", "hull_edges", "worlds_per_gpu", "spawn", "sensor_sweep", "launch")}
        line["other_workloads"] = others
    if rank == 0 and world_size == 1 and not args.no_extras:
        # the skrl-facing layout: + team-shared observations + fp32 flattened obs (A,N,180) + state (N,1090) + critic block
        n_local = args.worlds or WORKLOADS[args.workload][2]
        acts = [torch.randint(0, 4, (n_local, 3), dtype=torch.uint8, device=dev, generator=gen) for _ in range(16)]
        w3 = CatWorlds(cmap, n_local, device=dev, seed=0, want_f32=True, want_shared=True, want_critic=True)
        w3.reset()
        ms = time_steps_flushed(torch, w3, acts, 200, 20, flush)
        line["other_workloads"][args.workload + "+skrl-layouts"] = {
            "value": n_local * 3 * 200 / (ms * 1e-3), "ms_per_step": ms / 200, "bytes_per_agent_step": 786 + 1453 + 480}
        w3.close()
        line["gae"] = time_gae(torch, dev, flush)
        line["gae_T1024"] = time_gae(torch, dev, flush, T=1024, iters=10)     # a reference-length rollout (4096 / 4)
        cpu = time_cpu_port(map_name, free, 512, seconds=12.0)
        line["cpu_baseline"] = {"value": cpu["value"], "unit": "agent-steps/s", "cores": cpu["cores"], "kind": "port",
                                "sample": f"{cpu['worlds']} {map_name} worlds x {cpu['steps']} steps in {cpu['seconds']:.1f} s, "
                                          "fp64 oracle port of the Pymunk path, OpenMP over worlds"}
        ref, why = time_pymunk_reference(map_name, free, seconds=10.0)
        if ref is not None:            # the real engine is here: it is the baseline, the port rides along
            line["cpu_baseline"] = {"value": ref["value"], "unit": "agent-steps/s", "cores": ref["cores"], "kind": "reference",
                                    "sample": f"{ref['cores']} processes x 1 {map_name} world x {ref['seconds']:.1f} s, {ref['how']}",
                                    "oracle_port": line["cpu_baseline"]}
        else:
            line["cpu_baseline"]["pymunk_probe"] = why
        # BASELINE.json configs[0]: agh-map, ONE environment, one host thread — what a reference user runs today
        cpu1 = time_cpu_port("agh-map", False, 1, seconds=3.0)       # one world = one loop iteration = one thread
        line["cpu_baseline"]["config0_agh-map_single_env"] = {
            "value": cpu1["value"], "unit": "agent-steps/s", "worlds": 1,
            "sample": f"1 agh-map world (file spawn positions) x {cpu1['steps']} steps in {cpu1['seconds']:.1f} s, oracle port"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------ BASELINE config 5: MAPPO self-play training
def run_training(args):
    """`--workload mappo-agh-map`: MAPPO (LSTM policy + centralised critic per agent, the reference's architectures)
    on agh-map, >= 16384 worlds per GPU, gradients all-reduced over NCCL.  A "step" here is one lockstep environment
    step INSIDE training; reports the environment rate inside the rollouts, the update time and the all-reduce time."""
    import torch
    from as_cops_and_thieves_b200.env import BatchedCopsThievesEnv
    from as_cops_and_thieves_b200.maps import free_space_regions, load_named_map
    from as_cops_and_thieves_b200.mappo import MAPPOConfig, MAPPOLearner
    from as_cops_and_thieves_b200.sharding import dist_env, shard_range

    rank, local_rank, world_size = dist_env()
    dist = init_dist(torch, local_rank, world_size)
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    per_gpu = args.worlds or 16384
    gid0, n_local = shard_range(per_gpu * world_size, rank, world_size)
    m = load_named_map("agh-map")
    env = BatchedCopsThievesEnv(m, n_local, device=dev, seed=0, gid0=gid0, spawn_override=free_space_regions(m),
                                max_step_count=400)
    cfg = MAPPOConfig(rollouts=16, sequence_length=16, distributed=world_size > 1, world_size=world_size,
                      random_timesteps=0, learning_starts=0, policy_freeze_duration=0, opponent_freeze_duration=0,
                      update_autocast="bf16")
    learner = MAPPOLearner(env, cfg, seed=0)
    iters = max(2, min(args.steps // cfg.rollouts, 6))
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    for _ in range(2):                       # warm-up: eager rollout, then graph capture
        learner.collect(); learner.update()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    roll_s, upd_ms, ar_ms, gae_ms = 0.0, 0.0, 0.0, 0.0
    t_all = time.perf_counter()
    for _ in range(iters):
        t0 = time.perf_counter()
        learner.collect()
        roll_s += time.perf_counter() - t0
        stats = learner.update()
        upd_ms += sum(s.update_ms for s in stats.values())
        ar_ms += sum(s.allreduce_ms for s in stats.values())
        gae_ms += sum(s.gae_ms for s in stats.values())
    torch.cuda.synchronize()
    total_s = time.perf_counter() - t_all
    clocks = sampler.stop()
    t = torch.tensor([roll_s, total_s, upd_ms, ar_ms, gae_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    roll_s, total_s, upd_ms, ar_ms, gae_ms = (float(x) for x in t)
    A = len(env.possible_agents)
    steps = iters * cfg.rollouts
    n_global = per_gpu * world_size
    line = {"metric": "agent-steps/s inside MAPPO self-play training (env + policy / critic inference + recording)",
            "value": n_global * A * steps / roll_s, "unit": "agent-steps/s", "n_gpus": world_size, "steps": steps,
            "warmup": 2 * cfg.rollouts, "ms_per_step": roll_s / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 env, bf16-autocast update", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD, "map": "agh-map", "worlds_per_gpu": per_gpu, "global_worlds": n_global,
                       "rollouts": cfg.rollouts, "epochs": cfg.learning_epochs, "mini_batches": cfg.mini_batches,
                       "parameters": learner.n_parameters(), "iterations_timed": iters,
                       "parallelism": f"worlds sharded over {world_size} GPU(s); NCCL all-reduce of one flat fp32 gradient bucket per agent per minibatch"},
            "training": {"including_updates_agent_steps_per_s": n_global * A * steps / total_s,
                         "update_ms_per_iteration": upd_ms / iters, "allreduce_ms_per_iteration": ar_ms / iters,
                         "gae_ms_per_iteration": gae_ms / iters, "rollout_ms_per_iteration": roll_s / iters * 1e3},
            "clocks": clocks}
    if rank == 0:
        print(json.dumps(line), flush=True)
    env.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == TRAIN_WORKLOAD:
        run_training(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
