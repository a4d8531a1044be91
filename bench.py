#!/usr/bin/env python
"""Benchmark of the batched cops-and-thieves environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--worlds n]

A "step" is one lockstep transition of every world on a rank: ONE launch of `cat_world_kernel`
(termination test, action impulses, 90-ray sensor sweep per agent, float16 observation chain, rewards,
rigid-body step, auto-reset).  Default workload = BASELINE.json configs[1]: squarinth, 4096 worlds per
GPU, native observation dtypes.  Weak scaling: the worlds per GPU are fixed, ranks shard the global
world range, there is no data-path collective.

Prints ONE JSON line (rank 0).  `value` = agent-steps/s over all ranks with everything resident in
HBM; `e2e` = the same metric through the host-buffer API (pinned actions in, observations / rewards /
flags out, copies inside the timed region).  Cold L2: one step's working set is a few MB, far smaller than the
126 MB L2, so the timed steps rotate over enough independent world sets that their buffers exceed L2 ("inputs
larger than L2"); K steps are timed back to back between one CUDA-event pair.  The flush-between-steps method is
run as a cross-check (config.cross_check_flush_method).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALG_BYTES_PER_AGENT_STEP = 336          # native dtypes, step + raycast obs
FALLBACK_HBM_GBS = 6650.0               # /opt/skills/guides/B200_PROFILING.md fallback
WORKLOADS = {                           # name -> (map, free-space spawns, worlds per GPU)
    "squarinth-4096": ("squarinth", False, 4096),       # BASELINE.json configs[1]  (default)
    "lbirinth-4096": ("lbirinth", False, 4096),         # the fifth named map
    "labyrinth-8192": ("labyrinth", True, 8192),        # configs[2] per-GPU share
    "grandbyrinth-16384": ("grandbyrinth", False, 16384),  # configs[3]
    "agh-map-16384": ("agh-map", True, 16384),          # the real large-segment case (496 edges)
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="squarinth-4096", choices=list(WORKLOADS))
    ap.add_argument("--worlds", type=int, default=None, help="worlds per GPU (overrides the workload's)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other workloads / cpu baseline")
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


# ------------------------------------------------------------------ CPU arm (oracle port, host cores)
def time_cpu_port(map_name, free, n_worlds, seconds=12.0, min_steps=3):
    """The reference's algorithm on the host cores: the fp64 oracle port with OpenMP over worlds.
    (Pymunk / PettingZoo are not installable here, SURVEY.md §8c, so the reference itself cannot run.)"""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (libgomp reads this at load)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    map_name, free, _ = WORKLOADS[args.workload]
    sample_worlds = 512
    K, W = max(1, args.steps), max(0, args.warmup)
    import numpy as np
    from oracle.cat_oracle import Oracle, num_threads
    cmap = build_cmap(map_name, free)
    orc = Oracle(cmap, seed=0)
    st = orc.new_state(sample_worlds)
    out = orc.new_out(sample_worlds)
    orc.reset(st, out=out)
    rng = np.random.default_rng(1)
    acts = [rng.integers(0, 4, (sample_worlds, orc.A)).astype(np.int32) for _ in range(8)]
    # bounded: cap the number of CPU steps so that the whole run ends within a few minutes
    K = min(K, 400)
    W = min(W, 20)
    for i in range(W):
        orc.step(st, acts[i % 8], out)
    t0 = time.perf_counter()
    for i in range(K):
        orc.step(st, acts[i % 8], out)
    dt = time.perf_counter() - t0
    value = sample_worlds * orc.A * K / dt
    sample = f"{sample_worlds} {map_name} worlds x {K} steps, OpenMP over worlds (oracle port of the Pymunk path)"
    line = {
        "impl": "reference", "metric": "agent-steps/s (physics+raycast obs)", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "map": map_name, "worlds_per_step": sample_worlds,
                   "note": "reference Pymunk/PettingZoo path is not installable offline; CPU port of the same algorithm"},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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
    and writes its observations to HBM — "inputs larger than L2" — and K steps can be timed back to back with
    ONE CUDA-event pair, as they run in use, instead of one event pair per step around a flush kernel (which
    adds ~5 us of event / launch-gap overhead to every 40 us step).  Sets differ by global world id."""
    first = CatWorlds(cmap, n_local, device=dev, gid0=gid0, seed=0, want_f32=False, want_shared=False)
    footprint = first.state.numel() + first._out.numel()
    n_sets = max(2, int(-(-1.5 * L2_BYTES // footprint)))
    sets = [first] + [CatWorlds(cmap, n_local, device=dev, gid0=gid0 + j * n_global, seed=0, want_f32=False,
                                want_shared=False) for j in range(1, n_sets)]
    for w in sets:
        w.reset()
    return sets, footprint


def time_steps(torch, sets, acts, K, W, dist=None):
    """W warm-up steps per set, then EXACTLY K steps (round-robin over the sets) between one pair of CUDA events
    on the launching stream, barrier + synchronize on both sides.  Returns the timed milliseconds on this rank."""
    n = len(sets)
    for i in range(W * n):
        sets[i % n].step(acts[i % len(acts)])
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        sets[i % n].step(acts[i % len(acts)])
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


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


def time_e2e(torch, cw, K, W, dist=None, mode="pipelined"):
    """The host-buffer API: pinned uint8 actions in, observations / rewards / flags back in pinned host memory
    when each call returns (CatWorlds.step_host; modes: pipelined chunks + DMA, zero-copy stores, staged copies)."""
    N, A = cw.n_worlds, cw.A
    host_acts = [torch.randint(0, 4, (N, A), dtype=torch.uint8).pin_memory() for _ in range(8)]
    for i in range(W):
        cw.step_host(host_acts[i % 8], mode=mode)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        cw.step_host(host_acts[i % 8], mode=mode)   # synchronises: results are in host memory on return
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    return e0.elapsed_time(e1), cw.h2d_bytes_per_step, cw.d2h_bytes_per_step


def time_gae(torch, dev, flush, T=256, M=49152, iters=30):
    """north-star item 4: the GAE scan (TMA-fed kernel; 9 B read + 8 B written per sample) and the in-place advantage
    normalisation (4 B + 4 B), each timed with its own CUDA events.  L2 is flushed before every launch by READING a 192 MiB buffer
    (a write flush would leave 126 MB of dirty lines whose write-back then competes with the timed kernel)."""
    from as_cops_and_thieves_b200 import _lib
    L = _lib.load()
    g = torch.Generator(device=dev).manual_seed(7)
    r = torch.randn((T, M), device=dev, generator=g)
    v = torch.randn((T, M), device=dev, generator=g)
    d = (torch.rand((T, M), device=dev, generator=g) < 0.01).to(torch.uint8)
    lv = torch.randn((M,), device=dev, generator=g)
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    stats = torch.zeros(2, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    t_gae, t_norm = [], []
    for i in range(-3, iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.sum()      # read-only L2 flush: leaves clean lines, so no write-back of flush data inside the timed span
        e[0].record()
        _lib.check(L.cat_gae(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ret.data_ptr(), adv.data_ptr(),
                             stats.data_ptr(), T, M, 0.99, 0.95, stream), "cat_gae")
        e[1].record()
        flush.sum()
        e[2].record()
        _lib.check(L.cat_adv_normalize(adv.data_ptr(), adv.numel(), stats.data_ptr(), T * M, stream), "cat_adv_normalize")
        e[3].record()
        torch.cuda.synchronize()
        if i >= 0:
            t_gae.append(e[0].elapsed_time(e[1]))       # includes the 16-B stats memset
            t_norm.append(e[2].elapsed_time(e[3]))
    ms_gae, ms_norm = sorted(t_gae)[iters // 2], sorted(t_norm)[iters // 2]     # medians
    n = T * M
    peak, _ = measured_hbm_peak()
    gbs_gae, gbs_norm = 17 * n / ms_gae / 1e6, 8 * n / ms_norm / 1e6
    return {"T": T, "columns": M, "samples": n, "timing": f"median of {iters} launches, CUDA events, inputs > L2 and L2 read-flushed",
            "gae_ms": ms_gae, "normalize_ms": ms_norm,
            "gae_gbs": gbs_gae, "gae_frac_of_hbm_peak": gbs_gae / peak, "normalize_gbs": gbs_norm,
            "normalize_frac_of_hbm_peak": gbs_norm / peak, "combined_gbs": 25 * n / (ms_gae + ms_norm) / 1e6,
            "combined_frac_of_hbm_peak": 25 * n / (ms_gae + ms_norm) / 1e6 / peak,
            "bytes_per_sample": {"gae": 17, "normalize": 8}, "samples_per_s": n / (ms_gae + ms_norm) * 1e3}


def run_b200(args):
    import torch
    from as_cops_and_thieves_b200.sharding import dist_env, shard_range
    from as_cops_and_thieves_b200.worlds import CatWorlds

    rank, local_rank, world_size = dist_env()
    dist = None
    if world_size > 1:
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
        dist = dist_mod
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    numa_cpus = ()
    if world_size > 1 and os.environ.get("CAT_BENCH_NUMA_BIND", "1") == "1":
        from as_cops_and_thieves_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)   # before any pinned allocation (first touch)

    map_name, free, per_gpu = WORKLOADS[args.workload]
    per_gpu = args.worlds or per_gpu
    n_global = per_gpu * world_size
    gid0, n_local = shard_range(n_global, rank, world_size)
    K, W = max(1, args.steps), max(3, args.warmup)

    cmap = build_cmap(map_name, free)
    sets, footprint = make_world_sets(CatWorlds, cmap, n_local, gid0, n_global, dev)
    cw = sets[0]
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    acts = [torch.randint(0, 4, (n_local, cw.A), dtype=torch.uint8, device=dev, generator=g) for _ in range(16)]
    flush = torch.zeros(192 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # 192 MiB > 126 MB L2

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    sampler.start()
    ms_total = time_steps(torch, sets, acts, K, W, dist)
    clocks = sampler.stop()
    k_flush = min(K, 500)
    ms_flushed = time_steps_flushed(torch, cw, acts, k_flush, 10, flush) / k_flush
    e2e_steps = min(K, 1000)
    e2e_ms, h2d, d2h = time_e2e(torch, cw, e2e_steps, 5, dist, mode="zero_copy")
    e2e_pipe_ms, _, _ = time_e2e(torch, cw, e2e_steps, 5, dist, mode="pipelined")
    e2e_staged_ms, _, _ = time_e2e(torch, cw, e2e_steps, 5, dist, mode="staged")

    t = torch.tensor([ms_total, e2e_ms, e2e_pipe_ms, e2e_staged_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_pipe_ms, e2e_staged_ms = (float(x) for x in t)

    A = cw.A
    value = n_global * A * K / (ms_total * 1e-3)
    e2e_value = n_global * A * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_kind = measured_hbm_peak()
    launch_ms = ms_total / K
    alg_bytes = n_local * A * ALG_BYTES_PER_AGENT_STEP
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"          # dram bytes per launch from the committed ncu --set full capture
    if tf.exists():
        try:
            traffic = json.load(open(tf)).get(args.workload if not args.worlds else "", None)
        except Exception:
            traffic = None

    issue = None
    ti = ROOT / "profiles" / "thread_instr_per_world_step.json"   # from the committed ncu capture of this workload
    if ti.exists() and not args.worlds:
        try:
            per_world = json.load(open(ti)).get(args.workload)
            if per_world:
                n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
                peak_ti = n_sm * 128 * mhz * 1e6                 # one instruction per lane per clock
                ach_ti = per_world * n_local / (launch_ms * 1e-3)
                issue = {"what": "thread-level instructions per second of cat_world_kernel vs. 128 lanes x SMs x clock",
                         "thread_instr_per_world_step": per_world, "achieved": ach_ti, "peak": peak_ti,
                         "frac": ach_ti / peak_ti, "source": "profiles/thread_instr_per_world_step.json (ncu) x live rate"}
        except Exception:
            issue = None

    line = {
        "metric": "agent-steps/s (physics+raycast obs)", "value": value, "unit": "agent-steps/s",
        "n_gpus": world_size, "steps": K, "warmup": W, "ms_per_step": launch_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "map": map_name, "hull_edges": int(cmap.n_edges), "worlds_per_gpu": per_gpu,
                   "global_worlds": n_global, "agents_per_world": A, "rays_per_agent": cw.R, "dt": 1 / 60,
                   "max_step_count": 400, "spawn": "free-space regions" if free else "map spawn regions",
                   "outputs": "f16 distance + u8 type + f32 reward + u8 flags (native dtypes)",
                   "l2": f"inputs larger than L2: the timed steps rotate over {len(sets)} independent sets of {n_local} worlds "
                         f"({footprint * len(sets) / 1e6:.0f} MB of state + output buffers > 126 MB L2), one launch = one set; "
                         "K steps timed back to back with one CUDA-event pair",
                   "cross_check_flush_method": {"ms_per_step": ms_flushed, "steps": k_flush,
                                                "how": "one world set, 192 MiB write between steps, each step timed with "
                                                       "its own event pair (adds ~5 us of event / launch-gap overhead per step)"},
                   "parallelism": f"worlds sharded over {world_size} GPU(s), no data-path collective",
                   "cpu_affinity": f"rank 0 bound to the {len(numa_cpus)} cores local to its GPU" if numa_cpus else "unbound"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "api": "CatWorlds.step_host(mode='zero_copy'): ONE launch; the kernel reads pinned u8 actions and stores f16/u8 "
                       "observations, f32 rewards, u8 flags straight into mapped pinned host memory with 16-byte stores "
                       "(transfer overlaps compute); one stream sync per step",
                "pipelined": {"value": n_global * A * e2e_steps / (e2e_pipe_ms * 1e-3), "ms_per_step": e2e_pipe_ms / e2e_steps,
                              "chunks": cw.default_chunks(),
                              "api": "step_host(mode='pipelined') -> cat_env_step_host: chunked launches, per-chunk DMA on a copy stream"},
                "staged_copy": {"value": n_global * A * e2e_steps / (e2e_staged_ms * 1e-3), "ms_per_step": e2e_staged_ms / e2e_steps,
                                "api": "step_host(mode='staged'): H2D copy, launch, one D2H copy of the output blob"}},
        "gpu_launches": K,          # one cat_world_kernel launch per timed step (the e2e legs launch their own)
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                     "algorithmic_bytes_per_agent_step": ALG_BYTES_PER_AGENT_STEP, "kernel": "cat_world_kernel",
                     "instruction_throughput": issue,
                     "note": "ALU/latency-bound path (SURVEY.md §8d): HBM fraction is reported as asked, not a target; "
                             "what bounds the kernel is instruction issue (ncu: 68-76 % of issue slots busy, 23-25 of 32 "
                             "lanes active; profiles/r1_cat_world_kernel_*.txt)"},
    }

    if not args.no_extras:
        # every named map, on every rank (weak scaling: the same worlds per GPU), max over ranks like the headline
        others = {}
        for wl, (mn, fr, nw) in WORKLOADS.items():
            if wl == args.workload:
                continue
            c2 = build_cmap(mn, fr)
            g0, nl = shard_range(nw * world_size, rank, world_size)
            sets2, _ = make_world_sets(CatWorlds, c2, nl, g0, nw * world_size, dev)
            w2 = sets2[0]
            a2 = [torch.randint(0, 4, (nl, w2.A), dtype=torch.uint8, device=dev, generator=g) for _ in range(8)]
            ms = time_steps(torch, sets2, a2, 300, 5, dist)
            if dist is not None:
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt[0])
            others[wl] = {"value": nw * world_size * w2.A * 300 / (ms * 1e-3), "ms_per_step": ms / 300,
                          "hull_edges": int(c2.n_edges), "worlds_per_gpu": nw}
            for w_ in sets2:
                w_.close()
        line["other_workloads"] = others
    if rank == 0 and world_size == 1 and not args.no_extras:
        # the skrl-facing layout: + team-shared observations + fp32 flattened obs (A,N,180) + state (N,1090)
        w3 = CatWorlds(cmap, n_local, device=dev, seed=0, want_f32=True, want_shared=True)
        w3.reset()
        ms = time_steps_flushed(torch, w3, acts, 300, 20, flush)
        line["other_workloads"][args.workload + "+skrl-layouts"] = {"value": n_local * A * 300 / (ms * 1e-3), "ms_per_step": ms / 300,
                                                                    "bytes_per_agent_step": 786 + 1453}
        w3.close()
        line["gae"] = time_gae(torch, dev, flush)
        line["gae_T1024"] = time_gae(torch, dev, flush, T=1024, iters=10)     # a reference-length rollout (4096 / 4)
        cpu = time_cpu_port(map_name, free, 512, seconds=12.0)
        line["cpu_baseline"] = {"value": cpu["value"], "unit": "agent-steps/s", "cores": cpu["cores"], "kind": "port",
                                "sample": f"{cpu['worlds']} {map_name} worlds x {cpu['steps']} steps in {cpu['seconds']:.1f} s, "
                                          "fp64 oracle port of the Pymunk path, OpenMP over worlds"}
        # BASELINE.json configs[0]: agh-map, ONE environment, one host thread — what a reference user runs today
        cpu1 = time_cpu_port("agh-map", False, 1, seconds=3.0)       # one world = one loop iteration = one thread
        line["cpu_baseline"]["config0_agh-map_single_env"] = {
            "value": cpu1["value"], "unit": "agent-steps/s", "worlds": 1,
            "sample": f"1 agh-map world (file spawn positions) x {cpu1['steps']} steps in {cpu1['seconds']:.1f} s, oracle port"}
    elif rank == 0:
        line["cpu_baseline"] = None
    for w_ in sets:
        w_.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
