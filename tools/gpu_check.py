#!/usr/bin/env python
"""Developer diagnostic (GPU box): CUDA step vs CPU oracle on a few hundred worlds per map, with a
verbose mismatch report, plus a quick timing.  The real parity suite lives in tests/."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from as_cops_and_thieves_b200.worlds import CatWorlds  # noqa: E402
from oracle.cat_oracle import Oracle  # noqa: E402
import parity_utils as pu  # noqa: E402


def compare_once(name, cw, orc, steps_before, rng, label):
    N, A = cw.n_worlds, cw.A
    for _ in range(steps_before):
        cw.step(torch.from_numpy(rng.integers(0, 4, (N, A)).astype(np.uint8)).to(cw.device))
    st = cw.get_state()
    torch.cuda.synchronize()
    ost = pu.cuda_state_to_oracle(orc, st)
    acts = rng.integers(0, 4, (N, A))
    base_obs = orc.observe(ost)  # observation the oracle sees after the action impulses? no: pre-action == same positions
    cw.step(torch.from_numpy(acts.astype(np.uint8)).to(cw.device))
    torch.cuda.synchronize()
    oout = orc.step(ost, acts)
    st2 = cw.get_state()
    torch.cuda.synchronize()

    done = oout.terminated.astype(bool)
    nd = ~done
    # --- rays (worlds that did not reset, so both sides observed the pre-physics state)
    hp = cw.hit_point.cpu().numpy()
    otype = cw.obs_type.cpu().numpy()
    odist = cw.obs_dist.cpu().numpy()
    unstable = pu.ray_unstable_mask(orc, pu.cuda_state_to_oracle(orc, st), base_obs.hit_alpha, base_obs.obs_type)
    type_mis = (otype != oout.obs_type) & nd[:, None, None]
    dd = np.linalg.norm(hp - oout.hit_point, axis=-1)
    hit = oout.obs_type != 4
    dist_mis = (dd > pu.RAY_ATOL) & hit & nd[:, None, None]
    print(f"[{name}/{label}] worlds {N} done {done.sum()} rays {otype.size} unstable {unstable.mean():.4%}")
    print(f"   type mismatches: {type_mis.sum()} (stable: {(type_mis & ~unstable).sum()})  "
          f"hit-point >{pu.RAY_ATOL}: {dist_mis.sum()} (stable: {(dist_mis & ~unstable).sum()})  "
          f"max stable hit err {dd[hit & nd[:, None, None] & ~unstable & ~type_mis].max() if (hit & nd[:, None, None] & ~unstable & ~type_mis).any() else 0:.2e}")
    bad = np.argwhere((type_mis | dist_mis) & ~unstable)
    for w, a, r in bad[:6]:
        print(f"     w{w} a{a} r{r}: cuda type {otype[w, a, r]} d {odist[w, a, r]} hp {hp[w, a, r]} | oracle type "
              f"{oout.obs_type[w, a, r]} d {oout.obs_dist[w, a, r]} hp {oout.hit_point[w, a, r]} alpha {oout.hit_alpha[w, a, r]:.6f}"
              f" pos {ost.pos[w, a] if False else st['pos'][w, a].cpu().numpy()}")
    # f16 chain bit-exactness given CUDA's own fp32 hit points
    pos_before = st["pos"].cpu().numpy()
    chain = pu.f16_chain_numpy(hp, pos_before)
    chain = np.where(otype == 4, np.float16(orc.params["ray_length"]), chain)
    exact = (chain.view(np.uint16) == odist.view(np.uint16)) | ~nd[:, None, None]
    print(f"   f16 chain bit-exact vs numpy on CUDA hit points: {exact.mean():.6%} ({(~exact).sum()} differ)")
    f16_vs_oracle = (odist.view(np.uint16) != oout.obs_dist.view(np.uint16)) & nd[:, None, None] & ~unstable
    print(f"   f16 distances differing from oracle (stable rays): {f16_vs_oracle.sum()} ({f16_vs_oracle.mean():.4%})")
    # --- flags / rewards
    for k in ("terminated", "truncated", "winner"):
        c = getattr(cw, k).cpu().numpy()
        o = getattr(oout, k)
        print(f"   {k}: mismatches {(c != o).sum()} (cuda sum {int((c != 0).sum())})")
    rw = cw.reward.cpu().numpy()
    print(f"   reward max abs diff vs oracle: {np.abs(rw - oout.reward).max():.3e}")
    # --- physics
    for k in ("pos", "vel", "vbias"):
        c = st2[k].cpu().numpy().astype(np.float64)
        o = getattr(ost, k)
        err = np.abs(c - o) / np.maximum(1.0, np.abs(o))
        errn = err[nd] if k != "vbias" else err
        print(f"   {k}: max rel err {errn.max() if errn.size else 0:.3e}  (>1e-4: {(errn > 1e-4).sum()})")
        if (errn > 1e-4).any() and k != "vbias":
            w, a, c_ = np.argwhere((err > 1e-4) & nd[:, None, None])[0]
            print(f"      e.g. w{w} a{a}: cuda {st2[k][w].cpu().numpy().tolist()} oracle {o[w].tolist()}")
            print(f"           before pos {pos_before[w].tolist()} vel {st['vel'][w].cpu().numpy().tolist()} acts {acts[w]}")
    sc = st2["step_count"].cpu().numpy()
    print(f"   step_count mismatches {(sc != ost.step_count).sum()}; episode mismatches "
          f"{(st2['episode'].cpu().numpy().astype(np.uint32) != ost.episode).sum()}")
    if done.any():
        c = st2["pos"].cpu().numpy().astype(np.float64)[done]
        print(f"   reset worlds: pos max abs diff {np.abs(c - ost.pos[done]).max():.3e}")
    ncon = (st2["wall_hull"].cpu().numpy() >= 0).sum()
    print(f"   cached wall arbiters {ncon} (oracle {(ost.wall_age >= 0).sum()}), pair {(st2['pair_age'].cpu().numpy() >= 0).sum()} "
          f"(oracle {np.triu(ost.pair_age >= 0, 1).sum() if False else (ost.pair_age >= 0).sum()})")


def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    rng = np.random.default_rng(0)
    for name, free in (("squarinth", False), ("lbirinth", False), ("grandbyrinth", False), ("labyrinth", True),
                       ("agh-map", True), ("agh-map", False)):
        cmap = pu.named_cmap(name, free_spawn=free)
        N = 256
        cw = CatWorlds(cmap, N, want_hits=True, seed=3)
        orc = Oracle(cmap, seed=3)
        info = cw.info
        print(f"== {name} free={free}: H {info.n_hulls} E {info.n_edges} blob {info.map_blob_bytes} B smem/CTA "
              f"{info.smem_bytes_per_cta} B grid {info.grid} rec_words {info.record_words} S {info.state_dim}")
        cw.reset()
        torch.cuda.synchronize()
        compare_once(name, cw, orc, 0, rng, "after-reset")
        compare_once(name, cw, orc, 60, rng, "t=60")
        compare_once(name, cw, orc, 300, rng, "t=360")
        cw.close()
    # quick timing
    for name, free, N in (("squarinth", False, 4096), ("squarinth", False, 16384), ("agh-map", True, 16384),
                          ("labyrinth", True, 8192), ("grandbyrinth", False, 16384)):
        cmap = pu.named_cmap(name, free_spawn=free)
        cw = CatWorlds(cmap, N, want_f32=False)
        cw.reset()
        acts = [torch.randint(0, 4, (N, cw.A), dtype=torch.uint8, device="cuda") for _ in range(8)]
        for i in range(20):
            cw.step(acts[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 200
        for i in range(K):
            cw.step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"timing {name} N={N}: {ms * 1e3:.1f} us/step -> {N * cw.A / ms * 1e3:.3e} agent-steps/s (grid {cw.info.grid})")
        cw.close()


if __name__ == "__main__":
    main()
