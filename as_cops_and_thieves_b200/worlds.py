"""``CatWorlds`` — N lockstep worlds on one GPU behind the C ABI (``include/cat_b200.h``).

Thin host object: owns the PyTorch tensors (packed state buffer + outputs), builds the
``CatStepIO`` pointer table once, and enqueues one kernel launch per ``reset`` / ``step`` /
``observe`` on the current torch CUDA stream.  No compute happens in Python and there is no CPU
path: constructing it without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import CatEnvInfo, CatMapDesc, CatParams, CatStateView, CatStepIO, CAT_WALL_SLOTS
from .maps import CompiledMap
from .params import EnvParams

ActionsLike = Union[torch.Tensor, Mapping[str, torch.Tensor], Sequence[torch.Tensor]]


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda" or not torch.cuda.is_available():
        raise _lib.CatError(
            "as_cops_and_thieves_b200 needs a CUDA device (sm_100a); there is no CPU fallback "
            f"(requested device={device}, torch.cuda.is_available()={torch.cuda.is_available()})")


class CatWorlds:
    def __init__(self, cmap: CompiledMap, n_worlds: int, *, device: Union[str, torch.device] = "cuda:0",
                 gid0: int = 0, params: Optional[EnvParams] = None, want_f32: bool = True,
                 want_shared: bool = True, want_hits: bool = False, pinned_outputs: bool = False, **overrides):
        """``pinned_outputs=True`` places every output tensor in mapped pinned HOST memory: the kernel stores its
        results there directly and the host reads them (``tensor.numpy()``) after one stream synchronisation —
        the layout for callers that consume every step on the CPU, like the single-world PettingZoo face."""
        self.device = torch.device(device)
        _require_cuda(self.device)
        self.L = _lib.load()
        self.cmap = cmap
        p = (params or EnvParams()).as_dict()
        p.update(overrides)
        self.params = p
        self.n_worlds = int(n_worlds)
        self.gid0 = int(gid0)

        self._keep = dict(
            hull_off=np.ascontiguousarray(cmap.hull_off, np.int32),
            vert=np.ascontiguousarray(cmap.vert, np.float64),
            normal=np.ascontiguousarray(cmap.normal, np.float64),
            edge_len=np.ascontiguousarray(cmap.edge_len, np.float64),
            hull_bb=np.ascontiguousarray(cmap.hull_bb, np.float64),
            init_pos=np.ascontiguousarray(cmap.init_pos, np.float64),
            region_off=np.ascontiguousarray(cmap.region_off, np.int32),
            regions=np.ascontiguousarray(cmap.regions if len(cmap.regions) else np.zeros((1, 4)), np.float64),
            con_cell_off=np.ascontiguousarray(cmap.con_cell_off, np.int32),
            con_cell_hulls=np.ascontiguousarray(np.append(cmap.con_cell_hulls, 0), np.int32),
            view_cell_off=np.ascontiguousarray(cmap.view_cell_off, np.int32),
            view_cell_edges=np.ascontiguousarray(np.append(cmap.view_cell_edges, 0), np.int32),
        )
        k = self._keep
        md = CatMapDesc(
            cmap.n_hulls, cmap.n_edges, _lib.np_ptr(k["hull_off"]), _lib.np_ptr(k["vert"]), _lib.np_ptr(k["normal"]),
            _lib.np_ptr(k["edge_len"]), _lib.np_ptr(k["hull_bb"]), cmap.n_cops, cmap.n_thieves,
            _lib.np_ptr(k["init_pos"]), _lib.np_ptr(k["region_off"]), _lib.np_ptr(k["regions"]),
            cmap.grid_x0, cmap.grid_y0, cmap.cell, cmap.nx, cmap.ny,
            _lib.np_ptr(k["con_cell_off"]), _lib.np_ptr(k["con_cell_hulls"]),
            _lib.np_ptr(k["view_cell_off"]) if len(k["view_cell_off"]) == cmap.nx * cmap.ny + 1 else None,
            _lib.np_ptr(k["view_cell_edges"]), float(cmap.view_range))
        cp = CatParams(**{name: p[name] for name, _ in CatParams._fields_})
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self.L.cat_env_create(C.byref(md), C.byref(cp), self.n_worlds, self.gid0, dev_index,
                                         C.byref(handle)), "cat_env_create")
        self._h = handle
        info = CatEnvInfo()
        _lib.check(self.L.cat_env_info(self._h, C.byref(info)), "cat_env_info")
        self.info = info
        self.A, self.R, self.S, self.P = info.n_agents, info.n_rays, info.state_dim, info.n_pairs
        self.n_cops, self.n_thieves = info.n_cops, info.n_thieves

        N, A, R, dev = self.n_worlds, self.A, self.R, self.device
        nbytes = int(self.L.cat_env_state_bytes(self._h))
        self.state = torch.zeros(nbytes, dtype=torch.uint8, device=dev)   # packed per-world records
        # The per-step results live in ONE device buffer so the host-facing path needs a single D2H copy.
        def carve(nbytes_, off=[0]):
            o = off[0]
            off[0] = (o + nbytes_ + 255) // 256 * 256
            return o
        sizes = dict(obs_dist=N * A * R * 2, obs_type=N * A * R, reward=N * A * 4, terminated=N, truncated=N, winner=N)
        offs = {k_: carve(v) for k_, v in sizes.items()}
        total = carve(0)
        self.pinned_outputs = bool(pinned_outputs)

        def alloc(shape, dtype):
            if self.pinned_outputs:
                return torch.zeros(shape, dtype=dtype).pin_memory()
            return torch.zeros(shape, dtype=dtype, device=dev)
        self._out = alloc(total, torch.uint8)

        def view(name, dtype, shape):
            return self._out[offs[name]:offs[name] + sizes[name]].view(dtype).view(shape)
        self._out_layout = (offs, sizes, total)
        self.obs_dist = view("obs_dist", torch.float16, (N, A, R))
        self.obs_type = view("obs_type", torch.uint8, (N, A, R))
        self.reward = view("reward", torch.float32, (N, A))
        self.terminated = view("terminated", torch.uint8, (N,))
        self.truncated = view("truncated", torch.uint8, (N,))
        self.winner = view("winner", torch.int8, (N,))
        self.winner.fill_(-1)
        self.shared_dist = alloc((N, 2, R), torch.float16) if want_shared else None
        self.shared_type = alloc((N, 2, R), torch.uint8) if want_shared else None
        self.team_pos = alloc((N, A, 2), torch.float16) if want_shared else None
        self.obs_f32 = alloc((A, N, 2 * R), torch.float32) if want_f32 else None
        self.state_f32 = alloc((N, self.S), torch.float32) if want_f32 else None
        self.hit_point = alloc((N, A, R, 2), torch.float32) if want_hits else None
        self._host = None
        self._io = self._make_io()
        self._ptr_table = (C.c_void_p * 8)()
        _lib.check(self.L.cat_env_init_state(self._h, self.state.data_ptr(), self._stream()), "cat_env_init_state")

    # ------------------------------------------------------------------ plumbing
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _make_io(self) -> CatStepIO:
        def dp(t):
            return None if t is None else t.data_ptr()
        return CatStepIO(None, 0, None, dp(self.obs_dist), dp(self.obs_type), dp(self.reward), dp(self.terminated),
                         dp(self.truncated), dp(self.winner), dp(self.shared_dist), dp(self.shared_type),
                         dp(self.team_pos), dp(self.obs_f32), dp(self.state_f32), dp(self.hit_point), 0, 0)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.L.cat_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_seed(self, seed: int) -> None:
        _lib.check(self.L.cat_env_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), "cat_env_set_seed")
        self.params["seed"] = int(seed)

    # ------------------------------------------------------------------ the three launches
    def reset(self, mask: Optional[torch.Tensor] = None) -> None:
        """``BaseEnv.reset`` for the masked worlds (all if ``mask`` is None); fills the obs outputs."""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert mask.numel() == self.n_worlds
        self._io.reset_mask = None if mask is None else mask.data_ptr()
        _lib.check(self.L.cat_env_reset(self._h, self.state.data_ptr(), C.byref(self._io), self._stream()),
                   "cat_env_reset")
        self._io.reset_mask = None

    def step(self, actions: ActionsLike) -> None:
        """``BaseEnv.step`` for every world: one kernel launch on the current stream."""
        io = self._io
        if isinstance(actions, torch.Tensor):
            on_device = actions.device == self.device or (actions.device.type == "cpu" and actions.is_pinned())
            if not on_device or not actions.is_contiguous() or actions.numel() != self.n_worlds * self.A:
                raise ValueError("actions must be a contiguous (N, A) tensor on the env's device (or in pinned host memory)")
            kind = {torch.uint8: 0, torch.int32: 1, torch.int64: 2}.get(actions.dtype)
            if kind is None:
                raise ValueError(f"unsupported action dtype {actions.dtype}")
            io.actions, io.actions_kind = actions.data_ptr(), kind
        else:
            seq = list(actions.values()) if isinstance(actions, Mapping) else list(actions)
            if len(seq) != self.A:
                raise ValueError(f"need {self.A} per-agent action tensors")
            for a, t in enumerate(seq):
                if t.dtype != torch.int64 or t.device != self.device or not t.is_contiguous() or t.numel() != self.n_worlds:
                    raise ValueError("per-agent actions must be contiguous int64 tensors of N elements on the env's device")
                self._ptr_table[a] = t.data_ptr()
            io.actions, io.actions_kind = C.cast(self._ptr_table, C.c_void_p), 3
        _lib.check(self.L.cat_env_step(self._h, self.state.data_ptr(), C.byref(io), self._stream()), "cat_env_step")

    # ------------------------------------------------------------------ host-buffer face (what a CPU caller sees)
    @property
    def h2d_bytes_per_step(self) -> int:
        return self.n_worlds * self.A

    @property
    def d2h_bytes_per_step(self) -> int:
        """Bytes of results that cross to the host per ``step_host`` (observations, rewards, flags)."""
        return self.n_worlds * (self.A * self.R * 3 + self.A * 4 + 3)

    def _host_buffers(self, zero_copy: bool) -> Dict[str, torch.Tensor]:
        key = "zc" if zero_copy else "staged"
        if self._host is None:
            self._host = {}
        if key in self._host:
            return self._host[key]
        N, A, R = self.n_worlds, self.A, self.R
        if not zero_copy:
            offs, sizes, total = self._out_layout
            hb = torch.zeros(total, dtype=torch.uint8).pin_memory()

            def hview(name, dtype, shape):
                return hb[offs[name]:offs[name] + sizes[name]].view(dtype).view(shape)
            h = dict(blob=hb, actions_dev=torch.zeros((N, A), dtype=torch.uint8, device=self.device),
                     actions_pinned=torch.zeros((N, A), dtype=torch.uint8).pin_memory(),
                     obs_dist=hview("obs_dist", torch.float16, (N, A, R)),
                     obs_type=hview("obs_type", torch.uint8, (N, A, R)),
                     reward=hview("reward", torch.float32, (N, A)),
                     terminated=hview("terminated", torch.uint8, (N,)),
                     truncated=hview("truncated", torch.uint8, (N,)),
                     winner=hview("winner", torch.int8, (N,)))
            h["io"] = CatStepIO(None, 0, None, h["obs_dist"].data_ptr(), h["obs_type"].data_ptr(), h["reward"].data_ptr(),
                                h["terminated"].data_ptr(), h["truncated"].data_ptr(), h["winner"].data_ptr(),
                                None, None, None, None, None, None, 0, 0)
            h["dev_io"] = CatStepIO(None, 0, None, self.obs_dist.data_ptr(), self.obs_type.data_ptr(), self.reward.data_ptr(),
                                    self.terminated.data_ptr(), self.truncated.data_ptr(), self.winner.data_ptr(),
                                    None, None, None, None, None, None, 0, 0)
        else:
            # Pinned host memory is mapped into the device's address space (unified virtual addressing), so the
            # kernel can read the actions from it and store its results straight into it.  Each world's
            # observation block starts on a 16-byte boundary (world stride rounded up), so the kernel ships it
            # with 512-byte warp stores, which the PCIe root port sees as full-size writes; the (N, A, R)
            # tensors handed back are strided views of that buffer.
            layout = getattr(self, "_zc_layout", "record128")
            d16, t16 = (A * R * 2 + 15) // 16 * 16, (A * R + 15) // 16 * 16
            if layout.startswith("record"):
                # one record per world: [distance f16 | type u8 | pad], record size a multiple of 128 B
                al = int(layout[6:] or 128)
                ds = ts = (d16 + t16 + al - 1) // al * al
                sizes = dict(obs=N * ds, reward=N * A * 4, terminated=N, truncated=N, winner=N)
            else:                                        # "split<align>": two arrays, world stride rounded up
                al = int(layout[5:] or 16)
                ds, ts = (d16 + al - 1) // al * al, (t16 + al - 1) // al * al
                sizes = dict(obs_dist=N * ds, obs_type=N * ts, reward=N * A * 4, terminated=N, truncated=N, winner=N)
            offs, total = {}, 0
            for name, nb in sizes.items():
                offs[name] = total
                total = (total + nb + 255) // 256 * 256
            hb = torch.zeros(total, dtype=torch.uint8).pin_memory()

            def flat(name, dtype, skip=0):
                return hb[offs[name] + skip:offs[name] + sizes[name]].view(dtype)
            if layout.startswith("record"):
                od = hb[offs["obs"]:offs["obs"] + sizes["obs"]].view(torch.float16).as_strided((N, A, R), (ds // 2, R, 1))
                ot = hb[offs["obs"]:offs["obs"] + sizes["obs"]].as_strided((N, A, R), (ts, R, 1), d16)
            else:
                od = flat("obs_dist", torch.float16).as_strided((N, A, R), (ds // 2, R, 1))
                ot = flat("obs_type", torch.uint8).as_strided((N, A, R), (ts, R, 1))
            h = dict(blob=hb, actions_pinned=torch.zeros((N, A), dtype=torch.uint8).pin_memory(),
                     obs_dist=od, obs_type=ot,
                     reward=flat("reward", torch.float32).view(N, A),
                     terminated=flat("terminated", torch.uint8), truncated=flat("truncated", torch.uint8),
                     winner=flat("winner", torch.int8))
            h["io"] = CatStepIO(None, 0, None, h["obs_dist"].data_ptr(), h["obs_type"].data_ptr(), h["reward"].data_ptr(),
                                h["terminated"].data_ptr(), h["truncated"].data_ptr(), h["winner"].data_ptr(),
                                None, None, None, None, None, None, ds, ts)
        self._host[key] = h
        return h

    def default_chunks(self) -> int:
        """Chunks for the pipelined host path.  Every chunk costs six small DMA copies (~8 us each), so chunking
        only pays once a chunk's kernel is long; measured on B200 (gpurun_out/zc_layouts2.log): never below 8192
        worlds, 2 chunks at 16384 agh-map worlds."""
        return 2 if self.n_worlds >= 16384 else 1

    def step_host(self, host_actions: torch.Tensor, mode: str = "zero_copy", chunks: Optional[int] = None,
                  zero_copy: Optional[bool] = None) -> Dict[str, torch.Tensor]:
        """``step`` for a caller whose buffers live in host memory (the reference's own calling convention):
        uint8 actions ``(N, A)`` in, and the step's observations, rewards and flags in pinned host memory
        when the call returns.  Three ways to move the results, same bytes, same values:

        * ``"zero_copy"`` (default, fastest measured: 116 us for 4096 squarinth worlds, 435 us for 16384 agh-map
          worlds) — one launch whose 16-byte stores go straight into mapped pinned host memory, so the
          transfer of one world's results overlaps the computation of the others; SM-issued PCIe writes
          sustain ~30 GB/s.
        * ``"pipelined"`` — ``cat_env_step_host``: the worlds are stepped in ``chunks`` launches and each chunk's
          results are DMA-copied (~50 GB/s, but ~8 us fixed cost per copy, six arrays per chunk) on a second
          stream while the next chunk computes; the kernel reads the actions straight from the pinned buffer
          (143 us / 506 us on the same two workloads).
        * ``"staged"`` — H2D copy, one launch, one D2H copy of the output blob, strictly in sequence
          (135 us / 544 us).
        """
        if zero_copy is not None:                       # older spelling
            mode = "zero_copy" if zero_copy else "staged"
        if mode not in ("pipelined", "zero_copy", "staged"):
            raise ValueError(f"unknown step_host mode {mode!r}")
        h = self._host_buffers(mode == "zero_copy")
        if mode != "staged":
            if not host_actions.is_pinned():
                h["actions_pinned"].copy_(host_actions)
                host_actions = h["actions_pinned"]
            if host_actions.dtype != torch.uint8 or not host_actions.is_contiguous() or host_actions.numel() != self.n_worlds * self.A:
                raise ValueError("host actions must be a contiguous uint8 (N, A) tensor")
            io = h["io"]
            io.actions, io.actions_kind = host_actions.data_ptr(), 0
            if mode == "zero_copy":
                _lib.check(self.L.cat_env_step(self._h, self.state.data_ptr(), C.byref(io), self._stream()), "cat_env_step")
            else:
                _lib.check(self.L.cat_env_step_host(self._h, self.state.data_ptr(), C.byref(h["dev_io"]), C.byref(io),
                                                    int(chunks or self.default_chunks()), self._stream()), "cat_env_step_host")
        else:
            h["actions_dev"].copy_(host_actions, non_blocking=True)
            self.step(h["actions_dev"])
            h["blob"].copy_(self._out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h

    def synchronize(self) -> None:
        """Wait for the launches enqueued so far (needed before reading ``pinned_outputs`` tensors on the host)."""
        torch.cuda.current_stream(self.device).synchronize()

    def observe(self) -> None:
        _lib.check(self.L.cat_env_observe(self._h, self.state.data_ptr(), C.byref(self._io), self._stream()),
                   "cat_env_observe")

    # ------------------------------------------------------------------ state access (parity tests, checkpoints)
    def _view_tensors(self) -> Dict[str, torch.Tensor]:
        N, A, P, dev, K = self.n_worlds, self.A, self.P, self.device, CAT_WALL_SLOTS
        return dict(
            pos=torch.zeros((N, A, 2), dtype=torch.float32, device=dev),
            vel=torch.zeros((N, A, 2), dtype=torch.float32, device=dev),
            vbias=torch.zeros((N, A, 2), dtype=torch.float32, device=dev),
            tc=torch.zeros((N, A, 2), dtype=torch.float32, device=dev),
            step_count=torch.zeros(N, dtype=torch.int32, device=dev),
            episode=torch.zeros(N, dtype=torch.int32, device=dev),
            wall_hull=torch.full((N, A, K), -1, dtype=torch.int32, device=dev),
            wall_age=torch.full((N, A, K), -1, dtype=torch.int32, device=dev),
            wall_jn=torch.zeros((N, A, K), dtype=torch.float32, device=dev),
            pair_age=torch.full((N, max(P, 1)), -1, dtype=torch.int32, device=dev),
            pair_jn=torch.zeros((N, max(P, 1)), dtype=torch.float32, device=dev),
        )

    def get_state(self) -> Dict[str, torch.Tensor]:
        t = self._view_tensors()
        view = CatStateView(*[t[name].data_ptr() for name, _ in CatStateView._fields_])
        _lib.check(self.L.cat_env_get_state(self._h, self.state.data_ptr(), C.byref(view), self._stream()),
                   "cat_env_get_state")
        return t

    def set_state(self, **fields: torch.Tensor) -> None:
        """Overwrite selected state fields (``pos``, ``vel``, ``vbias``, ``tc``, ``step_count``, ``episode``,
        ``wall_hull``+``wall_age``+``wall_jn``, ``pair_age``+``pair_jn``)."""
        proto = self._view_tensors()
        ptrs, keep = [], []
        for name, _ in CatStateView._fields_:
            if name in fields:
                t = torch.as_tensor(fields[name]).to(device=self.device, dtype=proto[name].dtype).contiguous()
                if t.numel() != proto[name].numel():
                    raise ValueError(f"{name}: expected {tuple(proto[name].shape)}, got {tuple(t.shape)}")
                keep.append(t)
                ptrs.append(t.data_ptr())
            else:
                ptrs.append(None)
        unknown = set(fields) - {n for n, _ in CatStateView._fields_}
        if unknown:
            raise ValueError(f"unknown state fields {sorted(unknown)}")
        view = CatStateView(*ptrs)
        _lib.check(self.L.cat_env_set_state(self._h, self.state.data_ptr(), C.byref(view), self._stream()),
                   "cat_env_set_state")
        torch.cuda.current_stream(self.device).synchronize()  # `keep` tensors must outlive the copy

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Checkpoint = the packed state buffer (everything else is recomputed by ``observe``)."""
        return {"state": self.state.clone(), "seed": torch.tensor(self.params["seed"])}

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]) -> None:
        self.state.copy_(sd["state"].to(self.device))
        self.set_seed(int(sd["seed"]))
