"""MAPPO GAE + rollout-buffer advantage normalisation on the GPU (north-star item 4).

Restates what skrl's ``MAPPO._update`` does per agent (``compute_gae`` — reverse scan over the
rollout, ``returns = advantages + values``, then ``(adv - mean) / (std + 1e-8)`` with torch's
unbiased std; call site ``/root/reference/src/utils/agent_learning_utils.py:198-199``,
hyper-parameters ``/root/reference/src/configs/mappo_config.py:5-50``) as two
memory-bound CUDA kernels behind ``cat_gae`` / ``cat_adv_normalize``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


_STATS = {}


def _stats_buffer(device) -> torch.Tensor:
    """The per-(device, stream) statistics workspace of cat_gae (include/cat_b200.h: zeroed once by its owner)."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    buf = _STATS.get(key)
    if buf is None:
        buf = _STATS[key] = torch.zeros(6, dtype=torch.float64, device=device)
    return buf


def compute_gae(rewards: torch.Tensor, dones: torch.Tensor, values: torch.Tensor, last_values: torch.Tensor,
                discount_factor: float = 0.99, lambda_coefficient: float = 0.95, normalize: bool = True,
                group: Optional["torch.distributed.ProcessGroup"] = None,
                distributed: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``rewards``/``values`` float32 and ``dones`` bool/uint8 of shape (T, ...); ``last_values`` (...).

    Returns ``(returns, advantages)`` shaped like ``rewards``.  With ``distributed=True`` the
    normalisation statistics are all-reduced (3 doubles over NCCL) so every rank normalises with the
    global mean / std — the only collective the environment-side path ever issues.
    """
    if rewards.device.type != "cuda":
        raise _lib.CatError("compute_gae needs CUDA tensors; there is no CPU fallback")
    L = _lib.load()
    T = rewards.shape[0]
    M = rewards[0].numel()
    r = rewards.reshape(T, M).contiguous().float()
    v = values.reshape(T, M).contiguous().float()
    d = dones.reshape(T, M).contiguous()
    d = d.view(torch.uint8) if d.dtype == torch.bool else d.to(torch.uint8)
    lv = last_values.reshape(M).contiguous().float()
    ret = torch.empty_like(r)
    adv = torch.empty_like(r)
    stats = _stats_buffer(r.device)        # CAT_GAE_STATS_DOUBLES doubles, zeroed once; the kernel re-zeroes its part
    stream = torch.cuda.current_stream(r.device).cuda_stream
    _lib.check(L.cat_gae(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ret.data_ptr(), adv.data_ptr(),
                         stats.data_ptr(), T, M, float(discount_factor), float(lambda_coefficient), stream), "cat_gae")
    if normalize:
        count = T * M
        sums = stats
        if distributed:
            import torch.distributed as dist
            calls = stats.view(torch.int64)[4:5]                       # the latest call's sums are in slot (calls & 1)
            red = torch.zeros(6, dtype=torch.float64, device=r.device)
            red[:2] = stats[:4].view(2, 2).index_select(0, calls & 1)[0]
            red[2] = float(count)
            dist.all_reduce(red, group=group)
            count = int(round(float(red[2].item())))
            red[2] = 0.0                                               # back to cat_gae's layout: sums in slot 0, calls = 0
            sums = red
        _lib.check(L.cat_adv_normalize(adv.data_ptr(), adv.numel(), sums.data_ptr(), count, stream),
                   "cat_adv_normalize")
    return ret.reshape(rewards.shape), adv.reshape(rewards.shape)
