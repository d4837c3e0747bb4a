"""Batch sharding across the GPUs of one box and the single collective of the sampling path.

Mirrors the reference's data-parallel sampling driver (tools/utils_uvit.py:258-281): every rank integrates its
own slice of the batch with no communication inside the ODE loop, then ONE all-gather concatenates the results
in rank order (``accelerator.gather`` semantics, tools/utils_uvit.py:277).  Differences, by design:
  * the gather moves final *latents* ([B/N,4,32,32] fp32, 16 KiB/image), not decoded images;
  * noise is drawn from one global seed and sliced, so per-sample outputs are identical for any world size
    (the reference seeds per rank, dissect_lfm.py:44).
One process per GPU (torchrun / accelerate); backend "nccl" on GPUs (NVLink 5 / NVSwitch), "gloo" in CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def amortize(n_samples: int, batch_size: int) -> List[int]:
    """tools/utils_uvit.py:258-261."""
    k, r = divmod(n_samples, batch_size)
    return k * [batch_size] if r == 0 else k * [batch_size] + [r]


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous partition of ``n`` samples: the first ``n % world_size`` ranks get one extra."""
    q, r = divmod(n, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def global_noise(n: int, shape=(4, 32, 32), seed: int = 1230) -> torch.Tensor:
    """The whole batch's initial noise from ONE seed (CPU generator: identical on every rank and world size)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((n,) + tuple(shape), generator=g, dtype=torch.float32)


def shard(t: Optional[torch.Tensor], rank: Optional[int] = None, world_size: Optional[int] = None):
    if t is None:
        return None
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def gather_latents(local: torch.Tensor, n_total: Optional[int] = None) -> torch.Tensor:
    """Single all-gather of the final latents, rank-order concatenation, truncated to ``n_total``.

    Ragged shards (n_total % world != 0) are padded to the largest shard so one fixed-size collective suffices.
    """
    rank, w = world()
    if w == 1:
        return local if n_total is None else local[:n_total]
    n_total = n_total if n_total is not None else local.shape[0] * w
    sizes = [shard_bounds(n_total, r, w) for r in range(w)]
    mx = max(hi - lo for lo, hi in sizes)
    buf = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf = torch.cat([local, pad], dim=0)
    out = torch.empty((w * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    parts = [out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def sample_sharded(sample_fn, z_global: torch.Tensor, cond: Optional[torch.Tensor] = None, device=None):
    """Shard ``z_global`` (and ``cond``) over ranks, run ``sample_fn(z_local, cond_local)``, all-gather latents."""
    n = z_global.shape[0]
    z = shard(z_global)
    c = shard(cond)
    if device is not None:
        z = z.to(device)
        c = None if c is None else c.to(device)
    out = sample_fn(z, c)
    return gather_latents(out, n)
