"""uspace_b200 — Blackwell (sm_100a) flow-matching sampler for the U-ViT velocity field of dongzhuoyao/uspace.

Public surface (mirrors the reference's own names for this path):
  get_nnet("uvit" | "uvit_t2i", **cfg)      tools/utils_uvit.py:27-41
  UViT, UViTT2I                              libs/uvit.py, libs/uvit_t2i.py
  CNF(net).decode / .encode / .forward       flow_matching.py, flow_matching_t2i.py
  amortize, shard_batch, gather_latents      tools/utils_uvit.py:258-281 (batch sharding + the one all-gather)
"""
from .uvit import UViT, UViTT2I, get_nnet  # noqa: F401

__all__ = ["UViT", "UViTT2I", "get_nnet"]
