"""uspace_b200 — Blackwell (sm_100a) flow-matching sampler for the U-ViT velocity field of dongzhuoyao/uspace.

Public surface (mirrors the reference's own names for this path; submodules are imported on demand):
  get_nnet("uvit" | "uvit_t2i", **cfg)                      tools/utils_uvit.py:27-41
  UViT, UViTT2I                                              libs/uvit.py, libs/uvit_t2i.py
  flow_matching.CNF / CNFT2I (.decode / .encode / .forward)  flow_matching.py, flow_matching_t2i.py
       fixed grid (euler, heun, midpoint, rk4), adaptive (dopri5, bosh3, adaptive_heun), fixadp, the write / read hook,
       the p2p attention edit
  sweep.sample_write_scales                                  tools/utils_vis.py:189-201 (all write_scales in one batch)
  attr_delta.extract_deltas_by_attr                          tools/utils_attr.py:160-206 (read dumps -> delta files)
  autoencoder.get_model(path).decode / .encode               libs/autoencoder.py (latents <-> images)
  parallel.amortize / shard / gather_latents / sample_sharded  tools/utils_uvit.py:258-281 (sharding + one all-gather)
"""
from .uvit import UViT, UViTT2I, get_nnet  # noqa: F401

__all__ = ["UViT", "UViTT2I", "get_nnet"]
