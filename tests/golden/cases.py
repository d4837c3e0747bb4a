"""Golden-vector case table shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import torch

_L = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=20, num_heads=16, mlp_ratio=4,
          qkv_bias=False, mlp_time_embed=False, use_checkpoint=False)
_S16 = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=512, depth=16, num_heads=8, mlp_ratio=4,
            qkv_bias=False, mlp_time_embed=False, use_checkpoint=False)
_TINY = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=256, depth=4, num_heads=4, mlp_ratio=4,
             qkv_bias=False, mlp_time_embed=False, use_checkpoint=False)

CASES = {
    # BASELINE.json configs[0]: lfm_cm256_uvit_small_deep16 (configs/lfm_cm256_uvit_small_deep16_scratch.py:40-53)
    "small16_uncond": dict(cfg=dict(_S16, num_classes=-1), t2i=False, seed=0, B=2, in_seed=1230, euler_steps=None,
                           edit=dict(seed=1232, t=0.2, t_edit=0.4, write_scale=1.5, ith_attr=1)),
    "tiny_uncond": dict(cfg=dict(_TINY, num_classes=-1), t2i=False, seed=0, B=3, in_seed=7, euler_steps=5,
                        edit=dict(seed=5, t=0.4, t_edit=0.4, write_scale=-2.1, ith_attr="0_2")),
    "tiny_class": dict(cfg=dict(_TINY, num_classes=10), t2i=False, seed=1, B=2, in_seed=8, euler_steps=None),
    "tiny_qkvbias_noconv": dict(cfg=dict(_TINY, num_classes=-1, qkv_bias=True, conv=False), t2i=False, seed=2, B=2,
                                in_seed=9, euler_steps=None),
    "tiny_t2i": dict(cfg=dict(_TINY, clip_dim=768, num_clip_token=77), t2i=True, seed=3, B=2, in_seed=10,
                     euler_steps=4,
                     # dissect_lfm_t2i.py "p2p" mode: post-softmax re-weighting of context-token columns
                     p2p=dict(t=0.3, t_edit=0.4, block_id=[1, 3], multiplier=[2.0, -1.5], ids=[[3, 4], [10]])),
    # mlp_time_embed=True (libs/uvit.py:215-223; no shipped config sets it): class-conditional, forward + Euler loop
    "tiny_time_mlp": dict(cfg=dict(_TINY, num_classes=10, mlp_time_embed=True), t2i=False, seed=4, B=2, in_seed=11,
                          euler_steps=3),
    "tiny_t2i_time_mlp": dict(cfg=dict(_TINY, clip_dim=768, num_clip_token=77, mlp_time_embed=True), t2i=True, seed=5,
                              B=2, in_seed=12, euler_steps=None),
    # 64x64 latents (a 512^2 image, datasets.py:244-245): 1024 patches + the time token = 1025 tokens, beyond one
    # 384-key attention pass
    "tiny_long": dict(cfg=dict(_TINY, img_size=64, depth=2, num_classes=-1), t2i=False, seed=6, B=2, in_seed=13,
                      euler_steps=2),
    # configs/lfm_cm256_uvit_large.py:42-56 and configs/lfm_mmcelebahq256_uvit_large.py:43-58, one image
    "large_uncond": dict(cfg=dict(_L, num_classes=-1), t2i=False, seed=0, B=1, in_seed=1230, euler_steps=None),
    "large_t2i": dict(cfg=dict(_L, clip_dim=768, num_clip_token=77), t2i=True, seed=0, B=1, in_seed=1231,
                      euler_steps=None),
}


def build_inputs(case):
    """Deterministic inputs for a case: x ~ N(0,1), t ~ U(0,1) per sample, labels / context when conditional."""
    g = torch.Generator(device="cpu").manual_seed(case["in_seed"])
    B, cfg = case["B"], case["cfg"]
    x = torch.randn(B, cfg["in_chans"], cfg["img_size"], cfg["img_size"], generator=g)
    t = torch.rand(B, generator=g)
    y = None
    if cfg.get("num_classes", -1) > 0:
        y = torch.randint(0, cfg["num_classes"], (B,), generator=g)
    ctx = torch.randn(B, cfg["num_clip_token"], cfg["clip_dim"], generator=g) if case["t2i"] else None
    return x, t, y, ctx


def build_model(case, cls_uncond, cls_t2i):
    torch.manual_seed(case["seed"])
    return (cls_t2i if case["t2i"] else cls_uncond)(**case["cfg"]).eval()


def attr_delta_inputs():
    """Synthetic "read"-mode dump for the attribute-direction golden: features [B, T, C, W, H], encoded latents
    [B, C, W, H], FFHQ-style binary attributes [B, 11], the evaluation-time strings and the number of batches."""
    import numpy as np
    rng = np.random.default_rng(2024)
    B, times, batch_num = 12, ["0.25", "0.50", "1.00"], 3
    feats = rng.standard_normal((B, len(times), 4, 8, 8)).astype(np.float32)
    latent = rng.standard_normal((B, 4, 8, 8)).astype(np.float32)
    attr = (rng.random((B, 11)) < 0.5).astype(np.int64)
    attr[0], attr[1] = 1, 0          # every attribute has at least one positive and one negative sample
    return feats, latent, attr, times, batch_num


# ---- latent -> image decoder (libs/autoencoder.py) ----------------------------------------------------------
VAE_CASES = dict(seed=11, vae_small=dict(B=2, S=16, in_seed=21), vae_full=dict(B=1, S=32, in_seed=22))


def vae_model():
    """The mirror modules under the golden seed (== the reference's Decoder + post_quant_conv under that seed)."""
    from uspace_b200.autoencoder import DDCONFIG, Decoder
    torch.manual_seed(VAE_CASES["seed"])
    dec = Decoder(**DDCONFIG)
    pq = torch.nn.Conv2d(4, 4, 1)
    return dec, pq


def vae_state_dict():
    dec, pq = vae_model()
    sd = {f"decoder.{k}": v for k, v in dec.state_dict().items()}
    sd.update({f"post_quant_conv.{k}": v for k, v in pq.state_dict().items()})
    return sd


def vae_latents(name):
    c = VAE_CASES[name]
    g = torch.Generator().manual_seed(c["in_seed"])
    return 0.18215 * 4.0 * torch.randn(c["B"], 4, c["S"], c["S"], generator=g)    # scaled like encoded latents


VAE_ENC_CASES = dict(seed=12, vae_enc_small=dict(B=2, R=128, in_seed=31), vae_enc_full=dict(B=1, R=256, in_seed=32))


def vae_enc_state_dict():
    """Mirror Encoder + quant_conv under the golden seed (== the reference's modules under that seed)."""
    from uspace_b200.autoencoder import DDCONFIG, Encoder
    torch.manual_seed(VAE_ENC_CASES["seed"])
    enc = Encoder(**DDCONFIG)
    qc = torch.nn.Conv2d(8, 8, 1)
    sd = {f"encoder.{k}": v for k, v in enc.state_dict().items()}
    sd.update({f"quant_conv.{k}": v for k, v in qc.state_dict().items()})
    return sd


def vae_images(name):
    c = VAE_ENC_CASES[name]
    g = torch.Generator().manual_seed(c["in_seed"])
    return torch.rand(c["B"], 3, c["R"], c["R"], generator=g) * 2.0 - 1.0       # images in [-1, 1]
