"""Drop-in mirrors of the reference velocity-field modules, backed by the sm_100a C-ABI library.

``UViT`` mirrors ``libs/uvit.py:182-351`` and ``UViTT2I`` mirrors ``libs/uvit_t2i.py:192-342``: same constructor
keywords, same parameter names / shapes (``state_dict`` round-trips with reference checkpoints), same
``forward(x, timesteps, y=None | context, **kwargs) -> (pred, None)`` contract, arbitrary extra kwargs
tolerated (the reference splats ``config.dissection`` into every call).

Inference (``torch.no_grad`` / ``inference_mode``) runs ONLY through ``libuspace_b200.so``; there is no
CPU or eager fallback for it — a missing library or a non-CUDA tensor raises.  When autograd is enabled
(``train_lfm*.py``: training_losses -> backward) the same parameters are run through a plain
differentiable PyTorch graph, because the hand-written kernels are forward-only; that training path is
outside the accelerated scope.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Engine


def _trunc_normal_(tensor: torch.Tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    # same sampling recipe as timm 0.3.2 (libs/timm.py:11-62), so identical seeds give identical weights
    def norm_cdf(x):
        return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0

    with torch.no_grad():
        lo = norm_cdf((a - mean) / std)
        hi = norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.0))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, skip):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.skip_linear = nn.Linear(2 * dim, dim) if skip else None


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class _UViTBase(nn.Module):
    """Parameter container + engine lifecycle shared by the uncond/class and t2i models."""

    operand_dtype = "fp16"  # tensor-core operand type ("fp16" or "bf16"); fp32 accumulate either way
    # fold norm1 / norm2 into the qkv / fc1 GEMMs (exact algebra: gamma into the weights, their row means removed so that
    # the epilogue only scales by 1 / std and adds the folded bias; no LayerNorm kernels between the GEMMs).  +1.2-1.8 %
    # on B200; False keeps the stand-alone LayerNorm launches.
    fuse_layernorm = True

    def _build(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
               mlp_time_embed, conv, skip, extras):
        self.embed_dim = self.num_features = embed_dim
        self.in_chans = in_chans
        self.extras = extras
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        num_patches = (img_size // patch_size) ** 2
        # libs/uvit.py:215-223 (same construction order as the reference: the seeded initialisation must match)
        self.time_embed = (nn.Sequential(nn.Linear(embed_dim, 4 * embed_dim), nn.SiLU(), nn.Linear(4 * embed_dim, embed_dim))
                           if mlp_time_embed else nn.Identity())
        self._ctor_extra()
        self.pos_embed = nn.Parameter(torch.zeros(1, self.extras + num_patches, embed_dim))
        mk = lambda s: _Block(embed_dim, num_heads, mlp_ratio, qkv_bias, s)
        self.in_blocks = nn.ModuleList([mk(False) for _ in range(depth // 2)])
        self.mid_block = mk(False)
        self.out_blocks = nn.ModuleList([mk(skip) for _ in range(depth // 2)])
        self.norm = nn.LayerNorm(embed_dim)
        self.patch_dim = patch_size ** 2 * in_chans
        self.decoder_pred = nn.Linear(embed_dim, self.patch_dim, bias=True)
        self.final_layer = nn.Conv2d(in_chans, in_chans, 3, padding=1) if conv else nn.Identity()
        _trunc_normal_(self.pos_embed, std=0.02)
        self.apply(self._init_weights)
        self._engine: Optional[Engine] = None
        self._engine_versions = None

    def _ctor_extra(self):
        pass

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed"}

    # ---- engine lifecycle: rebuilt when the device changes, re-packed when any parameter changes ----
    def _versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self) -> Engine:
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError(
                "uspace_b200: the inference path needs the model on a CUDA (sm_100a) device; there is no CPU "
                "fallback — move the module with .to('cuda') / accelerator.prepare first")
        if (self._engine is None or self._engine.device != dev or self._engine.operand_dtype != self.operand_dtype
                or self._engine.fuse_layernorm != bool(self.fuse_layernorm)):
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self._ctor_kwargs, dev, self.operand_dtype, self.fuse_layernorm)
            self._engine_versions = None
        ver = self._versions()
        if ver != self._engine_versions:  # load_state_dict / optimizer step / .to() happened
            self._engine.load_state_dict(self.state_dict())
            self._engine_versions = ver
        return self._engine

    # ---- differentiable PyTorch graph for training only ------------------------------------------------
    def _block_autograd(self, blk: _Block, x, skip=None):
        if blk.skip_linear is not None:
            x = blk.skip_linear(torch.cat([x, skip], dim=-1))
        h = blk.norm1(x)
        B, L, D = h.shape
        H = blk.attn.num_heads
        qkv = blk.attn.qkv(h).reshape(B, L, 3, H, D // H).permute(2, 0, 3, 1, 4).float()
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, L, D)
        x = x + blk.attn.proj(a)
        x = x + blk.mlp.fc2(F.gelu(blk.mlp.fc1(blk.norm2(x))))
        return x

    def _trunk_autograd(self, x):
        x = x + self.pos_embed
        skips = []
        for blk in self.in_blocks:
            x = self._block_autograd(blk, x)
            skips.append(x)
        x = self._block_autograd(self.mid_block, x)
        for blk in self.out_blocks:
            x = self._block_autograd(blk, x, skips.pop())
        x = self.decoder_pred(self.norm(x))[:, self.extras:, :]
        B, n, _ = x.shape
        g = int(n ** 0.5)
        p = self.patch_embed.patch_size
        x = x.reshape(B, g, g, p, p, self.in_chans).permute(0, 5, 1, 3, 2, 4).reshape(B, self.in_chans, g * p, g * p)
        return self.final_layer(x)

    def _time_token(self, timesteps):
        half = self.embed_dim // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
        args = timesteps[:, None].float() * freqs[None]
        return self.time_embed(torch.cat([torch.cos(args), torch.sin(args)], dim=-1)).unsqueeze(1)

    @staticmethod
    def _wants_autograd(x):
        return torch.is_grad_enabled()


class UViT(_UViTBase):
    """Mirror of libs/uvit.py::UViT (unconditional / class-conditional)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, qk_scale=None, norm_layer=nn.LayerNorm, mlp_time_embed=False,
                 num_classes=-1, use_checkpoint=False, conv=True, skip=True, use_latent1d=0,
                 latent_1d_pooling=False):
        super().__init__()
        # qk_scale: accepted and WITHOUT EFFECT, exactly like the reference as it runs - libs/uvit.py:79 stores it in
        # Attention.scale, but with torch >= 2.0 ATTENTION_MODE is "flash" and F.scaled_dot_product_attention
        # (libs/uvit.py:95) always uses head_dim ** -0.5; only the dead "math" branch (:109) reads self.scale.
        self.qk_scale = qk_scale
        self.num_classes = num_classes
        self._ctor_kwargs = dict(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                 depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                 num_classes=num_classes, conv=conv, skip=skip, mlp_time_embed=mlp_time_embed)
        self._build(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                    mlp_time_embed, conv, skip, extras=2 if num_classes > 0 else 1)

    def _ctor_extra(self):
        if self.num_classes > 0:
            self.label_emb = nn.Embedding(self.num_classes, self.embed_dim)

    def _hook(self, timesteps, kwargs):
        """What dissect_helper_uvit (libs/dissection.py:115-186) would do inside this call (libs/uvit.py:313-314 head,
        :349-350 tail): -> (edit_loc, delta [C,S,S] | None, write_scale, read_path | None).  The hook keys on
        f"{timesteps[0].item():.2f}" (one host sync per call, like the reference); CNF.decode / encode of this package
        pre-gather the whole table instead and never come through here."""
        loc = kwargs.get("edit_loc")
        if loc == "mid":
            raise NotImplementedError("edit_loc='mid' is broken in the reference for U-ViT and is not built")
        if loc not in ("head", "tail") or kwargs.get("dissect_task") != "uspace_uvit":
            return None, None, 0.0, None
        from .flow_matching import _read_delta, should_edit
        import os
        name = kwargs.get("dissect_name")
        digit = f"{timesteps.reshape(-1)[0].item():.2f}"
        if name == "read":
            root = kwargs.get("read_path_root")
            os.makedirs(root, exist_ok=True)
            return loc, None, 0.0, os.path.join(root, f"{kwargs['batch_id']}_{digit}")
        if name not in ("write_attr", "write_pca"):
            raise ValueError(f"dissect_name should be read or write, here is {name}")
        if not should_edit(digit, kwargs.get("t_edit")):
            return None, None, 0.0, None
        root = kwargs.get("write_path_root")
        if name == "write_attr":
            delta = _read_delta(os.path.join(root, f"delta_{digit}.npy"), kwargs.get("ith_attr"))
        else:
            delta = _read_delta(os.path.join(root, f"pca{kwargs.get('pca_n')}_{digit}.npy"), kwargs.get("ith_component"))
        return loc, torch.from_numpy(delta.astype("float32")), float(kwargs.get("write_scale")), None

    def forward(self, x, timesteps, y=None, **kwargs):
        # kwargs.get: the reference indexes kwargs["edit_loc"] (libs/uvit.py:313) and raises KeyError without it;
        # accepting its absence is a strict superset.
        if self._wants_autograd(x):
            tok = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
            tok = torch.cat((self._time_token(timesteps), tok), dim=1)
            if y is not None:
                tok = torch.cat((self.label_emb(y).unsqueeze(1), tok), dim=1)
            return self._trunk_autograd(tok), None
        loc, delta, ws, read_path = self._hook(timesteps, kwargs)
        if read_path is not None:
            import numpy as np
            out, act = self.engine().forward(x, timesteps, y=y, edit_loc=loc, read=True)
            np.save(read_path, act.cpu().numpy())
            return out, None
        return self.engine().forward(x, timesteps, y=y, edit_loc=loc, delta=delta, write_scale=ws), None


class UViTT2I(_UViTBase):
    """Mirror of libs/uvit_t2i.py::UViT (77 CLIP context tokens)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, qk_scale=None, norm_layer=nn.LayerNorm, mlp_time_embed=False,
                 use_checkpoint=False, clip_dim=768, num_clip_token=77, conv=True, skip=True, use_latent1d=False):
        super().__init__()
        # qk_scale: without effect on the plain forward (flash path, libs/uvit_t2i.py:118-122, as in libs/uvit.py); the
        # reference reads it only in the attention-editing branch (:96-107), which is rejected below when it is set.
        self.qk_scale = qk_scale
        self._clip = (clip_dim, num_clip_token)
        self._ctor_kwargs = dict(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                 depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                 clip_dim=clip_dim, num_clip_token=num_clip_token, conv=conv, skip=skip,
                                 mlp_time_embed=mlp_time_embed)
        self._build(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, qkv_bias,
                    mlp_time_embed, conv, skip, extras=1 + num_clip_token)

    def _ctor_extra(self):
        self.context_embed = nn.Linear(self._clip[0], self.embed_dim)

    def forward(self, x, timesteps, context, **kwargs):
        if self._wants_autograd(x):
            tok = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
            tok = torch.cat((self._time_token(timesteps), self.context_embed(context.to(x.device)), tok), dim=1)
            return self._trunk_autograd(tok), None
        attn = None
        if kwargs.get("dissect_name") in ("p2p", "local_prompt", "sampled_image_editing"):
            # the reference's attention-editing branch (libs/uvit_t2i.py:91-107): active for t <= t_edit in decode
            from .flow_matching import build_attn_edit
            attn = build_attn_edit(x.shape[0], self.pos_embed.shape[1], **kwargs)
            if attn is not None and self.qk_scale is not None:
                raise NotImplementedError("attention editing with a qk_scale override is not built "
                                          "(the editing branch of libs/uvit_t2i.py:96-107 would use it)")
            if attn is not None and not float(f"{timesteps[0].item():.2f}") <= attn["t_edit"]:
                attn = None
        return self.engine().forward(x, timesteps, context=context, attn_edit=attn), None


def get_nnet(name, **kwargs):
    """Mirror of tools/utils_uvit.py:27-41 for the two U-ViT entries (the UNet backbone is out of scope)."""
    if name == "uvit":
        return UViT(**kwargs)
    if name == "uvit_t2i":
        return UViTT2I(**kwargs)
    raise NotImplementedError(f"{name}: only 'uvit' and 'uvit_t2i' are built (SURVEY.md §2.1 row 14)")
