"""CPU oracle for the latent -> image decoder (TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and bench's CPU
leg, never by the product path).

Restates ``FrozenAutoencoderKL.decode`` (libs/autoencoder.py:446-450) = 1/scale_factor, ``post_quant_conv`` (1x1),
``Decoder.forward`` (libs/autoencoder.py:376-409) with ``ResnetBlock`` (:114-134, temb is None), ``AttnBlock``
(:171-195), ``Upsample`` (:46-50, nearest x2 + 3x3 conv), GroupNorm(32, eps=1e-6) + swish (:26-32), on a flat
``state_dict`` in the reference's key layout (``decoder.*``, ``post_quant_conv.*``).  Pinned against the unmodified
reference module by tests/golden/vae_*.npz (tests/golden/make_golden_vae.py)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor

DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)      # libs/autoencoder.py:464-475
SCALE_FACTOR = 0.18215


def _gn(x: Tensor, sd, p: str) -> Tensor:
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def _conv(x: Tensor, sd, p: str, pad: int) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(x: Tensor, sd, p: str) -> Tensor:
    h = _conv(_swish(_gn(x, sd, p + ".norm1")), sd, p + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, p + ".norm2")), sd, p + ".conv2", 1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".nin_shortcut", 0)
    return x + h


def attn_block(x: Tensor, sd, p: str) -> Tensor:
    h = _gn(x, sd, p + ".norm")
    q, k, v = (_conv(h, sd, f"{p}.{n}", 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)          # b, hw, c
    k = k.reshape(b, c, hh * ww)                           # b, c, hw
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, p + ".proj_out", 0)


def decoder_forward(sd: Dict[str, Tensor], z: Tensor, prefix: str = "decoder") -> Tensor:
    n_res = len(DDCONFIG["ch_mult"])
    h = _conv(z, sd, f"{prefix}.conv_in", 1)
    h = resnet_block(h, sd, f"{prefix}.mid.block_1")
    h = attn_block(h, sd, f"{prefix}.mid.attn_1")
    h = resnet_block(h, sd, f"{prefix}.mid.block_2")
    for lvl in reversed(range(n_res)):
        for blk in range(DDCONFIG["num_res_blocks"] + 1):
            h = resnet_block(h, sd, f"{prefix}.up.{lvl}.block.{blk}")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"{prefix}.up.{lvl}.upsample.conv", 1)
    h = _swish(_gn(h, sd, f"{prefix}.norm_out"))
    return _conv(h, sd, f"{prefix}.conv_out", 1)


def decode(sd: Dict[str, Tensor], z: Tensor, scale_factor: float = SCALE_FACTOR) -> Tensor:
    """FrozenAutoencoderKL.decode: latents [B, 4, S, S] -> images [B, 3, 8S, 8S]."""
    z = (1.0 / scale_factor) * z
    z = _conv(z, sd, "post_quant_conv", 0)
    return decoder_forward(sd, z)


def flops_per_image(S: int = 32) -> float:
    """Multiply-add = 2 FLOPs over every convolution / attention product of one decode at latent side S."""
    ch, mult = DDCONFIG["ch"], DDCONFIG["ch_mult"]
    f = 0.0
    res = S
    c = ch * mult[-1]
    f += 2.0 * res * res * 4 * 4 + 2.0 * res * res * 9 * 4 * c          # post_quant_conv, conv_in
    f += 2 * (2 * 2.0 * res * res * 9 * c * c)                           # mid blocks
    f += 4 * 2.0 * res * res * c * c + 2 * 2.0 * (res * res) ** 2 * c    # attention: q, k, v, proj + two products
    cin = c
    for lvl in reversed(range(len(mult))):
        cout = ch * mult[lvl]
        for blk in range(DDCONFIG["num_res_blocks"] + 1):
            f += 2.0 * res * res * 9 * cin * cout + 2.0 * res * res * 9 * cout * cout
            if cin != cout:
                f += 2.0 * res * res * cin * cout
            cin = cout
        if lvl != 0:
            res *= 2
            f += 2.0 * res * res * 9 * cin * cin
    f += 2.0 * res * res * 9 * cin * 3
    return f


def encoder_forward(sd: Dict[str, Tensor], x: Tensor, prefix: str = "encoder") -> Tensor:
    """Encoder.forward (libs/autoencoder.py:275-300); Downsample = zero pad (0, 1, 0, 1) + 3x3 stride-2 conv (:65-69)."""
    n_res = len(DDCONFIG["ch_mult"])
    h = _conv(x, sd, f"{prefix}.conv_in", 1)
    for lvl in range(n_res):
        for blk in range(DDCONFIG["num_res_blocks"]):
            h = resnet_block(h, sd, f"{prefix}.down.{lvl}.block.{blk}")
        if lvl != n_res - 1:
            p = f"{prefix}.down.{lvl}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = resnet_block(h, sd, f"{prefix}.mid.block_1")
    h = attn_block(h, sd, f"{prefix}.mid.attn_1")
    h = resnet_block(h, sd, f"{prefix}.mid.block_2")
    h = _swish(_gn(h, sd, f"{prefix}.norm_out"))
    return _conv(h, sd, f"{prefix}.conv_out", 1)


def encode_moments(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """FrozenAutoencoderKL.encode_moments (libs/autoencoder.py:426-429): images [B, 3, R, R] -> [B, 8, R/8, R/8]."""
    return _conv(encoder_forward(sd, x), sd, "quant_conv", 0)
