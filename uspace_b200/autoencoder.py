"""Mirror of the reference's KL autoencoder (libs/autoencoder.py): same constructor keywords, same ``state_dict``
keys (``encoder.*``, ``decoder.*``, ``quant_conv.*``, ``post_quant_conv.*``), same ``decode(z)`` / ``encode(x)`` /
``encode_moments(x)`` calls - the steps on either side of the sampling path (dissect_lfm.py:86-98 decodes in chunks
of 50; real-image editing encodes first, dissect_lfm.py:150-160).

The modules below only HOLD parameters in the reference's layout (so reference checkpoints load unchanged and the
seeded constructor reproduces the reference initialisation); the convolutions run on the CUDA library (csrc/vae.cu:
implicit-GEMM convolutions on the tcgen05 kernels of the U-ViT path, GroupNorm/swish, 1024-token attention).  There is
no PyTorch fallback."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:   # asymmetric (0, 1, 0, 1) zero padding is applied by the caller (libs/autoencoder.py:65-69)
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        if temb_channels > 0 or conv_shortcut:
            raise NotImplementedError("temb / conv_shortcut are not used by the decoder and are not built")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)


class Encoder(nn.Module):
    """libs/autoencoder.py:209-272 (parameter creation order kept)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if use_linear_attn or attn_type != "vanilla" or not resamp_with_conv or not double_z:
            raise NotImplementedError("only the configuration of libs/autoencoder.py::get_model is built")
        self.num_resolutions = len(ch_mult)
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    raise NotImplementedError("attention inside the down path is not used by get_model and not built")
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels, kernel_size=3, stride=1, padding=1)


class Decoder(nn.Module):
    """libs/autoencoder.py:303-373 (parameter creation order kept, so a seeded constructor matches the reference)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or attn_type != "vanilla" or give_pre_end or tanh_out or not resamp_with_conv:
            raise NotImplementedError("only the configuration of libs/autoencoder.py::get_model is built")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.ch_mult = tuple(ch_mult)
        block_in = ch * ch_mult[-1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    raise NotImplementedError("attention inside the up path is not used by get_model and not built")
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)


DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)


# GEMM operand precision (include/uspace_b200.h USP_VAE_PRECISION_*): "fp16x3" splits every operand into hi + lo fp16
# parts (three products in one GEMM, ~1e-5 against fp64); "fp16" is one product: 3x less tensor work at ~2e-3, the
# precision of the TF32 convolutions the reference's torch modules run by default on a GPU.
PRECISIONS = {"fp16": 0, "fp16x3": 1}


class FrozenAutoencoderKL(nn.Module):
    """libs/autoencoder.py:412-460.  ``pretrained_path=None`` keeps the random initialisation."""

    def __init__(self, ddconfig=None, embed_dim=4, pretrained_path=None, scale_factor=0.18215, precision="fp16x3"):
        super().__init__()
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        ddconfig = dict(DDCONFIG if ddconfig is None else ddconfig)
        if (ddconfig["ch"], list(ddconfig["ch_mult"]), ddconfig["num_res_blocks"], ddconfig["z_channels"],
                ddconfig["out_ch"]) != (128, [1, 2, 4, 4], 2, 4, 3) or embed_dim != 4:
            raise NotImplementedError("only the Stable-Diffusion KL-f8 decoder of get_model() is built")
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim, self.scale_factor = embed_dim, scale_factor
        self._engine = None
        if pretrained_path is not None:
            m, u = self.load_state_dict(torch.load(pretrained_path, map_location="cpu"))
            assert len(m) == 0 and len(u) == 0
        self.eval().requires_grad_(False)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self):
        dev = self.post_quant_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("uspace_b200: the decoder runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        if self._engine is None:
            self._engine = VaeEngine(dev, self.state_dict(), self.scale_factor, self.precision)
        return self._engine

    @torch.no_grad()
    def decode(self, z):
        return self.engine().decode(z)

    @torch.no_grad()
    def encode_moments(self, x):
        """libs/autoencoder.py:426-429: images [B, 3, R, R] -> (mean, logvar) moments [B, 8, R/8, R/8]."""
        return self.engine().encode_moments(x)

    def sample(self, moments):
        """libs/autoencoder.py:431-437 (elementwise glue; the noise comes from torch's generator like the reference's)."""
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        std = torch.exp(0.5 * logvar)
        return self.scale_factor * (mean + std * torch.randn_like(mean))

    def encode(self, x):
        return self.sample(self.encode_moments(x))

    def forward(self, inputs, fn):
        if fn == "decode":
            return self.decode(inputs)
        if fn == "encode_moments":
            return self.encode_moments(inputs)
        if fn == "encode":
            return self.encode(inputs)
        raise NotImplementedError


def flops_per_image(S: int = 32) -> float:
    """Multiply-add = 2 FLOPs over every convolution and attention product of one decode at latent side S
    (0.62 TFLOP for the 256^2 models)."""
    ch, mult, n_blk = DDCONFIG["ch"], DDCONFIG["ch_mult"], DDCONFIG["num_res_blocks"] + 1
    res, c = S, DDCONFIG["ch"] * DDCONFIG["ch_mult"][-1]
    f = 2.0 * res * res * 4 * 4 + 2.0 * res * res * 9 * 4 * c + 4 * 2.0 * res * res * 9 * c * c
    f += 4 * 2.0 * res * res * c * c + 2 * 2.0 * (res * res) ** 2 * c
    cin = c
    for lvl in reversed(range(len(mult))):
        cout = ch * mult[lvl]
        for _ in range(n_blk):
            f += 2.0 * res * res * 9 * cin * cout + 2.0 * res * res * 9 * cout * cout
            if cin != cout:
                f += 2.0 * res * res * cin * cout
            cin = cout
        if lvl != 0:
            res *= 2
            f += 2.0 * res * res * 9 * cin * cin
    return f + 2.0 * res * res * 9 * cin * 3


def decode_large_batch(autoencoder, batch, chunk: int = 50):
    """dissect_lfm.py:86-98: decode in chunks of ``chunk`` latents and concatenate."""
    return torch.cat([autoencoder.decode(batch[i:i + chunk]) for i in range(0, batch.shape[0], chunk)], dim=0)


def get_model(pretrained_path=None, scale_factor=0.18215, precision="fp16x3"):
    return FrozenAutoencoderKL(DDCONFIG, 4, pretrained_path, scale_factor, precision)


class VaeEngine:
    """Owner of a ``usp_vae`` handle (C ABI): weights in the reference's state_dict layout, decode on the current stream."""

    def __init__(self, device, state_dict, scale_factor, precision="fp16x3"):
        self.lib = _lib.load()
        self.device = device
        self.handle = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.usp_vae_create(idx, C.c_float(scale_factor), C.byref(self.handle)), None, "usp_vae_create")
        n = self.lib.usp_vae_num_weights(self.handle)
        names = [self.lib.usp_vae_weight_name(self.handle, i).decode() for i in range(n)]
        missing = [k for k in names if k not in state_dict]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        for k in names:
            t = state_dict[k].detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.usp_vae_set_weight(self.handle, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                       self.handle, f"usp_vae_set_weight({k})", vae=True)
        _lib.check(self.lib.usp_vae_set_precision(self.handle, PRECISIONS[precision]), self.handle, "usp_vae_set_precision",
                   vae=True)
        _lib.check(self.lib.usp_vae_finalize(self.handle, self._stream()), self.handle, "usp_vae_finalize", vae=True)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode_moments(self, x):
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3] or x.shape[2] % 128 != 0:
            raise ValueError(f"images must be [B, 3, R, R] with R % 128 == 0, got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        B, R = x.shape[0], x.shape[2]
        out = torch.empty((B, 8, R // 8, R // 8), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.usp_vae_encode_moments(self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, R,
                                                   self._stream()), self.handle, "usp_vae_encode_moments", vae=True)
        return out

    def decode(self, z):
        if z.dim() != 4 or z.shape[1] != 4 or z.shape[2] != z.shape[3] or z.shape[2] % 4 != 0:
            raise ValueError(f"latents must be [B, 4, S, S] with S % 4 == 0, got {tuple(z.shape)}")
        z = z.to(self.device, torch.float32).contiguous()
        B, S = z.shape[0], z.shape[2]
        out = torch.empty((B, 3, 8 * S, 8 * S), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.usp_vae_decode(self.handle, C.c_void_p(z.data_ptr()), C.c_void_p(out.data_ptr()), B, S,
                                           self._stream()), self.handle, "usp_vae_decode", vae=True)
        return out

    def workspace_bytes(self):
        return int(self.lib.usp_vae_workspace_bytes(self.handle))

    def release_workspace(self):
        """Hand the activation slab back to the driver (synchronises); the next decode / encode allocates it again."""
        _lib.check(self.lib.usp_vae_release_workspace(self.handle), self.handle, "usp_vae_release_workspace", vae=True)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.usp_vae_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
