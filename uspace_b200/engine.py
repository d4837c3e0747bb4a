"""Thin owner of a ``usp_handle``: loads reference-format weights and enqueues forward / sample on the
caller's CUDA stream.  PyTorch is only used for device memory and streams here."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def config_from_kwargs(kw: dict, operand_dtype: str = "fp16", fuse_layernorm: bool = True) -> _lib.UspConfig:
    """Map the reference ctor kwargs (libs/uvit.py:183-202, libs/uvit_t2i.py:193-211) to usp_config."""
    t2i = "clip_dim" in kw or "num_clip_token" in kw
    c = _lib.UspConfig()
    c.img_size = kw.get("img_size", 224)
    c.patch_size = kw.get("patch_size", 16)
    c.in_chans = kw.get("in_chans", 3)
    c.embed_dim = kw.get("embed_dim", 768)
    c.depth = kw.get("depth", 12)
    c.num_heads = kw.get("num_heads", 12)
    c.mlp_hidden = int(c.embed_dim * kw.get("mlp_ratio", 4.0))
    c.num_classes = 0 if t2i else max(0, kw.get("num_classes", -1))
    c.clip_dim = kw.get("clip_dim", 768) if t2i else 0
    c.num_clip_token = kw.get("num_clip_token", 77) if t2i else 0
    c.qkv_bias = int(bool(kw.get("qkv_bias", False)))
    c.conv = int(bool(kw.get("conv", True)))
    c.skip = int(bool(kw.get("skip", True)))
    c.operand_dtype = _lib.OPERAND[operand_dtype]
    c.fuse_layernorm = int(bool(fuse_layernorm))
    c.mlp_time_embed = int(bool(kw.get("mlp_time_embed", False)))
    return c


class Engine:
    def __init__(self, ctor_kwargs: dict, device: torch.device, operand_dtype: str = "fp16",
                 fuse_layernorm: bool = True):
        if device.type != "cuda":
            raise RuntimeError("uspace_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.lib = _lib.load()
        self.device = device
        self.cfg = config_from_kwargs(ctor_kwargs, operand_dtype, fuse_layernorm)
        self.operand_dtype = operand_dtype
        self.fuse_layernorm = bool(fuse_layernorm)
        self.handle = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.usp_create(C.byref(self.cfg), idx, C.byref(self.handle)), None, "usp_create")
        self.names = [self.lib.usp_weight_name(self.handle, i).decode()
                      for i in range(self.lib.usp_num_weights(self.handle))]
        self.C, self.S = self.cfg.in_chans, self.cfg.img_size

    def close(self):
        if getattr(self, "handle", None):
            self.lib.usp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        missing = [n for n in self.names if n not in sd]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        for n in self.names:
            t = sd[n].detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.usp_set_weight(self.handle, n.encode(), _ptr(t), shape, t.dim()), self.handle,
                       f"usp_set_weight({n})")
        _lib.check(self.lib.usp_finalize_weights(self.handle, self._stream()), self.handle, "usp_finalize_weights")

    # ---- compute -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_latent(self, x: torch.Tensor):
        if x.device != self.device and x.device.type == "cuda" and x.device.index != self.device.index:
            raise RuntimeError(f"latent on {x.device}, engine on {self.device}")
        if x.dim() != 4 or x.shape[1:] != (self.C, self.S, self.S):
            raise ValueError(f"latent must be [B,{self.C},{self.S},{self.S}], got {tuple(x.shape)}")

    def _attn_edit(self, attn_edit, B):
        """attn_edit = dict(colscale=[B, L] tensor, block_mask=int, t_edit=float) -> (ctypes struct, keep-alive)."""
        if attn_edit is None:
            return None, None
        cs = attn_edit["colscale"].to(self.device, torch.float32).contiguous()
        if cs.shape[0] != B:
            raise ValueError(f"attention colscale must be [B={B}, L], got {tuple(cs.shape)}")
        e = _lib.UspAttnEdit(C.c_void_p(cs.data_ptr()), int(attn_edit.get("block_mask", (1 << 64) - 1)) & ((1 << 64) - 1),
                             float(attn_edit.get("t_edit", float("inf"))))
        return e, cs

    def forward(self, x, t, y=None, context=None, attn_edit=None, edit_loc=None, delta=None, write_scale=0.0,
                read=False):
        """One velocity evaluation.  ``edit_loc`` "head" / "tail" with ``delta`` [C,S,S] adds ``delta * write_scale`` to
        the latent before the embedding / to the velocity (libs/uvit.py:313-314,349-350); with ``read=True`` the call
        returns (out, activation at edit_loc) instead (the hook's "read" mode, libs/dissection.py:126-136)."""
        self._check_latent(x)
        B = x.shape[0]
        x = x.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.float32).expand(B).contiguous()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        if delta is not None:
            delta = torch.as_tensor(delta).to(self.device, torch.float32).contiguous()
            if tuple(delta.shape[-3:]) != (self.C, self.S, self.S) or delta.numel() != self.C * self.S * self.S:
                raise ValueError(f"delta must be [{self.C},{self.S},{self.S}], got {tuple(delta.shape)}")
        out = torch.empty_like(x)
        trace = torch.empty_like(x) if read else None
        edit, _keep = self._attn_edit(attn_edit, B)
        _lib.check(self.lib.usp_forward_hook(self.handle, _ptr(x), _ptr(t), _ptr(context), _ptr(y), _ptr(out), B,
                                             _lib.EDIT_LOC[edit_loc], _ptr(delta), float(write_scale), _ptr(trace),
                                             C.byref(edit) if edit is not None else None, self._stream()),
                   self.handle, "usp_forward")
        return (out, trace) if read else out

    def sample(self, z, t0=0.0, t1=1.0, step_size=0.02, method="euler", y=None, context=None, delta_table=None,
               write_scale=0.0, t_edit=0.0, edit_loc=None, attn_edit=None) -> torch.Tensor:
        """Device-resident fixed-grid integration; returns a new tensor (z is not modified)."""
        self._check_latent(z)
        B = z.shape[0]
        zz = z.to(self.device, torch.float32).contiguous().clone()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        if delta_table is not None:
            delta_table = delta_table.to(self.device, torch.float32).contiguous()
            n = self.lib.usp_grid_size(t0, t1, step_size)
            if delta_table.shape != (n, self.C, self.S, self.S):
                raise ValueError(f"delta_table must be [{n},{self.C},{self.S},{self.S}]")
        edit, _keep = self._attn_edit(attn_edit, B)
        _lib.check(self.lib.usp_sample_edit(self.handle, _ptr(zz), _ptr(context), _ptr(y), B, t0, t1, step_size,
                                            _lib.METHOD[method], _ptr(delta_table), write_scale, t_edit,
                                            _lib.EDIT_LOC[edit_loc], C.byref(edit) if edit is not None else None,
                                            self._stream()), self.handle, "usp_sample")
        return zz

    def sample_adaptive(self, z, t0=0.0, t1=1.0, rtol=1e-5, atol=1e-5, y=None, context=None, delta_digits=None,
                        write_scale=0.0, t_edit=0.0, edit_loc=None, attn_edit=None, max_steps=0, stats=None,
                        method="dopri5"):
        """Adaptive dopri5 / bosh3 / adaptive_heun (torchdiffeq semantics) from t0 to t1; returns a new tensor.  ``delta_digits`` rows are
        keyed by the "%.2f" digit of the evaluation time.  ``stats`` (dict) receives n_accept / n_reject / nfe.
        Synchronises the current stream (the host reads the step controller's verdict once per attempted step)."""
        self._check_latent(z)
        B = z.shape[0]
        zz = z.to(self.device, torch.float32).contiguous().clone()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        n_rows = 0
        if delta_digits is not None:
            delta_digits = delta_digits.to(self.device, torch.float32).contiguous()
            n_rows = delta_digits.shape[0]
            if delta_digits.shape[1:] != (self.C, self.S, self.S) or not 1 <= n_rows <= 128:
                raise ValueError(f"delta_digits must be [<=128,{self.C},{self.S},{self.S}]")
        edit, _keep = self._attn_edit(attn_edit, B)
        st = _lib.UspAdaptiveStats()
        _lib.check(self.lib.usp_sample_adaptive(self.handle, _ptr(zz), _ptr(context), _ptr(y), B, t0, t1,
                                                _lib.ADAPTIVE_METHOD[method], rtol, atol,
                                                _ptr(delta_digits), n_rows, write_scale, t_edit,
                                                _lib.EDIT_LOC[edit_loc], C.byref(edit) if edit is not None else None,
                                                int(max_steps), C.byref(st), self._stream()),
                   self.handle, "usp_sample_adaptive")
        if stats is not None:
            stats.update(n_accept=st.n_accept, n_reject=st.n_reject, nfe=st.nfe, last_ratio=st.last_ratio,
                         last_dt=st.last_dt)
        return zz

    def sample_adaptive_read(self, z, t0=0.0, t1=1.0, rtol=1e-5, atol=1e-5, y=None, context=None, edit_loc="tail",
                             max_steps=0, stats=None, method="dopri5", trace_cap=128):
        """dissect_name="read" under an adaptive solver: returns (result, trace [n_evals, B, C, S, S], times [n_evals]):
        the activation at ``edit_loc`` and the model time of every velocity evaluation, in evaluation order."""
        self._check_latent(z)
        B = z.shape[0]
        zz = z.to(self.device, torch.float32).contiguous().clone()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        trace = torch.zeros(trace_cap, B, self.C, self.S, self.S, device=self.device, dtype=torch.float32)
        times = torch.zeros(trace_cap, dtype=torch.float32)        # host
        n_evals = C.c_int(0)
        st = _lib.UspAdaptiveStats()
        _lib.check(self.lib.usp_sample_adaptive_read(self.handle, _ptr(zz), _ptr(context), _ptr(y), B, t0, t1,
                                                     _lib.ADAPTIVE_METHOD[method], rtol, atol, _lib.EDIT_LOC[edit_loc],
                                                     _ptr(trace), _ptr(times), int(trace_cap), C.byref(n_evals),
                                                     int(max_steps), C.byref(st), self._stream()),
                   self.handle, "usp_sample_adaptive_read")
        if stats is not None:
            stats.update(n_accept=st.n_accept, n_reject=st.n_reject, nfe=st.nfe, last_ratio=st.last_ratio,
                         last_dt=st.last_dt)
        return zz, trace[:n_evals.value], times[:n_evals.value]

    def sample_sweep(self, z, write_scales, t0=0.0, t1=1.0, step_size=0.02, method="euler", y=None, context=None,
                     delta_table=None, t_edit=0.0, edit_loc="tail") -> torch.Tensor:
        """All ``write_scales`` of the semantic-direction sweep in one batch: returns [B, len(write_scales), C, S, S]."""
        self._check_latent(z)
        B = z.shape[0]
        zz = z.to(self.device, torch.float32).contiguous()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        n = self.lib.usp_grid_size(t0, t1, step_size)
        if delta_table is None or delta_table.shape != (n, self.C, self.S, self.S):
            raise ValueError(f"delta_table must be [{n},{self.C},{self.S},{self.S}]")
        delta_table = delta_table.to(self.device, torch.float32).contiguous()
        scales = (C.c_float * len(write_scales))(*[float(s) for s in write_scales])
        out = torch.empty((B, len(write_scales), self.C, self.S, self.S), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.usp_sample_sweep(self.handle, _ptr(zz), _ptr(out), _ptr(context), _ptr(y), B, scales,
                                             len(write_scales), t0, t1, step_size, _lib.METHOD[method],
                                             _ptr(delta_table), t_edit, _lib.EDIT_LOC[edit_loc], self._stream()),
                   self.handle, "usp_sample_sweep")
        return out

    def sample_read(self, z, t0=0.0, t1=1.0, step_size=0.02, method="euler", y=None, context=None, edit_loc="tail"):
        """Integrate and also return the activation at ``edit_loc`` of every velocity evaluation:
        (z_end, trace [grid points, B, C, S, S]); trace[i] belongs to the evaluation at grid[i]."""
        self._check_latent(z)
        B = z.shape[0]
        zz = z.to(self.device, torch.float32).contiguous().clone()
        if y is not None:
            y = y.to(self.device, torch.int64).contiguous()
        if context is not None:
            context = context.to(self.device, torch.float32).contiguous()
        n = self.lib.usp_grid_size(t0, t1, step_size)
        if n < 2:
            raise ValueError(f"bad time grid: t0={t0} t1={t1} step_size={step_size}")
        trace = torch.empty((n, B, self.C, self.S, self.S), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.usp_sample_read(self.handle, _ptr(zz), _ptr(context), _ptr(y), B, t0, t1, step_size,
                                            _lib.METHOD[method], _lib.EDIT_LOC[edit_loc], _ptr(trace), self._stream()),
                   self.handle, "usp_sample_read")
        return zz, trace

    def sample_host(self, z_host, t0=0.0, t1=1.0, step_size=0.02, method="euler", y=None, context=None,
                    delta_table=None, write_scale=0.0, t_edit=0.0, edit_loc=None) -> torch.Tensor:
        """End-to-end call on HOST tensors (ideally pinned): H2D, sample, D2H, synchronise. In place on z_host."""
        for tns in (z_host, y, context, delta_table):
            if tns is not None and (tns.device.type != "cpu" or not tns.is_contiguous()):
                raise ValueError("sample_host takes contiguous CPU tensors")
        self._check_latent(z_host)
        _lib.check(self.lib.usp_sample_host(self.handle, _ptr(z_host), _ptr(context), _ptr(y), z_host.shape[0],
                                            t0, t1, step_size, _lib.METHOD[method], _ptr(delta_table),
                                            write_scale, t_edit, _lib.EDIT_LOC[edit_loc]), self.handle,
                   "usp_sample_host")
        return z_host

    def profile_forward(self, x, t, y=None, context=None, warmup=0, reps=1):
        """Eager forwards with an event between launches -> {kernel class: (ms, launches)} per evaluation (mean over
        ``reps`` evaluations enqueued back to back after ``warmup`` untimed ones)."""
        B = x.shape[0]
        x = x.to(self.device, torch.float32).contiguous()
        t = t.to(self.device, torch.float32).expand(B).contiguous()
        out = torch.empty_like(x)
        n = len(_lib.KERNEL_CLASSES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        _lib.check(self.lib.usp_profile_forward_n(self.handle, _ptr(x), _ptr(t), _ptr(context), _ptr(y), _ptr(out), B,
                                                  int(warmup), int(reps), ms, cnt, self._stream()),
                   self.handle, "usp_profile_forward_n")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def nonfinite(self) -> bool:
        """True if any velocity evaluation since the last call produced inf / NaN (fp16 operand overflow); reads and
        clears the device flag, synchronising the current stream."""
        return bool(self.status_flags() & 1)

    def status_flags(self) -> int:
        """Reads and clears the device status word (synchronises): bit 1 = a non-finite velocity, bit 2 = a token whose
        mean exceeded 4 standard deviations reached a folded LayerNorm (precision loss in its 16-bit operand)."""
        flag = C.c_int(0)
        _lib.check(self.lib.usp_nonfinite(self.handle, C.byref(flag), self._stream()), self.handle, "usp_nonfinite")
        return int(flag.value)

    # ---- introspection ---------------------------------------------------------------------------
    def last_ms(self) -> float:
        ms = C.c_float()
        _lib.check(self.lib.usp_last_forward_ms(self.handle, C.byref(ms)), self.handle, "usp_last_forward_ms")
        return ms.value

    def kernels_per_forward(self) -> int:
        return self.lib.usp_kernels_per_forward(self.handle)

    def flops_per_forward(self) -> float:
        return self.lib.usp_flops_per_forward(self.handle)

    def grid_size(self, t0, t1, step_size) -> int:
        return self.lib.usp_grid_size(t0, t1, step_size)


def time_grid(t0: float, t1: float, step_size: float):
    """The fp32 time grid the sampler will use (torchdiffeq fixed-grid construction), as a list of floats."""
    lib = _lib.load()
    n = lib.usp_grid_size(t0, t1, step_size)
    if n < 2:
        raise ValueError(f"bad time grid: t0={t0} t1={t1} step_size={step_size}")
    buf = (C.c_float * n)()
    if lib.usp_time_grid(t0, t1, step_size, buf, n) != n:
        raise RuntimeError("usp_time_grid failed")
    return [float(v) for v in buf]
