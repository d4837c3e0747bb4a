"""Mirror of the reference ``CNF`` wrappers (flow_matching.py:15-180, flow_matching_t2i.py:14-175).

``decode`` / ``encode`` keep the reference signatures and the ``solver_kwargs`` dictionary, but the
``torchdiffeq.odeint`` call (flow_matching.py:118-125,140-147) is replaced by the library's fixed-grid
Euler / Heun loop: one CUDA-graph-captured step replayed on the current stream, no host sync per NFE.

Built: ``solver="fixed"`` with ``solver_fix`` in {"euler", "heun", "midpoint", "rk4"}; ``solver="adaptive"`` / the non-dissection
default / the adaptive tail of ``"fixadp"`` with dopri5, bosh3 or adaptive_heun (step control and dense output on the device, torchdiffeq
semantics restated in csrc/ode.cu); the "write_attr" / "write_pca" edit hook at ``edit_loc`` head / tail
(libs/dissection.py:115-186) through a pre-loaded delta table, and its "read" mode (the activation at ``edit_loc``
of every evaluation, saved as ``{batch_id}_{t:.2f}.npy``) on the fixed grid; the "p2p_rescale" attention edit.
Not built (NotImplementedError): other torchdiffeq methods (dopri8, fehlberg2, implicit_adams, ...), "read" under an
adaptive solver, ``edit_loc="mid"`` (broken in the reference for U-ViT).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .engine import time_grid

_FIXED_METHODS = ("euler", "heun", "heun2", "midpoint", "rk4")   # "heun2" is torchdiffeq's name for the 2-stage Heun
_ADAPTIVE_METHODS = ("dopri5", "bosh3", "adaptive_heun")
_RTOL = 1e-5   # flow_matching.py:11-12
_ATOL = 1e-5


def should_edit(timestep_digit: str, t_edit) -> bool:
    """libs/dissection.py:21-34."""
    if timestep_digit == "0.00":
        return False
    if isinstance(t_edit, (float, int)):
        return float(timestep_digit) <= t_edit
    if isinstance(t_edit, str) and t_edit.startswith("every_"):
        return float(timestep_digit) % float(t_edit.replace("every_", "")) == 0.0
    raise ValueError(f"t_edit={t_edit!r}")


def _read_delta(path: str, ith) -> np.ndarray:
    """libs/dissection.py:55-70: row ``ith`` of a [n, C, W, H] .npy, or the mean of rows "a_b_c"."""
    arr = np.load(path)
    if isinstance(ith, (int, np.integer)):
        return arr[int(ith)]
    if isinstance(ith, str):
        ids = [int(s) for s in ith.split("_")]
        return sum(arr[i] for i in ids) / len(ids)
    raise TypeError(f"ith element must be int or 'a_b_c' string, got {ith!r}")


def build_delta_table(grid, shape, **kwargs):
    """Pre-gather what dissect_helper_uvit would np.load at each NFE (libs/dissection.py:139-183).

    Returns (table [len(grid), C, S, S] float32 with zero rows where no edit applies, edit_loc) or (None, None).
    """
    name = kwargs.get("dissect_name")
    if kwargs.get("dissect_task") != "uspace_uvit" or name in (None, "none", "read"):
        return None, None
    if name not in ("write_attr", "write_pca"):
        raise ValueError(f"dissect_name should be read or write, here is {name}")
    loc = kwargs.get("edit_loc")
    if loc == "mid":
        raise NotImplementedError("edit_loc='mid' is broken in the reference for U-ViT and is not built")
    if loc not in ("head", "tail"):
        return None, None
    root = kwargs["write_path_root"]
    table = np.zeros((len(grid),) + tuple(shape), dtype=np.float32)
    for i, t in enumerate(grid):
        digit = f"{t:.2f}"
        if not should_edit(digit, kwargs.get("t_edit")):
            continue
        if name == "write_attr":
            table[i] = _read_delta(os.path.join(root, f"delta_{digit}.npy"), kwargs.get("ith_attr"))
        else:
            table[i] = _read_delta(os.path.join(root, f"pca{kwargs.get('pca_n')}_{digit}.npy"),
                                   kwargs.get("ith_component"))
    return torch.from_numpy(table), loc


def build_delta_digits(shape, n_rows: int = 101, missing=None, **kwargs):
    """The same edit for a solver whose evaluation times are not known in advance: row i holds what the hook would
    load for timestep_digit f"{i/100:.2f}" (libs/dissection.py:139-183).  The reference raises FileNotFoundError when
    the solver evaluates at a digit without a file; an adaptive solver's evaluation times are not known here, so a
    missing file gives a zero row, its digit is appended to ``missing`` (reported in ``CNF.last_solver_stats`` with a
    one-time warning), and a root without ANY matching file raises FileNotFoundError like the reference would.
    Returns (table, edit_loc) or (None, None)."""
    name = kwargs.get("dissect_name")
    if kwargs.get("dissect_task") != "uspace_uvit" or name in (None, "none"):
        return None, None
    if name == "read":
        raise NotImplementedError("dissect_name='read' under an adaptive solver is not built")
    if name not in ("write_attr", "write_pca"):
        raise ValueError(f"dissect_name should be read or write, here is {name}")
    loc = kwargs.get("edit_loc")
    if loc == "mid":
        raise NotImplementedError("edit_loc='mid' is broken in the reference for U-ViT and is not built")
    if loc not in ("head", "tail"):
        return None, None
    root = kwargs["write_path_root"]
    table = np.zeros((n_rows,) + tuple(shape), dtype=np.float32)
    wanted = n_missing = 0
    for i in range(n_rows):
        digit = f"{i / 100:.2f}"
        if not should_edit(digit, kwargs.get("t_edit")):
            continue
        path = os.path.join(root, f"delta_{digit}.npy" if name == "write_attr" else f"pca{kwargs.get('pca_n')}_{digit}.npy")
        wanted += 1
        if os.path.exists(path):
            table[i] = _read_delta(path, kwargs.get("ith_attr") if name == "write_attr" else kwargs.get("ith_component"))
        elif missing is not None:
            missing.append(digit)
        else:
            n_missing += 1
    n_missing += len(missing) if missing is not None else 0
    if wanted and n_missing == wanted:
        raise FileNotFoundError(f"no delta / pca file for any edited timestep under {root!r}")
    if n_missing:
        import warnings
        warnings.warn(f"uspace_b200: {n_missing} of {wanted} edited timestep digits have no file under {root!r}; the "
                      "edit is skipped there (the reference raises FileNotFoundError if the solver evaluates at one)")
    return torch.from_numpy(table), loc


_ATTN_EDIT_NAMES = ("p2p", "local_prompt", "sampled_image_editing")


def block_mask_from_ids(block_id) -> int:
    """should_edit_attention_by_blockids (tools/utils_t2i.py:227-240) as a bit mask over executed blocks."""
    if block_id is None or (isinstance(block_id, str) and block_id == "all"):
        return (1 << 64) - 1
    if isinstance(block_id, int):
        return 1 << block_id
    if isinstance(block_id, (list, tuple)):
        m = 0
        for b in block_id:
            m |= 1 << int(b)
        return m
    raise ValueError(f"unknown target_block_id {block_id}")


def build_attn_edit(B: int, L: int, **kwargs):
    """The attention edit the reference would apply for these kwargs (libs/uvit_t2i.py:91-107 ->
    tools/utils_t2i.py:265-296,196-224), or None.  Only "p2p_rescale" changes the attention map ("lp_*" edits act on
    the prompt / context, not on the map; "p2p_replace" raises in the reference too)."""
    if kwargs.get("dissect_name") not in _ATTN_EDIT_NAMES:
        return None
    if kwargs.get("fm_direction") == "encode":
        return None
    tk = kwargs.get("token_kwargs") or {}
    mode = tk.get("token_dissect")
    if mode is None or str(mode).startswith("lp_"):
        return None
    if mode != "p2p_rescale":
        raise NotImplementedError(f"token_dissect={mode!r}")
    ids = kwargs["target_context_ids"]
    mult = tk["p2p_multiplier"]
    mults = [mult] * len(ids) if isinstance(mult, (int, float)) else list(mult)
    cs = torch.ones(B, L, dtype=torch.float32)
    for i, tid in enumerate(ids):
        tid = np.asarray(tid, dtype=np.int64)
        if tid.size > 0:
            cs[i, torch.from_numpy(tid) + 1] = float(mults[i])   # + TIME_TOKEN_NUM: [time, ctx x 77, patches]
    return dict(colscale=cs, block_mask=block_mask_from_ids(kwargs.get("block_id")), t_edit=float(kwargs["t_edit"]))


class _CNFBase(nn.Module):
    # decode / encode end with one 4-byte read of the library's non-finite flag (a stream synchronisation) and raise
    # when a velocity evaluation overflowed the fp16 operand range; set False to keep the call fully asynchronous
    check_overflow = True

    def __init__(self, net):
        super().__init__()
        self.net = net

    def _checked(self, engine, out: Tensor) -> Tensor:
        flags = engine.status_flags() if self.check_overflow else 0
        if flags & 1:
            raise FloatingPointError(
                "uspace_b200: a velocity evaluation produced inf / NaN - activations of this checkpoint exceed the "
                "fp16 tensor-core operand range (65504); set net.operand_dtype = 'bf16'")
        if flags & 2:
            raise FloatingPointError(
                "uspace_b200: a token's mean exceeded 4 standard deviations at a folded LayerNorm, where the 16-bit "
                "operand loses precision; set net.fuse_layernorm = False for this checkpoint")
        return out

    def _call_net(self, x, t, cond, **kwargs):
        raise NotImplementedError

    def _cond_kw(self, cond):
        raise NotImplementedError

    def _velocity(self, t: Tensor, x: Tensor, cond, **kwargs) -> Tensor:
        """flow_matching.py:23-36 (the t.item() logging — a host sync per NFE — is dropped)."""
        if t.numel() == 1:
            t = t.expand(x.size(0))
        pred, _ = self._call_net(x, t, cond, **kwargs)
        return pred

    def is_dissection_mode(self, kwargs):
        return "dissect_name" in kwargs and kwargs["dissect_name"] is not None

    def get_ode_kwargs(self, **kwargs):
        """flow_matching.py:38-85: the odeint keyword set (a (fixed, adaptive) pair for "fixadp")."""
        if not self.is_dissection_mode(kwargs):
            return dict(method="dopri5", rtol=_RTOL, atol=_ATOL)
        sk = kwargs["solver_kwargs"]
        if sk["solver"] == "fixed":
            return self._fixed_kwargs(sk)
        if sk["solver"] == "adaptive":
            return self._adaptive_kwargs(sk)
        if sk["solver"] == "fixadp":
            return self._fixed_kwargs(sk), self._adaptive_kwargs(sk)
        raise NotImplementedError(f"solver={sk['solver']}")

    @staticmethod
    def _adaptive_kwargs(sk):
        if sk["solver_adaptive"] not in _ADAPTIVE_METHODS:
            raise NotImplementedError(f"solver_adaptive={sk['solver_adaptive']!r}: built methods are {_ADAPTIVE_METHODS}")
        return dict(method=sk["solver_adaptive"], rtol=_RTOL, atol=_ATOL)

    @staticmethod
    def _fixed_kwargs(sk):
        if sk["solver_fix"] not in _FIXED_METHODS:
            raise NotImplementedError(f"solver_fix={sk['solver_fix']!r}: built methods are {_FIXED_METHODS}")
        method = "heun" if sk["solver_fix"] == "heun2" else sk["solver_fix"]
        return dict(method=method, rtol=_RTOL, atol=_ATOL, options=dict(step_size=float(sk["solver_fix_step"])))

    def training_losses(self, x, cond, sigma_min, **kwargs):
        """flow_matching.py:88-100 (runs the differentiable PyTorch graph of the mirror module)."""
        noise = torch.randn_like(x)
        t = torch.rand(len(x), device=x.device, dtype=x.dtype)
        t_ = t[:, None, None, None]
        x_new = t_ * x + (1 - (1 - sigma_min) * t_) * noise
        u = x - (1 - sigma_min) * noise
        return (self._velocity(t, x_new, cond, **kwargs) - u).square().mean(dim=(1, 2, 3))

    @torch.no_grad()
    def _integrate(self, z: Tensor, cond, t0: float, t1: float, ode_kwargs: dict, **kwargs) -> Tensor:
        """odeint(func, z, [t0, t1], **ode_kwargs)[-1] on the library: fixed grid if ode_kwargs carries a step_size,
        adaptive dopri5 otherwise."""
        net = self.net.module if hasattr(self.net, "module") else self.net  # DDP / accelerate wrapper
        engine = net.engine()
        attn = None
        if isinstance(self, CNFT2I):
            attn = build_attn_edit(z.shape[0], engine.cfg.num_clip_token + 1 + (engine.S // engine.cfg.patch_size) ** 2,
                                   **kwargs)
            if attn is not None and getattr(net, "qk_scale", None) is not None:
                raise NotImplementedError("attention editing with a qk_scale override is not built "
                                          "(the editing branch of libs/uvit_t2i.py:96-107 would use it)")
        dissect = self.is_dissection_mode(kwargs)
        reading = dissect and kwargs.get("dissect_task") == "uspace_uvit" and kwargs.get("dissect_name") == "read"
        if reading:
            return self._integrate_read(engine, z, cond, t0, t1, ode_kwargs, **kwargs)
        if "options" in ode_kwargs:
            h = ode_kwargs["options"]["step_size"]
            table, loc = build_delta_table(time_grid(t0, t1, h), z.shape[1:], **kwargs) if dissect else (None, None)
            ws = float(kwargs.get("write_scale") or 0.0) if table is not None else 0.0
            # rows where should_edit() is false are zero, so the library's own mask can stay wide open
            if table is not None and ode_kwargs["method"] in ("midpoint", "rk4"):
                # the reference's hook keys on f"{t:.2f}" at EVERY evaluation, so midpoint / rk4 would load (or fail to
                # find) delta files for their in-between stage times; the library's table is keyed by grid point only
                raise NotImplementedError(
                    f"write edits under solver_fix={ode_kwargs['method']!r}: the stages between grid points have no "
                    "row in the grid-keyed delta table (use euler / heun, or an adaptive solver)")
            return self._checked(engine, engine.sample(z, t0, t1, h, ode_kwargs["method"], delta_table=table,
                                                       write_scale=ws, t_edit=float("inf"), edit_loc=loc,
                                                       attn_edit=attn, **self._cond_kw(cond)))
        if ode_kwargs["method"] not in _ADAPTIVE_METHODS:
            raise NotImplementedError(f"method={ode_kwargs['method']!r}")
        missing = []
        table, loc = build_delta_digits(z.shape[1:], missing=missing, **kwargs) if dissect else (None, None)
        self._missing_digits = missing
        ws = float(kwargs.get("write_scale") or 0.0) if table is not None else 0.0
        self.last_solver_stats = {}
        if table is not None:
            self.last_solver_stats["delta_rows_missing"] = self._missing_digits
        return self._checked(engine, engine.sample_adaptive(
            z, t0, t1, ode_kwargs["rtol"], ode_kwargs["atol"], delta_digits=table, write_scale=ws,
            t_edit=float("inf"), edit_loc=loc, attn_edit=attn, stats=self.last_solver_stats,
            method=ode_kwargs["method"], **self._cond_kw(cond)))

    def _integrate_read(self, engine, z, cond, t0, t1, ode_kwargs, **kwargs) -> Tensor:
        """dissect_name="read" (libs/dissection.py:126-136): np.save(f"{read_path_root}/{batch_id}_{t:.2f}", x) for the
        activation x at edit_loc of every evaluation - gathered on the device, written after the last step."""
        loc = kwargs.get("edit_loc")
        if loc == "mid":
            raise NotImplementedError("edit_loc='mid' is broken in the reference for U-ViT and is not built")
        if "options" not in ode_kwargs:
            # adaptive solver: the evaluation times are the solver's; one file per "%.2f" digit, the last evaluation that
            # printed the digit wins (the reference overwrites {batch_id}_{digit}.npy the same way)
            self.last_solver_stats = {}
            kw = dict(rtol=ode_kwargs["rtol"], atol=ode_kwargs["atol"], method=ode_kwargs.get("method", "dopri5"),
                      stats=self.last_solver_stats, **self._cond_kw(cond))
            if loc not in ("head", "tail"):
                return engine.sample_adaptive(z, t0, t1, **kw)
            out, trace, times = engine.sample_adaptive_read(z, t0, t1, edit_loc=loc, **kw)
            root = kwargs.get("read_path_root")
            os.makedirs(root, exist_ok=True)
            last = {}                                   # digit -> last evaluation that printed it
            for i, t in enumerate(times.tolist()):
                last[f"{t:.2f}"] = i
            host = trace.cpu().numpy()
            for digit, i in last.items():
                np.save(os.path.join(root, f"{kwargs['batch_id']}_{digit}"), host[i])
            return out
        h, method = ode_kwargs["options"]["step_size"], ode_kwargs["method"]
        if loc not in ("head", "tail"):     # the hook is never reached: a plain integration
            return engine.sample(z, t0, t1, h, method, **self._cond_kw(cond))
        out, trace = engine.sample_read(z, t0, t1, h, method, edit_loc=loc, **self._cond_kw(cond))
        root = kwargs.get("read_path_root")
        os.makedirs(root, exist_ok=True)
        grid = time_grid(t0, t1, h)
        n_eval = len(grid) if method == "heun" else len(grid) - 1   # Euler never evaluates at the last grid point
        host = trace[:n_eval].cpu().numpy()
        for i in range(n_eval):
            np.save(os.path.join(root, f"{kwargs['batch_id']}_{grid[i]:.2f}"), host[i])
        return out

    def _decode(self, z: Tensor, cond, **kwargs) -> Tensor:
        """flow_matching.py:130-180 (decode + decode_fixadp)."""
        solver = kwargs["solver_kwargs"]["solver"]
        if solver in ("fixed", "adaptive"):
            return self._integrate(z, cond, 0.0, 1.0, self.get_ode_kwargs(**kwargs), **kwargs)
        if solver == "fixadp":
            t_mid = kwargs["t_edit"]
            assert t_mid >= 0 and t_mid <= 1, f"t_mid={t_mid}"
            fixed_kw, adaptive_kw = self.get_ode_kwargs(**kwargs)
            mid = self._integrate(z, cond, 0.0, float(t_mid), fixed_kw, **kwargs) if t_mid > 0 else z
            if t_mid >= 1:
                return mid
            return self._integrate(mid, cond, float(t_mid), 1.0, adaptive_kw, **kwargs)
        raise NotImplementedError(f"unknown solver {kwargs['solver_kwargs']}")


class CNF(_CNFBase):
    """Mirror of flow_matching.py::CNF (unconditional / class-conditional, ``y=``)."""

    def _call_net(self, x, t, y, **kwargs):
        return self.net(x, t, y, **kwargs)

    def _cond_kw(self, y):
        return dict(y=y)

    def forward(self, t: Tensor, x: Tensor, y: Tensor = None, **kwargs) -> Tensor:
        return self._velocity(t, x, y, **kwargs)

    def training_losses(self, x, y, sigma_min, **kwargs):
        return super().training_losses(x, y, sigma_min, **kwargs)

    def encode(self, x: Tensor, y: Tensor = None, **kwargs) -> Tensor:
        """flow_matching.py:102-125: always the fixed solver, t: 1 -> 0."""
        return self._integrate(x, y, 1.0, 0.0, self._fixed_kwargs(kwargs["solver_kwargs"]), **kwargs)

    def decode(self, z: Tensor, y: Tensor = None, **kwargs) -> Tensor:
        return self._decode(z, y, **kwargs)


class CNFT2I(_CNFBase):
    """Mirror of flow_matching_t2i.py::CNF (``context=``)."""

    def _call_net(self, x, t, context, **kwargs):
        return self.net(x, t, context=context, **kwargs)

    def _cond_kw(self, context):
        return dict(context=context)

    def forward(self, t: Tensor, x: Tensor, context: Tensor, **kwargs) -> Tensor:
        return self._velocity(t, x, context, **kwargs)

    def training_losses(self, x, context, sigma_min, **kwargs):
        return super().training_losses(x, context, sigma_min, **kwargs)

    def encode(self, x: Tensor, context: Tensor, **kwargs) -> Tensor:
        """flow_matching_t2i.py:105-122: goes through get_ode_kwargs, t: 1 -> 0."""
        kwargs.update({"fm_direction": "encode"})
        return self._integrate(x, context, 1.0, 0.0, self.get_ode_kwargs(**kwargs), **kwargs)

    def decode(self, z: Tensor, context: Tensor, **kwargs) -> Tensor:
        kwargs.update({"fm_direction": "decode"})
        return self._decode(z, context, **kwargs)
