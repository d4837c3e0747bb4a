"""CPU oracle for the U-ViT velocity field and the fixed-grid ODE loop.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (uspace_b200/) never does.

It is a restatement (not a copy) of the reference algorithm in plain torch fp32/fp64 tensor arithmetic
with explicit index maps, working on a flat ``state_dict`` (same keys as the reference modules):

  timestep_embedding   libs/uvit.py:26-46
  patch embedding      libs/uvit.py:165-179   (Conv2d k=s=p == (C,p1,p2) gather + matmul)
  token assembly       libs/uvit.py:316-327, libs/uvit_t2i.py:316-324
  attention            libs/uvit.py:86-118    (qkv split "(K H D)", softmax(QK^T/sqrt(hd))V, merge "(H D)")
  mlp                  libs/timm.py:96-112    (fc1 -> exact-erf GELU -> fc2)
  block                libs/uvit.py:157-162   (skip_linear(cat[x, skip]); pre-LN residuals, eps 1e-5)
  output head          libs/uvit.py:342-347   (LN, decoder_pred, drop extras, unpatchify (p1,p2,C), conv3x3)
  edit hook            libs/dissection.py:21-34,115-157 ("write_attr"/"write_pca": x + delta[t]*write_scale)
  ODE                  flow_matching.py:102-151 + torchdiffeq's documented FixedGridODESolver semantics
                       (torchdiffeq is an un-vendored, unpinned dependency: README.md:121)

Pinning: the reference ships no tests or golden vectors.  The model part is pinned against
the reference modules themselves (imported from /root/reference by tests/golden/make_golden.py, outputs
committed under tests/golden/).  The ODE loop is "parity unpinned": torchdiffeq is not installed anywhere
in this environment, so its fixed-grid algorithm is restated from its published behaviour (grid lengths
51/101/41 probe-verified in SURVEY.md §8c).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration helpers
# ----------------------------------------------------------------------------------------------
def model_dims(cfg: dict) -> dict:
    """Derived sizes for a reference ctor-kwargs dict (libs/uvit.py:183-202 / libs/uvit_t2i.py:193-211)."""
    D = cfg["embed_dim"]
    p = cfg["patch_size"]
    n_patch = (cfg["img_size"] // p) ** 2
    n_ctx = cfg.get("num_clip_token", 0) if "clip_dim" in cfg else 0
    if n_ctx:
        extras = 1 + n_ctx
    else:
        extras = 2 if cfg.get("num_classes", -1) > 0 else 1
    return dict(
        D=D, p=p, C=cfg["in_chans"], S=cfg["img_size"], n_patch=n_patch, extras=extras, L=extras + n_patch,
        H=cfg["num_heads"], hidden=int(D * cfg.get("mlp_ratio", 4.0)), n_in=cfg["depth"] // 2,
        n_ctx=n_ctx, P=p * p * cfg["in_chans"],
    )


def flops_per_forward(cfg: dict) -> float:
    """algorithmic FLOPs per image per forward (multiply-add = 2)."""
    d = model_dims(cfg)
    D, L = d["D"], d["L"]
    n_blk = 2 * d["n_in"] + 1
    n_skip = d["n_in"] if cfg.get("skip", True) else 0
    f = n_blk * (2 * L * D * 3 * D + 2 * L * D * D + 4 * L * D * d["hidden"] + 4 * L * L * D)
    f += n_skip * 4 * L * D * D
    f += 2 * d["n_patch"] * d["P"] * D + 2 * L * D * d["P"]
    if d["n_ctx"]:
        f += 2 * d["n_ctx"] * cfg["clip_dim"] * D
    if cfg.get("conv", True):
        f += 2 * d["C"] * d["C"] * 9 * d["S"] * d["S"]
    return float(f)


# ----------------------------------------------------------------------------------------------
# index maps (integer work: bit-exact targets)
# ----------------------------------------------------------------------------------------------
def patchify_index(C: int, S: int, p: int) -> Tensor:
    """[n_patch, C*p*p] flat indices into a [C,S,S] image; feature order (C, p1, p2), token = h*gw + w."""
    gw = S // p
    idx = torch.empty(gw * gw, C * p * p, dtype=torch.long)
    for ph in range(gw):
        for pw in range(gw):
            f = 0
            for c in range(C):
                for p1 in range(p):
                    for p2 in range(p):
                        idx[ph * gw + pw, f] = (c * S + ph * p + p1) * S + pw * p + p2
                        f += 1
    return idx


def unpatchify_index(C: int, S: int, p: int) -> Tensor:
    """[C*S*S] flat indices into a [n_patch, p*p*C] token matrix; feature order (p1, p2, C)."""
    gw = S // p
    P = p * p * C
    idx = torch.empty(C * S * S, dtype=torch.long)
    for c in range(C):
        for y in range(S):
            for x in range(S):
                tok = (y // p) * gw + (x // p)
                feat = ((y % p) * p + (x % p)) * C + c
                idx[(c * S + y) * S + x] = tok * P + feat
    return idx


# ----------------------------------------------------------------------------------------------
# model pieces
# ----------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(device=t.device, dtype=t.dtype)
    args = t[:, None] * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# FAST=True swaps the explicit formulas for the fused library calls the reference itself makes (F.linear,
# F.layer_norm, F.gelu, F.scaled_dot_product_attention).  Same arithmetic; used only so that the CPU baseline
# timed by bench.py is as fast as the reference's own torch-eager path (tests check FAST == explicit).
FAST = False


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    if FAST:
        return F.linear(x, w, b)
    y = x @ w.T
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    if FAST:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    if FAST:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def attention(x: Tensor, sd: Dict[str, Tensor], pre: str, H: int, vscale: Optional[Tensor] = None) -> Tensor:
    B, L, D = x.shape
    hd = D // H
    qkv = linear(x, sd[pre + ".attn.qkv.weight"], sd.get(pre + ".attn.qkv.bias"))
    q, k, v = (qkv[..., i * D:(i + 1) * D].reshape(B, L, H, hd).permute(0, 2, 1, 3) for i in range(3))
    if vscale is not None:  # post-softmax column re-weighting == scaling rows of V (tools/utils_t2i.py:196-224)
        v = v * vscale[:, None, :, None]
    if FAST:
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        o = torch.softmax(s, dim=-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(B, L, D)
    return linear(o, sd[pre + ".attn.proj.weight"], sd[pre + ".attn.proj.bias"])


def mlp(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    h = gelu_erf(linear(x, sd[pre + ".mlp.fc1.weight"], sd[pre + ".mlp.fc1.bias"]))
    return linear(h, sd[pre + ".mlp.fc2.weight"], sd[pre + ".mlp.fc2.bias"])


def block(x: Tensor, sd: Dict[str, Tensor], pre: str, H: int, skip: Optional[Tensor] = None,
          vscale: Optional[Tensor] = None) -> Tensor:
    if pre + ".skip_linear.weight" in sd:
        x = linear(torch.cat([x, skip], dim=-1), sd[pre + ".skip_linear.weight"], sd[pre + ".skip_linear.bias"])
    x = x + attention(layer_norm(x, sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"]), sd, pre, H, vscale)
    x = x + mlp(layer_norm(x, sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"]), sd, pre)
    return x


def should_edit(t: float, t_edit) -> bool:
    digit = f"{t:.2f}"
    if digit == "0.00":
        return False
    if isinstance(t_edit, (int, float)):
        return float(digit) <= t_edit
    if isinstance(t_edit, str) and t_edit.startswith("every_"):
        return float(digit) % float(t_edit.replace("every_", "")) == 0.0
    raise ValueError(t_edit)


def uvit_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, t: Tensor, y: Optional[Tensor] = None,
                 context: Optional[Tensor] = None, head_delta: Optional[Tensor] = None,
                 tail_delta: Optional[Tensor] = None, attn_colscale: Optional[Tensor] = None,
                 attn_blocks=None) -> Tensor:
    """UViT.forward on a flat state_dict.  head_delta / tail_delta are already scaled [C,S,S] edits or None.
    attn_colscale [B, L] re-weights post-softmax attention columns (p2p_rescale, tools/utils_t2i.py:196-224) in the
    executed blocks listed in attn_blocks (None = all)."""
    d = model_dims(cfg)
    dt = x.dtype
    if any(v.dtype != dt for v in sd.values()):
        sd = {k: v.to(dt) for k, v in sd.items()}
    B = x.shape[0]
    C, S, p, D = d["C"], d["S"], d["p"], d["D"]
    if head_delta is not None:
        x = x + head_delta.to(dt)
    pidx = patchify_index(C, S, p)
    feats = x.reshape(B, C * S * S)[:, pidx.reshape(-1)].reshape(B, d["n_patch"], d["P"])
    tok = feats @ sd["patch_embed.proj.weight"].reshape(D, d["P"]).T + sd["patch_embed.proj.bias"]
    time_tok = timestep_embedding(t.to(dt), D)
    if "time_embed.0.weight" in sd:       # mlp_time_embed=True: Linear -> SiLU -> Linear (libs/uvit.py:215-223, :320)
        hid = time_tok @ sd["time_embed.0.weight"].T + sd["time_embed.0.bias"]
        hid = hid * torch.sigmoid(hid)
        time_tok = hid @ sd["time_embed.2.weight"].T + sd["time_embed.2.bias"]
    time_tok = time_tok[:, None, :]
    if d["n_ctx"]:
        ctx = context.to(dt) @ sd["context_embed.weight"].T + sd["context_embed.bias"]
        h = torch.cat([time_tok, ctx, tok], dim=1)
    else:
        h = torch.cat([time_tok, tok], dim=1)
        if y is not None:
            h = torch.cat([sd["label_emb.weight"][y][:, None, :], h], dim=1)
    h = h + sd["pos_embed"]
    skips: List[Tensor] = []
    bid = 0

    def vs():
        on = attn_colscale is not None and (attn_blocks is None or bid in attn_blocks)
        return attn_colscale.to(dt) if on else None

    for i in range(d["n_in"]):
        h = block(h, sd, f"in_blocks.{i}", d["H"], None, vs())
        skips.append(h)
        bid += 1
    h = block(h, sd, "mid_block", d["H"], None, vs())
    bid += 1
    for i in range(d["n_in"]):
        h = block(h, sd, f"out_blocks.{i}", d["H"], skips.pop(), vs())
        bid += 1
    h = layer_norm(h, sd["norm.weight"], sd["norm.bias"])
    h = h @ sd["decoder_pred.weight"].T + sd["decoder_pred.bias"]
    h = h[:, d["extras"]:, :]
    uidx = unpatchify_index(C, S, p)
    img = h.reshape(B, -1)[:, uidx].reshape(B, C, S, S)
    if "final_layer.weight" in sd:
        img = F.conv2d(img, sd["final_layer.weight"], sd["final_layer.bias"], padding=1)
    if tail_delta is not None:
        img = img + tail_delta.to(dt)
    return img


# ----------------------------------------------------------------------------------------------
# fixed-grid ODE (torchdiffeq FixedGridODESolver semantics, restated)
# ----------------------------------------------------------------------------------------------
def fixed_grid(t0: float, t1: float, step_size: float, dtype=torch.float32) -> Tensor:
    """grid = arange(ceil((t1-t0)/h + 1)) * h + t0, last point clamped to t1; decreasing time via s = -t."""
    sgn = 1.0 if t1 >= t0 else -1.0
    s0 = torch.tensor(sgn * t0, dtype=dtype)
    s1 = torch.tensor(sgn * t1, dtype=dtype)
    h = torch.tensor(step_size, dtype=dtype)
    niters = int(torch.ceil((s1 - s0) / h + 1).item())
    grid = torch.arange(0, niters, dtype=dtype) * h + s0
    grid[-1] = s1
    return grid * sgn


def odeint_fixed(func: Callable[[Tensor, Tensor], Tensor], z: Tensor, t0: float, t1: float, step_size: float,
                 method: str = "euler") -> Tensor:
    grid = fixed_grid(t0, t1, step_size, torch.float32)
    y = z
    for i in range(len(grid) - 1):
        ta, tb = grid[i], grid[i + 1]
        dt = (tb - ta).to(z.dtype)
        if method == "euler":
            y = y + dt * func(ta, y)
        elif method == "heun":
            k1 = func(ta, y)
            k2 = func(tb, y + dt * k1)
            y = y + dt * 0.5 * (k1 + k2)
        elif method == "midpoint":      # torchdiffeq Midpoint._step_func
            half = 0.5 * (tb - ta)
            y = y + dt * func(ta + half, y + func(ta, y) * half.to(z.dtype))
        elif method == "rk4":           # torchdiffeq rk4_alt_step_func (3/8 rule), fp32 stage times
            h32 = tb - ta
            third, two_thirds = torch.tensor(1.0 / 3.0, dtype=ta.dtype), torch.tensor(2.0 / 3.0, dtype=ta.dtype)
            k1 = func(ta, y)
            k2 = func(ta + h32 * third, y + dt * k1 / 3.0)
            k3 = func(ta + h32 * two_thirds, y + dt * (k2 - k1 / 3.0))
            k4 = func(tb, y + dt * (k1 - k2 + k3))
            y = y + (k1 + 3.0 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise NotImplementedError(method)
    return y


def sample(sd: Dict[str, Tensor], cfg: dict, z: Tensor, t0: float = 0.0, t1: float = 1.0, step_size: float = 0.02,
           method: str = "euler", y: Optional[Tensor] = None, context: Optional[Tensor] = None,
           delta_table: Optional[Tensor] = None, write_scale: float = 0.0, t_edit: float = 0.0,
           edit_loc: Optional[str] = None, attn_colscale: Optional[Tensor] = None, attn_blocks=None,
           attn_t_edit: float = 0.0) -> Tensor:
    """CNF.decode / CNF.encode (flow_matching.py:102-151) with the write_attr edit hook; delta_table rows are
    indexed by grid point (== the delta_{t:.2f}.npy file the reference would load at that time)."""
    grid = fixed_grid(t0, t1, step_size, torch.float32)
    lookup = {float(g): i for i, g in enumerate(grid.tolist())}

    def func(t: Tensor, x: Tensor) -> Tensor:
        hd = td = None
        # stages between grid points (midpoint, rk4) have no row in a grid-keyed table: no edit there
        if (edit_loc is not None and delta_table is not None and float(t) in lookup
                and should_edit(float(t), t_edit)):
            dlt = delta_table[lookup[float(t)]] * write_scale
            hd, td = (dlt, None) if edit_loc == "head" else (None, dlt)
        # attention edit: float(f"{t:.2f}") <= t_edit, "0.00" included (tools/utils_t2i.py:284)
        cs = attn_colscale if (attn_colscale is not None and float(f"{float(t):.2f}") <= attn_t_edit) else None
        return uvit_forward(sd, cfg, x, t.expand(x.shape[0]), y=y, context=context, head_delta=hd, tail_delta=td,
                            attn_colscale=cs, attn_blocks=attn_blocks)

    return odeint_fixed(func, z, t0, t1, step_size, method)


# ----------------------------------------------------------------------------------------------
# adaptive dopri5 (torchdiffeq RKAdaptiveStepsizeODESolver semantics, restated; PARITY UNPINNED: torchdiffeq is
# an un-vendored, unpinned dependency (README.md:121) that is installed nowhere here.  The constants below are
# checked mathematically in tests/test_oracle.py: Dormand-Prince A/B/C against scipy's RK45, order conditions of
# the embedded pair, accuracy of the mid-point weights.)
# Call sites: flow_matching.py:79-84 (default sampling), :50-57 ("adaptive"), :172-179 ("fixadp" tail).
# ----------------------------------------------------------------------------------------------
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
# Shampine's 4th-order companion, the pair torchdiffeq's "dopri5" uses
DP_C_ERROR = [35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
              -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
DP_C_MID = [6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
            187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]
DP_SAFETY, DP_IFACTOR, DP_DFACTOR = 0.9, 10.0, 0.2


def _rms(x: Tensor) -> float:
    """torchdiffeq's default norm: sqrt(mean(x^2)) over the WHOLE state tensor (the batch shares one step size)."""
    return float(x.double().pow(2).mean().sqrt())


# torchdiffeq's other adaptive tableaus: Bogacki-Shampine 3(2) ("bosh3") and Heun 2(1) ("adaptive_heun")
ADAPTIVE_TABLEAUS = {
    "dopri5": dict(order=5, alpha=DP_ALPHA, beta=DP_BETA, c_sol=DP_C_SOL, c_error=DP_C_ERROR, c_mid=DP_C_MID),
    "bosh3": dict(order=3, alpha=[1 / 2, 3 / 4, 1.0], beta=[[1 / 2], [0.0, 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
                  c_sol=[2 / 9, 1 / 3, 4 / 9, 0.0], c_error=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
                  c_mid=[0.0, 0.5, 0.0, 0.0]),
    "adaptive_heun": dict(order=2, alpha=[1.0], beta=[[1.0]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.0]),
}


def rk_adaptive_step(func, s0: float, y0: Tensor, f0: Tensor, dt: float, tab: dict):
    """One step of an embedded pair: returns (y1, f1, error estimate, stage derivatives).  As in torchdiffeq's
    _runge_kutta_step, f1 is ALWAYS the last stage's derivative - exact FSAL only when c_sol[:-1] == beta[-1]."""
    n = len(tab["alpha"])
    k = [f0]
    yi = y0
    for i in range(n):
        si = s0 + dt if tab["alpha"][i] == 1.0 else s0 + tab["alpha"][i] * dt
        yi = y0 + sum(k[j] * (tab["beta"][i][j] * dt) for j in range(i + 1))
        k.append(func(si, yi))
    if not (tab["c_sol"][-1] == 0 and list(tab["c_sol"][:-1]) == list(tab["beta"][-1])):
        yi = y0 + sum(k[j] * (tab["c_sol"][j] * dt) for j in range(n + 1))
    err = sum(k[j] * (tab["c_error"][j] * dt) for j in range(n + 1))
    return yi, k[n], err, k


def dopri5_step(func, s0: float, y0: Tensor, f0: Tensor, dt: float):
    """One Dormand-Prince step: returns (y1, f1, error estimate, stage derivatives k[0..6])."""
    return rk_adaptive_step(func, s0, y0, f0, dt, ADAPTIVE_TABLEAUS["dopri5"])


def dopri5_initial_step(func, s0: float, y0: Tensor, f0: Tensor, rtol: float, atol: float, order: int = 5) -> float:
    """Hairer's starting step as torchdiffeq's _select_initial_step computes it (its order argument = order - 1)."""
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(s0 + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / order)
    return min(100 * h0, h1)


def dopri5_interp(y0, y1, k, dt, x: float, c_mid=None) -> Tensor:
    """4th-order dense output on the accepted step, x = (t - t0) / (t1 - t0)."""
    c_mid = DP_C_MID if c_mid is None else c_mid
    y_mid = y0 + sum(k[j] * (c_mid[j] * dt) for j in range(len(c_mid)))
    f0, f1 = k[0], k[-1]
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    return y0 + x * d + x ** 2 * c + x ** 3 * b + x ** 4 * a


def odeint_dopri5(func: Callable[[float, Tensor], Tensor], z: Tensor, t0: float, t1: float, rtol: float = 1e-5,
                  atol: float = 1e-5, max_steps: int = 100000, stats: Optional[dict] = None,
                  method: str = "dopri5") -> Tensor:
    """odeint(func, z, [t0, t1], method=method, rtol, atol)[-1] for "dopri5" / "bosh3" / "adaptive_heun".  Steps are never clipped to t1: the last accepted
    step overshoots and the result is the dense-output polynomial evaluated at t1.  Decreasing time integrates
    g(s, y) = -f(-s, y) over s = -t."""
    sgn = 1.0 if t1 >= t0 else -1.0
    g = (lambda s, y: func(s, y)) if sgn > 0 else (lambda s, y: -func(-s, y))
    s0, s_end = sgn * t0, sgn * t1
    tab = ADAPTIVE_TABLEAUS[method]
    order, n_stage = tab["order"], len(tab["alpha"])
    y0 = z
    f0 = g(s0, y0)
    dt = dopri5_initial_step(g, s0, y0, f0, rtol, atol, order)
    n_acc = n_rej = 0
    nfe = 2
    while True:
        if n_acc + n_rej >= max_steps:
            raise RuntimeError("max_num_steps exceeded")
        y1, f1, err, k = rk_adaptive_step(g, s0, y0, f0, dt, tab)
        nfe += n_stage
        tol = atol + rtol * torch.maximum(y0.abs(), y1.abs())
        ratio = _rms(err / tol)
        accept = ratio <= 1.0
        if ratio == 0.0:
            factor = DP_IFACTOR
        else:
            factor = min(DP_IFACTOR, max(DP_SAFETY / ratio ** (1.0 / order), 1.0 if ratio < 1.0 else DP_DFACTOR))
        if accept:
            n_acc += 1
            if s0 + dt >= s_end:
                if stats is not None:
                    stats.update(n_accept=n_acc, n_reject=n_rej, nfe=nfe)
                return dopri5_interp(y0, y1, k, dt, (s_end - s0) / dt, tab["c_mid"])
            s0, y0, f0 = s0 + dt, y1, f1
        else:
            n_rej += 1
        dt = dt * factor


def digit_index(t: float) -> int:
    """Index of the delta_{t:.2f}.npy file / mask entry that belongs to time t."""
    return int(round(float(f"{t:.2f}") * 100))


def sample_adaptive(sd: Dict[str, Tensor], cfg: dict, z: Tensor, t0: float = 0.0, t1: float = 1.0, rtol: float = 1e-5,
                    atol: float = 1e-5, y: Optional[Tensor] = None, context: Optional[Tensor] = None,
                    delta_digits: Optional[Tensor] = None, write_scale: float = 0.0, t_edit: float = 0.0,
                    edit_loc: Optional[str] = None, attn_colscale: Optional[Tensor] = None, attn_blocks=None,
                    attn_t_edit: float = 0.0, stats: Optional[dict] = None, method: str = "dopri5",
                    read_trace: Optional[dict] = None) -> Tensor:
    """CNF.decode with an adaptive method.  delta_digits[i] is the row delta_{i/100:.2f}.npy would supply; the model
    sees the time rounded to fp32, like the reference's fp32 stage times."""
    def func(t: float, x: Tensor) -> Tensor:
        tf = float(torch.tensor(t, dtype=torch.float32))
        hd = td = None
        if edit_loc is not None and delta_digits is not None and should_edit(tf, t_edit):
            i = digit_index(tf)
            if 0 <= i < delta_digits.shape[0]:
                dlt = delta_digits[i] * write_scale
                hd, td = (dlt, None) if edit_loc == "head" else (None, dlt)
        cs = attn_colscale if (attn_colscale is not None and float(f"{tf:.2f}") <= attn_t_edit) else None
        tt = torch.full((x.shape[0],), tf, dtype=torch.float32)
        # dissect_name="read" (libs/dissection.py:126-136): np.save(f"{batch_id}_{t:.2f}", x) at every evaluation - the
        # file of a digit is overwritten by later evaluations at the same digit; read_trace["%.2f"] plays the file
        if read_trace is not None and edit_loc == "head":
            read_trace[f"{tf:.2f}"] = x.detach().clone()
        v = uvit_forward(sd, cfg, x, tt, y=y, context=context, head_delta=hd, tail_delta=td,
                         attn_colscale=cs, attn_blocks=attn_blocks)
        if read_trace is not None and edit_loc == "tail":
            read_trace[f"{tf:.2f}"] = v.detach().clone()
        return v

    return odeint_dopri5(func, z, t0, t1, rtol, atol, stats=stats, method=method)
