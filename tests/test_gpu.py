"""GPU tier (-m gpu, run on a real B200): the CUDA path, called through the C ABI, against the CPU oracle and the
golden vectors generated from the reference.  Tolerances (norm-wise relative error  ||a-b|| / ||b||):

  integer / index work (patch gather, unpatchify scatter, batch independence)      bit-exact
  kernels on identical 16-bit inputs (GEMM fp32 accumulate, LN, attention)        see per-test bounds
  whole forward / sampler, fp16 tensor-core operands (the default)                 <= 1.5e-3  (measured 4.6e-4..7.2e-4)
  whole forward, bf16 tensor-core operands                                         <= 1.2e-2  (measured 3.6e-3..5.7e-3)

BASELINE.json's north star asks for 1e-3 relative to the fp32 reference; fp16 operands with fp32 accumulation,
fp32 residual stream / LayerNorm / softmax statistics meet it on every measured case, bf16 operands do not
(8 mantissa bits), which is why fp16 is the default operand type (same tcgen05 kind::f16 rate).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import uvit_oracle as O
from tests.golden.cases import CASES, build_inputs, build_model
from uspace_b200 import _lib, parallel
from uspace_b200.flow_matching import CNF, CNFT2I
from uspace_b200.uvit import UViT, UViTT2I

pytestmark = pytest.mark.gpu
TD = {"fp16": torch.float16, "bf16": torch.bfloat16}
FWD_TOL = {"fp16": 1.5e-3, "bf16": 1.2e-2}


def dev():
    return torch.device("cuda:0")


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"))


@pytest.fixture(scope="module")
def lib():
    lib = _lib.load()
    assert torch.cuda.is_available()
    return lib


_models = {}


def model(name, opd="fp16"):
    key = (name, opd)
    if key not in _models:
        m = build_model(CASES[name], UViT, UViTT2I)
        m.operand_dtype = opd
        _models[key] = m.to(dev())
    return _models[key]


# ---- kernels through the ABI ---------------------------------------------------------------------------
@pytest.mark.parametrize("opd", ["fp16", "bf16"])
@pytest.mark.parametrize("epi,M,N,K,K0", [
    ("bias_f32", 128, 256, 64, 64), ("bias_f32", 300, 384, 128, 128), ("bias_f32", 514, 1024, 2048, 1024),
    ("qkv", 514, 3072, 1024, 1024), ("qkv", 771, 768, 256, 256), ("bias_gelu", 514, 4096, 1024, 1024),
    # full-size problems (>= 148 tiles): the wide ones on CTA pairs, N = 1024 on clusters of 4 with multicast A loads
    ("qkv", 16448, 3072, 1024, 1024), ("bias_gelu", 10688, 4096, 1024, 1024), ("bias_resid", 16448, 1024, 4096, 4096),
    ("bias_f32", 10688, 1024, 2048, 1024),
    ("bias_resid", 514, 1024, 4096, 4096), ("bias_resid", 16448, 1024, 1024, 1024),
    ("bias_f32", 16448, 1024, 2048, 1024), ("qkv", 10688, 3072, 1024, 1024),
    # BASELINE configs[2] (t2i-L, batch 128): M = 128 * 334 = 42752
    ("bias_gelu", 42752, 4096, 1024, 1024), ("bias_resid", 42752, 1024, 4096, 4096), ("qkv", 42752, 3072, 1024, 1024)])
def test_gemm_epilogues(lib, opd, epi, M, N, K, K0):
    td, L = TD[opd], (334 if M % 334 == 0 else 257)
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev()).to(td)
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev()).to(td)
    bias = torch.randn(N, generator=g).to(dev())
    resid = torch.randn(M, N, generator=g).to(dev())
    ref = a.float() @ w.float().T  # same rounded operands: only accumulation order differs
    a0 = a[:, :K0].contiguous()
    a1 = a[:, K0:].contiguous() if K0 < K else None
    out32 = torch.zeros(M, N, device=dev())
    out16 = torch.zeros(M, N, device=dev(), dtype=td)
    eps16 = 2.0 ** -11 if opd == "fp16" else 2.0 ** -8
    e = _lib.EPI[epi]
    if epi == "qkv":
        H = N // 3 // 64
        _lib.check(lib.usp_op_gemm(e, P(a0), P(a1), P(w), None, None, None, P(out16), M, N, K, K0, L, H,
                                   _lib.OPERAND[opd], stream()))
        torch.cuda.synchronize()
        B = M // L
        got = out16.view(3, B, H, L, 64).float()
        want = ref.view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
        assert rel(got, want) < eps16
    elif epi == "bias_gelu":
        _lib.check(lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), None, None, P(out16), M, N, K, K0, L, 1,
                                   _lib.OPERAND[opd], stream()))
        torch.cuda.synchronize()
        assert rel(out16.float(), torch.nn.functional.gelu(ref + bias)) < eps16
    elif epi == "bias_resid":
        _lib.check(lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), P(resid), P(out32), P(out16), M, N, K, K0, L, 1,
                                   _lib.OPERAND[opd], stream()))
        torch.cuda.synchronize()
        want = ref + bias + resid
        assert rel(out32, want) < 1e-5
        assert rel(out16.float(), want) < eps16
    else:
        _lib.check(lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), None, P(out32), None, M, N, K, K0, L, 1,
                                   _lib.OPERAND[opd], stream()))
        torch.cuda.synchronize()
        assert rel(out32, ref + bias) < 1e-5


@pytest.mark.parametrize("opd", ["fp16", "bf16"])
@pytest.mark.parametrize("B,H,L", [(1, 1, 17), (1, 1, 128), (1, 2, 256), (2, 8, 257), (3, 4, 258), (2, 16, 334),
                                   (1, 1, 384), (64, 16, 257), (128, 16, 334), (3, 2, 288), (2, 2, 320), (2, 3, 352),
                                   (5, 16, 129), (2, 4, 64), (1, 3, 1),
                                   # beyond one 384-key pass: key passes with a running maximum (attention.cu);
                                   # 1025 = a 512^2 image's 1024 patches + the time token
                                   (2, 3, 385), (1, 2, 400), (2, 4, 768), (1, 16, 1025), (1, 1, 1153), (1, 2, 2049)])
def test_attention(lib, opd, B, H, L):
    td = TD[opd]
    g = torch.Generator().manual_seed(B * 1000 + L)
    q, k, v = (torch.randn(B * H, L, 64, generator=g).to(dev()).to(td) for _ in range(3))
    out = torch.zeros(B * L, H * 64, device=dev(), dtype=td)
    _lib.check(lib.usp_op_attention(P(q), P(k), P(v), P(out), B, H, L, _lib.OPERAND[opd], stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    want = ref.view(B, H, L, 64).permute(0, 2, 1, 3).reshape(B * L, H * 64)
    # P and O are rounded to 16 bits once each
    assert rel(out.float(), want) < (6e-4 if opd == "fp16" else 5e-3)


@pytest.mark.parametrize("B,H,L,scale", [(2, 4, 257, 6.0), (3, 16, 334, 5.0), (64, 16, 257, 4.0), (2, 2, 200, 8.0),
                                         (2, 2, 1025, 6.0), (1, 3, 800, 8.0)])
def test_attention_large_scores_running_max(lib, B, H, L, scale):
    """Scores whose block maxima differ by far more than 2^8: the lazy running-maximum path of attention3.cu (O rescaled
    in TMEM between key blocks) against fp32 SDPA (libs/uvit.py:95)."""
    g = torch.Generator().manual_seed(7 * B + L)
    q = (scale * torch.randn(B * H, L, 64, generator=g)).to(dev()).half()
    k, v = (torch.randn(B * H, L, 64, generator=g).to(dev()).half() for _ in range(2))
    out = torch.zeros(B * L, H * 64, device=dev(), dtype=torch.float16)
    _lib.check(lib.usp_op_attention(P(q), P(k), P(v), P(out), B, H, L, _lib.OPERAND["fp16"], stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    want = ref.view(B, H, L, 64).permute(0, 2, 1, 3).reshape(B * L, H * 64)
    assert torch.isfinite(out).all()
    assert rel(out.float(), want) < 6e-4


def test_attention_rejects_sequences_beyond_its_limit(lib):
    q = torch.zeros(1, 16385, 64, device=dev(), dtype=torch.float16)
    assert lib.usp_op_attention(P(q), P(q), P(q), P(q), 1, 1, 16385, 1, stream()) != 0


@pytest.mark.parametrize("M,D", [(7, 256), (514, 512), (16448, 1024)])
def test_layernorm(lib, M, D):
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, D, generator=g) * 3 + 1).to(dev())
    gam, bet = torch.randn(D, generator=g).to(dev()), torch.randn(D, generator=g).to(dev())
    out = torch.zeros(M, D, device=dev(), dtype=torch.float16)
    _lib.check(lib.usp_op_layernorm(P(x), P(gam), P(bet), P(out), M, D, 1, stream()))
    torch.cuda.synchronize()
    want = O.layer_norm(x.cpu().double(), gam.cpu().double(), bet.cpu().double())
    assert rel(out.float(), want) < 2.0 ** -11


def test_patch_gather_and_unpatchify_scatter_bit_exact(lib):
    """Integer-valued data through identity weights: the index maps must reproduce the oracle's bit for bit."""
    B, Cc, S, p, D = 3, 4, 32, 2, 128
    n_patch, Pd = (S // p) ** 2, Cc * p * p
    x = torch.arange(B * Cc * S * S, dtype=torch.float32).reshape(B, Cc, S, S) % 4099.0
    w = torch.zeros(D, Pd)
    w[:Pd, :Pd] = torch.eye(Pd)
    out = torch.full((B, 1 + n_patch, D), -1.0, device=dev())
    t = torch.zeros(B, device=dev())
    xd, wd = x.to(dev()), w.to(dev())   # keep every device buffer alive across the asynchronous launch
    bd, pd = torch.zeros(D, device=dev()), torch.zeros(1 + n_patch, D, device=dev())
    _lib.check(lib.usp_op_patch_embed(P(xd), P(t), P(wd), P(bd), P(pd), P(out), B, Cc, S, p, D, stream()))
    torch.cuda.synchronize()
    want = x.reshape(B, -1)[:, O.patchify_index(Cc, S, p).reshape(-1)].reshape(B, n_patch, Pd)
    assert torch.equal(out[:, 1:, :Pd].cpu(), want)
    assert torch.equal(out[:, 1:, Pd:].cpu(), torch.zeros(B, n_patch, D - Pd))
    # time token at t=0: cos(0)=1 for the first half, sin(0)=0 for the second
    assert torch.equal(out[:, 0, :D // 2].cpu(), torch.ones(B, D // 2))
    assert torch.equal(out[:, 0, D // 2:].cpu(), torch.zeros(B, D // 2))

    pf = (torch.arange(B * n_patch * Pd, dtype=torch.float32) % 8191.0).reshape(B, n_patch, Pd)
    img = torch.zeros(B, Cc, S, S, device=dev())
    pfd = pf.to(dev())
    _lib.check(lib.usp_op_unpatchify_conv(P(pfd), None, None, P(img), B, Cc, S, p, stream()))
    torch.cuda.synchronize()
    want = pf.reshape(B, -1)[:, O.unpatchify_index(Cc, S, p)].reshape(B, Cc, S, S)
    assert torch.equal(img.cpu(), want)
    cw = torch.randn(Cc, Cc, 3, 3)
    cb = torch.randn(Cc)
    cwd, cbd = cw.to(dev()), cb.to(dev())
    _lib.check(lib.usp_op_unpatchify_conv(P(pfd), P(cwd), P(cbd), P(img), B, Cc, S, p, stream()))
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(want.double(), cw.double(), cb.double(), padding=1)
    assert rel(img, want) < 1e-6


# ---- whole forward against the reference goldens ---------------------------------------------------------
@pytest.mark.parametrize("opd", ["fp16", "bf16"])
@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(golden_dir, name, opd):
    case = CASES[name]
    m = model(name, opd)
    x, t, y, ctx = build_inputs(case)
    with torch.no_grad():
        if case["t2i"]:
            out, aux = m(x.to(dev()), t.to(dev()), context=ctx.to(dev()), dissect_name=None, extra_kwarg=1)
        else:
            out, aux = m(x.to(dev()), t.to(dev()), y if y is None else y.to(dev()), edit_loc=None)
    assert aux is None and out.shape == x.shape and out.dtype == torch.float32
    want = golden(golden_dir, name)["forward"]
    assert rel(out, want) < FWD_TOL[opd]


@pytest.mark.parametrize("name", ["small16_uncond", "tiny_class", "tiny_t2i", "large_uncond"])
def test_forward_with_stand_alone_layernorm_kernels(golden_dir, name):
    """usp_config.fuse_layernorm = 0: norm1 / norm2 as LayerNorm launches of their own instead of the default fold
    into the qkv / fc1 GEMMs (gamma in the weights, rstd and the folded bias in the epilogue)."""
    case = CASES[name]
    m = build_model(case, UViT, UViTT2I)
    assert m.fuse_layernorm is True
    m.fuse_layernorm = False
    m = m.to(dev())
    x, t, y, ctx = build_inputs(case)
    with torch.no_grad():
        if case["t2i"]:
            out = m(x.to(dev()), t.to(dev()), context=ctx.to(dev()))[0]
            ref = model(name)(x.to(dev()), t.to(dev()), context=ctx.to(dev()))[0]
        else:
            out = m(x.to(dev()), t.to(dev()), y if y is None else y.to(dev()))[0]
            ref = model(name)(x.to(dev()), t.to(dev()), y if y is None else y.to(dev()))[0]
    assert m.engine().kernels_per_forward() > model(name).engine().kernels_per_forward()  # the LayerNorm launches
    want = golden(golden_dir, name)["forward"]
    assert rel(out, want) < 1e-3 and rel(ref, want) < 1e-3
    assert not torch.equal(out, ref)                       # two different computations of the same function


@pytest.mark.parametrize("embed_dim,heads,fold", [(384, 6, False), (768, 12, True), (1536, 24, True)])
def test_forward_other_widths_against_oracle(embed_dim, heads, fold):
    """Widths no reference config uses but the constructor accepts: 768 / 1536 take the folded-LayerNorm pair-kernel
    path, 384 (not a multiple of 256) falls back to stand-alone LayerNorm and the 128-wide GEMM tiles."""
    cfg = dict(CASES["tiny_class"]["cfg"], embed_dim=embed_dim, num_heads=heads, depth=2)
    torch.manual_seed(21)
    m = UViT(**cfg).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(22)
    x, t = torch.randn(3, 4, 32, 32, generator=g), torch.rand(3, generator=g)
    y = torch.randint(0, 10, (3,), generator=g)
    want = O.uvit_forward(sd, cfg, x, t, y=y)
    m = m.to(dev())
    with torch.no_grad():
        got = m(x.to(dev()), t.to(dev()), y.to(dev()))[0]
    assert rel(got, want) < 1e-3
    plain = UViT(**cfg).eval()
    plain.fuse_layernorm = False
    plain.load_state_dict(sd)
    plain = plain.to(dev())
    with torch.no_grad():
        plain(x.to(dev()), t.to(dev()), y.to(dev()))
    # the fold removes the LayerNorm launches only where every GEMM shape is served by the pair kernel
    assert (m.engine().kernels_per_forward() < plain.engine().kernels_per_forward()) == fold


def test_forward_meets_1e3_on_north_star_model(golden_dir):
    for name in ("large_uncond", "large_t2i", "small16_uncond"):
        case = CASES[name]
        m = model(name)
        x, t, y, ctx = build_inputs(case)
        with torch.no_grad():
            out = m(x.to(dev()), t.to(dev()), context=ctx.to(dev()))[0] if case["t2i"] else m(x.to(dev()), t.to(dev()))[0]
        assert rel(out, golden(golden_dir, name)["forward"]) < 1e-3


def test_batch_independence_at_full_size_bit_exact(golden_dir):
    """BASELINE config 2 size (B=64, M=16448): a sample's velocity does not depend on its batch neighbours."""
    m = model("large_uncond")
    z = parallel.global_noise(64).to(dev())
    t = torch.full((64,), 0.5, device=dev())
    with torch.no_grad():
        big = m(z, t)[0]
        small = m(z[61:64].clone(), t[:3])[0]
    assert torch.equal(big[61:64], small)
    # and the first sample of that batch matches the oracle
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    want = O.uvit_forward(sd, CASES["large_uncond"]["cfg"], z[:1].cpu(), t[:1].cpu())
    assert rel(big[:1], want) < 1e-3


def test_load_state_dict_repacks_weights():
    case = CASES["tiny_uncond"]
    m = build_model(case, UViT, UViTT2I).to(dev())
    x, t, _, _ = build_inputs(case)
    with torch.no_grad():
        a = m(x.to(dev()), t.to(dev()))[0]
        torch.manual_seed(123)
        other = UViT(**case["cfg"]).state_dict()
        m.load_state_dict(other)
        b = m(x.to(dev()), t.to(dev()))[0]
    want = O.uvit_forward(other, case["cfg"], x, t)
    assert rel(b, want) < 1.5e-3 and rel(a, want) > 0.1


# ---- sampler -------------------------------------------------------------------------------------------
FIXED = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=0.2))


@pytest.mark.parametrize("name", ["tiny_uncond", "tiny_t2i", "tiny_long"])
def test_decode_matches_reference_driven_euler_golden(golden_dir, name):
    case = CASES[name]
    m = model(name)
    x, _, y, ctx = build_inputs(case)
    kw = dict(FIXED, solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=1.0 / case["euler_steps"]))
    if case["t2i"]:
        out = CNFT2I(m).decode(x.to(dev()), context=ctx.to(dev()), **kw)
    else:
        out = CNF(m).decode(x.to(dev()), y=None, **kw)
    assert rel(out, golden(golden_dir, name)["euler"]) < 1e-3


@pytest.mark.parametrize("method,h", [("euler", 0.1), ("heun", 0.25)])
def test_sampler_against_oracle(method, h):
    case = CASES["tiny_class"]
    m = model("tiny_class")
    x, _, y, _ = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, method, y=y)
    kw = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix=method, solver_fix_step=h))
    got = CNF(m).decode(x.to(dev()), y=y.to(dev()), **kw)
    assert rel(got, want) < 1e-3
    # encode runs the same loop on the reversed grid (flow_matching.py:102-125)
    want = O.sample(sd, case["cfg"], x.double(), 1.0, 0.0, h, method, y=y)
    got = CNF(m).encode(x.to(dev()), y=y.to(dev()), **kw)
    assert rel(got, want) < 1e-3


@pytest.mark.parametrize("name", ["tiny_uncond", "small16_uncond"])
@pytest.mark.parametrize("loc", ["head", "tail"])
def test_edit_hook_against_reference_golden(golden_dir, tmp_path, name, loc):
    """One Euler step starting at the edit time: (z1 - z0)/h is the edited velocity the reference produced."""
    case = CASES[name]
    e = case["edit"]
    g = golden(golden_dir, name)
    m = model(name)
    x, _, _, _ = build_inputs(case)
    h = 0.1
    np.save(tmp_path / f"delta_{e['t']:.2f}.npy", g["edit_delta"])
    np.save(tmp_path / f"delta_{e['t'] + h:.2f}.npy", g["edit_delta"])
    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path),
              ith_attr=e["ith_attr"], t_edit=e["t_edit"], write_scale=e["write_scale"], edit_loc=loc,
              solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=h))
    eng = m.engine()
    from uspace_b200.flow_matching import build_delta_table
    from uspace_b200.engine import time_grid
    grid = time_grid(e["t"], e["t"] + h, h)
    assert len(grid) == 2
    tab, tloc = build_delta_table(grid, (4, 32, 32), **kw)
    z1 = eng.sample(x.to(dev()), e["t"], e["t"] + h, h, "euler", delta_table=tab, write_scale=e["write_scale"],
                    t_edit=float("inf"), edit_loc=tloc)
    v = (z1.cpu().double() - x.double()) / (np.float32(grid[1]) - np.float32(grid[0]))
    assert rel(v, g[f"edit_{loc}"]) < 2e-3
    # without the edit the velocity is measurably different
    z0 = eng.sample(x.to(dev()), e["t"], e["t"] + h, h, "euler")
    assert rel((z0.cpu().double() - x.double()) / h, g[f"edit_{loc}"]) > 1e-2


def test_edit_sweep_via_cnf_decode_against_oracle(tmp_path):
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x, _, _, _ = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    h = 0.1
    rng = torch.Generator().manual_seed(3)
    table = torch.zeros(11, 4, 32, 32)
    for i, t in enumerate(O.fixed_grid(0.0, 1.0, h).tolist()):
        if O.should_edit(t, 0.4):
            d = 0.1 * torch.randn(2, 4, 32, 32, generator=rng)
            np.save(tmp_path / f"delta_{t:.2f}.npy", d.numpy())
            table[i] = d[1]
    for scale in (-2.1, 0.0, 1.5):
        kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path), ith_attr=1,
                  t_edit=0.4, write_scale=scale, edit_loc="tail",
                  solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=h))
        got = CNF(m).decode(x.to(dev()), y=None, **kw)
        want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, "euler", delta_table=table.double(),
                        write_scale=scale, t_edit=0.4, edit_loc="tail")
        assert rel(got, want) < 1e-3


def test_library_mask_follows_should_edit():
    """usp_sample's own t_edit mask (used when the caller passes a dense table) == libs/dissection.py:21-26."""
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x, _, _, _ = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    table = 0.1 * torch.randn(6, 4, 32, 32, generator=g)  # dense: every grid point has a row
    got = m.engine().sample(x.to(dev()), 0.0, 1.0, 0.2, "euler", delta_table=table, write_scale=2.0, t_edit=0.4,
                            edit_loc="head")
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, 0.2, "euler", delta_table=table.double(), write_scale=2.0,
                    t_edit=0.4, edit_loc="head")
    assert rel(got, want) < 1e-3


@pytest.mark.parametrize("loc", ["head", "tail"])
def test_module_call_applies_the_edit_hook_like_the_reference(golden_dir, tmp_path, loc):
    """nnet(x, t, y, **config.dissection) itself (libs/uvit.py:313-314,349-350 -> dissect_helper_uvit): the velocity
    the reference module returned with edit_loc head / tail, write_attr."""
    case = CASES["tiny_uncond"]
    e = case["edit"]
    g = golden(golden_dir, "tiny_uncond")
    m = model("tiny_uncond")
    x, _, _, _ = build_inputs(case)
    np.save(tmp_path / f"delta_{e['t']:.2f}.npy", g["edit_delta"])
    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path),
              ith_attr=e["ith_attr"], t_edit=e["t_edit"], write_scale=e["write_scale"], edit_loc=loc, batch_id=0)
    t = torch.full((x.shape[0],), e["t"], device=dev())
    with torch.no_grad():
        got = m(x.to(dev()), t, None, **kw)[0]
        plain = m(x.to(dev()), t, None, edit_loc=None)[0]
        # t > t_edit: should_edit() is false, no file is touched (libs/dissection.py:21-34)
        late = m(x.to(dev()), torch.full_like(t, 0.9), None, **kw)[0]
        late_plain = m(x.to(dev()), torch.full_like(t, 0.9), None, edit_loc=None)[0]
    assert rel(got, g[f"edit_{loc}"]) < 2e-3
    assert rel(plain, g[f"edit_{loc}"]) > 1e-2
    assert torch.equal(late, late_plain)
    # "read": the hook dumps the activation at edit_loc as {batch_id}_{t:.2f}.npy (libs/dissection.py:126-136)
    rd = dict(kw, dissect_name="read", read_path_root=str(tmp_path / "read"))
    with torch.no_grad():
        out = m(x.to(dev()), t, None, **rd)[0]
    dumped = np.load(tmp_path / "read" / f"0_{e['t']:.2f}.npy")
    assert torch.equal(out, plain)
    want = x.numpy() if loc == "head" else plain.cpu().numpy()
    assert np.array_equal(dumped, want)
    with torch.no_grad():
        with pytest.raises(ValueError):
            m(x.to(dev()), t, None, **dict(kw, dissect_name="bogus"))
        with pytest.raises(NotImplementedError):
            m(x.to(dev()), t, None, **dict(kw, edit_loc="mid"))


# ---- the BASELINE configurations as whole trajectories against the oracle (fp32 FAST mode: the same torch CPU calls
#      the reference makes; about a minute of host time each) -------------------------------------------------------
def _oracle_fast(fn):
    O.FAST = True
    try:
        return fn()
    finally:
        O.FAST = False


@pytest.mark.parametrize("method,nfe", [("euler", 50), ("heun", 100)])
def test_uvit_large_50_step_trajectory_against_oracle(method, nfe):
    """BASELINE configs[1] (50 Euler steps) and the Heun variant of configs[3], U-ViT-L, two images: the final latents
    against the CPU oracle driving the same 50-interval grid (flow_matching.py:130-151)."""
    case = CASES["large_uncond"]
    m = model("large_uncond")
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    z = parallel.global_noise(2)
    want = _oracle_fast(lambda: O.sample(sd, case["cfg"], z, 0.0, 1.0, 0.02, method))
    got = m.engine().sample(z.to(dev()), 0.0, 1.0, 0.02, method)
    assert len(O.fixed_grid(0.0, 1.0, 0.02)) == 51
    assert rel(got, want) < 1e-3


def test_t2i_large_trajectory_against_oracle():
    """BASELINE configs[2] model (t2i U-ViT-L, 77 context tokens, L = 334): 10 Euler steps, two images."""
    case = CASES["large_t2i"]
    m = model("large_t2i")
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    z = parallel.global_noise(2)
    ctx = torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(1231))
    want = _oracle_fast(lambda: O.sample(sd, case["cfg"], z, 0.0, 1.0, 0.1, "euler", context=ctx))
    got = m.engine().sample(z.to(dev()), 0.0, 1.0, 0.1, "euler", context=ctx.to(dev()))
    assert rel(got, want) < 1e-3


def test_fp16_operand_range_scaled_weights():
    """fp16 tensor-core operands saturate at 65504.  A checkpoint whose MLP hidden activations exceed that must either
    stay within tolerance or fail loudly - never return silently wrong numbers.  The library's contract: GELU outputs
    beyond the fp16 range make the forward non-finite (inf -> NaN through fc2 / LayerNorm), which the caller can test
    with one isfinite(); with bf16 operands (8 exponent bits) the same weights stay finite and within the bf16 bound."""
    case = CASES["tiny_uncond"]
    x, t, _, _ = build_inputs(case)
    torch.manual_seed(case["seed"])
    sd = UViT(**case["cfg"]).state_dict()
    # 1) large but representable: |fc1 output| ~ 2e3 (100 x the random-init scale) - still exact to tolerance
    sd_ok = {k: (v * 100.0 if k.endswith("mlp.fc1.weight") else v * (0.01 if k.endswith("mlp.fc2.weight") else 1.0))
             for k, v in sd.items()}
    m = UViT(**case["cfg"]).eval()
    m.load_state_dict(sd_ok)
    m = m.to(dev())
    with torch.no_grad():
        got = m(x.to(dev()), t.to(dev()))[0]
    want = O.uvit_forward(sd_ok, case["cfg"], x, t)
    hidden_max = max((O.layer_norm(torch.randn(4, 256), sd_ok["in_blocks.0.norm2.weight"], sd_ok["in_blocks.0.norm2.bias"])
                      @ sd_ok["in_blocks.0.mlp.fc1.weight"].T).abs().max().item(), 0.0)
    assert hidden_max > 50.0
    assert torch.isfinite(got).all() and rel(got, want) < 1.5e-3
    # 2) beyond the fp16 range: fc1 pre-activations ~ N(0, (2.6e4)^2), i.e. GELU outputs far above 65504 in every block
    #    (fc2 is scaled back so that the fp32 oracle stays finite; its fp16 copy underflows, which no longer matters:
    #    the run must be flagged.  A 4000x scale - used by an earlier version of this test - does NOT overflow (hidden
    #    maximum ~6e3) and only measures the fp16-subnormal rounding of the shrunken fc2 weights, 1.8e-3.)
    sd_big = {k: (v * 8.0e4 if k.endswith("mlp.fc1.weight") else v * (1.25e-5 if k.endswith("mlp.fc2.weight") else 1.0))
              for k, v in sd.items()}
    want = O.uvit_forward(sd_big, case["cfg"], x, t)
    assert torch.isfinite(want).all()
    m16 = UViT(**case["cfg"]).eval()
    m16.load_state_dict(sd_big)
    m16 = m16.to(dev())
    with torch.no_grad():
        got16 = m16(x.to(dev()), t.to(dev()))[0]
    overflowed = m16.engine().nonfinite()          # the library's own sticky flag (usp_nonfinite)
    assert overflowed == (not torch.isfinite(got16).all().item())
    assert overflowed                                 # loud, never quietly wrong
    if overflowed:
        kw = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=0.5))
        with pytest.raises(FloatingPointError, match="fp16"):
            CNF(m16).decode(x.to(dev()), y=None, **kw)
        zh = x.clone().pin_memory()
        with pytest.raises(RuntimeError, match="inf / NaN"):
            m16.engine().sample_host(zh, 0.0, 1.0, 0.5, "euler")
    assert not m.engine().nonfinite()
    mb = UViT(**case["cfg"]).eval()
    mb.operand_dtype = "bf16"
    mb.load_state_dict(sd_big)
    mb = mb.to(dev())
    with torch.no_grad():
        gotb = mb(x.to(dev()), t.to(dev()))[0]
    assert torch.isfinite(gotb).all() and rel(gotb, want) < 1.2e-2


def test_folded_layernorm_flags_tokens_far_from_zero_mean():
    """The default fold feeds the GEMMs fp16(x) instead of fp16(LN(x)): harmless while a token's mean is of the order of
    its standard deviation (every golden), lossy when the whole token is shifted.  A residual stream pushed to
    mean = 40 std must raise the library's flag (CNF.decode turns it into an exception); the stand-alone LayerNorm
    kernels (fuse_layernorm = False) stay within tolerance on the same weights and raise nothing."""
    case = CASES["tiny_uncond"]
    sd = {k: v.clone() for k, v in build_model(case, UViT, UViTT2I).state_dict().items()}
    sd["pos_embed"] = sd["pos_embed"] + 40.0
    x, t, _, _ = build_inputs(case)
    want = O.uvit_forward(sd, case["cfg"], x, t)
    fold = UViT(**case["cfg"]).eval()
    fold.load_state_dict(sd)
    fold = fold.to(dev())
    with torch.no_grad():
        got_fold = fold(x.to(dev()), t.to(dev()))[0]
    assert fold.engine().status_flags() & 2
    kw = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=0.5))
    with pytest.raises(FloatingPointError, match="fuse_layernorm"):
        CNF(fold).decode(x.to(dev()), y=None, **kw)
    plain = UViT(**case["cfg"]).eval()
    plain.fuse_layernorm = False
    plain.load_state_dict(sd)
    plain = plain.to(dev())
    with torch.no_grad():
        got_plain = plain(x.to(dev()), t.to(dev()))[0]
    assert plain.engine().status_flags() == 0
    assert rel(got_plain, want) < 1e-3
    assert rel(got_fold, want) > rel(got_plain, want)        # the loss the flag announces is real
    assert model("tiny_uncond").engine().status_flags() == 0  # and ordinary weights never raise it


def test_encode_decode_round_trip_full_size_model():
    """Size-independent property at the north-star model: decode(encode(x)) returns to x up to O(h) Euler error,
    and the error shrinks with the step (flow_matching.py vis_reversible idea, dissect_lfm.py:171-195)."""
    m = model("large_uncond")
    eng = m.engine()
    x = parallel.global_noise(4, seed=5).to(dev())
    errs = []
    for h in (0.1, 0.05):
        z = eng.sample(x, 1.0, 0.0, h, "heun")
        back = eng.sample(z, 0.0, 1.0, h, "heun")
        errs.append(rel(back, x))
    assert errs[0] < 5e-2 and errs[1] < errs[0]


def test_sample_host_equals_device_path_bit_exact():
    m = model("tiny_uncond")
    eng = m.engine()
    z = parallel.global_noise(5, seed=9)
    a = eng.sample(z.to(dev()), 0.0, 1.0, 0.25, "heun").cpu()
    zh = z.clone().pin_memory()
    eng.sample_host(zh, 0.0, 1.0, 0.25, "heun")
    assert torch.equal(a, zh)
    assert eng.last_ms() > 0 and eng.kernels_per_forward() > 10


def test_error_paths():
    m = model("tiny_uncond")
    eng = m.engine()
    z = torch.zeros(1, 4, 32, 32, device=dev())
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 3, 32, 32, device=dev()), torch.zeros(1, device=dev()))
    with pytest.raises(RuntimeError, match="y must be given exactly"):
        eng.forward(z, torch.zeros(1, device=dev()), y=torch.zeros(1, dtype=torch.long, device=dev()))
    with pytest.raises(RuntimeError, match="bad time grid"):
        eng.sample(z, 0.0, 0.0, 0.1)
    with pytest.raises(RuntimeError, match="delta_table"):
        eng.sample(z, 0.0, 1.0, 0.5, edit_loc="tail")
    with pytest.raises(KeyError):
        eng.load_state_dict({})
    lib, h = eng.lib, eng.handle
    st = _lib.UspAdaptiveStats()
    # C ABI, invalid arguments: status < 0 and a message, nothing launched
    assert lib.usp_sample_adaptive(h, P(z), None, None, 1, 0.0, 1.0, 4, 0.0, 1e-5, None, 0, 0.0, 0.0, 0, None, 0,
                                   C.byref(st), stream()) < 0
    assert b"rtol" in lib.usp_last_error(h)
    assert lib.usp_sample_adaptive(h, P(z), None, None, 1, 0.0, 1.0, 4, 1e-5, 1e-5, P(z), 200, 1.0, 0.4, 2, None, 0,
                                   C.byref(st), stream()) < 0
    assert b"n_rows" in lib.usp_last_error(h)
    assert lib.usp_sample_adaptive(h, P(z), None, None, 1, 0.0, 1.0, 1, 1e-5, 1e-5, None, 0, 0.0, 0.0, 0, None, 0,
                                   C.byref(st), stream()) < 0
    assert b"adaptive method" in lib.usp_last_error(h)
    sc = (C.c_float * 2)(1.0, 2.0)
    assert lib.usp_sample_sweep(h, P(z), P(z), None, None, 1, sc, 2, 0.0, 1.0, 0.5, 0, None, 0.4, 0, stream()) < 0
    assert b"edit_loc" in lib.usp_last_error(h)
    assert lib.usp_sample_sweep(h, P(z), P(z), None, None, 1, None, 0, 0.0, 1.0, 0.5, 0, P(z), 0.4, 2, stream()) < 0
    assert lib.usp_sample_read(h, P(z), None, None, 1, 0.0, 1.0, 0.5, 0, 2, None, stream()) < 0
    assert lib.usp_sample_read(h, P(z), None, None, 1, 0.0, 1.0, 0.5, 0, 0, P(z), stream()) < 0
    assert b"read mode" in lib.usp_last_error(h)
    assert lib.usp_sample_edit(h, P(z), None, None, 1, 0.0, 1.0, 0.5, 7, None, 0.0, 0.0, 0, None, stream()) < 0
    assert b"unknown method" in lib.usp_last_error(h)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sampling_matches_single_gpu():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(root, "tests", "mp_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MP_CHECK_OK" in r.stdout


def _p2p_kwargs(case):
    pp = case["p2p"]
    return dict(dissect_name="p2p", fm_direction="decode", t_edit=pp["t_edit"], block_id=pp["block_id"],
                token_kwargs=dict(token_dissect="p2p_rescale", p2p_multiplier=pp["multiplier"]),
                target_context_ids=[np.array(i) for i in pp["ids"]])


def test_p2p_attention_rescale_forward_matches_reference_golden(golden_dir):
    """dissect_lfm_t2i.py "p2p" mode through the module call, against the reference's own editing branch."""
    case = CASES["tiny_t2i"]
    pp = case["p2p"]
    g = golden(golden_dir, "tiny_t2i")
    m = model("tiny_t2i")
    x, _, _, ctx = build_inputs(case)
    kw = _p2p_kwargs(case)
    with torch.no_grad():
        t_in = torch.full((x.shape[0],), pp["t"], device=dev())
        out = m(x.to(dev()), t_in, context=ctx.to(dev()), **kw)[0]
        out_all = m(x.to(dev()), t_in, context=ctx.to(dev()), **dict(kw, block_id="all"))[0]
        t_late = torch.full((x.shape[0],), 0.9, device=dev())
        late_edit = m(x.to(dev()), t_late, context=ctx.to(dev()), **kw)[0]
        late_plain = m(x.to(dev()), t_late, context=ctx.to(dev()))[0]
    assert rel(out, g["p2p_forward"]) < 1.5e-3
    assert rel(out_all, g["p2p_all_blocks"]) < 1.5e-3
    assert rel(out, g["p2p_plain"]) > 1e-3
    assert torch.equal(late_edit, late_plain)   # t > t_edit: the hook is inactive (tools/utils_t2i.py:284)


@pytest.mark.parametrize("method", ["euler", "heun"])
def test_p2p_attention_rescale_sampling_against_oracle(method):
    case = CASES["tiny_t2i"]
    pp = case["p2p"]
    m = model("tiny_t2i")
    x, _, _, ctx = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    # a strong edit (every context token, every block) so that the edited trajectory is far from the plain one
    ids, mult = [list(range(77)), list(range(77))], [30.0, -20.0]
    cs = torch.ones(case["B"], 334)
    for i, tid in enumerate(ids):
        cs[i, torch.tensor(tid) + 1] = mult[i]
    h = 0.2
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, method, context=ctx.double(), attn_colscale=cs.double(),
                    attn_blocks=None, attn_t_edit=pp["t_edit"])
    kw = dict(dissect_name="p2p", t_edit=pp["t_edit"], block_id="all",
              token_kwargs=dict(token_dissect="p2p_rescale", p2p_multiplier=mult),
              target_context_ids=[np.array(i) for i in ids],
              solver_kwargs=dict(solver="fixed", solver_fix=method, solver_fix_step=h))
    got = CNFT2I(m).decode(x.to(dev()), context=ctx.to(dev()), **kw)
    assert rel(got, want) < 1e-3
    plain = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, method, context=ctx.double())
    assert rel(want, plain) > 5e-2
    assert rel(got.double().cpu() - plain, want - plain) < 2e-2   # the edit itself, not just the trajectory
    # encode never edits the attention map (fm_direction == "encode", tools/utils_t2i.py:276-277)
    enc = CNFT2I(m).encode(x.to(dev()), context=ctx.to(dev()), **kw)
    enc_want = O.sample(sd, case["cfg"], x.double(), 1.0, 0.0, h, method, context=ctx.double())
    assert rel(enc, enc_want) < 1e-3


# ---- adaptive dopri5 (csrc/ode.cu) -----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_uncond", "tiny_class", "tiny_t2i"])
def test_adaptive_dopri5_against_oracle(name):
    case = CASES[name]
    m = model(name)
    x, _, y, ctx = build_inputs(case)
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    so, sg = {}, {}
    want = O.sample_adaptive(sd, case["cfg"], x.double(), 0.0, 1.0, 1e-5, 1e-5, y=y,
                             context=None if ctx is None else ctx.double(), stats=so)
    got = m.engine().sample_adaptive(x.to(dev()), 0.0, 1.0, 1e-5, 1e-5, y=y, context=ctx, stats=sg)
    assert rel(got, want) < 1e-3
    assert sg["nfe"] == 2 + 6 * (sg["n_accept"] + sg["n_reject"])
    # 16-bit operand noise in the velocity reads as extra local error: the device run may take more steps
    assert so["n_accept"] <= sg["n_accept"] <= 4 * so["n_accept"] + 4


def test_adaptive_reversed_time_and_mid_interval():
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0]
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    for t0, t1 in ((1.0, 0.0), (0.4, 1.0)):
        want = O.sample_adaptive(sd, case["cfg"], x.double(), t0, t1, 1e-5, 1e-5)
        got = m.engine().sample_adaptive(x.to(dev()), t0, t1, 1e-5, 1e-5)
        assert rel(got, want) < 1e-3
    # round trip through the two directions
    z = m.engine().sample_adaptive(x.to(dev()), 1.0, 0.0, 1e-5, 1e-5)
    back = m.engine().sample_adaptive(z, 0.0, 1.0, 1e-5, 1e-5)
    assert rel(back, x) < 1e-3


def test_adaptive_is_deterministic_and_tighter_tolerance_takes_more_steps():
    m = model("tiny_uncond")
    x = build_inputs(CASES["tiny_uncond"])[0].to(dev())
    a, b, sa, sb = None, None, {}, {}
    a = m.engine().sample_adaptive(x, 0.0, 1.0, 1e-5, 1e-5, stats=sa)
    b = m.engine().sample_adaptive(x, 0.0, 1.0, 1e-5, 1e-5, stats=sb)
    assert torch.equal(a, b) and sa == sb
    st = {}
    c = m.engine().sample_adaptive(x, 0.0, 1.0, 1e-3, 1e-3, stats=st)
    assert st["n_accept"] <= sa["n_accept"] and rel(c, a) < 1e-2
    with pytest.raises(RuntimeError, match="max_steps"):
        m.engine().sample_adaptive(x, 0.0, 1.0, 1e-5, 1e-5, max_steps=1)
    with pytest.raises(RuntimeError, match="t0 != t1"):
        m.engine().sample_adaptive(x, 0.5, 0.5)


def test_cnf_default_adaptive_and_fixadp_against_oracle(tmp_path):
    """flow_matching.py:79-84 (non-dissection default = dopri5) and :153-180 (decode_fixadp with the tail edit)."""
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0]
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    cnf = CNF(m)
    got = cnf.decode(x.to(dev()), y=None, solver_kwargs=dict(solver="adaptive"))
    want = O.sample_adaptive(sd, case["cfg"], x.double(), 0.0, 1.0, 1e-5, 1e-5)
    assert rel(got, want) < 1e-3 and cnf.last_solver_stats["n_accept"] >= 2
    # fixadp: Euler on [0, t_edit] with the edit at every grid point, dopri5 on [t_edit, 1] where the hook still
    # fires for evaluations that print as "0.40"
    # (a large edit: the discontinuity it puts between f(0.40) and the later stages forces rejected steps)
    x = x[:1]
    h, t_mid, scale = 0.1, 0.4, 40.0
    rng = torch.Generator().manual_seed(5)
    grid_table = torch.zeros(5, 4, 32, 32)
    digits = torch.zeros(101, 4, 32, 32)
    for i in range(1, 5):
        d = 0.3 * torch.randn(2, 4, 32, 32, generator=rng)
        np.save(tmp_path / f"delta_{i * h:.2f}.npy", d.numpy())
        grid_table[i] = d[1]
        digits[10 * i] = d[1]
    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path), ith_attr=1,
              t_edit=t_mid, write_scale=scale, edit_loc="tail",
              solver_kwargs=dict(solver="fixadp", solver_fix="euler", solver_fix_step=h, solver_adaptive="dopri5"))
    got = cnf.decode(x.to(dev()), y=None, **kw)
    mid = O.sample(sd, case["cfg"], x.double(), 0.0, t_mid, h, "euler", delta_table=grid_table.double(),
                   write_scale=scale, t_edit=t_mid, edit_loc="tail")
    want = O.sample_adaptive(sd, case["cfg"], mid, t_mid, 1.0, 1e-5, 1e-5, delta_digits=digits.double(),
                             write_scale=scale, t_edit=t_mid, edit_loc="tail")
    assert rel(got, want) < 1e-3
    assert cnf.last_solver_stats["n_reject"] >= 3
    plain = O.sample_adaptive(sd, case["cfg"], mid, t_mid, 1.0, 1e-5, 1e-5)
    assert rel(want, plain) > 1e-2            # the "0.40" evaluations of the adaptive phase were edited
    assert rel(got.double().cpu() - plain, want - plain) < 3e-2


@pytest.mark.parametrize("loc", ["head", "tail"])
def test_cnf_read_under_adaptive_solver_against_oracle(tmp_path, loc):
    """dissect_name="read" with solver="adaptive" (libs/dissection.py:126-136 runs under any solver): one
    {batch_id}_{t:.2f}.npy per evaluation digit, the last evaluation at a digit wins.  The oracle replays dopri5 with
    the same hook.  The two controllers need not take the same steps (at rtol = 1e-5 the embedded error estimate of
    the 16-bit path carries its rounding noise), so the digit sets are compared where they overlap: the rows of common
    digits must agree up to the solution's own tolerance plus the drift of x over the difference of the two evaluation
    times inside one digit."""
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0][:2]
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    cnf = CNF(m)
    root = tmp_path / "dump"
    kw = dict(dissect_task="uspace_uvit", dissect_name="read", read_path_root=str(root), batch_id=3, edit_loc=loc,
              t_edit=1.0, write_scale=0.0, solver_kwargs=dict(solver="adaptive", solver_adaptive="dopri5"))
    got = cnf.decode(x.to(dev()), y=None, **kw)
    trace = {}
    want = O.sample_adaptive(sd, case["cfg"], x.double(), 0.0, 1.0, 1e-5, 1e-5, edit_loc=loc, read_trace=trace)
    assert rel(got, want) < 1e-3
    files = sorted(os.listdir(root))
    assert all(f.startswith("3_") and f.endswith(".npy") for f in files)
    digits_got = [f[len("3_"):-len(".npy")] for f in files]
    assert "0.00" in digits_got and 5 <= len(digits_got) <= cnf.last_solver_stats["nfe"]
    common = sorted(set(digits_got) & set(trace))
    assert "0.00" in common and len(common) >= 3, (digits_got, sorted(trace))
    for d in digits_got:
        assert np.load(root / f"3_{d}.npy").shape == (2, 4, 32, 32)
    # (torchdiffeq does not clip the last step to t1: evaluation times beyond 1.00 are dumped like any other.)
    # Inside a digit the two evaluation times differ by up to 0.01
    for d in common:
        row = torch.from_numpy(np.load(root / f"3_{d}.npy"))
        assert rel(row, trace[d]) < 5e-2, (d, rel(row, trace[d]))
    if loc == "head":       # no later evaluation prints "0.00" here: the file holds what the net saw near t0
        assert rel(torch.from_numpy(np.load(root / "3_0.00.npy")), trace["0.00"]) < 1e-4


def test_adaptive_agrees_with_fine_fixed_grid_on_north_star_model():
    """Size-independent property on U-ViT-L: dopri5(1e-5) and Heun(h = 0.01) integrate the same field."""
    m = model("large_uncond")
    g = torch.Generator().manual_seed(1230)
    x = torch.randn(8, 4, 32, 32, generator=g).to(dev())
    st = {}
    ada = m.engine().sample_adaptive(x, 0.0, 1.0, 1e-5, 1e-5, stats=st)
    fine = m.engine().sample(x, 0.0, 1.0, 0.01, "heun")
    assert rel(ada, fine) < 1e-3 and st["n_accept"] >= 2


# ---- scale sweep in one batch, "read" mode ------------------------------------------------------------------
@pytest.mark.parametrize("name,method,loc", [("tiny_uncond", "euler", "tail"), ("tiny_class", "heun", "head"),
                                             ("tiny_t2i", "euler", "tail")])
def test_scale_sweep_rows_equal_separate_runs_bit_exact(name, method, loc):
    case = CASES[name]
    m = model(name)
    x, _, y, ctx = build_inputs(case)
    eng = m.engine()
    h = 0.2
    g = torch.Generator().manual_seed(21)
    table = 0.2 * torch.randn(6, 4, 32, 32, generator=g)
    scales = [-2.1, 0.0, 0.5, 2.0]
    out = eng.sample_sweep(x.to(dev()), scales, 0.0, 1.0, h, method, y=y, context=ctx, delta_table=table, t_edit=0.4,
                           edit_loc=loc)
    assert out.shape == (case["B"], len(scales), 4, 32, 32)
    for s, scale in enumerate(scales):
        one = eng.sample(x.to(dev()), 0.0, 1.0, h, method, y=y, context=ctx, delta_table=table, write_scale=scale,
                         t_edit=0.4, edit_loc=loc)
        assert torch.equal(out[:, s], one), (name, scale)
    assert not torch.equal(out[:, 0], out[:, 1])


def test_sweep_driver_mirror_against_oracle(tmp_path):
    """tools/utils_vis.py:189-201 through uspace_b200.sweep.sample_write_scales: "(b s)" order, oracle per scale."""
    from uspace_b200.sweep import sample_write_scales
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0]
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    h = 0.1
    rng = torch.Generator().manual_seed(3)
    table = torch.zeros(11, 4, 32, 32)
    for i in range(1, 5):
        d = 0.1 * torch.randn(3, 4, 32, 32, generator=rng)
        np.save(tmp_path / f"delta_{i * h:.2f}.npy", d.numpy())
        table[i] = (d[0] + d[2]) / 2
    scales = [-2.0, 0.0, 1.0]
    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path), ith_attr="0_2",
              t_edit=0.4, edit_loc="tail", solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=h))
    got = sample_write_scales(CNF(m), x.to(dev()), scales, None, **kw)
    assert got.shape == (case["B"] * 3, 4, 32, 32)
    got = got.reshape(case["B"], 3, 4, 32, 32)
    for s, scale in enumerate(scales):
        want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, "euler", delta_table=table.double(),
                        write_scale=scale, t_edit=0.4, edit_loc="tail")
        assert rel(got[:, s], want) < 1e-3
    # an adaptive solver falls back to the reference's loop (its step size depends on the batch)
    kw["solver_kwargs"] = dict(solver="adaptive", solver_adaptive="dopri5")
    loop = sample_write_scales(CNF(m), x.to(dev()), scales[:2], None, **kw)
    assert loop.shape == (case["B"] * 2, 4, 32, 32) and torch.isfinite(loop).all()


def test_read_mode_trace_and_files(tmp_path):
    """dissect_name="read" (libs/dissection.py:126-136): the activation at edit_loc of every evaluation."""
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0]
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    eng = m.engine()
    h = 0.25
    z_h, states = eng.sample_read(x.to(dev()), 1.0, 0.0, h, "euler", edit_loc="head")
    z_t, vels = eng.sample_read(x.to(dev()), 1.0, 0.0, h, "euler", edit_loc="tail")
    plain = eng.sample(x.to(dev()), 1.0, 0.0, h, "euler")
    assert torch.equal(z_h, plain) and torch.equal(z_t, plain)     # reading does not disturb the trajectory
    assert states.shape == (5, 3, 4, 32, 32)
    assert torch.equal(states[0].cpu(), x) and not states[4].any() and not vels[4].any()
    grid = O.fixed_grid(1.0, 0.0, h)
    for i in range(4):
        v = O.uvit_forward(sd, case["cfg"], states[i].cpu().double(), grid[i].expand(3))
        assert rel(vels[i], v) < 1.5e-3
        nxt = states[i + 1] if i < 3 else plain
        assert (states[i] + (grid[i + 1] - grid[i]).item() * vels[i] - nxt).abs().max() < 1e-5
    # through the CNF mirror: the files the reference would have written during encode (dissect_lfm.py:216-228)
    kw = dict(dissect_task="uspace_uvit", dissect_name="read", read_path_root=str(tmp_path / "dump"), batch_id=7,
              edit_loc="tail", solver_kwargs=dict(solver="fixed", solver_fix="euler", solver_fix_step=h))
    enc = CNF(m).encode(x.to(dev()), y=None, **kw)
    assert torch.equal(enc, plain)
    assert sorted(os.listdir(tmp_path / "dump")) == ["7_0.25.npy", "7_0.50.npy", "7_0.75.npy", "7_1.00.npy"]
    assert np.array_equal(np.load(tmp_path / "dump" / "7_0.75.npy"), vels[1].cpu().numpy())
    # Heun also evaluates at the last grid point
    _, tr = eng.sample_read(x.to(dev()), 0.0, 1.0, 0.5, "heun", edit_loc="head")
    assert tr.shape[0] == 3 and tr[2].any()


def test_read_delta_write_closed_loop(tmp_path):
    """The reference's three-stage workflow on this library: (1) "read" the tail activation while encoding labelled
    latents (dissect_lfm.py:216-228), (2) delta_t = mean(attr = 1) - mean(attr = 0) (tools/utils_attr.py:160-206),
    (3) decode with "write_attr" reading those files (libs/dissection.py:138-157) - checked against the oracle."""
    from uspace_b200 import attr_delta
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    cnf = CNF(m)
    root = str(tmp_path / "feat")
    h = 0.25
    sk = dict(solver="fixed", solver_fix="euler", solver_fix_step=h)
    g = torch.Generator().manual_seed(77)
    attrs = (torch.rand(6, 11, generator=g) < 0.5).long()
    attrs[0], attrs[1] = 1, 0
    lat = []
    for b in range(2):                                   # two batches of three "real" latents
        x = torch.randn(3, 4, 32, 32, generator=g)
        z = cnf.encode(x.to(dev()), y=None, dissect_task="uspace_uvit", dissect_name="read", read_path_root=root,
                       batch_id=b, edit_loc="tail", solver_kwargs=sk)
        lat.append(z.cpu())
    attr_delta.save_latents(root, torch.cat(lat).numpy(), attrs.numpy())
    times = attr_delta.extract_deltas_by_attr(root, batch_num=2)
    assert times == ["0.25", "0.50", "0.75", "1.00"]
    delta = np.load(os.path.join(root, "delta_0.25.npy"))
    assert delta.shape == (11, 4, 32, 32) and np.abs(delta).max() > 0
    # decode fresh noise with attribute 3 pushed: grid 0, .25, .5, .75 -> edits at .25 and .5 (t_edit = 0.5)
    x = torch.randn(2, 4, 32, 32, generator=g)
    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=root, ith_attr=3, t_edit=0.5,
              write_scale=4.0, edit_loc="tail", solver_kwargs=sk)
    got = cnf.decode(x.to(dev()), y=None, **kw)
    table = torch.zeros(5, 4, 32, 32)
    table[1] = torch.from_numpy(np.load(os.path.join(root, "delta_0.25.npy"))[3])
    table[2] = torch.from_numpy(np.load(os.path.join(root, "delta_0.50.npy"))[3])
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, "euler", delta_table=table.double(), write_scale=4.0,
                    t_edit=0.5, edit_loc="tail")
    assert rel(got, want) < 1e-3
    plain = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, "euler")
    assert rel(want, plain) > 1e-3


# ---- torchdiffeq's other fixed-grid methods ----------------------------------------------------------------
@pytest.mark.parametrize("name,method,h", [("tiny_uncond", "midpoint", 0.25), ("tiny_uncond", "rk4", 0.25),
                                           ("tiny_class", "rk4", 0.5), ("tiny_t2i", "midpoint", 0.2)])
def test_midpoint_and_rk4_against_oracle(name, method, h):
    case = CASES[name]
    m = model(name)
    x, _, y, ctx = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, method, y=y, context=None if ctx is None else ctx.double())
    got = m.engine().sample(x.to(dev()), 0.0, 1.0, h, method, y=y, context=ctx)
    assert rel(got, want) < 1e-3
    euler = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, h, "euler", y=y, context=None if ctx is None else ctx.double())
    assert rel(got.double().cpu() - euler, want - euler) < 5e-2     # the higher-order correction itself
    # reversed time through the CNF mirror
    kw = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix=method, solver_fix_step=h))
    if ctx is None:
        enc = CNF(m).encode(x.to(dev()), y=y, **kw)
    else:
        enc = CNFT2I(m).encode(x.to(dev()), context=ctx.to(dev()), **kw)
    enc_want = O.sample(sd, case["cfg"], x.double(), 1.0, 0.0, h, method, y=y,
                        context=None if ctx is None else ctx.double())
    assert rel(enc, enc_want) < 1e-3


def test_rk4_with_grid_point_edit_and_attention_rule():
    """Edits under rk4: the write edit fires at the stages that sit on grid points (k1 at t_i, k4 at t_{i+1}); the
    attention edit follows its "%.2f" rule at every stage, in-between ones included."""
    case = CASES["tiny_t2i"]
    m = model("tiny_t2i")
    x, _, _, ctx = build_inputs(case)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    table = 0.3 * torch.randn(3, 4, 32, 32, generator=g)      # grid 0, 0.5, 1.0
    cs = torch.ones(case["B"], 334)
    cs[:, 1:78] = 25.0
    got = m.engine().sample(x.to(dev()), 0.0, 1.0, 0.5, "rk4", context=ctx, delta_table=table, write_scale=2.0,
                            t_edit=0.5, edit_loc="tail",
                            attn_edit=dict(colscale=cs, block_mask=(1 << 64) - 1, t_edit=0.2))
    want = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, 0.5, "rk4", context=ctx.double(),
                    delta_table=table.double(), write_scale=2.0, t_edit=0.5, edit_loc="tail",
                    attn_colscale=cs.double(), attn_blocks=None, attn_t_edit=0.2)
    assert rel(got, want) < 1e-3
    # attention edit active at t = 0 and t = 1/6 ("0.17" <= 0.2) only: dropping the in-between rule must be visible
    no_mid = O.sample(sd, case["cfg"], x.double(), 0.0, 1.0, 0.5, "rk4", context=ctx.double(),
                      delta_table=table.double(), write_scale=2.0, t_edit=0.5, edit_loc="tail",
                      attn_colscale=cs.double(), attn_blocks=None, attn_t_edit=0.1)
    assert rel(want, no_mid) > 5e-3
    with pytest.raises(RuntimeError, match="read mode"):
        m.engine().sample_read(x.to(dev()), 0.0, 1.0, 0.5, "rk4", context=ctx, edit_loc="tail")


def test_cnf_rejects_write_edits_under_midpoint_and_rk4(tmp_path):
    """The reference's hook keys on f"{t:.2f}" at every evaluation (libs/dissection.py:120), in-between stages of
    midpoint / rk4 included; the CNF mirror's table is keyed by grid point, so it refuses rather than diverge."""
    m = model("tiny_uncond")
    x, _, _, _ = build_inputs(CASES["tiny_uncond"])
    for t in ("0.25", "0.50"):
        np.save(tmp_path / f"delta_{t}.npy", np.zeros((2, 4, 32, 32), np.float32))
    for method in ("midpoint", "rk4"):
        kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=str(tmp_path), ith_attr=1,
                  t_edit=0.5, write_scale=1.0, edit_loc="tail",
                  solver_kwargs=dict(solver="fixed", solver_fix=method, solver_fix_step=0.25))
        with pytest.raises(NotImplementedError, match="write edits"):
            CNF(m).decode(x.to(dev()), y=None, **kw)
    # torchdiffeq's name for the two-stage Heun
    kw = dict(dissect_name="none", solver_kwargs=dict(solver="fixed", solver_fix="heun2", solver_fix_step=0.25))
    a = CNF(m).decode(x.to(dev()), y=None, **kw)
    b = m.engine().sample(x.to(dev()), 0.0, 1.0, 0.25, "heun")
    assert torch.equal(a, b)


def test_scale_sweep_full_size_model_bit_exact():
    """U-ViT-L, 8 latents x 9 write_scales (configs/lfm_cm256_uvit_large.py:84) in one batch of 72: every row equals
    the separate batch-8 run bit for bit - across batch sizes AND across GEMM variants (the batch of 72 runs proj / fc2 /
    skip_linear on clusters of 4 with multicast loads, the batch of 8 on CTA pairs)."""
    m = model("large_uncond")
    eng = m.engine()
    g = torch.Generator().manual_seed(1230)
    x = torch.randn(8, 4, 32, 32, generator=g).to(dev())
    table = 0.1 * torch.randn(5, 4, 32, 32, generator=torch.Generator().manual_seed(1232))
    scales = [-2.1, -1.5, -1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
    out = eng.sample_sweep(x, scales, 0.0, 1.0, 0.25, "euler", delta_table=table, t_edit=0.5, edit_loc="tail")
    assert out.shape == (8, 9, 4, 32, 32) and torch.isfinite(out).all()
    for s in (0, 4, 8):
        one = eng.sample(x, 0.0, 1.0, 0.25, "euler", delta_table=table, write_scale=scales[s], t_edit=0.5,
                         edit_loc="tail")
        assert torch.equal(out[:, s], one)
    assert torch.equal(out[:, 4], eng.sample(x, 0.0, 1.0, 0.25, "euler"))     # scale 0 == no edit


@pytest.mark.parametrize("method", ["bosh3", "adaptive_heun"])
def test_other_adaptive_methods_against_oracle(method):
    case = CASES["tiny_uncond"]
    m = model("tiny_uncond")
    x = build_inputs(case)[0][:1]
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    tol = 1e-5      # the two runs may accept / reject differently; both stay within a few tol of the true solution
    so, sg = {}, {}
    want = O.sample_adaptive(sd, case["cfg"], x.double(), 0.0, 1.0, tol, tol, stats=so, method=method)
    got = m.engine().sample_adaptive(x.to(dev()), 0.0, 1.0, tol, tol, stats=sg, method=method)
    assert rel(got, want) < 1e-3
    n_stage = {"bosh3": 3, "adaptive_heun": 1}[method]
    assert sg["nfe"] == 2 + n_stage * (sg["n_accept"] + sg["n_reject"])
    assert so["n_accept"] <= sg["n_accept"] <= 4 * so["n_accept"] + 4
    # through the CNF mirror, reversed direction
    kw = dict(dissect_name="none", solver_kwargs=dict(solver="adaptive", solver_adaptive=method))
    dec = CNF(m).decode(x.to(dev()), y=None, **kw)
    assert rel(dec, want) < 1e-3


# ---- latent -> image decoder (csrc/vae.cu) -----------------------------------------------------------------------
VAE_TOL = 2e-4        # default "fp16x3" operands (hi + lo parts, three products per GEMM); measured 2.7e-5 - 5.7e-5
VAE_TOL_FP16 = 5e-3   # precision="fp16": 11-bit operands through 37 convolutions + GroupNorms (~2e-3, like TF32)

_vae = {}


def vae_model_gpu(precision="fp16x3"):
    if precision not in _vae:
        from tests.golden.cases import vae_enc_state_dict, vae_state_dict
        from uspace_b200.autoencoder import get_model
        m = get_model(precision=precision)
        m.load_state_dict({**vae_state_dict(), **vae_enc_state_dict()})
        _vae[precision] = m.to(dev())
    return _vae[precision]


def test_vae_fp16_operand_mode_against_reference_golden(golden_dir):
    """precision="fp16": the single-product path (3x less tensor work) stays within its looser bound and is a
    different computation from the default."""
    from tests.golden.cases import vae_images, vae_latents
    m = vae_model_gpu("fp16")
    z = vae_latents("vae_small").to(dev())
    got = m.decode(z)
    assert rel(got, golden(golden_dir, "vae_small")["decode"]) < VAE_TOL_FP16
    assert not torch.equal(got, vae_model_gpu().decode(z))
    x = vae_images("vae_enc_small").to(dev())
    assert rel(m.encode_moments(x), golden(golden_dir, "vae_enc_small")["moments"]) < VAE_TOL_FP16
    with pytest.raises(ValueError):
        from uspace_b200.autoencoder import get_model
        get_model(precision="fp8")
    # C ABI: an unknown mode is rejected, and switching the mode un-finalises the handle (weights are packed per mode)
    eng = m.engine()
    assert eng.lib.usp_vae_set_precision(eng.handle, 7) != 0
    assert b"precision" in eng.lib.usp_vae_last_error(eng.handle)
    assert eng.lib.usp_vae_set_precision(eng.handle, 1) == 0
    out = torch.empty(1, 3, 128, 128, device=dev())
    assert eng.lib.usp_vae_decode(eng.handle, P(z[:1].contiguous()), P(out), 1, 16, stream()) != 0      # not finalised
    assert eng.lib.usp_vae_finalize(eng.handle, stream()) == 0
    assert eng.lib.usp_vae_decode(eng.handle, P(z[:1].contiguous()), P(out), 1, 16, stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, vae_model_gpu().decode(z[:1]))
    _vae.pop("fp16")                                             # that handle now runs the split mode


@pytest.mark.parametrize("name", ["vae_small", "vae_full"])
def test_vae_decode_matches_reference_golden(golden_dir, name):
    from tests.golden.cases import vae_latents
    m = vae_model_gpu()
    z = vae_latents(name)
    got = m.decode(z.to(dev()))
    want = golden(golden_dir, name)["decode"]
    assert tuple(got.shape) == want.shape
    assert rel(got, want) < VAE_TOL
    assert torch.equal(m(z.to(dev()), "decode"), got)         # forward(inputs, fn) dispatch, deterministic


def test_vae_decode_batch_independence_and_chunking():
    """Chunks of 16 inside the library, GroupNorm statistics per sample: every image equals its own batch-1 decode."""
    m = vae_model_gpu()
    g = torch.Generator().manual_seed(5)
    z = (0.7 * torch.randn(19, 4, 16, 16, generator=g)).to(dev())
    out = m.decode(z)
    assert out.shape == (19, 3, 128, 128) and torch.isfinite(out).all()
    for i in (0, 15, 16, 18):
        assert torch.equal(out[i:i + 1], m.decode(z[i:i + 1]))
    from oracle import vae_oracle as V
    from tests.golden.cases import vae_state_dict
    want = V.decode({k: v.double() for k, v in vae_state_dict().items()}, z[16:18].cpu().double())
    assert rel(out[16:18], want) < VAE_TOL
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1, 4, 10, 10, device=dev()))


@pytest.mark.parametrize("name", ["vae_enc_small", "vae_enc_full"])
def test_vae_encode_moments_matches_reference_golden(golden_dir, name):
    from tests.golden.cases import vae_images
    m = vae_model_gpu()
    x = vae_images(name)
    got = m.encode_moments(x.to(dev()))
    want = golden(golden_dir, name)["moments"]
    assert tuple(got.shape) == want.shape and rel(got, want) < VAE_TOL
    assert torch.equal(m(x.to(dev()), "encode_moments"), got)
    # sample() (libs/autoencoder.py:431-437): scale * (mean + std * eps) with std = exp(clamp(logvar) / 2)
    torch.manual_seed(3)
    z = m.encode(x.to(dev()))
    torch.manual_seed(3)
    eps = torch.randn_like(got[:, :4])
    assert torch.allclose(z, 0.18215 * (got[:, :4] + torch.exp(0.5 * got[:, 4:].clamp(-30, 20)) * eps), atol=1e-6)


def test_vae_round_trip_and_chunked_decode():
    """Image -> moments -> decode on random weights is no identity, but every stage must be finite, batch-independent
    and chunk-independent (decode_large_batch, dissect_lfm.py:86-98)."""
    from uspace_b200.autoencoder import decode_large_batch
    m = vae_model_gpu()
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(18, 3, 128, 128, generator=g) * 2 - 1).to(dev())
    mo = m.encode_moments(x)
    assert mo.shape == (18, 8, 16, 16) and torch.isfinite(mo).all()
    assert torch.equal(mo[16:17], m.encode_moments(x[16:17]))               # second library chunk == batch of one
    z = 0.18215 * mo[:, :4]
    img = m.decode(z)
    assert torch.equal(decode_large_batch(m, z, chunk=3), img)
    # the workspace is kept between calls, can be handed back, and comes back on demand
    eng = m.engine()
    assert eng.workspace_bytes() > 0
    eng.release_workspace()
    assert eng.workspace_bytes() == 0
    assert torch.equal(m.decode(z[:2]), img[:2]) and eng.workspace_bytes() > 0
    with pytest.raises(ValueError):
        m.encode_moments(torch.zeros(1, 3, 100, 100, device=dev()))


def test_vae_other_resolutions_against_oracle():
    """The decoder / encoder are resolution-agnostic up to the GEMM shape rules (latent side 16, 32, 48, 64)."""
    from oracle import vae_oracle as V
    from tests.golden.cases import vae_enc_state_dict, vae_state_dict
    m = vae_model_gpu()
    sd = {k: v.double() for k, v in {**vae_state_dict(), **vae_enc_state_dict()}.items()}
    g = torch.Generator().manual_seed(17)
    z = 0.7 * torch.randn(1, 4, 48, 48, generator=g)
    assert rel(m.decode(z.to(dev())), V.decode(sd, z.double())) < VAE_TOL
    x = torch.rand(1, 3, 384, 384, generator=g) * 2 - 1
    assert rel(m.encode_moments(x.to(dev())), V.encode_moments(sd, x.double())) < VAE_TOL
