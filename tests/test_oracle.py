"""CPU tier: the oracle restatement against the golden vectors generated from the reference modules."""
import os

import einops
import numpy as np
import pytest
import torch

from oracle import uvit_oracle as O
from tests.golden.cases import CASES, build_inputs, build_model
from uspace_b200.uvit import UViT, UViTT2I


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"))


FAST = [n for n in CASES if not n.startswith("large")]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_forward_matches_reference_golden(golden_dir, name):
    case = CASES[name]
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, t, y, ctx = build_inputs(case)
    got = O.uvit_forward(sd, case["cfg"], x, t, y=y, context=ctx)
    want = load(golden_dir, name)["forward"]
    assert got.shape == want.shape
    assert rel(got, want) < 2e-6  # same fp32 arithmetic, different op order only


@pytest.mark.parametrize("name", [n for n in FAST if CASES[n].get("euler_steps")])
def test_oracle_euler_matches_reference_driven_loop(golden_dir, name):
    case = CASES[name]
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, t, y, ctx = build_inputs(case)
    got = O.sample(sd, case["cfg"], x, 0.0, 1.0, 1.0 / case["euler_steps"], "euler", y=y, context=ctx)
    assert rel(got, load(golden_dir, name)["euler"]) < 5e-6


@pytest.mark.parametrize("name", [n for n in FAST if CASES[n].get("edit")])
@pytest.mark.parametrize("loc", ["head", "tail"])
def test_oracle_edit_hook_matches_reference_dissect_helper(golden_dir, name, loc):
    case = CASES[name]
    e = case["edit"]
    g = load(golden_dir, name)
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, _, y, _ = build_inputs(case)
    delta = torch.from_numpy(g["edit_delta"])
    ith = e["ith_attr"]
    row = delta[ith] if isinstance(ith, int) else sum(delta[int(i)] for i in ith.split("_")) / len(ith.split("_"))
    assert O.should_edit(e["t"], e["t_edit"])
    scaled = row * e["write_scale"]
    t = torch.full((x.shape[0],), e["t"])
    got = O.uvit_forward(sd, case["cfg"], x, t, y=y, head_delta=scaled if loc == "head" else None,
                         tail_delta=scaled if loc == "tail" else None)
    assert rel(got, g[f"edit_{loc}"]) < 2e-6


def test_index_maps_bit_exact_against_einops():
    C, S, p = 4, 32, 2
    img = torch.arange(C * S * S, dtype=torch.int64).reshape(1, C, S, S)
    # PatchEmbed == Conv2d(k=s=p): feature order of the flattened conv weight is (C, p1, p2)  (libs/uvit.py:171-178)
    want = einops.rearrange(img, "B C (h p1) (w p2) -> B (h w) (C p1 p2)", p1=p, p2=p)[0]
    got = img.reshape(-1)[O.patchify_index(C, S, p)]
    assert torch.equal(got, want)
    # unpatchify uses (p1, p2, C)  (libs/uvit.py:56-63)
    tok = torch.arange((S // p) ** 2 * p * p * C, dtype=torch.int64).reshape(1, (S // p) ** 2, p * p * C)
    want = einops.rearrange(tok, "B (h w) (p1 p2 C) -> B C (h p1) (w p2)", h=S // p, p1=p, p2=p)[0]
    got = tok.reshape(-1)[O.unpatchify_index(C, S, p)].reshape(C, S, S)
    assert torch.equal(got, want)


def test_patch_embed_equals_conv2d():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(4, 64, 2, 2)
    x = torch.randn(2, 4, 32, 32)
    want = conv(x).flatten(2).transpose(1, 2)
    feats = x.reshape(2, -1)[:, O.patchify_index(4, 32, 2).reshape(-1)].reshape(2, 256, 16)
    got = feats @ conv.weight.reshape(64, 16).T + conv.bias
    assert (got - want).abs().max() < 1e-5


@pytest.mark.parametrize("t0,t1,h,n", [(0.0, 1.0, 0.02, 51), (0.0, 1.0, 0.01, 101), (0.0, 0.4, 0.01, 41),
                                        (1.0, 0.0, 0.02, 51), (0.0, 1.0, 0.3, 5)])
def test_fixed_grid_lengths(t0, t1, h, n):
    g = O.fixed_grid(t0, t1, h)
    assert len(g) == n  # 51 / 101 / 41: SURVEY.md §8c probe of torchdiffeq's grid constructor
    assert g[0].item() == pytest.approx(t0) and g[-1].item() == np.float32(t1)
    d = (g[1:] - g[:-1]) * (1 if t1 > t0 else -1)
    assert (d > 0).all()


def test_should_edit_semantics():
    assert not O.should_edit(0.0, 0.4)          # "0.00" never edits (libs/dissection.py:22-23)
    assert not O.should_edit(0.004, 0.4)        # rounds to "0.00"
    assert O.should_edit(0.02, 0.4) and O.should_edit(0.4, 0.4)
    assert not O.should_edit(0.42, 0.4)
    assert O.should_edit(0.404, 0.4)            # compared after rounding to 2 digits
    assert O.should_edit(0.5, "every_0.25") and not O.should_edit(0.3, "every_0.25")


def test_heun_is_second_order_on_a_linear_field():
    # dx/dt = a*x  ->  x(1) = exp(a); Euler error O(h), Heun O(h^2)
    a = 0.7
    z = torch.ones(1, dtype=torch.float64)
    f = lambda t, x: a * x
    exact = np.exp(a)
    e1 = abs(O.odeint_fixed(f, z, 0.0, 1.0, 0.1, "euler").item() - exact)
    e2 = abs(O.odeint_fixed(f, z, 0.0, 1.0, 0.1, "heun").item() - exact)
    e2h = abs(O.odeint_fixed(f, z, 0.0, 1.0, 0.05, "heun").item() - exact)
    assert e2 < e1 / 10 and e2h < e2 / 3.5


def test_flop_model_matches_baseline_md():
    assert O.flops_per_forward(CASES["large_uncond"]["cfg"]) / 1e9 == pytest.approx(152.298, abs=2e-3)
    assert O.flops_per_forward(CASES["large_t2i"]["cfg"]) / 1e9 == pytest.approx(200.258, abs=2e-3)
    assert O.flops_per_forward(CASES["small16_uncond"]["cfg"]) / 1e9 == pytest.approx(31.952, abs=2e-3)


def test_fast_library_mode_equals_explicit_formulas():
    case = CASES["tiny_t2i"]
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, t, y, ctx = build_inputs(case)
    a = O.uvit_forward(sd, case["cfg"], x, t, context=ctx)
    O.FAST = True
    try:
        b = O.uvit_forward(sd, case["cfg"], x, t, context=ctx)
    finally:
        O.FAST = False
    assert rel(b, a) < 2e-6


def test_config0_single_euler_step_small_deep16(golden_dir):
    """BASELINE.json configs[0]: lfm_cm256_uvit_small_deep16, one Euler step, batch 2, CPU (plumbing check)."""
    case = CASES["small16_uncond"]
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, _, _, _ = build_inputs(case)
    h = 0.02
    z1 = O.odeint_fixed(lambda t, z: O.uvit_forward(sd, case["cfg"], z, t.expand(z.shape[0])), x, 0.0, h, h, "euler")
    v0 = O.uvit_forward(sd, case["cfg"], x, torch.zeros(2))
    assert z1.shape == x.shape and torch.isfinite(z1).all()
    assert rel(z1, x + np.float32(h) * v0) < 1e-6


def _p2p_colscale(case, L):
    pp = case["p2p"]
    cs = torch.ones(case["B"], L)
    for i, ids in enumerate(pp["ids"]):
        cs[i, torch.tensor(ids) + 1] = pp["multiplier"][i]
    return cs


def test_oracle_p2p_attention_rescale_matches_reference(golden_dir):
    """Post-softmax column re-weighting == the reference's editing_attention_map_vit / _p2p_rescale branch."""
    case = CASES["tiny_t2i"]
    pp = case["p2p"]
    g = load(golden_dir, "tiny_t2i")
    sd = build_model(case, UViT, UViTT2I).state_dict()
    x, _, _, ctx = build_inputs(case)
    t = torch.full((x.shape[0],), pp["t"])
    cs = _p2p_colscale(case, 334)
    got = O.uvit_forward(sd, case["cfg"], x, t, context=ctx, attn_colscale=cs, attn_blocks=pp["block_id"])
    assert rel(got, g["p2p_forward"]) < 2e-6
    got_all = O.uvit_forward(sd, case["cfg"], x, t, context=ctx, attn_colscale=cs, attn_blocks=None)
    assert rel(got_all, g["p2p_all_blocks"]) < 2e-6
    plain = O.uvit_forward(sd, case["cfg"], x, t, context=ctx)
    assert rel(plain, g["p2p_plain"]) < 2e-6
    assert rel(g["p2p_forward"], g["p2p_plain"]) > 1e-3   # the edit is visible


# ---- adaptive dopri5 (torchdiffeq is absent: the restated constants are pinned mathematically) ----------------
def test_dopri5_tableau_against_scipy_and_order_conditions():
    from scipy.integrate._ivp.rk import RK45
    assert np.allclose(RK45.C[1:], O.DP_ALPHA[:5], rtol=0, atol=1e-15)
    assert np.allclose(RK45.B, O.DP_C_SOL[:6], rtol=0, atol=1e-15)
    for i in range(5):
        assert np.allclose(RK45.A[i + 1][:i + 1], O.DP_BETA[i], rtol=0, atol=1e-15)
    assert O.DP_BETA[5] == O.DP_C_SOL[:6]          # FSAL: the last stage state is the 5th-order solution
    # the embedded companion b^ = c_sol - c_error is a 4th-order method: all 8 rooted-tree conditions up to order 4
    c = np.array([0.0] + O.DP_ALPHA)
    A = np.zeros((7, 7))
    for i, row in enumerate(O.DP_BETA):
        A[i + 1, :len(row)] = row
    for b, order in ((np.array(O.DP_C_SOL), 5), (np.array(O.DP_C_SOL) - np.array(O.DP_C_ERROR), 4)):
        assert abs(b.sum() - 1) < 1e-14
        assert abs(b @ c - 1 / 2) < 1e-14
        assert abs(b @ c ** 2 - 1 / 3) < 1e-14 and abs(b @ A @ c - 1 / 6) < 1e-14
        assert abs(b @ c ** 3 - 1 / 4) < 1e-14 and abs((b * c) @ A @ c - 1 / 8) < 1e-14
        assert abs(b @ A @ c ** 2 - 1 / 12) < 1e-14 and abs(b @ A @ A @ c - 1 / 24) < 1e-14
        if order == 5:
            assert abs(b @ c ** 4 - 1 / 5) < 1e-14
    # the companion really is only 4th order (so the difference is an error estimate, not zero)
    assert abs((np.array(O.DP_C_SOL) - np.array(O.DP_C_ERROR)) @ c ** 4 - 1 / 5) > 1e-4
    assert abs(sum(O.DP_C_ERROR)) < 1e-15


def test_dopri5_dense_output_and_step():
    f = lambda t, y: y * np.cos(t)                       # y = y0 exp(sin t)
    y0 = torch.tensor([1.0, -2.0], dtype=torch.float64)
    errs = []
    for dt in (0.2, 0.1):
        y1, f1, err, k = O.dopri5_step(f, 0.0, y0, f(0.0, y0), dt)
        assert torch.allclose(f1, f(dt, y1))
        errs.append(float((y1 - y0 * np.exp(np.sin(dt))).abs().max()))
        for x in (0.0, 0.25, 0.5, 1.0):
            mid = O.dopri5_interp(y0, y1, k, dt, x)
            assert float((mid - y0 * np.exp(np.sin(x * dt))).abs().max()) < 2e-6 * (dt / 0.2) ** 5 + 1e-15
    assert errs[0] / errs[1] > 40                        # local error ~ dt^6


def test_dopri5_against_closed_form_and_scipy():
    from scipy.integrate import solve_ivp
    f = lambda t, y: -2.0 * y + torch.sin(5.0 * torch.as_tensor(t)) * torch.ones_like(y)
    y0 = torch.tensor([1.0, 0.5, -3.0], dtype=torch.float64)
    ref = solve_ivp(lambda t, y: -2.0 * y + np.sin(5.0 * t), (0.0, 1.0), y0.numpy(), rtol=1e-13, atol=1e-13).y[:, -1]
    for tol, bound in ((1e-5, 1e-4), (1e-8, 1e-7)):    # global error ~ a few local tolerances
        st = {}
        y = O.odeint_dopri5(f, y0, 0.0, 1.0, tol, tol, stats=st)
        assert np.abs(y.numpy() - ref).max() < bound
        sp = solve_ivp(lambda t, y: -2.0 * y + np.sin(5.0 * t), (0.0, 1.0), y0.numpy(), rtol=tol, atol=tol, method="RK45")
        # same family of controller (Hairer start, 0.9 safety, [0.2, 10] clamp): step counts agree within a few
        assert abs((st["n_accept"] + st["n_reject"]) - (sp.nfev - 2) // 6) <= 3
        assert st["nfe"] == 2 + 6 * (st["n_accept"] + st["n_reject"])
    # reversed time returns to the start
    back = O.odeint_dopri5(f, torch.as_tensor(ref), 1.0, 0.0, 1e-9, 1e-9)
    assert (back - y0).abs().max() < 1e-6


def test_oracle_adaptive_sampling_agrees_with_a_fine_fixed_grid():
    case = CASES["tiny_uncond"]
    sd = {k: v.double() for k, v in build_model(case, UViT, UViTT2I).state_dict().items()}
    x = build_inputs(case)[0][:1].double()
    st = {}
    ada = O.sample_adaptive(sd, case["cfg"], x, 0.0, 1.0, 1e-5, 1e-5, stats=st)
    fine = O.sample(sd, case["cfg"], x, 0.0, 1.0, 0.01, "heun")
    assert rel(ada, fine) < 5e-5 and st["n_accept"] >= 2
    assert O.digit_index(0.125) == 12 and O.digit_index(0.4049999) == 40 and O.digit_index(0.405001) == 41


def test_fixed_midpoint_and_rk4_orders_of_accuracy():
    """torchdiffeq "midpoint" (2nd order) and "rk4" (3/8 rule, 4th order) as restated in odeint_fixed."""
    f = lambda t, y: y * torch.cos(t.double())
    y0 = torch.tensor([1.0, -0.5], dtype=torch.float64)
    exact = y0 * np.exp(np.sin(1.0))
    errs = {}
    for method in ("euler", "heun", "midpoint", "rk4"):
        errs[method] = [float((O.odeint_fixed(f, y0, 0.0, 1.0, h, method) - exact).abs().max()) for h in (0.1, 0.05)]
    assert 1.7 < errs["euler"][0] / errs["euler"][1] < 2.3
    assert 3.4 < errs["heun"][0] / errs["heun"][1] < 4.6
    assert 3.4 < errs["midpoint"][0] / errs["midpoint"][1] < 4.6
    assert 12.0 < errs["rk4"][0] / errs["rk4"][1] < 20.0      # fp32 grid times put a floor under the finest error
    assert errs["rk4"][0] < 1e-5 < errs["midpoint"][0]
    # one rk4 step against the textbook 3/8-rule combination
    h = 0.5
    k1 = f(torch.tensor(0.0), y0)
    k2 = f(torch.tensor(h / 3), y0 + h * k1 / 3)
    k3 = f(torch.tensor(2 * h / 3), y0 + h * (k2 - k1 / 3))
    k4 = f(torch.tensor(h), y0 + h * (k1 - k2 + k3))
    one = O.odeint_fixed(f, y0, 0.0, h, h, "rk4")
    assert (one - (y0 + h * (k1 + 3 * k2 + 3 * k3 + k4) / 8)).abs().max() < 1e-7


def test_bosh3_and_adaptive_heun_tableaus_and_accuracy():
    """Order conditions of the restated Bogacki-Shampine 3(2) and Heun 2(1) pairs, and end-to-end accuracy."""
    for name, orders in (("bosh3", (3, 2)), ("adaptive_heun", (2, 1))):
        tab = O.ADAPTIVE_TABLEAUS[name]
        n = len(tab["alpha"]) + 1
        c = np.array([0.0] + list(tab["alpha"]))
        A = np.zeros((n, n))
        for i, row in enumerate(tab["beta"]):
            A[i + 1, :len(row)] = row
        assert np.allclose(A.sum(1), c)                       # row-sum condition
        for b, order in ((np.array(tab["c_sol"]), orders[0]), (np.array(tab["c_sol"]) - np.array(tab["c_error"]), orders[1])):
            assert abs(b.sum() - 1) < 1e-14
            if order >= 2:
                assert abs(b @ c - 1 / 2) < 1e-14
            if order >= 3:
                assert abs(b @ c ** 2 - 1 / 3) < 1e-14 and abs(b @ A @ c - 1 / 6) < 1e-14
        assert abs(sum(tab["c_error"])) < 1e-15
        # the mid-point weights reproduce y(t0 + dt/2) to second order at least
        assert abs(sum(tab["c_mid"]) - 0.5) < 1e-15
    from scipy.integrate._ivp.rk import RK23
    assert np.allclose(RK23.C[1:], O.ADAPTIVE_TABLEAUS["bosh3"]["alpha"][:2])
    assert np.allclose(RK23.B, O.ADAPTIVE_TABLEAUS["bosh3"]["c_sol"][:3])
    assert np.allclose(RK23.E, -np.array(O.ADAPTIVE_TABLEAUS["bosh3"]["c_error"]))     # scipy: E = b^ - b
    f = lambda t, y: -2.0 * y + torch.sin(5.0 * torch.as_tensor(t)) * torch.ones_like(y)
    y0 = torch.tensor([1.0, 0.5, -3.0], dtype=torch.float64)
    from scipy.integrate import solve_ivp
    ref = solve_ivp(lambda t, y: -2.0 * y + np.sin(5.0 * t), (0.0, 1.0), y0.numpy(), rtol=1e-13, atol=1e-13).y[:, -1]
    steps = {}
    for method in ("dopri5", "bosh3", "adaptive_heun"):
        st = {}
        y = O.odeint_dopri5(f, y0, 0.0, 1.0, 1e-6, 1e-6, stats=st, method=method)
        assert np.abs(y.numpy() - ref).max() < 2e-4, method      # global error: a few hundred local tolerances
        steps[method] = st["n_accept"] + st["n_reject"]
        assert st["nfe"] == 2 + len(O.ADAPTIVE_TABLEAUS[method]["alpha"]) * steps[method]
    assert steps["dopri5"] < steps["bosh3"] < steps["adaptive_heun"]     # lower order, more steps


# ---- latent -> image decoder ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["vae_small", "vae_full"])
def test_vae_oracle_matches_reference_golden(golden_dir, name):
    from oracle import vae_oracle as V
    from tests.golden.cases import vae_latents, vae_state_dict
    sd = vae_state_dict()
    got = V.decode(sd, vae_latents(name))
    want = load(golden_dir, name)["decode"]
    assert got.shape == want.shape and rel(got, want) < 5e-6


def test_vae_mirror_state_dict_layout_and_flops():
    from oracle import vae_oracle as V
    from uspace_b200.autoencoder import FrozenAutoencoderKL, get_model
    m = get_model()
    sd = m.state_dict()
    assert len(sd) == 248 and sum(v.numel() for v in sd.values()) == 83653863         # the SD KL-f8 autoencoder
    dec = {k: v for k, v in sd.items() if k.startswith(("decoder.", "post_quant_conv."))}
    assert len(dec) == 140 and sum(v.numel() for v in dec.values()) == 49490179 + 20
    assert sd["decoder.up.1.block.0.nin_shortcut.weight"].shape == (256, 512, 1, 1)
    assert sd["decoder.up.3.upsample.conv.weight"].shape == (512, 512, 3, 3)
    assert "decoder.up.0.upsample.conv.weight" not in sd and sd["post_quant_conv.weight"].shape == (4, 4, 1, 1)
    assert sd["encoder.down.2.downsample.conv.weight"].shape == (512, 512, 3, 3) and sd["quant_conv.weight"].shape == (8, 8, 1, 1)
    assert sd["encoder.conv_out.weight"].shape == (8, 512, 3, 3) and "encoder.down.3.downsample.conv.weight" not in sd
    m.load_state_dict(sd)
    with pytest.raises(RuntimeError):
        m.load_state_dict(dict(sd, bogus=torch.zeros(1)))
    with pytest.raises(RuntimeError, match="no CPU"):
        m.decode(torch.zeros(1, 4, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU"):
        m.encode(torch.zeros(1, 3, 256, 256))
    with pytest.raises(NotImplementedError):
        FrozenAutoencoderKL(dict(V.DDCONFIG, ch=64))
    assert V.flops_per_image(32) / 1e9 == pytest.approx(622.19, rel=1e-3)      # 0.62 TFLOP per 256^2 image
    from uspace_b200.autoencoder import flops_per_image
    assert flops_per_image(32) == V.flops_per_image(32) and flops_per_image(16) == V.flops_per_image(16)


@pytest.mark.parametrize("name", ["vae_enc_small", "vae_enc_full"])
def test_vae_encoder_oracle_matches_reference_golden(golden_dir, name):
    from oracle import vae_oracle as V
    from tests.golden.cases import vae_enc_state_dict, vae_images
    got = V.encode_moments(vae_enc_state_dict(), vae_images(name))
    want = load(golden_dir, name)["moments"]
    assert got.shape == want.shape and rel(got, want) < 5e-6
