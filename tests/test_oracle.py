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
