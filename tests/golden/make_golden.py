"""Generate golden vectors from the UNMODIFIED reference modules (run in the build container only).

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.npz

The reference ships no tests or fixtures, so these files are produced by importing
libs.uvit.UViT / libs.uvit_t2i.UViT from /root/reference on CPU (torch fp32 eager) with fixed seeds.
Weights are NOT stored: they are re-created from the seed by uspace_b200.uvit (this script asserts the mirror
constructor reproduces the reference state_dict bit-for-bit), so each fixture is only inputs' seed + outputs.
The GPU box has no /root/reference; tests read only the committed .npz files.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# libs.uvit_t2i -> tools.utils_t2i -> tools.ptp_utils imports IPython (absent here): stub it
if "IPython" not in sys.modules:
    ip = types.ModuleType("IPython")
    ipd = types.ModuleType("IPython.display")
    ipd.display = lambda *a, **k: None
    ip.display = ipd
    sys.modules["IPython"] = ip
    sys.modules["IPython.display"] = ipd

from libs.uvit import UViT as RefUViT  # noqa: E402
from libs.uvit_t2i import UViT as RefUViTT2I  # noqa: E402
from tests.golden.cases import CASES, build_inputs  # noqa: E402
from uspace_b200.uvit import UViT, UViTT2I  # noqa: E402
from oracle.uvit_oracle import fixed_grid  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, case in CASES.items():
        cfg, t2i = case["cfg"], case["t2i"]
        torch.manual_seed(case["seed"])
        ref = (RefUViTT2I if t2i else RefUViT)(**cfg).eval()
        torch.manual_seed(case["seed"])
        mine = (UViTT2I if t2i else UViT)(**cfg)
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys()), name
        assert all(torch.equal(a[k], b[k]) for k in a), f"{name}: mirror ctor does not reproduce reference weights"
        x, t, y, ctx = build_inputs(case)
        out = {}
        with torch.no_grad():
            if t2i:
                out["forward"] = ref(x, t, ctx)[0].numpy()
            else:
                out["forward"] = ref(x, t, y, edit_loc=None)[0].numpy()
            if case.get("euler_steps"):
                # reference net driven by the restated fixed-grid Euler loop (torchdiffeq is not installed)
                h = 1.0 / case["euler_steps"]
                grid = fixed_grid(0.0, 1.0, h)
                z = x.clone()
                for i in range(len(grid) - 1):
                    tt = grid[i].expand(z.shape[0])
                    v = ref(z, tt, ctx)[0] if t2i else ref(z, tt, y, edit_loc=None)[0]
                    z = z + (grid[i + 1] - grid[i]) * v
                out["euler"] = z.numpy()
            if case.get("p2p"):
                # the reference's attention-editing branch (libs/uvit_t2i.py:91-107 + tools/utils_t2i.py:265-296)
                pp = case["p2p"]
                kw = dict(dissect_name="p2p", fm_direction="decode", t_edit=pp["t_edit"], block_id=pp["block_id"],
                          token_kwargs=dict(token_dissect="p2p_rescale", p2p_multiplier=pp["multiplier"]),
                          target_context_ids=[np.array(i) for i in pp["ids"]])
                tt = torch.full((x.shape[0],), pp["t"])
                out["p2p_forward"] = ref(x, tt, ctx, **kw)[0].numpy()
                out["p2p_plain"] = ref(x, tt, ctx)[0].numpy()
                kw_all = dict(kw, block_id="all")
                out["p2p_all_blocks"] = ref(x, tt, ctx, **kw_all)[0].numpy()
            if case.get("edit"):
                # the reference's own edit hook (libs/dissection.py:115-157) reading delta_{t:.2f}.npy from disk
                e = case["edit"]
                g = torch.Generator().manual_seed(e["seed"])
                delta = 0.1 * torch.randn(3, *x.shape[1:], generator=g)  # [n_attr, C, S, S]
                with tempfile.TemporaryDirectory() as d:
                    np.save(os.path.join(d, f"delta_{e['t']:.2f}.npy"), delta.numpy())
                    kw = dict(dissect_task="uspace_uvit", dissect_name="write_attr", write_path_root=d,
                              ith_attr=e["ith_attr"], t_edit=e["t_edit"], write_scale=e["write_scale"])
                    tt = torch.full((x.shape[0],), e["t"])
                    for loc in ("head", "tail"):
                        out[f"edit_{loc}"] = ref(x, tt, y, edit_loc=loc, **kw)[0].numpy()
                out["edit_delta"] = delta.numpy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        o = out["forward"]
        print(f"{name}: forward mean {o.mean():+.6f} std {o.std():.6f} -> {path} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main()
