"""Golden for uspace_b200.attr_delta from the UNMODIFIED reference (tools/utils_attr.py), build container only.

    python tests/golden/make_golden_attr.py     # needs /root/reference; writes tests/golden/attr_delta.npz

Inputs are re-created from the seed by the test (tests/golden/cases.py::attr_delta_inputs); the fixture holds only the
arrays the reference wrote (delta_{t}.npy per timestep and delta_latentz.npy)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from tools.utils_attr import extract_hspace_feat_unet_by_attr  # noqa: E402
from tests.golden.cases import attr_delta_inputs  # noqa: E402


def main():
    feats, latent, attr, times, batch_num = attr_delta_inputs()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        np.savez(os.path.join(d, "latents.npy"), latent=latent, attr=attr)
        per = feats.shape[0] // batch_num
        for ti, ts in enumerate(times):
            for b in range(batch_num):
                np.save(os.path.join(d, f"{b}_{ts}"), feats[b * per:(b + 1) * per, ti])
        extract_hspace_feat_unet_by_attr(d, batch_num)
        extract_hspace_feat_unet_by_attr(d, batch_num, cal_latentz_delta_only=True)
        for ts in times:
            out[f"delta_{ts}"] = np.load(os.path.join(d, f"delta_{ts}.npy"))
        out["delta_latentz"] = np.load(os.path.join(d, "delta_latentz.npy"))
    path = os.path.join(HERE, "attr_delta.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, os.path.getsize(path), "B ->", path)


if __name__ == "__main__":
    main()
