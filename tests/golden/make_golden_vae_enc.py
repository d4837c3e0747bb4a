"""Golden vectors for the image -> latent-moments encoder from the UNMODIFIED reference (libs/autoencoder.py).

    python tests/golden/make_golden_vae_enc.py    # build container only; writes tests/golden/vae_enc_*.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from libs.autoencoder import Encoder  # noqa: E402
from tests.golden.cases import VAE_ENC_CASES, vae_enc_state_dict, vae_images  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    torch.manual_seed(VAE_ENC_CASES["seed"])
    enc = Encoder(**dd).eval()
    qc = torch.nn.Conv2d(8, 8, 1)
    sd = vae_enc_state_dict()
    ref = {f"encoder.{k}": v for k, v in enc.state_dict().items()}
    ref.update({f"quant_conv.{k}": v for k, v in qc.state_dict().items()})
    assert list(ref) == list(sd) and all(torch.equal(ref[k], sd[k]) for k in ref), "mirror ctor != reference weights"
    for name in ("vae_enc_small", "vae_enc_full"):
        with torch.no_grad():
            out = qc(enc(vae_images(name)))        # FrozenAutoencoderKL.encode_moments, libs/autoencoder.py:426-429
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, moments=out.numpy().astype(np.float32))
        print(name, tuple(out.shape), float(out.std()), os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
