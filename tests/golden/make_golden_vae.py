"""Golden vectors for the latent -> image decoder from the UNMODIFIED reference (libs/autoencoder.py), build container only.

    python tests/golden/make_golden_vae.py      # needs /root/reference; writes tests/golden/vae_small.npz, vae_full.npz

FrozenAutoencoderKL.__init__ loads a checkpoint file, which does not exist here: the reference's own Decoder and
post_quant_conv are instantiated directly with random weights under a fixed seed, exactly as uspace_b200.autoencoder
re-creates them (this script asserts the two state_dicts are identical), so the fixtures hold outputs only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from libs.autoencoder import Decoder  # noqa: E402
from tests.golden.cases import VAE_CASES, vae_state_dict, vae_latents  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    torch.manual_seed(VAE_CASES["seed"])
    dec = Decoder(**dd).eval()
    pq = torch.nn.Conv2d(4, 4, 1)
    sd = vae_state_dict()
    ref = {f"decoder.{k}": v for k, v in dec.state_dict().items()}
    ref.update({f"post_quant_conv.{k}": v for k, v in pq.state_dict().items()})
    assert list(ref) == list(sd) and all(torch.equal(ref[k], sd[k]) for k in ref), "mirror ctor != reference weights"
    for name in ("vae_small", "vae_full"):
        z = vae_latents(name)
        with torch.no_grad():
            out = dec(pq((1.0 / 0.18215) * z))           # FrozenAutoencoderKL.decode, libs/autoencoder.py:446-450
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, decode=out.numpy().astype(np.float32))
        print(name, tuple(out.shape), float(out.std()), os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
