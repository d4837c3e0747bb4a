"""Developer probe (run under gpurun): decoder accuracy against the goldens and stand-alone throughput."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vae_oracle as V  # noqa: E402  (developer probe: torch-library timing of the restatement)
from tests.golden.cases import vae_enc_state_dict, vae_latents, vae_state_dict  # noqa: E402
from uspace_b200.autoencoder import get_model  # noqa: E402

dev = torch.device("cuda:0")
m = get_model()
m.load_state_dict({**vae_state_dict(), **vae_enc_state_dict()})
m = m.to(dev)
for name in ("vae_small", "vae_full"):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))["decode"]
    out = m.decode(vae_latents(name).to(dev)).cpu().double()
    r = torch.from_numpy(g).double()
    print(name, "rel err", float((out - r).norm() / r.norm()), "max abs", float((out - r).abs().max()), flush=True)
for B in (1, 8, 50, 64):
    z = (0.7 * torch.randn(B, 4, 32, 32)).to(dev)
    m.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 3
    e0.record()
    for _ in range(n):
        m.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"decode B={B}: {ms:.2f} ms  {B / ms * 1e3:.1f} img/s  {V.flops_per_image(32) * B / ms / 1e9:.0f} TFLOP/s", flush=True)
# the same decoder through torch library kernels (cuDNN), for context
sd = {k: v.to(dev) for k, v in vae_state_dict().items()}
z = (0.7 * torch.randn(8, 4, 32, 32)).to(dev)
for dt in (torch.float32, torch.float16):
    sdd = {k: v.to(dt) for k, v in sd.items()}
    with torch.no_grad():
        V.decode(sdd, z.to(dt))
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(3):
            V.decode(sdd, z.to(dt))
        torch.cuda.synchronize()
    print(f"torch {dt} B=8: {(time.time() - t) / 3 * 1e3:.1f} ms", flush=True)
for B in (8, 64):
    x = (torch.rand(B, 3, 256, 256) * 2 - 1).to(dev)
    m.encode_moments(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        m.encode_moments(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"encode_moments B={B}: {ms:.2f} ms  {B / ms * 1e3:.1f} img/s", flush=True)
