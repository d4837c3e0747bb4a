"""torchrun helper: sharded sampling on N GPUs must reproduce the single-GPU result sample for sample.
Launched by tests/test_gpu.py (2 GPUs) and usable by hand: torchrun --nproc-per-node N tests/mp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden.cases import CASES, build_model  # noqa: E402
from uspace_b200 import parallel  # noqa: E402
from uspace_b200.uvit import UViT, UViTT2I  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    case = CASES["tiny_t2i"]
    m = build_model(case, UViT, UViTT2I).to(dev)
    eng = m.engine()
    n = 2 * world + 1  # ragged on purpose
    z = parallel.global_noise(n)
    ctx = torch.randn(n, 77, 768, generator=torch.Generator().manual_seed(1231))
    out = parallel.sample_sharded(lambda zl, cl: eng.sample(zl, 0.0, 1.0, 0.25, "heun", context=cl), z, ctx, device=dev)
    assert out.shape == (n, 4, 32, 32)
    if rank == 0:
        ref = eng.sample(z.to(dev), 0.0, 1.0, 0.25, "heun", context=ctx.to(dev))
        assert torch.equal(out, ref), (out - ref).abs().max().item()
        print("MP_CHECK_OK", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
