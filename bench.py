#!/usr/bin/env python
"""Headline benchmark: images/sec at 50 ODE steps, U-ViT-L 256 (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one complete 50-NFE Euler sampling (CNF.decode, flow_matching.py:130-151) of one batch of 64 images
per GPU (BASELINE.json configs[1], lfm_cm256_uvit_large), synthetic N(0,1) latents and reference-initialised
random weights.  `value` times K steps with the latents already in HBM; `e2e` times the same K steps through
usp_sample_host() with pinned HOST latents (H2D + sampling + D2H inside the timed region).  N > 1 shards the
global batch (64 per GPU, weak scaling) and ends every step with the one NCCL all-gather of final latents.

--impl reference times the reference algorithm's CPU implementation (the torch-fp32 oracle port; the reference
itself is pure Python and cannot travel to the GPU box) on all host cores, on a bounded sample of the same work.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_L = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=20, num_heads=16, mlp_ratio=4,
             qkv_bias=False, mlp_time_embed=False, num_classes=-1, use_checkpoint=False)  # configs/lfm_cm256_uvit_large.py:42-56
CFG_L_T2I = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=20, num_heads=16, mlp_ratio=4,
                 qkv_bias=False, mlp_time_embed=False, clip_dim=768, num_clip_token=77, use_checkpoint=False)
WORKLOADS = {
    "c2": dict(name="lfm_cm256_uvit_large: 50-step Euler sampling, batch=64 per GPU, U-ViT-L 256 (configs[1])",
               cfg=CFG_L, t2i=False, batch=64, method="euler", step=0.02),
    "c3": dict(name="lfm_mmcelebahq256_uvit_large (t2i): 50-step Euler, 77-token context, batch=128 per GPU (configs[2])",
               cfg=CFG_L_T2I, t2i=True, batch=128, method="euler", step=0.02),
    "c4": dict(name="lfm_mscoco_uvit_from_in256 (t2i): 50-step Heun, batch=32 per GPU (configs[3])",
               cfg=CFG_L_T2I, t2i=True, batch=32, method="heun", step=0.02),
    "c5": dict(name="dissect sweep: U-ViT-L, 50-step Euler, tail write_attr edit, batch=64 per GPU (configs[4])",
               cfg=CFG_L, t2i=False, batch=64, method="euler", step=0.02, edit=True),
    # the whole sweep of configs/lfm_cm256_uvit_large.py:84 as one batch: 64 latents x 9 write_scales per GPU
    "c5sweep": dict(name="dissect sweep: U-ViT-L, 50-step Euler, tail write_attr edit, 64 latents x 9 write_scales per GPU "
                         "in one batch (configs[4], tools/utils_vis.py:189-201)",
                    cfg=CFG_L, t2i=False, batch=64, method="euler", step=0.02, edit=True,
                    scales=[-2.1, -1.5, -1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]),
}
METRIC = "images/sec at 50 ODE steps, U-ViT-L 256"


def decoder_line(latents, dev):
    """Context, not the headline: the latent -> image decoder (csrc/vae.cu) on the latents the timed run produced."""
    import torch

    from uspace_b200.autoencoder import flops_per_image, get_model
    try:
        torch.manual_seed(0)
        vae = get_model().to(dev)
        z = latents.contiguous()
        vae.decode(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            img = vae.decode(z)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        alg = flops_per_image(32) * z.shape[0] / ms / 1e9
        return {"images_per_s": z.shape[0] / ms * 1e3, "batch": z.shape[0], "ms": ms, "precision": vae.precision,
                "achieved_tflops": alg, "issued_tflops": alg * (3 if vae.precision == "fp16x3" else 1),
                "finite": bool(torch.isfinite(img).all().item()),
                "note": "FrozenAutoencoderKL.decode, random-init weights; follows sampling in the reference's pipeline; "
                        "fp16x3 = split hi + lo operands, three products per GEMM (achieved counts the algorithm's FLOPs once)"}
    except Exception as e:   # context only: never fail the headline line
        return {"error": str(e)[:200]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for ts, r in self.rows if t_begin <= ts <= t_end and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=float(rows[0][1]) if rows[0][1].isdigit() else None,
                    reasons=reasons, power_w_max=max(pw) if pw else None, samples=len(rows))


def cpu_port_forward_time(cfg, t2i, B, max_seconds=25.0, warmup=1, reps=3):
    """Times the CPU oracle (torch fp32, all host threads) for one velocity evaluation at batch B."""
    import torch

    from oracle import uvit_oracle as O
    from uspace_b200.uvit import UViT, UViTT2I
    torch.set_num_threads(os.cpu_count() or 1)
    O.FAST = True  # the same fused library calls the reference's torch-eager path makes
    torch.manual_seed(0)
    sd = (UViTT2I if t2i else UViT)(**cfg).state_dict()
    g = torch.Generator().manual_seed(1230)
    x = torch.randn(B, 4, 32, 32, generator=g)
    t = torch.full((B,), 0.5)
    ctx = torch.randn(B, 77, 768, generator=g) if t2i else None
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            O.uvit_forward(sd, cfg, x, t, context=ctx)
        t_start = time.perf_counter()
        for _ in range(reps):
            t0 = time.perf_counter()
            O.uvit_forward(sd, cfg, x, t, context=ctx)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > max_seconds:
                break
    return statistics.median(times), len(times)


_STAGED_NETS = {}


def staged_reference_forward_time(cfg, t2i, B, max_seconds=25.0, warmup=1, reps=3):
    """Times the reference's OWN module (libs.uvit.UViT / libs.uvit_t2i.UViT from baseline/_ref, staged unmodified by
    baseline/stage_reference.py) for one velocity evaluation at batch B on the host cores.  Returns None when the
    staged copy is absent or cannot be imported here (then the oracle port is timed instead)."""
    import importlib
    import types

    import torch
    ref_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_root, "libs", "uvit.py")):
        return None
    try:
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
        if "IPython" not in sys.modules:      # tools.ptp_utils imports it for notebook display only
            ip, ipd = types.ModuleType("IPython"), types.ModuleType("IPython.display")
            ipd.display = lambda *a, **k: None
            ip.display = ipd
            sys.modules["IPython"], sys.modules["IPython.display"] = ip, ipd
        import contextlib
        key = ("t2i" if t2i else "uvit", json.dumps(cfg, sort_keys=True, default=str))
        net = _STAGED_NETS.get(key)
        if net is None:
            with contextlib.redirect_stdout(sys.stderr):     # the reference modules print at import time
                mod = importlib.import_module("libs.uvit_t2i" if t2i else "libs.uvit")
            torch.set_num_threads(os.cpu_count() or 1)
            torch.manual_seed(0)
            net = mod.UViT(**cfg).eval()
            _STAGED_NETS[key] = net
    except Exception as ex:  # missing third-party dependency of the t2i closure, ...
        sys.stderr.write(f"bench: staged reference not importable ({type(ex).__name__}: {str(ex)[:100]}); timing the port\n")
        return None
    g = torch.Generator().manual_seed(1230)
    x = torch.randn(B, 4, 32, 32, generator=g)
    t = torch.full((B,), 0.5)
    ctx = torch.randn(B, 77, 768, generator=g) if t2i else None
    y = torch.zeros(B, dtype=torch.long) if (not t2i and cfg.get("num_classes", -1) > 0) else None
    times = []
    with torch.no_grad():
        def fwd():
            return net(x, t, ctx) if t2i else net(x, t, y, edit_loc=None)
        for _ in range(warmup):
            fwd()
        t_start = time.perf_counter()
        for _ in range(reps):
            t0 = time.perf_counter()
            fwd()
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > max_seconds:
                break
    return statistics.median(times), len(times)


def torch_eager_on_gpu(wl, dev, nfe):
    """Context only (not the product, not the reference arm): the reference's algorithm as plain torch-eager library
    calls on the same B200 (fp32 as shipped - Attention.forward forces .float(), libs/uvit.py:93 - and bf16 autocast),
    timed on a few velocity evaluations at the workload's batch, extrapolated like the CPU baseline."""
    import torch

    from oracle import uvit_oracle as O
    from uspace_b200.uvit import UViT, UViTT2I
    O.FAST = True
    torch.manual_seed(0)
    sd = {k: v.to(dev) for k, v in (UViTT2I if wl["t2i"] else UViT)(**wl["cfg"]).state_dict().items()}
    B = wl["batch"]
    x = torch.randn(B, 4, 32, 32, device=dev)
    t = torch.full((B,), 0.5, device=dev)
    ctx = torch.randn(B, 77, 768, device=dev) if wl["t2i"] else None
    out = {}
    with torch.no_grad():
        for name, ac in (("fp32", False), ("bf16_autocast", True)):
            try:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    for _ in range(2):
                        O.uvit_forward(sd, wl["cfg"], x, t, context=ctx)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        O.uvit_forward(sd, wl["cfg"], x, t, context=ctx)
                    e1.record()
                    torch.cuda.synchronize()
                out[name] = {"images_per_s": B / (e0.elapsed_time(e1) / 3 * 1e-3 * nfe), "ms_per_forward": e0.elapsed_time(e1) / 3}
            except Exception as ex:  # context only: never fail the benchmark because of it
                out[name] = {"error": str(ex)[:120]}
    out["note"] = "oracle restatement through torch library kernels on the GPU; extrapolated from 3 forwards; context only"
    return out


def run_reference(args, wl):
    """Reference arm: the reference's own modules (baseline/_ref, staged by baseline/stage_reference.py) on the host
    cores, bounded sample per step; the oracle port only when the staged copy is missing / not importable."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    nfe = 50 * (2 if wl["method"] == "heun" else 1)
    Bs = 8
    cores = os.cpu_count() or 1
    sec = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        r = staged_reference_forward_time(wl["cfg"], wl["t2i"], Bs, max_seconds=60, warmup=0, reps=1) if kind == "reference" else None
        if r is None:
            kind = "port"
            r = cpu_port_forward_time(wl["cfg"], wl["t2i"], Bs, max_seconds=60, warmup=0, reps=1)
        tf = r[0]
        if i >= args.warmup:
            sec.append(tf)
    tf = sum(sec) / len(sec)
    value = Bs / (tf * nfe)
    sample = f"each step = 1 velocity evaluation at batch {Bs} (of {nfe} x {wl['batch']}); images/s = {Bs}/(t_fwd*{nfe}), extrapolated"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tf * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": wl["name"], "nfe": nfe, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample,
                         "torch_threads": torch.get_num_threads(),
                         "what": ("libs.uvit.UViT.forward of the staged, unmodified reference (baseline/_ref)" if kind == "reference"
                                  else "oracle/uvit_oracle.py in FAST mode (staged reference absent)")},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fuse-ln", type=int, default=-1, help="override usp_config.fuse_layernorm (0 / 1)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist

    from uspace_b200 import parallel
    from uspace_b200.uvit import UViT, UViTT2I

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the uspace_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = (UViTT2I if wl["t2i"] else UViT)(**wl["cfg"]).eval()
    model.operand_dtype = args.operand
    if args.fuse_ln >= 0:
        model.fuse_layernorm = bool(args.fuse_ln)
    model = model.to(dev)
    eng = model.engine()
    Bl = wl["batch"]
    Bg = Bl * world
    nfe_per_step = 2 if wl["method"] == "heun" else 1
    n_grid = eng.grid_size(0.0, 1.0, wl["step"])
    nfe = (n_grid - 1) * nfe_per_step

    z_glob = parallel.global_noise(Bg, seed=1230)
    z_loc = parallel.shard(z_glob, rank, world).contiguous()
    ctx_loc = None
    if wl["t2i"]:
        ctx = torch.randn(Bg, 77, 768, generator=torch.Generator().manual_seed(1231))
        ctx_loc = parallel.shard(ctx, rank, world).contiguous()
    delta = None
    ekw = {}
    if wl.get("edit"):
        delta = 0.1 * torch.randn(n_grid, 4, 32, 32, generator=torch.Generator().manual_seed(1232))
        ekw = dict(delta_table=delta.to(dev), write_scale=1.5, t_edit=0.4, edit_loc="tail")
    z_dev = z_loc.to(dev)
    ctx_dev = None if ctx_loc is None else ctx_loc.to(dev)

    scales = wl.get("scales")
    n_rep = len(scales) if scales else 1

    def step_device():
        if scales:
            out = eng.sample_sweep(z_dev, scales, 0.0, 1.0, wl["step"], wl["method"], delta_table=ekw["delta_table"],
                                   t_edit=ekw["t_edit"], edit_loc=ekw["edit_loc"]).flatten(0, 1)
            return parallel.gather_latents(out, Bg * n_rep) if world > 1 else out
        out = eng.sample(z_dev, 0.0, 1.0, wl["step"], wl["method"], context=ctx_dev, **ekw)
        return parallel.gather_latents(out, Bg) if world > 1 else out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    sync_all()
    t_end = time.perf_counter()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    finite = bool(torch.isfinite(out).all().item())

    # ---- end to end: pinned host latents in, host latents out, every step ----
    zh_src = z_loc.clone().pin_memory()
    zh = torch.empty_like(zh_src).pin_memory()
    ctx_h = None if ctx_loc is None else ctx_loc.clone().pin_memory()
    ekw_h = dict(ekw)
    if delta is not None:
        ekw_h["delta_table"] = delta.clone().pin_memory()

    if scales:
        zh = torch.empty((Bl * n_rep, 4, 32, 32)).pin_memory()

    def step_host():
        if scales:   # no host-buffer variant of the sweep in the ABI: the copies are done here, inside the timed region
            o = eng.sample_sweep(zh_src.to(dev, non_blocking=True), scales, 0.0, 1.0, wl["step"], wl["method"],
                                 delta_table=ekw_h["delta_table"].to(dev, non_blocking=True), t_edit=ekw["t_edit"],
                                 edit_loc=ekw["edit_loc"])
            zh.copy_(o.flatten(0, 1), non_blocking=True)
            torch.cuda.synchronize()
            return
        zh.copy_(zh_src)
        eng.sample_host(zh, 0.0, 1.0, wl["step"], wl["method"], context=ctx_h, **ekw_h)

    step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = e2e_s.item()
    same = bool(torch.equal(zh.to(dev), out[rank * Bl * n_rep:(rank + 1) * Bl * n_rep] if world > 1 else out))
    h2d = zh_src.numel() * 4 + (ctx_h.numel() * 4 if ctx_h is not None else 0) + (delta.numel() * 4 if delta is not None else 0)
    d2h = zh.numel() * 4

    # ---- per-kernel-class device time of one velocity evaluation: events between launches, mean over 12 evaluations
    #      enqueued back to back right after the timed loops (the GPU is still at the power-capped clock of the run) ----
    t_half = torch.full((Bl * n_rep,), 0.5, device=dev)
    z_prof = z_dev if n_rep == 1 else z_dev.repeat_interleave(n_rep, 0)
    prof = eng.profile_forward(z_prof, t_half, context=ctx_dev, warmup=2, reps=12)
    torch.cuda.synchronize()

    if rank == 0:
        pk = peaks()
        D, n_in = wl["cfg"]["embed_dim"], wl["cfg"]["depth"] // 2
        Hd = int(D * wl["cfg"]["mlp_ratio"])
        L = (wl["cfg"]["img_size"] // wl["cfg"]["patch_size"]) ** 2 + (78 if wl["t2i"] else 1)
        M = Bl * n_rep * L
        # algorithmic FLOPs / bytes (16-bit operands in, result out; fp32 residual in+out where the epilogue has one)
        # of ONE launch of each GEMM class
        shapes = {"gemm_qkv": (3 * D, D, 2 * M * 3 * D), "gemm_proj": (D, D, 8 * M * D + 2 * M * D),
                  "gemm_fc1": (Hd, D, 2 * M * Hd), "gemm_fc2": (D, Hd, 8 * M * D + 2 * M * D),
                  "gemm_skip": (D, 2 * D, 4 * M * D + 2 * M * D)}
        classes = {}
        for k, (N, K, out_bytes) in shapes.items():
            ms_k, n_k = prof[k]
            if n_k == 0:
                continue
            fl = 2.0 * M * N * K
            by = 2.0 * M * K + 2.0 * N * K + out_bytes
            classes[k] = {"launches": n_k, "avg_launch_ms": ms_k / n_k, "flops_per_launch": fl, "algorithmic_bytes": by,
                          "tflops": fl / (ms_k / n_k * 1e-3) / 1e12}
        dom = max(classes, key=lambda k: prof[k][0])          # the class with the largest share of the evaluation
        gemm_ms = sum(prof[k][0] for k in classes)
        gemm_flops = sum(c["flops_per_launch"] * c["launches"] for c in classes.values())
        fwd_ms = sum(v[0] for v in prof.values())
        achieved = classes[dom]["tflops"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by profiles/summarize.py from the ncu capture
        if os.path.exists(tpath) and args.workload == "c2" and Bl == 64:
            tj = json.load(open(tpath))
            if dom in tj.get("kernels", {}):
                traffic = tj["kernels"][dom]["dram_bytes"]
                classes[dom]["traffic_source"] = tj.get("source")
        kernel_names = {"gemm_qkv": "gemm2_kernel<EPI_QKV>", "gemm_proj": "gemm2_kernel<EPI_BIAS_RESID> (proj)",
                        "gemm_fc1": "gemm2_kernel<EPI_BIAS_GELU> (fc1 + erf-GELU)",
                        "gemm_fc2": "gemm2_kernel<EPI_BIAS_RESID, LONGK> (fc2)", "gemm_skip": "gemm2_kernel<EPI_BIAS_F32> (skip_linear)"}
        flops_img = eng.flops_per_forward()
        sec = ms_total * 1e-3
        value = Bg * n_rep * args.steps / sec
        res = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.operand, "data": "synthetic",
            "config": {"workload": wl["name"], "global_batch": Bg, "per_gpu_batch": Bl, "nfe": nfe, "method": wl["method"],
                       "parallelism": f"dp{world}", "weights": "random-init (reference ctor, seed 0)",
                       "accumulate": "fp32 (TMEM), fp32 residual stream / LayerNorm / softmax",
                       "l2": "activations per velocity evaluation (~1.9 GB at batch 64) exceed the 126 MB L2; no flush needed"},
            "achieved_tflops_per_gpu": flops_img * Bl * n_rep * nfe * args.steps / sec / 1e12,
            "tensor_peak_frac_burst": flops_img * Bl * n_rep * nfe * args.steps / sec / 1e12 / pk["burst"],
            "tensor_peak_frac_sustained": flops_img * Bl * n_rep * nfe * args.steps / sec / 1e12 / pk["sustained"],
            "e2e": {"value": Bg * n_rep * args.steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "matches_device_path": same},
            "gpu_launches": args.steps * (n_grid - 1) * nfe_per_step * (eng.kernels_per_forward() + 1),
            "roofline": {"bound": "tensor", "kernel": "usp::" + kernel_names[dom] + ": the dominant kernel of the step (2-CTA tcgen05 GEMM)",
                         "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                         "frac_burst": achieved / pk["burst"],
                         "peak_kind": f"bf16_tflops_sustained of {pk['src']} (kernel timed inside a long run at the power cap); burst {pk['burst']}",
                         "launches": classes[dom]["launches"], "avg_launch_ms": classes[dom]["avg_launch_ms"],
                         "share_of_forward": prof[dom][0] / fwd_ms,
                         "how": "CUDA events between launches on the launching stream, mean over 12 evaluations enqueued back "
                                "to back right after the timed loops",
                         "traffic": traffic, "algorithmic_bytes": classes[dom]["algorithmic_bytes"],
                         "all_gemms": {"tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12,
                                       "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / pk["sustained"],
                                       "frac_burst": gemm_flops / (gemm_ms * 1e-3) / 1e12 / pk["burst"],
                                       "share_of_forward": gemm_ms / fwd_ms},
                         "classes": {k: {"tflops": round(c["tflops"], 1), "avg_launch_us": round(c["avg_launch_ms"] * 1e3, 2),
                                         "launches": c["launches"]} for k, c in classes.items()}},
            "kernel_ms_per_forward": {k: round(v[0], 4) for k, v in prof.items()},
            "ms_per_forward": {"in_graph": ms_total / args.steps / nfe, "eager_with_events": fwd_ms},
            "clocks": clocks, "finite": finite,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = staged_reference_forward_time(wl["cfg"], wl["t2i"], 8)
            kind = "reference" if r is not None else "port"
            tf, n = r if r is not None else cpu_port_forward_time(wl["cfg"], wl["t2i"], 8)
            res["cpu_baseline"] = {"value": 8 / (tf * nfe), "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
                                   "sample": f"{n} velocity evaluations at batch 8 on the host cores ("
                                             + ("the staged reference's own libs.uvit.UViT, torch fp32 eager" if kind == "reference"
                                                else "torch fp32 oracle port")
                                             + f"), images/s = 8/(t_fwd*{nfe}) extrapolated from per-forward time"}
        if world == 1 and not args.no_cpu_baseline:
            res["torch_eager_b200"] = torch_eager_on_gpu(wl, dev, nfe)
            res["decoder"] = decoder_line(out[:Bl] if not scales else out[:Bl], dev)
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
