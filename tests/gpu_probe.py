"""Developer probe run under gpurun: prints error metrics and timings per section (no asserts).
usage: python tests/gpu_probe.py <gemm|attn|ln|fwd|sample|perf> [opd]
"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uspace_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
TD = {"fp16": torch.float16, "bf16": torch.bfloat16}


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def gemm_case(lib, opd, epi, M, N, K, K0=None, L=257, H=None):
    K0 = K0 or K
    td = TD[opd]
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev).to(td)
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).to(td)
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev)
    ref = a.float() @ w.float().T
    a0 = a[:, :K0].contiguous()
    a1 = a[:, K0:].contiguous() if K0 < K else None
    out32 = torch.zeros(M, N, device=dev)
    out16 = torch.zeros(M, N, device=dev, dtype=td)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    e = _lib.EPI[epi]
    if epi == "qkv":
        D = N // 3
        H = D // 64
        rc = lib.usp_op_gemm(e, P(a0), P(a1), P(w), None, None, None, P(out16), M, N, K, K0, L, H, _lib.OPERAND[opd], s)
        torch.cuda.synchronize()
        B = M // L
        got = out16.view(3, B, H, L, 64).float()
        want = ref.view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
        return rc, rel(got, want)
    if epi == "bias_gelu":
        rc = lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), None, None, P(out16), M, N, K, K0, L, 1, _lib.OPERAND[opd], s)
        torch.cuda.synchronize()
        return rc, rel(out16.float(), torch.nn.functional.gelu(ref + bias))
    if epi == "bias_resid":
        rc = lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), P(resid), P(out32), P(out16), M, N, K, K0, L, 1,
                             _lib.OPERAND[opd], s)
        torch.cuda.synchronize()
        want = ref + bias + resid
        return rc, max(rel(out32, want), rel(out16.float(), want) / 50)
    rc = lib.usp_op_gemm(e, P(a0), P(a1), P(w), P(bias), None, P(out32), None, M, N, K, K0, L, 1, _lib.OPERAND[opd], s)
    torch.cuda.synchronize()
    return rc, rel(out32, ref + bias)


def sec_gemm(lib, opd):
    cases = [("bias_f32", 128, 256, 64), ("bias_f32", 128, 256, 256), ("bias_f32", 514, 1024, 1024),
             ("bias_f32", 514, 1024, 2048, 1024), ("bias_f32", 514, 128, 64), ("bias_f32", 300, 384, 128),
             ("qkv", 514, 3072, 1024), ("qkv", 514, 1536, 512), ("bias_gelu", 514, 4096, 1024),
             ("bias_resid", 514, 1024, 4096), ("bias_resid", 16448, 1024, 1024),
             ("bias_f32", 16448, 1024, 2048, 1024), ("qkv", 16448, 3072, 1024), ("bias_gelu", 42752, 4096, 1024)]
    for c in cases:
        try:
            rc, err = gemm_case(lib, opd, *c)
            print(f"gemm {opd} {c}: rc={rc} rel_err={err:.3e}", flush=True)
        except Exception as ex:
            print(f"gemm {opd} {c}: EXC {ex}", flush=True)
            break


def sec_attn(lib, opd):
    td = TD[opd]
    for (B, H, L) in [(1, 1, 128), (1, 1, 64), (1, 2, 256), (2, 8, 257), (2, 16, 334), (3, 4, 258), (1, 1, 384), (1, 1, 17)]:
        g = torch.Generator(device="cpu").manual_seed(B * 1000 + L)
        q, k, v = (torch.randn(B * H, L, 64, generator=g).to(dev).to(td) for _ in range(3))
        out = torch.zeros(B * L, H * 64, device=dev, dtype=td)
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        try:
            rc = lib.usp_op_attention(P(q), P(k), P(v), P(out), B, H, L, _lib.OPERAND[opd], s)
            torch.cuda.synchronize()
            ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
            want = ref.view(B, H, L, 64).permute(0, 2, 1, 3).reshape(B * L, H * 64)
            print(f"attn {opd} B{B} H{H} L{L}: rc={rc} rel_err={rel(out.float(), want):.3e} "
                  f"max_abs={(out.float() - want).abs().max().item():.3e}", flush=True)
        except Exception as ex:
            print(f"attn {opd} B{B} H{H} L{L}: EXC {ex}", flush=True)
            break


def sec_ln(lib, opd):
    td = TD[opd]
    for (M, D) in [(514, 1024), (16448, 1024), (514, 512), (7, 256)]:
        x = torch.randn(M, D, device=dev) * 3 + 1
        g = torch.randn(D, device=dev)
        b = torch.randn(D, device=dev)
        out = torch.zeros(M, D, device=dev, dtype=td)
        rc = lib.usp_op_layernorm(P(x), P(g), P(b), P(out), M, D, _lib.OPERAND[opd], C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        want = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
        print(f"ln {opd} M{M} D{D}: rc={rc} rel_err={rel(out.float(), want):.3e}", flush=True)


CFG_SMALL = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=512, depth=16, num_heads=8, mlp_ratio=4,
                 qkv_bias=False, mlp_time_embed=False, num_classes=-1, use_checkpoint=False)
CFG_L = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=20, num_heads=16, mlp_ratio=4,
             qkv_bias=False, mlp_time_embed=False, num_classes=-1, use_checkpoint=False)
CFG_L_T2I = dict(img_size=32, patch_size=2, in_chans=4, embed_dim=1024, depth=20, num_heads=16, mlp_ratio=4,
                 qkv_bias=False, mlp_time_embed=False, clip_dim=768, num_clip_token=77, use_checkpoint=False)


def build(cfg, opd, t2i=False):
    from uspace_b200.uvit import UViT, UViTT2I
    torch.manual_seed(0)
    m = (UViTT2I if t2i else UViT)(**cfg).eval()
    m.operand_dtype = opd
    m.fuse_layernorm = os.environ.get("USP_FUSE", "1") == "1"    # the default: LayerNorm folded into the GEMMs
    return m


def sec_fwd(lib, opd):
    from oracle import uvit_oracle as O
    for name, cfg, t2i, B in [("small16", CFG_SMALL, False, 2), ("L", CFG_L, False, 3), ("L_t2i", CFG_L_T2I, True, 2)]:
        m = build(cfg, opd, t2i)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        g = torch.Generator().manual_seed(1230)
        x = torch.randn(B, 4, 32, 32, generator=g)
        t = torch.rand(B, generator=g)
        ctx = torch.randn(B, 77, 768, generator=g) if t2i else None
        ref64 = O.uvit_forward(sd, cfg, x.double(), t.double(), context=None if ctx is None else ctx.double())
        ref32 = O.uvit_forward(sd, cfg, x, t, context=ctx)
        m = m.to(dev)
        with torch.no_grad():
            out = (m(x.to(dev), t.to(dev), context=ctx.to(dev)) if t2i else m(x.to(dev), t.to(dev)))[0]
        torch.cuda.synchronize()
        out = out.cpu()
        print(f"fwd {name} {opd} B{B}: rel(cuda,fp64)={rel(out, ref64):.3e} rel(fp32oracle,fp64)={rel(ref32, ref64):.3e} "
              f"maxabs={(out.double() - ref64).abs().max().item():.3e} refmax={ref64.abs().max().item():.3f} "
              f"nan={torch.isnan(out).any().item()} kernels={m.engine().kernels_per_forward()}", flush=True)
        del m


def sec_sample(lib, opd):
    from oracle import uvit_oracle as O
    cfg = CFG_SMALL
    m = build(cfg, opd)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1230)
    z = torch.randn(2, 4, 32, 32, generator=g)
    m = m.to(dev)
    eng = m.engine()
    for method, h in [("euler", 0.1), ("heun", 0.2)]:
        ref = O.sample(sd, cfg, z.double(), 0.0, 1.0, h, method)
        got = eng.sample(z.to(dev), 0.0, 1.0, h, method).cpu()
        print(f"sample small16 {opd} {method} h={h}: rel={rel(got, ref):.3e}", flush=True)
    delta = 0.1 * torch.randn(11, 4, 32, 32, generator=g)
    for loc in ["tail", "head"]:
        ref = O.sample(sd, cfg, z.double(), 0.0, 1.0, 0.1, "euler", delta_table=delta.double(), write_scale=1.5, t_edit=0.4, edit_loc=loc)
        got = eng.sample(z.to(dev), 0.0, 1.0, 0.1, "euler", delta_table=delta, write_scale=1.5, t_edit=0.4, edit_loc=loc).cpu()
        print(f"sample small16 {opd} euler edit={loc}: rel={rel(got, ref):.3e}", flush=True)
    ref = O.sample(sd, cfg, z.double(), 1.0, 0.0, 0.1, "euler")
    got = eng.sample(z.to(dev), 1.0, 0.0, 0.1, "euler").cpu()
    print(f"encode small16 {opd} euler: rel={rel(got, ref):.3e}", flush=True)
    zh = z.clone().pin_memory()
    eng.sample_host(zh, 0.0, 1.0, 0.1, "euler")
    ref = O.sample(sd, cfg, z.double(), 0.0, 1.0, 0.1, "euler")
    print(f"sample_host small16 {opd}: rel={rel(zh, ref):.3e}", flush=True)


def sec_perf(lib, opd):
    m = build(CFG_L, opd).to(dev)
    eng = m.engine()
    flops = eng.flops_per_forward()
    for B in [64, 256]:
        z = torch.randn(B, 4, 32, 32, device=dev)
        t = torch.full((B,), 0.5, device=dev)
        for _ in range(3):
            eng.forward(z, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            eng.forward(z, t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"perf L {opd} B{B}: fwd {ms:.3f} ms -> {flops * B / ms / 1e9:.1f} TFLOP/s, {B / (ms * 50) * 1e3:.1f} img/s@50", flush=True)
        eng.sample(z, 0.0, 1.0, 0.02, "euler")
        torch.cuda.synchronize()
        t0 = time.time()
        eng.sample(z, 0.0, 1.0, 0.02, "euler")
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"perf L {opd} B{B}: 50-step sample {dt * 1e3:.1f} ms ({eng.last_ms():.1f} dev) -> {B / dt:.1f} img/s, "
              f"{flops * B * 50 / dt / 1e12:.1f} TFLOP/s", flush=True)


def sec_prof(lib, opd):
    """Per-kernel-class device time of one U-ViT-L velocity evaluation (events between launches)."""
    for cfg, t2i, B in [(CFG_L, False, 64), (CFG_L_T2I, True, 128)]:
        m = build(cfg, opd, t2i).to(dev)
        eng = m.engine()
        z = torch.randn(B, 4, 32, 32, device=dev)
        t = torch.full((B,), 0.5, device=dev)
        ctx = torch.randn(B, 77, 768, device=dev) if t2i else None
        for _ in range(3):
            p = eng.profile_forward(z, t, context=ctx)
        tot = sum(v[0] for v in p.values())
        print(f"prof {'t2i' if t2i else 'uncond'} B{B} total {tot:.3f} ms: " +
              " ".join(f"{k}={v[0]:.3f}/{v[1]}" for k, v in p.items()), flush=True)
        del m, eng


def sec_gemmperf(lib, opd):
    """Stand-alone timing of the five U-ViT-L GEMM shapes (M = 64*257), 20 back-to-back launches each."""
    td = TD[opd]
    M, D = 16448, 1024
    shapes = [("qkv", 3 * D, D, D), ("bias_resid", D, D, D), ("bias_gelu", 4 * D, D, D),
              ("bias_resid", D, 4 * D, 4 * D), ("bias_f32", D, 2 * D, D)]
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for epi, N, K, K0 in shapes:
        a0 = torch.randn(M, K0, device=dev).to(td)
        a1 = torch.randn(M, K - K0, device=dev).to(td) if K0 < K else None
        w = (torch.randn(N, K, device=dev) * 0.05).to(td)
        bias = torch.randn(N, device=dev)
        x32 = torch.randn(M, N, device=dev) if epi in ("bias_resid", "bias_f32") else None
        o16 = torch.empty(M, N, device=dev, dtype=td) if epi in ("qkv", "bias_gelu", "bias_resid") else None
        H = D // 64 if epi == "qkv" else 1

        def run():
            return lib.usp_op_gemm(_lib.EPI[epi], P(a0), P(a1), P(w), None if epi == "qkv" else P(bias),
                                   P(x32) if epi == "bias_resid" else None, P(x32), P(o16), M, N, K, K0, 257, H,
                                   _lib.OPERAND[opd], s)
        for _ in range(3):
            rc = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"gemmperf {epi:10s} N={N} K={K}: rc={rc} {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)


def sec_gemmsustained(lib, opd):
    """Power-capped regime: each U-ViT-L GEMM shape back to back for ~2 s (0.5 s warm-up + 1.5 s timed), this library's
    fused-epilogue kernel against torch.matmul (cuBLAS, plain 16-bit output, no epilogue) on the same operands."""
    td = TD[opd]
    M, D = 16448, 1024
    shapes = [("qkv", 3 * D, D, D), ("bias_resid", D, D, D), ("bias_gelu", 4 * D, D, D),
              ("bias_resid", D, 4 * D, 4 * D), ("bias_f32", D, 2 * D, D)]
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def sustained(fn, est_us):
        n_warm, n = max(int(0.5e6 / est_us), 10), max(int(1.5e6 / est_us), 20)
        for _ in range(n_warm):
            fn()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    for epi, N, K, K0 in shapes:
        a0 = torch.randn(M, K0, device=dev).to(td)
        a1 = torch.randn(M, K - K0, device=dev).to(td) if K0 < K else None
        a_full = a0 if a1 is None else torch.cat([a0, a1], 1).contiguous()
        w = (torch.randn(N, K, device=dev) * 0.05).to(td)
        bias = torch.randn(N, device=dev)
        x32 = torch.randn(M, N, device=dev) if epi in ("bias_resid", "bias_f32") else None
        o16 = torch.empty(M, N, device=dev, dtype=td)
        H = D // 64 if epi == "qkv" else 1
        mine = lambda: lib.usp_op_gemm(_lib.EPI[epi], P(a0), P(a1), P(w), None if epi == "qkv" else P(bias),
                                       P(x32) if epi == "bias_resid" else None, P(x32),
                                       P(o16) if epi in ("qkv", "bias_gelu", "bias_resid") else None, M, N, K, K0, 257, H,
                                       _lib.OPERAND[opd], s)
        wt = w.t()
        ref = lambda: torch.matmul(a_full, wt, out=o16)
        fl = 2.0 * M * N * K
        us_m = sustained(mine, 100.0)
        us_r = sustained(ref, 100.0)
        print(f"gemmsustained {epi:10s} N={N} K={K}: this {us_m:7.1f} us {fl / us_m / 1e6:7.1f} TFLOP/s | "
              f"torch.matmul {us_r:7.1f} us {fl / us_r / 1e6:7.1f} TFLOP/s", flush=True)


def sec_attnperf(lib, opd):
    """Stand-alone timing of the attention kernel at the U-ViT-L shapes."""
    td = TD[opd]
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, H, L) in [(64, 16, 256), (64, 16, 257), (128, 16, 334), (16, 16, 1025)] if not os.environ.get("USP_ATTN_TRACE") else [(64, 16, 256)]:
        q, k, v = (torch.randn(B * H, L, 64, device=dev).to(td) for _ in range(3))
        out = torch.zeros(B * L, H * 64, device=dev, dtype=td)
        for _ in range(3):
            rc = lib.usp_op_attention(P(q), P(k), P(v), P(out), B, H, L, _lib.OPERAND[opd], s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            lib.usp_op_attention(P(q), P(k), P(v), P(out), B, H, L, _lib.OPERAND[opd], s)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"attnperf B{B} H{H} L{L}: rc={rc} {us:7.1f} us  {4.0 * B * H * L * L * 64 / us / 1e6:7.1f} TFLOP/s", flush=True)
        if os.environ.get("USP_ATTN_TRACE"):
            continue
        # the library kernel the reference would run (libs/uvit.py:95), same operands in its [B, H, L, 64] layout
        q4, k4, v4 = (t.view(B, H, L, 64) for t in (q, k, v))
        for _ in range(3):
            torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"attnperf B{B} H{H} L{L}: torch SDPA {us:7.1f} us  {4.0 * B * H * L * L * 64 / us / 1e6:7.1f} TFLOP/s", flush=True)


def sec_one(lib, opd):
    """Two eager velocity evaluations of U-ViT-L at batch 64 (profile the second one under ncu)."""
    m = build(CFG_L, opd).to(dev)
    eng = m.engine()
    z = torch.randn(64, 4, 32, 32, device=dev)
    t = torch.full((64,), 0.5, device=dev)
    for _ in range(2):
        eng.forward(z, t)
    torch.cuda.synchronize()
    print("kernels per forward", eng.kernels_per_forward())


def sec_vae(lib, opd):
    """Two decodes of 8 latents through the autoencoder (profile the second one under ncu)."""
    from uspace_b200.autoencoder import get_model
    torch.manual_seed(0)
    m = get_model().to(dev)
    z = 0.7 * torch.randn(8, 4, 32, 32, device=dev)
    for _ in range(2):
        img = m.decode(z)
    torch.cuda.synchronize()
    print("decoded", tuple(img.shape), bool(torch.isfinite(img).all()))


def sec_vaeprec(lib, opd):
    """Error against the reference goldens and decode / encode time per image for both operand precisions."""
    import numpy as np
    from tests.golden.cases import vae_enc_state_dict, vae_images, vae_latents, vae_state_dict
    from uspace_b200.autoencoder import get_model
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    rel = lambda a, b: ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
    for prec in ("fp16", "fp16x3"):
        m = get_model(precision=prec)
        m.load_state_dict({**vae_state_dict(), **vae_enc_state_dict()})
        m = m.to(dev)
        for name in ("vae_small", "vae_full"):
            want = torch.from_numpy(np.load(os.path.join(gd, name + ".npz"))["decode"])
            print(prec, name, "decode rel err %.3e" % rel(m.decode(vae_latents(name).to(dev)), want), flush=True)
        for name in ("vae_enc_small", "vae_enc_full"):
            want = torch.from_numpy(np.load(os.path.join(gd, name + ".npz"))["moments"])
            print(prec, name, "moments rel err %.3e" % rel(m.encode_moments(vae_images(name).to(dev)), want), flush=True)
        z = 0.7 * torch.randn(64, 4, 32, 32, device=dev)
        x = torch.rand(64, 3, 256, 256, device=dev) * 2 - 1
        for fn, inp, what in ((m.decode, z, "decode"), (m.encode_moments, x, "encode")):
            fn(inp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(inp)
            e1.record()
            torch.cuda.synchronize()
            print(prec, what, "%.3f ms per image (64 images, 256^2)" % (e0.elapsed_time(e1) / 64), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    lib = _lib.load()
    sec = sys.argv[1]
    opd = sys.argv[2] if len(sys.argv) > 2 else "fp16"
    print(f"== {sec} {opd} on {torch.cuda.get_device_name(0)}", flush=True)
    globals()["sec_" + sec](lib, opd)
