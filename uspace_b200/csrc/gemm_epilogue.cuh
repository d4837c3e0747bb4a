// Fused GEMM epilogues shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.
// A thread owns one accumulator row (TMEM lane) and works on 32 consecutive columns at a time.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace usp {

// exact-erf GELU (nn.GELU() default, libs/timm.py:101-108) as  gelu(v) = max(v, 0) - |v| * Phi(-|v|)  with
// Phi(-a) = 0.5 erfc(a / sqrt2) = 2^q(a): q is the degree-5 minimax fit of log2(0.5 erfc(a / sqrt2)) on [0, 5.5]
// (weighted by a * Phi(-a), the sensitivity of the result); its leading coefficient is negative, so 2^q(a) keeps
// decaying beyond the fitted range and no clamp is needed.  |abs err| <= 1.4e-6 for every finite v (checked on a
// 5e-5 grid over [-60, 60] against scipy's erf in fp64, coefficients rounded to fp32) - far below the 16-bit
// rounding of the stored activation.  1 MUFU + 7 FMA-class ops: the Abramowitz-Stegun form used before
// (2 MUFU + ~14 ops) made fc1's epilogue longer than its main loop (MUFU alone: 2 x 128 x 8 clk x 2 warps per
// scheduler = 4096 clk per tile = the tile's MMA time).
__device__ __forceinline__ float gelu_erf_fast(float v) {
    const float a = fabsf(v);
    float q = fmaf(-0.0003865355101879686f, a, 0.006509846542030573f);
    q = fmaf(q, a, -0.050480134785175323f);
    q = fmaf(q, a, -0.46135035157203674f);
    q = fmaf(q, a, -1.1502251625061035f);
    q = fmaf(q, a, -1.0001084804534912f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
    return fmaf(-a, e, fmaxf(v, 0.f));
}

__device__ __forceinline__ uint32_t pack16(int opd, float a, float b) {
    return opd == OPD_FP16 ? Op16<OPD_FP16>::pack(a, b) : Op16<OPD_BF16>::pack(a, b);
}

struct EpiRow {          // per-row bookkeeping, computed once per tile
    int m;               // global row
    bool ok;             // m < M
    long long qkv_row;   // EPI_QKV: (b*H*L + l) * 64
    float mean, rstd;    // folded LayerNorm (consumer GEMMs)
    float s1, s2;        // partial row statistics being accumulated (producer GEMMs)
};

__device__ __forceinline__ EpiRow epi_row(const GemmArgs& g, int epi, int m) {
    EpiRow r;
    r.m = m;
    r.ok = m < g.M;
    r.qkv_row = 0;
    r.mean = 0.f;
    r.rstd = 1.f;
    r.s1 = r.s2 = 0.f;
    if (g.ln_stats != nullptr && r.ok) {
        // fixed-order sum of the producer's partials: deterministic (no atomics anywhere)
        const float2* sp = reinterpret_cast<const float2*>(g.ln_stats) + static_cast<long long>(m) * g.ln_np;
        float a = 0.f, b = 0.f;
        for (int i = 0; i < g.ln_np; ++i) {
            const float2 p = __ldg(sp + i);
            a += p.x;
            b += p.y;
        }
        r.mean = a * g.ln_inv_d;
        const float var = fmaxf(b * g.ln_inv_d - r.mean * r.mean, 0.f);
        r.rstd = rsqrtf(var + 1e-5f);
        if (g.ln_flag != nullptr && r.mean * r.mean > 16.f * (var + 1e-5f)) atomicOr(g.ln_flag, 2);
    }
    if (epi == EPI_QKV) {
        const int b = m / g.L;
        const int l = m - b * g.L;
        r.qkv_row = (static_cast<long long>(b) * g.H * g.L + l) * 64;
    }
    return r;
}

// residual prefetch for EPI_BIAS_RESID: 32 fp32 = 8 x float4 of row m, columns [n, n+32)
__device__ __forceinline__ void epi_load_resid(const GemmArgs& g, const EpiRow& row, int n, float4* buf) {
    if (row.ok) {
        const float4* rp = reinterpret_cast<const float4*>(g.resid + static_cast<long long>(row.m) * g.N + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) buf[j] = rp[j];
    }
}

// the arithmetic of one 32-column chunk: accumulator -> value to store
template <int EPI>
__device__ __forceinline__ void epi_math(const GemmArgs& g, EpiRow& row, int n, const uint32_t* r,
                                         const float4* resid, float* v) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (g.ln_stats != nullptr && g.ln_c == nullptr) {
        // centred folded weights (zero row sums): the mean term is gone, rstd and the folded bias remain
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(g.ln_d + n + j));
            v[j] = fmaf(row.rstd, v[j], d4.x);
            v[j + 1] = fmaf(row.rstd, v[j + 1], d4.y);
            v[j + 2] = fmaf(row.rstd, v[j + 2], d4.z);
            v[j + 3] = fmaf(row.rstd, v[j + 3], d4.w);
        }
    } else if (g.ln_stats != nullptr) {
        const float nm = -row.mean;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(g.ln_c + n + j));
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(g.ln_d + n + j));
            v[j] = fmaf(row.rstd, fmaf(nm, c4.x, v[j]), d4.x);
            v[j + 1] = fmaf(row.rstd, fmaf(nm, c4.y, v[j + 1]), d4.y);
            v[j + 2] = fmaf(row.rstd, fmaf(nm, c4.z, v[j + 2]), d4.z);
            v[j + 3] = fmaf(row.rstd, fmaf(nm, c4.w, v[j + 3]), d4.w);
        }
    } else if (g.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
    }
    if (EPI == EPI_BIAS_GELU && !(g.diag & 8)) {   // diag 8: skip the GELU arithmetic (diagnostic)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
    }
    if (EPI == EPI_BIAS_RESID) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[4 * j] += resid[j].x; v[4 * j + 1] += resid[j].y; v[4 * j + 2] += resid[j].z; v[4 * j + 3] += resid[j].w;
        }
    }
    if ((EPI == EPI_BIAS_RESID || EPI == EPI_BIAS_F32) && g.stats_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            row.s1 += v[j];
            row.s2 = fmaf(v[j], v[j], row.s2);
        }
    }
}

template <int EPI>
__device__ __forceinline__ void epi_chunk(const GemmArgs& g, EpiRow& row, int n, const uint32_t* r,
                                          const float4* resid) {
    if (!row.ok) return;
    float v[32];
    epi_math<EPI>(g, row, n, r, resid, v);
    if (EPI == EPI_BIAS_RESID || EPI == EPI_BIAS_F32) {
        if (g.out32 != nullptr && !(g.diag & 4)) {   // diag 4: skip the fp32 store (diagnostic)
            float* op = g.out32 + static_cast<long long>(row.m) * g.N + n;
#pragma unroll
            for (int j = 0; j < 4; ++j) st_global_v8_f32(op + 8 * j, v + 8 * j);
        }
    }
    uint16_t* o16 = nullptr;
    if (EPI == EPI_QKV) {
        const int Dm = g.H * 64;
        const int which = n / Dm;
        const int rem = n - which * Dm;
        o16 = reinterpret_cast<uint16_t*>(g.out16) + which * g.qkv_stride + row.qkv_row +
              static_cast<long long>(rem >> 6) * g.L * 64 + (rem & 63);
    } else if (g.out16 != nullptr) {
        o16 = reinterpret_cast<uint16_t*>(g.out16) + static_cast<long long>(row.m) * g.N + n;
    }
    if (o16 != nullptr && !(g.diag & 16)) {        // diag 16: skip the 16-bit store (diagnostic)
        uint32_t u[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) u[j] = pack16(g.opd, v[2 * j], v[2 * j + 1]);
        st_global_v8_b32(o16, u);
        st_global_v8_b32(o16 + 16, u + 8);
    }
}

}  // namespace usp

namespace usp {
// after a tile: publish this thread's (row, 128-column group) partial statistics
__device__ __forceinline__ void epi_store_stats(const GemmArgs& g, const EpiRow& row, int group) {
    if (g.stats_out != nullptr && row.ok)
        reinterpret_cast<float2*>(g.stats_out)[static_cast<long long>(row.m) * (g.N / 128) + group] =
            make_float2(row.s1, row.s2);
}
}  // namespace usp
