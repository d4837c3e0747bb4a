// HBM-bound glue kernels around the tensor-core GEMMs: token embedding, LayerNorm -> 16-bit operand,
// output head (final LN + decoder_pred), unpatchify + 3x3 conv + edit + ODE update, dtype conversion,
// and the device-side ODE step bookkeeping that lets one captured CUDA graph serve every step.
#include "common.cuh"
#include "kernels.h"

namespace usp {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// embed: PatchEmbed.forward (libs/uvit.py:175-179) as an explicit (C,p1,p2) gather + [P]x[P,D] product,
//        timestep_embedding (libs/uvit.py:26-46; t used raw, cos half then sin half),
//        token order [label?, time, ctx..., patches] (libs/uvit.py:320-327, libs/uvit_t2i.py:320-324), + pos_embed.
//        Optional "head" edit x + delta[t]*write_scale (libs/dissection.py:157, libs/uvit.py:313-314).
// grid = (ceil(L / EMB_TOK), B), block = 256.  Each block handles EMB_TOK consecutive tokens of one sample so
// that the [D, P] projection weights are read once per block (registers), not once per token.
// ------------------------------------------------------------------------------------------------
constexpr int EMB_TOK = 16;
constexpr int EMB_MAXP = 64;   // C*p*p upper bound

template <bool EXTRA>   // EXTRA: also emit the 16-bit copy and the partial row statistics (folded-LayerNorm path)
__global__ void __launch_bounds__(256) embed_kernel(const EmbedArgs a) {
    pdl_wait();
    pdl_launch();
    const int l0 = blockIdx.x * EMB_TOK;
    const int b = blockIdx.y;
    const int D = a.D;
    const int t_tok = a.has_label ? 1 : 0;
    const int first_patch = t_tok + 1 + a.n_ctx;
    const int p = a.p, C = a.C, S = a.S;
    const int P = C * p * p;
    const int gw = S / p;
    const int l1 = min(l0 + EMB_TOK, a.L);

    __shared__ __align__(16) float feat[EMB_TOK][EMB_MAXP];
    // gather the (C,p1,p2) features of every patch token in this block
    for (int i = threadIdx.x; i < EMB_TOK * P; i += blockDim.x) {
        const int tk = i / P, f = i % P;
        const int l = l0 + tk;
        if (l >= first_patch && l < a.L) {
            const int pi = l - first_patch;
            const int ph = pi / gw, pw = pi % gw;
            const int c = f / (p * p);
            const int p1 = (f / p) % p;
            const int p2 = f % p;
            const int idx = (c * S + ph * p + p1) * S + pw * p + p2;
            float v = a.x[static_cast<long long>(b) * C * S * S + idx];
            // sampling: row / scale of the current step; plain forward with a hook (usp_forward_hook): row 0
            const int didx = a.st ? a.st->didx : 0;
            if (a.delta != nullptr) {
                float sc = a.st ? a.st->edit : a.hook_scale;
                if (sc != 0.f && a.sscale != nullptr) sc = a.sscale[b];
                if (sc != 0.f) v += a.delta[static_cast<long long>(didx) * C * S * S + idx] * sc;
            }
            if (a.trace != nullptr && didx >= 0)   // dissect_name="read" (libs/dissection.py:126-136)
                a.trace[(static_cast<long long>(didx) * a.B + b) * C * S * S + idx] = v;
            feat[tk][f] = v;
        }
    }
    __syncthreads();

    float st1[EMB_TOK], st2[EMB_TOK];   // this thread's partial row statistics (folded-LayerNorm path)
#pragma unroll
    for (int i = 0; i < EMB_TOK; ++i) st1[i] = st2[i] = 0.f;
    const bool fast = P == 16 && (D & 3) == 0;
    if (fast) {
        // every reference config (patch 2, 4 channels): four consecutive output channels per thread - 128-bit position
        // loads and stores, the four W rows in registers, one set of broadcast feature reads for four outputs.  The
        // arithmetic per output (fma chain over f = 0..15, + bias, + pos) is the scalar path's, so the bits are too.
        for (int d0 = threadIdx.x * 4; d0 < D; d0 += blockDim.x * 4) {
            float wr[4][16];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.w + static_cast<long long>(d0 + k) * 16) + q);
                    wr[k][4 * q] = w4.x; wr[k][4 * q + 1] = w4.y; wr[k][4 * q + 2] = w4.z; wr[k][4 * q + 3] = w4.w;
                }
            const float4 bias4 = *reinterpret_cast<const float4*>(a.bias + d0);
#pragma unroll 4
            for (int i = 0; i < EMB_TOK; ++i) {
                const int l = l0 + i;
                if (l >= l1) break;
                const float4 pos4 = __ldg(reinterpret_cast<const float4*>(a.pos + static_cast<long long>(l) * D + d0));
                float v[4];
                if (l >= first_patch) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 f4 = *reinterpret_cast<const float4*>(&feat[i][4 * q]);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            acc[k] = fmaf(wr[k][4 * q], f4.x, acc[k]);
                            acc[k] = fmaf(wr[k][4 * q + 1], f4.y, acc[k]);
                            acc[k] = fmaf(wr[k][4 * q + 2], f4.z, acc[k]);
                            acc[k] = fmaf(wr[k][4 * q + 3], f4.w, acc[k]);
                        }
                    }
                    v[0] = acc[0] + bias4.x; v[1] = acc[1] + bias4.y; v[2] = acc[2] + bias4.z; v[3] = acc[3] + bias4.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int d = d0 + k;
                        if (a.has_label && l == 0) {
                            v[k] = a.label[a.y[b] * D + d];
                        } else if (l == t_tok && a.ttok != nullptr) {
                            v[k] = a.ttok[static_cast<long long>(a.st ? 0 : b) * D + d];
                        } else if (l == t_tok) {
                            const float t = a.st ? a.st->t : a.tvec[b];
                            const int half = D / 2;
                            v[k] = 0.f;
                            if (d < half) v[k] = cosf(t * a.freqs[d]);
                            else if (d < 2 * half) v[k] = sinf(t * a.freqs[d - half]);
                        } else {
                            v[k] = a.ctxemb[(static_cast<long long>(b) * a.n_ctx + (l - t_tok - 1)) * D + d];
                        }
                    }
                }
                const float4 o4 = make_float4(v[0] + pos4.x, v[1] + pos4.y, v[2] + pos4.z, v[3] + pos4.w);
                *reinterpret_cast<float4*>(a.out32 + (static_cast<long long>(b) * a.L + l) * D + d0) = o4;
                if (EXTRA && a.out16 != nullptr) {
                    uint2 u;
                    u.x = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(o4.x, o4.y) : Op16<OPD_BF16>::pack(o4.x, o4.y);
                    u.y = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(o4.z, o4.w) : Op16<OPD_BF16>::pack(o4.z, o4.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.out16) +
                                              (static_cast<long long>(b) * a.L + l) * D + d0) = u;
                }
                if (EXTRA) {
                    st1[i] += (o4.x + o4.y) + (o4.z + o4.w);
                    st2[i] += (o4.x * o4.x + o4.y * o4.y) + (o4.z * o4.z + o4.w * o4.w);
                }
            }
        }
    }
    for (int d = threadIdx.x; !fast && d < D; d += blockDim.x) {
        const float* wrow = a.w + static_cast<long long>(d) * P;
        const float bias = a.bias[d];
        // P == 16 (patch 2, 4 channels: every reference config): keep this row of W in registers, read once as
        // 4 x 128-bit (scalar per-token re-reads cost 32 L1 sectors per warp instruction and dominated the kernel)
        float wr[16];
        const bool p16 = (P == 16);
        if (p16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow) + q);
                wr[4 * q] = w4.x; wr[4 * q + 1] = w4.y; wr[4 * q + 2] = w4.z; wr[4 * q + 3] = w4.w;
            }
        }
#pragma unroll
        for (int i = 0; i < EMB_TOK; ++i) {
            const int l = l0 + i;
            if (l >= l1) break;
            float* out = a.out32 + (static_cast<long long>(b) * a.L + l) * D;
            const float pos = a.pos[static_cast<long long>(l) * D + d];
            float v;
            if (a.has_label && l == 0) {
                v = a.label[a.y[b] * D + d];
            } else if (l == t_tok && a.ttok != nullptr) {
                v = a.ttok[static_cast<long long>(a.st ? 0 : b) * D + d];    // mlp_time_embed: from launch_time_mlp
            } else if (l == t_tok) {
                const float t = a.st ? a.st->t : a.tvec[b];
                const int half = D / 2;
                v = 0.f;
                if (d < half) v = cosf(t * a.freqs[d]);
                else if (d < 2 * half) v = sinf(t * a.freqs[d - half]);
            } else if (l < first_patch) {
                v = a.ctxemb[(static_cast<long long>(b) * a.n_ctx + (l - t_tok - 1)) * D + d];
            } else {
                float acc = 0.f;
                if (p16) {
                    // 4 x 128-bit broadcast reads instead of 16 scalar ones: the scalar version was LDS-issue bound
                    // (8192 warp-level LDS per block, ~40 us of the kernel's 83)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 f4 = *reinterpret_cast<const float4*>(&feat[i][4 * q]);
                        acc = fmaf(wr[4 * q], f4.x, acc);
                        acc = fmaf(wr[4 * q + 1], f4.y, acc);
                        acc = fmaf(wr[4 * q + 2], f4.z, acc);
                        acc = fmaf(wr[4 * q + 3], f4.w, acc);
                    }
                } else {
                    for (int f = 0; f < P; ++f) acc = fmaf(__ldg(wrow + f), feat[i][f], acc);
                }
                v = acc + bias;
            }
            v += pos;
            out[d] = v;
            if (EXTRA && a.out16 != nullptr)
                reinterpret_cast<uint16_t*>(a.out16)[(static_cast<long long>(b) * a.L + l) * D + d] =
                    a.opd == OPD_FP16 ? Op16<OPD_FP16>::one(v) : Op16<OPD_BF16>::one(v);
            if (EXTRA) {
                st1[i] += v;
                st2[i] = fmaf(v, v, st2[i]);
            }
        }
    }
    if (EXTRA && a.stats != nullptr) {
        // one (sum, sum of squares) slot per (row, warp): the consumer adds the 8 slots in fixed order
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < EMB_TOK; ++i) {
            const float s1 = warp_sum(st1[i]), s2 = warp_sum(st2[i]);
            if (lane == 0 && l0 + i < a.L)
                reinterpret_cast<float2*>(a.stats)[(static_cast<long long>(b) * a.L + l0 + i) * 8 + warp] =
                    make_float2(s1, s2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, affine; libs/uvit.py:160-161) fp32 in -> 16-bit GEMM operand out. One warp per row,
// the row lives in registers (two-pass mean / biased variance in fp32).
// ------------------------------------------------------------------------------------------------
template <int V4>  // float4 per lane: D = 128 * V4
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ bta, uint16_t* __restrict__ out,
                                                        int M, int opd, int reverse) {
    constexpr int D = 128 * V4;
    pdl_wait();
    pdl_launch();
    // rows are visited from the END: the GEMM that produced x wrote its last rows last, and with ~170 MB passing
    // through the 126 MB L2 during that GEMM those are the rows still resident
    const int row_fwd = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row_fwd >= M) return;
    const int row = reverse ? M - 1 - row_fwd : row_fwd;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
    float4 v[V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        v[i] = xr[i * 32 + lane];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<long long>(row) * D);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + i * 32 + lane);
        const float y0 = (v[i].x - mean) * rstd * gg.x + bb.x;
        const float y1 = (v[i].y - mean) * rstd * gg.y + bb.y;
        const float y2 = (v[i].z - mean) * rstd * gg.z + bb.z;
        const float y3 = (v[i].w - mean) * rstd * gg.w + bb.w;
        uint2 u;
        if (opd == OPD_FP16) {
            u.x = Op16<OPD_FP16>::pack(y0, y1);
            u.y = Op16<OPD_FP16>::pack(y2, y3);
        } else {
            u.x = Op16<OPD_BF16>::pack(y0, y1);
            u.y = Op16<OPD_BF16>::pack(y2, y3);
        }
        orow[i * 32 + lane] = u;
    }
}

// ------------------------------------------------------------------------------------------------
// head: final LayerNorm + decoder_pred Linear(D -> P) on the patch tokens only (libs/uvit.py:342-345), fp32 SIMT
// (P = 16: 0.03 % of the FLOPs, kept exact).  One warp per patch token.
// ------------------------------------------------------------------------------------------------
constexpr int HEAD_TOK = 2;   // patch tokens per warp: every decoder_pred weight vector read feeds both (the 64 KB of
                              // weights per token, re-read from L1 for each of 16 k tokens, bound the kernel)
template <int V4>
__global__ void __launch_bounds__(256) head_kernel(const HeadArgs a) {
    constexpr int D = 128 * V4;
    pdl_wait();
    pdl_launch();
    const int n_patch = a.L - a.extras;
    const int n_tok = a.B * n_patch;
    const int tok0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * HEAD_TOK;
    const int lane = threadIdx.x & 31;
    if (tok0 >= n_tok) return;
    float4 v[HEAD_TOK][V4];
#pragma unroll
    for (int t = 0; t < HEAD_TOK; ++t) {
        const int tok = tok0 + t < n_tok ? tok0 + t : n_tok - 1;   // a trailing odd token is computed twice, stored once
        const int b = tok / n_patch, pi = tok % n_patch;
        const float4* xr =
            reinterpret_cast<const float4*>(a.x32 + (static_cast<long long>(b) * a.L + a.extras + pi) * D);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            v[t][i] = xr[i * 32 + lane];
            s += (v[t][i].x + v[t][i].y) + (v[t][i].z + v[t][i].w);
        }
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float a0 = v[t][i].x - mean, a1 = v[t][i].y - mean, a2 = v[t][i].z - mean, a3 = v[t][i].w - mean;
            q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float4 gg = __ldg(reinterpret_cast<const float4*>(a.ng) + i * 32 + lane);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.nb) + i * 32 + lane);
            v[t][i].x = (v[t][i].x - mean) * rstd * gg.x + bb.x;
            v[t][i].y = (v[t][i].y - mean) * rstd * gg.y + bb.y;
            v[t][i].z = (v[t][i].z - mean) * rstd * gg.z + bb.z;
            v[t][i].w = (v[t][i].w - mean) * rstd * gg.w + bb.w;
        }
    }
    for (int j = 0; j < a.P; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(a.w + static_cast<long long>(j) * D);
        float acc[HEAD_TOK];
#pragma unroll
        for (int t = 0; t < HEAD_TOK; ++t) acc[t] = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float4 w4 = __ldg(wr + i * 32 + lane);
#pragma unroll
            for (int t = 0; t < HEAD_TOK; ++t)
                acc[t] += (v[t][i].x * w4.x + v[t][i].y * w4.y) + (v[t][i].z * w4.z + v[t][i].w * w4.w);
        }
        const float bj = a.bias[j];
#pragma unroll
        for (int t = 0; t < HEAD_TOK; ++t) {
            const float r = warp_sum(acc[t]);
            if (lane == 0 && tok0 + t < n_tok) a.pf[static_cast<long long>(tok0 + t) * a.P + j] = r + bj;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// final: unpatchify with feature order (p1,p2,C) (libs/uvit.py:56-63) + final_layer Conv2d 3x3 pad 1
// (libs/uvit.py:346-347) + optional "tail" edit v + delta[t]*write_scale (libs/uvit.py:349-350) +
// fixed-grid ODE update (torchdiffeq Euler: y1 = y0 + dt*f; Heun: y0 + dt/2*(k1+k2)).
// One thread per output element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_kernel(const FinalArgs a) {
    pdl_wait();
    pdl_launch();
    const int C = a.C, S = a.S, p = a.p;
    const long long n = static_cast<long long>(a.B) * C * S * S;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int xw = static_cast<int>(i % S);
    const int yh = static_cast<int>((i / S) % S);
    const int co = static_cast<int>((i / (static_cast<long long>(S) * S)) % C);
    const int b = static_cast<int>(i / (static_cast<long long>(S) * S * C));
    const int gw = S / p;
    const int P = C * p * p;
    const float* pf = a.pf + static_cast<long long>(b) * gw * gw * P;
    float v;
    if (a.cw != nullptr) {
        v = a.cb[co];
        for (int ci = 0; ci < C; ++ci) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = yh + ky - 1;
                if (yy < 0 || yy >= S) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = xw + kx - 1;
                    if (xx < 0 || xx >= S) continue;
                    const float img = pf[((yy / p) * gw + (xx / p)) * P + ((yy % p) * p + (xx % p)) * C + ci];
                    v = fmaf(a.cw[((co * C + ci) * 3 + ky) * 3 + kx], img, v);
                }
            }
        }
    } else {
        v = pf[((yh / p) * gw + (xw / p)) * P + ((yh % p) * p + (xw % p)) * C + co];
    }
    // an MLP hidden activation beyond the fp16 operand range turns into inf and reaches every output of its sample
    // as NaN (fc2 -> LayerNorm -> attention): one check here sees any overflow of the whole evaluation
    if (a.nonfinite != nullptr && !isfinite(v)) atomicOr(a.nonfinite, 1);
    const int chw = static_cast<int>(i % (static_cast<long long>(C) * S * S));
    const int didx = a.st ? a.st->didx : 0;   // plain forward with a hook (usp_forward_hook): row 0
    if (a.delta != nullptr) {
        float sc = a.st ? a.st->edit : a.hook_scale;
        if (sc != 0.f && a.sscale != nullptr) sc = a.sscale[b];
        if (sc != 0.f) v += a.delta[static_cast<long long>(didx) * C * S * S + chw] * sc;
    }
    if (a.trace != nullptr && didx >= 0) a.trace[static_cast<long long>(didx) * n + i] = v;
    if (a.st == nullptr) {
        a.out[i] = v;
        return;
    }
    if (a.base == nullptr) {   // general Runge-Kutta stage: keep the (direction-signed) derivative only
        a.out[i] = a.m1 * v;
        return;
    }
    const float dt = a.st->dt;
    float inc = a.m1 * v;
    if (a.aux != nullptr) inc += a.m2 * a.aux[i];
    if (a.vstore != nullptr) a.vstore[i] = a.vs_b != 0.f ? fmaf(a.vs_b, a.vstore[i], a.vs_a * v) : a.vs_a * v;
    if (a.acc2 != nullptr) a.acc2[i] = a.a2_b != 0.f ? fmaf(a.a2_b, a.acc2[i], a.a2_a * v) : a.a2_a * v;
    a.out[i] = a.base[i] + dt * inc;
}

__global__ void convert16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n, int opd) {
    pdl_wait();
    pdl_launch();
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride)
        out[i] = opd == OPD_FP16 ? Op16<OPD_FP16>::one(in[i]) : Op16<OPD_BF16>::one(in[i]);
}

// one block per output row n of W[N,K]: W'[n,k] = W[n,k] gamma[k], optionally CENTRED (its row mean removed, so that
// sum_k x[k] W'[n,k] no longer contains the row mean of x and the epilogue needs rstd and d only), rounded to 16 bits;
// c[n] = sum_k of the rounded row (what the mean multiplies: ~0 when centred), d[n] = sum_k beta[k] W[n,k] + bias[n]
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                                      uint16_t* __restrict__ w16, float* __restrict__ c,
                                                      float* __restrict__ d, int K, int opd, int centre) {
    const int n = blockIdx.x;
    __shared__ float sc[8], sd[8], sm[8];
    float wmean = 0.f;
    if (centre) {
        float ms = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) ms += W[static_cast<long long>(n) * K + k] * gamma[k];
        ms = warp_sum(ms);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ms;
        __syncthreads();
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += sm[i];
        wmean = a / static_cast<float>(K);
    }
    float cs = 0.f, ds = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float w = W[static_cast<long long>(n) * K + k];
        const float wg = w * gamma[k] - wmean;
        float wr;
        uint16_t h;
        if (opd == OPD_FP16) {
            h = Op16<OPD_FP16>::one(wg);
            wr = __half2float(*reinterpret_cast<__half*>(&h));
        } else {
            h = Op16<OPD_BF16>::one(wg);
            wr = __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&h));
        }
        w16[static_cast<long long>(n) * K + k] = h;
        cs += wr;
        ds = fmaf(beta[k], w, ds);
    }
    cs = warp_sum(cs);
    ds = warp_sum(ds);
    if ((threadIdx.x & 31) == 0) {
        sc[threadIdx.x >> 5] = cs;
        sd[threadIdx.x >> 5] = ds;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) {
            a += sc[i];
            b += sd[i];
        }
        c[n] = a;
        d[n] = b + (bias != nullptr ? bias[n] : 0.f);
    }
}

__global__ void step_kernel(StepState* st, const float* __restrict__ grid, const unsigned char* __restrict__ mask,
                            const unsigned char* __restrict__ amask, int stage, float frac) {
    pdl_wait();
    pdl_launch();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (stage == 0) {
        const int i = st->next;
        st->cur = i;
        st->next = i + 1;
        st->t = grid[i];
        st->dt = grid[i + 1] - grid[i];
        st->edit = mask[i] ? st->write_scale : 0.f;
        st->didx = i;
        st->attn_on = amask[i];
    } else if (stage == 1) {
        const int i = st->cur;
        st->t = grid[i + 1];
        st->edit = mask[i + 1] ? st->write_scale : 0.f;
        st->didx = i + 1;
        st->attn_on = amask[i + 1];
    } else {
        const int i = st->cur;
        const float t = __fadd_rn(grid[i], __fmul_rn(st->dt, frac));   // t0 + dt * frac in the grid's precision
        st->t = t;
        st->edit = 0.f;
        st->didx = i;
        // float(f"{t:.2f}") <= t_edit with both sides in fp32 (see digit_leq in api.cu)
        const float digit = static_cast<float>(rint(static_cast<double>(t) * 100.0) / 100.0);
        st->attn_on = (st->attn_t_edit >= 0.f && digit <= st->attn_t_edit) ? 1 : 0;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// mlp_time_embed: sinusoid -> Linear(D, 4D) -> SiLU -> Linear(4D, D)   (libs/uvit.py:215-223, :320)
// Three small fp32 launches; one warp per output neuron, the weight row read once and reused for every row of t.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void time_sincos_kernel(const TimeMlpArgs a) {
    const int half = a.D / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.rows * a.D; i += gridDim.x * blockDim.x) {
        const int n = i / a.D, d = i - n * a.D;
        const float t = a.st ? a.st->t : a.tvec[n];
        float v = 0.f;
        if (d < half) v = cosf(t * a.freqs[d]);
        else if (d < 2 * half) v = sinf(t * a.freqs[d - half]);
        a.sincos[i] = v;
    }
}

template <bool SILU>
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int rows, int K, int N) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);     // output neuron of this warp
    const int lane = threadIdx.x & 31;
    if (j >= N) return;
    const float* wr = w + static_cast<long long>(j) * K;
    for (int n0 = 0; n0 < rows; n0 += 4) {                  // four rows of t per pass over the weight row
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane * 4; k < K; k += 128) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (n0 + r < rows) {
                    const float4 x4 = *reinterpret_cast<const float4*>(in + static_cast<long long>(n0 + r) * K + k);
                    acc[r] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[r]))));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
            if (lane == 0 && n0 + r < rows) {
                float v = acc[r] + bias[j];
                if (SILU) v = v / (1.0f + expf(-v));
                out[static_cast<long long>(n0 + r) * N + j] = v;
            }
        }
    }
}

}  // namespace

cudaError_t launch_time_mlp(const TimeMlpArgs& a, cudaStream_t s) {
    if (a.D % 4 != 0) return cudaErrorInvalidValue;
    const int n = a.rows * a.D;
    time_sincos_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
    small_linear_kernel<true><<<(4 * a.D + 7) / 8, 256, 0, s>>>(a.sincos, a.w1, a.b1, a.hidden, a.rows, a.D, 4 * a.D);
    small_linear_kernel<false><<<(a.D + 7) / 8, 256, 0, s>>>(a.hidden, a.w2, a.b2, a.ttok, a.rows, 4 * a.D, a.D);
    return cudaGetLastError();
}

cudaError_t launch_embed(const EmbedArgs& a, cudaStream_t s) {
    if (a.C * a.p * a.p > 64) return cudaErrorInvalidValue;
    if (a.out16 != nullptr || a.stats != nullptr)
        return launch_pdl(embed_kernel<true>, dim3((a.L + EMB_TOK - 1) / EMB_TOK, a.B), dim3(256), 0, s, a);
    return launch_pdl(embed_kernel<false>, dim3((a.L + EMB_TOK - 1) / EMB_TOK, a.B), dim3(256), 0, s, a);
}

cudaError_t launch_layernorm(const float* x, const float* g, const float* b, void* out16, int M, int D, int opd,
                             cudaStream_t s) {
    static int rev = -1;
    if (rev < 0) {   // USP_LN_REVERSE=0 restores ascending row order (A/B comparison)
        const char* e = getenv("USP_LN_REVERSE");
        rev = e ? atoi(e) : 1;
    }
    const int rows_per_block = 8;
    const int grid = (M + rows_per_block - 1) / rows_per_block;
    uint16_t* o = reinterpret_cast<uint16_t*>(out16);
    switch (D) {
        case 256: return launch_pdl(layernorm_kernel<2>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        case 384: return launch_pdl(layernorm_kernel<3>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        case 512: return launch_pdl(layernorm_kernel<4>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        case 768: return launch_pdl(layernorm_kernel<6>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        case 1024: return launch_pdl(layernorm_kernel<8>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        case 1536: return launch_pdl(layernorm_kernel<12>, dim3(grid), dim3(256), 0, s, x, g, b, o, M, opd, rev);
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_head(const HeadArgs& a, cudaStream_t s) {
    const int n_tok = a.B * (a.L - a.extras);
    const int grid = (n_tok + 8 * HEAD_TOK - 1) / (8 * HEAD_TOK);
    switch (a.D) {
        case 256: return launch_pdl(head_kernel<2>, dim3(grid), dim3(256), 0, s, a);
        case 384: return launch_pdl(head_kernel<3>, dim3(grid), dim3(256), 0, s, a);
        case 512: return launch_pdl(head_kernel<4>, dim3(grid), dim3(256), 0, s, a);
        case 768: return launch_pdl(head_kernel<6>, dim3(grid), dim3(256), 0, s, a);
        case 1024: return launch_pdl(head_kernel<8>, dim3(grid), dim3(256), 0, s, a);
        case 1536: return launch_pdl(head_kernel<12>, dim3(grid), dim3(256), 0, s, a);
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_final(const FinalArgs& a, cudaStream_t s) {
    const long long n = static_cast<long long>(a.B) * a.C * a.S * a.S;
    return launch_pdl(final_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, a);
}

cudaError_t launch_convert16(const float* in, void* out16, long long n, int opd, cudaStream_t s) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return launch_pdl(convert16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, in,
                      reinterpret_cast<uint16_t*>(out16), n, opd);
}

cudaError_t launch_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* w16,
                           float* c, float* d, int N, int K, int opd, cudaStream_t s, int centre) {
    fold_ln_kernel<<<N, 256, 0, s>>>(W, gamma, beta, bias, reinterpret_cast<uint16_t*>(w16), c, d, K, opd, centre);
    return cudaGetLastError();
}

cudaError_t launch_step(StepState* st, const float* grid, const unsigned char* mask, const unsigned char* amask,
                        int stage, cudaStream_t s, float frac) {
    return launch_pdl(step_kernel, dim3(1), dim3(32), 0, s, st, grid, mask, amask, stage, frac);
}

}  // namespace usp
