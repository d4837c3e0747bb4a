// Latent -> image decoder (FrozenAutoencoderKL.decode, libs/autoencoder.py:446-450 -> Decoder.forward :376-409):
// the step right after the sampling path (dissect_lfm.py:86-98).  SURVEY.md section 8f "next-4".
//
// Layout: activations are NHWC, fp32 "residual" copies [B*H*W, C] plus fp16 GEMM operands.  Every convolution is a
// GEMM on the tcgen05 kernels of the U-ViT path (csrc/gemm2.cu / gemm.cu through gemm_raw):
//   3x3 conv  = IMPLICIT GEMM: the A tiles are loaded straight from the fp16 NHWC activation through a TMA im2col
//               tensor map (128 output pixels x 64 channels of one filter tap per load, padding zero-filled by the
//               TMA unit) x W[Cout, 9C] with K ordered (ky, kx, c).  conv_in (C = 4) and USP_VAE_IM2COL=explicit
//               materialise the im2col matrix instead.
//   1x1 conv  = the fp16 activation itself x W[Cout, Cin]
// with bias (+ residual for the second conv of a ResnetBlock and for the attention's proj_out) in the GEMM epilogue,
// which writes the NHWC fp32 result directly (row m = pixel, column n = output channel).  GroupNorm(32, eps 1e-6) +
// swish produce the next fp16 operand (two-stage deterministic statistics, no atomics).  The single 1024-token
// attention block is three GEMMs per image (q k^T, softmax rows, P v) - 1 GFLOP of 620.
// Everything here is either a GEMM (tensor pipe) or a streaming pass (HBM).
//
// Operand precision (usp_vae_set_precision).  fp16 operands carry 11 significant bits - like the TF32 convolutions the
// reference runs by default on a GPU - and 37 convolutions of that measure ~2e-3 against the fp64 oracle.  The default
// mode "fp16x3" therefore splits every GEMM operand into hi = fp16(x) and lo = fp16(x - hi) and folds the three
// significant products into ONE GEMM over a 3x longer K:
//     A' = [hi | lo | hi]  (3C channels per pixel),   W' = [W_hi | W_hi | W_lo] per filter tap
//     A' W'^T = hi W_hi + lo W_hi + hi W_lo  =  x W  up to the 2^-22 lo x lo term,
// accumulated in fp32 in tensor memory - no kernel changes, the implicit-GEMM im2col map simply sees 3C channels.
// GroupNorm then reads the fp32 activation.  3x the tensor work for ~1e-5 instead of 2e-3; USP_VAE_PRECISION_FP16 keeps
// the single-product path.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/uspace_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace usp {
cudaError_t gemm_configure();
namespace {

constexpr int VT = 256;   // threads per block of the streaming kernels

// ---- weights ------------------------------------------------------------------------------------------------
// conv weight [Co, Ci, kh, kw] fp32 -> fp16 [Np, Kp] with k = (ky*kw + kx)*Cip + ci (Cip >= Ci: the activation's
// channel count after padding, 4 for the encoder's RGB input), zero padded
// P = 3 (split operands): k = (tap*3 + part)*Cip + ci with parts [W_hi | W_hi | W_lo]
__global__ void pack_conv_kernel(const float* __restrict__ w, __half* __restrict__ out, int Co, int Ci, int Cip, int ks,
                                 int Np, int Kp, int P) {
    const long long n = static_cast<long long>(Np) * Kp;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int co = static_cast<int>(i / Kp), k = static_cast<int>(i % Kp);
        __half r = __float2half_rn(0.f);
        if (co < Co && k < ks * ks * P * Cip && k % Cip < Ci) {
            const int tap = k / (P * Cip), part = (k / Cip) % P, ci = k % Cip;
            const float v = w[((static_cast<long long>(co) * Ci + ci) * ks + tap / ks) * ks + tap % ks];
            r = __float2half_rn(v);
            if (part == 2) r = __float2half_rn(v - __half2float(r));
        }
        out[i] = r;
    }
}

// ---- split operands -----------------------------------------------------------------------------------------------------
// four consecutive values -> the hi and lo fp16 quads
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}
// one row quad of an operand with P parts: row-major [rows, P*C]; A side [hi | lo | hi], W side [hi | hi | lo]
__device__ __forceinline__ void store_parts(__half* out, long long row, int C, int q, int P, bool wside, const float4 v) {
    uint2 hi, lo;
    split4(v, hi, lo);
    uint2* o = reinterpret_cast<uint2*>(out + row * P * C) + q;
    o[0] = hi;
    if (P == 3) {
        o[C / 4] = wside ? hi : lo;
        o[C / 2] = wside ? lo : hi;
    }
}
// fp32 [rows, C] (row stride ld) -> operand [rows, P*C]
__global__ void to_operand_kernel(const float* __restrict__ in, __half* __restrict__ out, long long rows, int C, int ld,
                                  int P, int wside) {
    const int qpr = C / 4;
    const long long n = rows * qpr;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long r = i / qpr;
        const int q = static_cast<int>(i % qpr);
        store_parts(out, r, C, q, P, wside != 0, *reinterpret_cast<const float4*>(in + r * ld + q * 4));
    }
}
// out[c][r] = in[r][c], fp32
__global__ void transpose32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (r0 + i < rows && c0 + threadIdx.x < cols) tile[i][threadIdx.x] = in[static_cast<long long>(r0 + i) * cols + c0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (c0 + i < cols && r0 + threadIdx.x < rows) out[static_cast<long long>(c0 + i) * rows + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// ---- input: z NCHW / scale_factor -> post_quant_conv (1x1, 4 -> 4) -> fp16 NHWC ----------------------------------
__global__ void vae_in_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                              __half* __restrict__ out, int B, int S, float inv_scale, int P) {
    const long long n = static_cast<long long>(B) * S * S;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bi = static_cast<int>(i / (S * S)), p = static_cast<int>(i % (S * S));
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = z[(static_cast<long long>(bi) * 4 + c) * S * S + p] * inv_scale;
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float acc = b[o];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc = fmaf(w[o * 4 + c], v[c], acc);
        r[o] = acc;
    }
    store_parts(out, i, 4, 0, P, false, make_float4(r[0], r[1], r[2], r[3]));
}

// ---- im2col -------------------------------------------------------------------------------------------------------
// in16 [B, Hin, Win, C] -> A [B*H*W, Kp] (H = Hin*up), A[m][(ky*3+kx)*C + c] = in(y+ky-1, x+kx-1) of the (nearest-
// upsampled) image, zero outside.  VEC = channels per thread (8: one 16-byte vector; 4 for conv_in's C = 4).
template <int VEC>
__global__ void im2col_kernel(const __half* __restrict__ in, __half* __restrict__ A, int B, int Hin, int Win, int C,
                              int up, int Kp) {
    const int H = Hin * up, W = Win * up;
    const int cv = C / VEC;                                  // vectors per tap
    const int kv = Kp / VEC;                                 // vectors per row of A (padding columns included)
    const long long n = static_cast<long long>(B) * H * W * kv;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long m = i / kv;
        const int j = static_cast<int>(i % kv);
        const int tap = j / cv, c0 = (j % cv) * VEC;
        const int x = static_cast<int>(m % W), y = static_cast<int>((m / W) % H), b = static_cast<int>(m / (static_cast<long long>(W) * H));
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (VEC == 8) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (tap < 9 && yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * Hin + yy / up) * Win + xx / up) * C + c0);
            *reinterpret_cast<uint4*>(A + m * Kp + static_cast<long long>(j) * 8) = v;
        } else {
            uint2 v = make_uint2(0, 0);
            if (tap < 9 && yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = *reinterpret_cast<const uint2*>(in + ((static_cast<long long>(b) * Hin + yy / up) * Win + xx / up) * C + c0);
            *reinterpret_cast<uint2*>(A + m * Kp + static_cast<long long>(j) * 4) = v;
        }
    }
}

// Downsample (libs/autoencoder.py:65-69): zero pad (0, 1, 0, 1), 3x3 conv, stride 2, no padding -> the im2col row of
// output pixel (y, x) gathers input pixels (2y + ky, 2x + kx), zero beyond the last row / column
__global__ void im2col_s2_kernel(const __half* __restrict__ in, __half* __restrict__ A, int B, int Hin, int Win, int C) {
    const int H = Hin / 2, W = Win / 2, cv = C / 8, kv = 9 * cv;
    const long long n = static_cast<long long>(B) * H * W * kv;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long m = i / kv;
        const int j = static_cast<int>(i % kv), tap = j / cv, c0 = (j % cv) * 8;
        const int x = static_cast<int>(m % W), y = static_cast<int>((m / W) % H), b = static_cast<int>(m / (static_cast<long long>(W) * H));
        const int yy = 2 * y + tap / 3, xx = 2 * x + tap % 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (yy < Hin && xx < Win) v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * Hin + yy) * Win + xx) * C + c0);
        reinterpret_cast<uint4*>(A)[i] = v;
    }
}

// encoder input: image NCHW fp32 [B, 3, R, R] -> fp16 NHWC with a zero 4th channel
__global__ void vae_enc_in_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int HW, int P) {
    const long long n = static_cast<long long>(B) * HW;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = static_cast<int>(i / HW), p = static_cast<int>(i % HW);
    store_parts(out, i, 4, 0, P, false,
                make_float4(x[(static_cast<long long>(b) * 3 + 0) * HW + p], x[(static_cast<long long>(b) * 3 + 1) * HW + p],
                            x[(static_cast<long long>(b) * 3 + 2) * HW + p], 0.f));
}
// encoder output: conv_out result fp32 NHWC [M, Np] (8 channels) -> quant_conv (1x1, 8 -> 8) -> moments NCHW
__global__ void vae_enc_out_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bq,
                                   float* __restrict__ out, int B, int HW, int Np) {
    const long long n = static_cast<long long>(B) * HW;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = static_cast<int>(i / HW), p = static_cast<int>(i % HW);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = h[i * Np + c];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float acc = bq[o];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = fmaf(w[o * 8 + c], v[c], acc);
        out[(static_cast<long long>(b) * 8 + o) * HW + p] = acc;
    }
}

// nearest x2 upsample of an fp16 NHWC tensor (torch.nn.functional.interpolate(scale_factor=2, mode="nearest"))
__global__ void upsample2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int Hin, int Win, int C) {
    const int cv = C / 8, H = 2 * Hin, W = 2 * Win;
    const long long n = static_cast<long long>(B) * H * W * cv;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c8 = static_cast<int>(i % cv);
        const long long m = i / cv;
        const int x = static_cast<int>(m % W), y = static_cast<int>((m / W) % H), b = static_cast<int>(m / (static_cast<long long>(W) * H));
        reinterpret_cast<uint4*>(out)[i] =
            *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * Hin + y / 2) * Win + x / 2) * C + c8 * 8);
    }
}

// ---- GroupNorm(32 groups, eps 1e-6, affine) [+ swish]: fp16 copy of the activation in, fp16 GEMM operand out ----------
// (every producing GEMM writes the fp16 copy next to its fp32 result; reading 2 instead of 4 bytes twice takes a third
// off the GroupNorm time, which was 29 % of a decode)
// pass 1: per (sample, pixel chunk, group) partial sum / sum of squares; a thread always sees the same 4 channels
__device__ __forceinline__ float4 ld_half4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_half4(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename TIn>
__global__ void __launch_bounds__(VT) gn_partial_kernel(const TIn* __restrict__ x, double2* __restrict__ part, int HW,
                                                        int C, int chunks) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int qpp = C / 4;                       // float4 per pixel
    const int ppi = VT / qpp;                    // pixels per block iteration (C <= 1024)
    const int q = threadIdx.x % qpp, sub = threadIdx.x / qpp;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    float s = 0.f, ss = 0.f;
    if (sub < ppi) {
        for (int p = p0 + sub; p < p1; p += ppi) {
            const float4 v = ld_half4(x + (static_cast<long long>(b) * HW + p) * C + q * 4);
            s += (v.x + v.y) + (v.z + v.w);
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    }
    __shared__ float sh_s[VT], sh_q[VT];
    sh_s[threadIdx.x] = s;
    sh_q[threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        // group g owns quads [g*cpg/4, (g+1)*cpg/4) of every pixel slot; fixed summation order
        const int qpg = qpp / 32;
        double a = 0.0, c = 0.0;
        for (int sb = 0; sb < ppi; ++sb)
            for (int k = 0; k < qpg; ++k) {
                const int t = sb * qpp + threadIdx.x * qpg + k;
                a += sh_s[t];
                c += sh_q[t];
            }
        part[(static_cast<long long>(b) * chunks + chunk) * 32 + threadIdx.x] = make_double2(a, c);
    }
}
// pass 2: every block first folds its sample's partials into (mean, rstd) per group - fixed order, so all blocks and
// all runs agree - then normalises, applies the affine and optional swish x*sigmoid(x) (libs/autoencoder.py:26-28)
// and writes fp16.  grid = (blocks per sample, B)
template <typename TIn>
__global__ void __launch_bounds__(VT) gn_apply_kernel(const TIn* __restrict__ x, const double2* __restrict__ part,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      __half* __restrict__ out, int HW, int C, int chunks, double inv_n,
                                                      int swish, int P) {
    __shared__ float2 sh_stats[32];
    const int b = blockIdx.y;
    if (threadIdx.x < 32) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < chunks; ++k) {
            const double2 p = part[(static_cast<long long>(b) * chunks + k) * 32 + threadIdx.x];
            a += p.x;
            c += p.y;
        }
        const double mean = a * inv_n;
        const double var = fmax(c * inv_n - mean * mean, 0.0);
        sh_stats[threadIdx.x] = make_float2(static_cast<float>(mean), static_cast<float>(rsqrt(var + 1e-6)));
    }
    __syncthreads();
    const int qpp = C / 4, cpg = C / 32;
    const long long n4 = static_cast<long long>(HW) * qpp;          // float4 per sample
    const long long base = static_cast<long long>(b) * n4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long ii = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; ii < n4; ii += stride) {
        const long long i = base + ii;
        const int q = static_cast<int>(ii % qpp);
        const float2 st = sh_stats[(q * 4) / cpg];
        const float4 v = ld_half4(x + 4 * i);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + q);
        float y[4] = {(v.x - st.x) * st.y * g.x + be.x, (v.y - st.x) * st.y * g.y + be.y,
                      (v.z - st.x) * st.y * g.z + be.z, (v.w - st.x) * st.y * g.w + be.w};
        if (swish) {
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = y[k] / (1.f + __expf(-y[k]));
        }
        store_parts(out, i / qpp, C, q, P, false, make_float4(y[0], y[1], y[2], y[3]));
    }
}

// ---- attention helpers ------------------------------------------------------------------------------------------------
// softmax over the keys of one query row (libs/autoencoder.py:183-184): one warp per row
__global__ void __launch_bounds__(VT) softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ p, int rows,
                                                           int cols, float scale, int P) {
    const int row = blockIdx.x * (VT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* r = s + static_cast<long long>(row) * cols;
    float mx = -INFINITY;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < cols; j += 32) sum += P == 3 ? expf((r[j] - mx) * scale) : __expf((r[j] - mx) * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    __half* o = p + static_cast<long long>(row) * cols * P;
    for (int j = lane; j < cols; j += 32) {
        const float v = (P == 3 ? expf((r[j] - mx) * scale) : __expf((r[j] - mx) * scale)) * inv;
        const __half hi = __float2half_rn(v);
        o[j] = hi;
        if (P == 3) {
            o[cols + j] = __float2half_rn(v - __half2float(hi));
            o[2 * cols + j] = hi;
        }
    }
}
// out[c][r] = in[r][c]
__global__ void transpose16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows, int cols) {
    __shared__ __half tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (r0 + i < rows && c0 + threadIdx.x < cols) tile[i][threadIdx.x] = in[static_cast<long long>(r0 + i) * cols + c0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (c0 + i < cols && r0 + threadIdx.x < rows) out[static_cast<long long>(c0 + i) * rows + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// ---- output: fp32 NHWC [M, Np] (first 3 channels) -> NCHW image ---------------------------------------------------
__global__ void vae_out_kernel(const float* __restrict__ o, float* __restrict__ img, int B, int HW, int Np) {
    const long long n = static_cast<long long>(B) * 3 * HW;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = static_cast<int>(i % HW), c = static_cast<int>((i / HW) % 3), b = static_cast<int>(i / (3LL * HW));
    img[i] = o[(static_cast<long long>(b) * HW + p) * Np + c];
}

inline int grid_for(long long n) {
    long long b = (n + VT - 1) / VT;
    if (b > 148LL * 32) b = 148LL * 32;
    return static_cast<int>(b < 1 ? 1 : b);
}

struct VWeight {
    std::string name;
    std::vector<int64_t> shape;
    long long numel = 0;
    float* d32 = nullptr;
    __half* d16 = nullptr;      // packed GEMM operand (conv weights; [W_hi | W_hi | W_lo] per tap in the split mode)
    float* bias_pad = nullptr;  // bias zero-padded to the GEMM's N (conv biases whose Cout is not a multiple of 128)
    int Np = 0, Kp = 0, Cip = 0;
    bool set = false;
};

}  // namespace
}  // namespace usp

using namespace usp;

struct usp_vae {
    int device = 0, num_sms = 148;
    int P = 3;                  // operand parts: 3 = split hi / lo operands (default), 1 = plain fp16
    float scale = 0.18215f;
    std::vector<VWeight> w;
    std::map<std::string, int> idx;
    bool finalized = false;
    std::string err;
    // workspace for (chunk batch, latent side)
    int ws_B = 0, ws_S = 0;
    size_t ws_bytes = 0;
    void* slab = nullptr;
    float *f0 = nullptr, *f1 = nullptr, *f2 = nullptr;   // fp32 NHWC activations: x, y ping-pong, nin_shortcut result
    float *f3 = nullptr;                                  // split mode: conv1 result (GroupNorm reads fp32 there)
    float *q32 = nullptr, *k32 = nullptr, *v32 = nullptr, *vt32 = nullptr;   // split mode: attention projections
    __half *h0 = nullptr, *h1 = nullptr, *t16 = nullptr;  // fp16 copies of x / y, conv1 result (split mode: t16 = the
                                                          // operand form of x for 1x1 / resampling convolutions)
    __half *a16 = nullptr, *col = nullptr, *q16 = nullptr, *k16 = nullptr, *v16 = nullptr, *p16 = nullptr, *vt16 = nullptr;
    float* s32 = nullptr;
    double2* gn_part = nullptr;
};

namespace {

std::string g_vae_error;

int vfail(usp_vae* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    else g_vae_error = msg;
    return code;
}
#define VTRY(h, expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return vfail(h, USP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

int add(usp_vae* h, const std::string& name, std::vector<int64_t> shape) {
    VWeight w;
    w.name = name;
    w.shape = shape;
    w.numel = 1;
    for (auto s : shape) w.numel *= s;
    h->w.push_back(w);
    h->idx[name] = static_cast<int>(h->w.size()) - 1;
    return static_cast<int>(h->w.size()) - 1;
}
void add_conv(usp_vae* h, const std::string& p, int ci, int co, int ks) {
    add(h, p + ".weight", {co, ci, ks, ks});
    add(h, p + ".bias", {co});
}
void add_norm(usp_vae* h, const std::string& p, int c) {
    add(h, p + ".weight", {c});
    add(h, p + ".bias", {c});
}
void add_res(usp_vae* h, const std::string& p, int ci, int co) {
    add_norm(h, p + ".norm1", ci);
    add_conv(h, p + ".conv1", ci, co, 3);
    add_norm(h, p + ".norm2", co);
    add_conv(h, p + ".conv2", co, co, 3);
    if (ci != co) add_conv(h, p + ".nin_shortcut", ci, co, 1);
}

constexpr int CH = 128;
constexpr int MULT[4] = {1, 2, 4, 4};
constexpr int NRES = 2;

const VWeight& W(const usp_vae* h, const std::string& n) { return h->w[h->idx.at(n)]; }

// one convolution as im2col + GEMM.  in16: fp16 NHWC [B, Hin, Win, Ci]; result fp32 NHWC [B*H*W, Np] (+ optional fp16)
int conv(usp_vae* h, const std::string& p, const __half* in16, int B, int Hin, int Win, int up, const float* resid,
         float* out32, __half* out16, int epi, cudaStream_t s) {
    const VWeight& w = W(h, p + ".weight");
    const VWeight& b = W(h, p + ".bias");
    const int Ci = h->P * w.Cip, ks = static_cast<int>(w.shape[2]);   // channels of the (padded, split) operand
    const long long M = static_cast<long long>(B) * Hin * up * Win * up;
    const __half* A = in16;
    // USP_VAE_IM2COL=explicit materialises the im2col matrix instead of loading through a TMA im2col map
    static const bool explicit_cols = [] { const char* e = getenv("USP_VAE_IM2COL"); return e && e[0] == 'e'; }();
    if (ks == 3 && Ci % 64 == 0 && M % 256 == 0 && !explicit_cols) {
        // implicit GEMM: the TMA unit gathers the 3x3 neighbourhood (and zero-fills the padding) while loading A
        if (up == 2) {
            upsample2_kernel<<<grid_for(M * (Ci / 8)), VT, 0, s>>>(in16, h->col, B, Hin, Win, Ci);
            VTRY(h, cudaGetLastError());
            A = h->col;
        }
        const char* e = gemm_raw(epi, A, w.d16, b.bias_pad, resid, out32, out16, static_cast<int>(M), w.Np, w.Kp, OPD_FP16,
                                 h->num_sms, s, Ci, Hin * up, Win * up);
        if (e) return vfail(h, USP_ERR_CUDA, "conv " + p + ": " + e);
        return USP_OK;
    }
    if (ks == 3) {
        if (Ci % 8 == 0)
            im2col_kernel<8><<<grid_for(M * (w.Kp / 8)), VT, 0, s>>>(in16, h->col, B, Hin, Win, Ci, up, w.Kp);
        else
            im2col_kernel<4><<<grid_for(M * (w.Kp / 4)), VT, 0, s>>>(in16, h->col, B, Hin, Win, Ci, up, w.Kp);
        VTRY(h, cudaGetLastError());
        A = h->col;
    }
    const char* e = gemm_raw(epi, A, w.d16, b.bias_pad, resid, out32, out16, static_cast<int>(M), w.Np, w.Kp, OPD_FP16,
                             h->num_sms, s);
    if (e) return vfail(h, USP_ERR_CUDA, "conv " + p + ": " + e);
    return USP_OK;
}

// an activation: fp32 NHWC (residual path) + its fp16 copy (GroupNorm / 1x1 conv / resampling input)
struct Act {
    float* f;
    __half* h;
};

// GroupNorm (+ swish) of the NHWC activation x [B, HW, C] (its fp16 copy; the fp32 tensor in the split mode) -> operand
int group_norm(usp_vae* h, const std::string& p, Act x, int B, int HW, int C, bool swish, __half* out, cudaStream_t s) {
    int chunks = HW / 64;
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    const long long n4 = static_cast<long long>(HW) * C / 4;     // per sample
    int per_sample = static_cast<int>((n4 + VT * 8 - 1) / (VT * 8));   // ~8 float4 per thread
    if (per_sample < 1) per_sample = 1;
    if (per_sample > 1024) per_sample = 1024;
    const float* g = W(h, p + ".weight").d32;
    const float* b = W(h, p + ".bias").d32;
    const double inv_n = 1.0 / (static_cast<double>(HW) * (C / 32));
    if (h->P == 3) {
        gn_partial_kernel<float><<<dim3(chunks, B), VT, 0, s>>>(x.f, h->gn_part, HW, C, chunks);
        gn_apply_kernel<float><<<dim3(per_sample, B), VT, 0, s>>>(x.f, h->gn_part, g, b, out, HW, C, chunks, inv_n, swish ? 1 : 0, 3);
    } else {
        gn_partial_kernel<__half><<<dim3(chunks, B), VT, 0, s>>>(x.h, h->gn_part, HW, C, chunks);
        gn_apply_kernel<__half><<<dim3(per_sample, B), VT, 0, s>>>(x.h, h->gn_part, g, b, out, HW, C, chunks, inv_n, swish ? 1 : 0, 1);
    }
    VTRY(h, cudaGetLastError());
    return USP_OK;
}

// split mode: the operand form [hi | lo | hi] of an fp32 activation (what the epilogue's fp16 copy is in the fp16 mode)
int to_operand(usp_vae* h, const float* x, __half* out, long long rows, int C, bool wside, cudaStream_t s) {
    to_operand_kernel<<<grid_for(rows * (C / 4)), VT, 0, s>>>(x, out, rows, C, C, h->P, wside ? 1 : 0);
    VTRY(h, cudaGetLastError());
    return USP_OK;
}

// ResnetBlock.forward (libs/autoencoder.py:114-134, temb is None): x [B, H*W, Ci] -> y [B, H*W, Co]
int res_block(usp_vae* h, const std::string& p, Act x, Act y, int B, int H, int Ci, int Co, cudaStream_t s) {
    int rc;
    const bool split = h->P == 3;
    if ((rc = group_norm(h, p + ".norm1", x, B, H * H, Ci, true, h->a16, s))) return rc;
    // conv1's result only feeds norm2: fp16 is all that is written (fp32 in the split mode)
    Act t = {split ? h->f3 : nullptr, split ? nullptr : h->t16};
    if ((rc = conv(h, p + ".conv1", h->a16, B, H, H, 1, nullptr, t.f, t.h, EPI_BIAS_F32, s))) return rc;
    const float* resid = x.f;
    if (Ci != Co) {
        const __half* xop = x.h;
        if (split) {
            if ((rc = to_operand(h, x.f, h->t16, static_cast<long long>(B) * H * H, Ci, false, s))) return rc;
            xop = h->t16;
        }
        if ((rc = conv(h, p + ".nin_shortcut", xop, B, H, H, 1, nullptr, h->f2, nullptr, EPI_BIAS_F32, s))) return rc;
        resid = h->f2;
    }
    if ((rc = group_norm(h, p + ".norm2", t, B, H * H, Co, true, h->a16, s))) return rc;
    return conv(h, p + ".conv2", h->a16, B, H, H, 1, resid, y.f, y.h, EPI_BIAS_RESID, s);
}

// AttnBlock.forward (libs/autoencoder.py:171-195): single head over H*W tokens of width C
int attn_block(usp_vae* h, const std::string& p, Act x, Act y, int B, int H, int C, cudaStream_t s) {
    const int T = H * H;
    int rc;
    if ((rc = group_norm(h, p + ".norm", x, B, T, C, false, h->a16, s))) return rc;
    if (h->P == 3) {
        // split mode: fp32 projections; every product below runs on [hi | lo | hi] x [hi | hi | lo] operands
        if ((rc = conv(h, p + ".q", h->a16, B, H, H, 1, nullptr, h->q32, nullptr, EPI_BIAS_F32, s))) return rc;
        if ((rc = conv(h, p + ".k", h->a16, B, H, H, 1, nullptr, h->k32, nullptr, EPI_BIAS_F32, s))) return rc;
        if ((rc = conv(h, p + ".v", h->a16, B, H, H, 1, nullptr, h->v32, nullptr, EPI_BIAS_F32, s))) return rc;
        const float scale = 1.0f / sqrtf(static_cast<float>(C));
        for (int b = 0; b < B; ++b) {
            const long long o = static_cast<long long>(b) * T * C;
            if ((rc = to_operand(h, h->q32 + o, h->q16, T, C, false, s))) return rc;
            if ((rc = to_operand(h, h->k32 + o, h->k16, T, C, true, s))) return rc;
            const char* e = gemm_raw(EPI_BIAS_F32, h->q16, h->k16, nullptr, nullptr, h->s32, nullptr, T, T, 3 * C, OPD_FP16,
                                     h->num_sms, s);
            if (e) return vfail(h, USP_ERR_CUDA, std::string("attention q k^T: ") + e);
            softmax_rows_kernel<<<(T + VT / 32 - 1) / (VT / 32), VT, 0, s>>>(h->s32, h->p16, T, T, scale, 3);
            transpose32_kernel<<<dim3((C + 31) / 32, (T + 31) / 32), dim3(32, 8), 0, s>>>(h->v32 + o, h->vt32, T, C);
            VTRY(h, cudaGetLastError());
            if ((rc = to_operand(h, h->vt32, h->vt16, C, T, true, s))) return rc;
            e = gemm_raw(EPI_BIAS_F32, h->p16, h->vt16, nullptr, nullptr, h->f3 + o, nullptr, T, C, 3 * T, OPD_FP16, h->num_sms, s);
            if (e) return vfail(h, USP_ERR_CUDA, std::string("attention P v: ") + e);
        }
        if ((rc = to_operand(h, h->f3, h->a16, static_cast<long long>(B) * T, C, false, s))) return rc;
        return conv(h, p + ".proj_out", h->a16, B, H, H, 1, x.f, y.f, nullptr, EPI_BIAS_RESID, s);
    }
    if ((rc = conv(h, p + ".q", h->a16, B, H, H, 1, nullptr, nullptr, h->q16, EPI_BIAS_F32, s))) return rc;
    if ((rc = conv(h, p + ".k", h->a16, B, H, H, 1, nullptr, nullptr, h->k16, EPI_BIAS_F32, s))) return rc;
    if ((rc = conv(h, p + ".v", h->a16, B, H, H, 1, nullptr, nullptr, h->v16, EPI_BIAS_F32, s))) return rc;
    const float scale = 1.0f / sqrtf(static_cast<float>(C));
    for (int b = 0; b < B; ++b) {
        const __half* qb = h->q16 + static_cast<long long>(b) * T * C;
        const __half* kb = h->k16 + static_cast<long long>(b) * T * C;
        const __half* vb = h->v16 + static_cast<long long>(b) * T * C;
        const char* e = gemm_raw(EPI_BIAS_F32, qb, kb, nullptr, nullptr, h->s32, nullptr, T, T, C, OPD_FP16, h->num_sms, s);
        if (e) return vfail(h, USP_ERR_CUDA, std::string("attention q k^T: ") + e);
        softmax_rows_kernel<<<(T + VT / 32 - 1) / (VT / 32), VT, 0, s>>>(h->s32, h->p16, T, T, scale, 1);
        transpose16_kernel<<<dim3((C + 31) / 32, (T + 31) / 32), dim3(32, 8), 0, s>>>(vb, h->vt16, T, C);
        VTRY(h, cudaGetLastError());
        // h_[i, c] = sum_j P[i, j] v[j, c]: A = P [T, T], W = v^T [C, T]; the fp16 result is the operand of proj_out
        e = gemm_raw(EPI_BIAS_F32, h->p16, h->vt16, nullptr, nullptr, nullptr, h->a16 + static_cast<long long>(b) * T * C, T, C, T,
                     OPD_FP16, h->num_sms, s);
        if (e) return vfail(h, USP_ERR_CUDA, std::string("attention P v: ") + e);
    }
    return conv(h, p + ".proj_out", h->a16, B, H, H, 1, x.f, y.f, y.h, EPI_BIAS_RESID, s);
}

int ensure_workspace(usp_vae* h, int B, int S) {
    if (h->slab && h->ws_B >= B && h->ws_S == S) return USP_OK;
    const bool split = h->P == 3;
    const long long P = h->P;
    if (h->slab) cudaFree(h->slab);
    h->slab = nullptr;
    const long long px = static_cast<long long>(S) * 8 * S * 8;                 // output pixels per image
    const long long act = static_cast<long long>(B) * px * 256;                 // largest activation: 256 channels at full size
    // scratch shared by the nearest-x2 upsampled operand (256 channels at full size) and the explicit im2col matrices:
    // only conv_in (K = 36 padded to 64) is materialised on the default implicit-GEMM path, so 9 * 256 columns are
    // reserved only for the USP_VAE_IM2COL=explicit fallback (4.8 GB less at 16 images per pass)
    static const bool explicit_ws = [] { const char* e = getenv("USP_VAE_IM2COL"); return e && e[0] == 'e'; }();
    // (the implicit path needs M % 256 == 0 at every level; the coarsest one has B * S * S rows)
    const bool all_implicit = !explicit_ws && (static_cast<long long>(B) * S * S) % 256 == 0;
    const long long colb = static_cast<long long>(B) * px * (all_implicit ? 256 : 2304) * P;
    const long long T = static_cast<long long>(S) * S, C = 512;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
    const size_t o_f0 = carve(act * 4), o_f1 = carve(act * 4), o_f2 = carve(act * 4);
    // fp16 mode: fp16 copies of x / y and of conv1's result; split mode: conv1's fp32 result and one operand buffer
    const size_t o_f3 = carve(split ? act * 4 : 0);
    const size_t o_h0 = carve(split ? 0 : act * 2), o_h1 = carve(split ? 0 : act * 2), o_t16 = carve(act * 2 * P);
    const size_t o_a = carve(act * 2 * P), o_col = carve(colb * 2);
    const size_t o_q = carve(B * T * C * 2 * P), o_k = carve(B * T * C * 2 * P), o_v = carve(split ? 0 : B * T * C * 2);
    const size_t o_p = carve(T * T * 2 * P), o_vt = carve(T * C * 2 * P), o_s = carve(T * T * 4);
    const size_t o_q32 = carve(split ? B * T * C * 4 : 0), o_k32 = carve(split ? B * T * C * 4 : 0);
    const size_t o_v32 = carve(split ? B * T * C * 4 : 0), o_vt32 = carve(split ? T * C * 4 : 0);
    const size_t o_gp = carve(static_cast<size_t>(B) * 64 * 32 * sizeof(double2));
    VTRY(h, cudaMalloc(&h->slab, off));
    h->ws_bytes = off;
    char* base = static_cast<char*>(h->slab);
    h->f0 = reinterpret_cast<float*>(base + o_f0); h->f1 = reinterpret_cast<float*>(base + o_f1);
    h->f2 = reinterpret_cast<float*>(base + o_f2);
    h->f3 = reinterpret_cast<float*>(base + o_f3);
    h->q32 = reinterpret_cast<float*>(base + o_q32); h->k32 = reinterpret_cast<float*>(base + o_k32);
    h->v32 = reinterpret_cast<float*>(base + o_v32); h->vt32 = reinterpret_cast<float*>(base + o_vt32);
    h->h0 = reinterpret_cast<__half*>(base + o_h0); h->h1 = reinterpret_cast<__half*>(base + o_h1);
    h->t16 = reinterpret_cast<__half*>(base + o_t16);
    h->a16 = reinterpret_cast<__half*>(base + o_a); h->col = reinterpret_cast<__half*>(base + o_col);
    h->q16 = reinterpret_cast<__half*>(base + o_q); h->k16 = reinterpret_cast<__half*>(base + o_k);
    h->v16 = reinterpret_cast<__half*>(base + o_v); h->p16 = reinterpret_cast<__half*>(base + o_p);
    h->vt16 = reinterpret_cast<__half*>(base + o_vt); h->s32 = reinterpret_cast<float*>(base + o_s);
    h->gn_part = reinterpret_cast<double2*>(base + o_gp);
    h->ws_B = B; h->ws_S = S;
    return USP_OK;
}

// Downsample.forward: x fp32 [B, H, H, C] -> y [B, H/2, H/2, C]
int downsample(usp_vae* h, const std::string& p, Act x, Act y, int B, int H, int C, cudaStream_t s) {
    const VWeight& w = W(h, p + ".weight");
    const VWeight& b = W(h, p + ".bias");
    const long long M = static_cast<long long>(B) * (H / 2) * (H / 2);
    static const bool explicit_cols = [] { const char* e = getenv("USP_VAE_IM2COL"); return e && e[0] == 'e'; }();
    const char* err;
    const __half* xop = x.h;
    if (h->P == 3) {
        int rc = to_operand(h, x.f, h->t16, static_cast<long long>(B) * H * H, C, false, s);
        if (rc) return rc;
        xop = h->t16;
        C *= 3;
    }
    if (C % 64 == 0 && M % 256 == 0 && !explicit_cols) {
        // implicit GEMM with a stride-2 im2col map (every second base pixel, zero row / column at the far edge)
        err = gemm_raw(EPI_BIAS_F32, xop, w.d16, b.bias_pad, nullptr, y.f, y.h, static_cast<int>(M), w.Np, w.Kp, OPD_FP16,
                       h->num_sms, s, C, H / 2, H / 2, 2);
    } else {
        im2col_s2_kernel<<<grid_for(M * 9 * (C / 8)), VT, 0, s>>>(xop, h->col, B, H, H, C);
        VTRY(h, cudaGetLastError());
        err = gemm_raw(EPI_BIAS_F32, h->col, w.d16, b.bias_pad, nullptr, y.f, y.h, static_cast<int>(M), w.Np, w.Kp, OPD_FP16,
                       h->num_sms, s);
    }
    if (err) return vfail(h, USP_ERR_CUDA, "conv " + p + ": " + err);
    return USP_OK;
}

// Encoder.forward + quant_conv for a chunk of B images of side R
int encode_chunk(usp_vae* h, const float* img, float* moments, int B, int R, cudaStream_t s) {
    int rc;
    const std::string d = "encoder.";
    const long long n_in = static_cast<long long>(B) * R * R;
    vae_enc_in_kernel<<<static_cast<unsigned>((n_in + VT - 1) / VT), VT, 0, s>>>(img, h->a16, B, R * R, h->P);
    VTRY(h, cudaGetLastError());
    const bool split = h->P == 3;
    Act x = {h->f0, split ? nullptr : h->h0}, y = {h->f1, split ? nullptr : h->h1};
    int C = CH, H = R;
    if ((rc = conv(h, d + "conv_in", h->a16, B, H, H, 1, nullptr, x.f, x.h, EPI_BIAS_F32, s))) return rc;
    for (int lvl = 0; lvl < 4; ++lvl) {
        const int Co = CH * MULT[lvl];
        for (int blk = 0; blk < NRES; ++blk) {
            if ((rc = res_block(h, d + "down." + std::to_string(lvl) + ".block." + std::to_string(blk), x, y, B, H, C, Co, s)))
                return rc;
            std::swap(x, y);
            C = Co;
        }
        if (lvl != 3) {
            if ((rc = downsample(h, d + "down." + std::to_string(lvl) + ".downsample.conv", x, y, B, H, C, s))) return rc;
            std::swap(x, y);
            H /= 2;
        }
    }
    if ((rc = res_block(h, d + "mid.block_1", x, y, B, H, C, C, s))) return rc;
    std::swap(x, y);
    if ((rc = attn_block(h, d + "mid.attn_1", x, y, B, H, C, s))) return rc;
    std::swap(x, y);
    if ((rc = res_block(h, d + "mid.block_2", x, y, B, H, C, C, s))) return rc;
    std::swap(x, y);
    if ((rc = group_norm(h, d + "norm_out", x, B, H * H, C, true, h->a16, s))) return rc;
    if ((rc = conv(h, d + "conv_out", h->a16, B, H, H, 1, nullptr, y.f, nullptr, EPI_BIAS_F32, s))) return rc;
    const long long n_out = static_cast<long long>(B) * H * H;
    vae_enc_out_kernel<<<static_cast<unsigned>((n_out + VT - 1) / VT), VT, 0, s>>>(
        y.f, W(h, "quant_conv.weight").d32, W(h, "quant_conv.bias").d32, moments, B, H * H, W(h, d + "conv_out.weight").Np);
    VTRY(h, cudaGetLastError());
    return USP_OK;
}

// images per pass through the network: bounds the workspace (fp32 activations sized for 256 channels at full
// resolution: 67 MB per image and buffer) and sets the GEMMs' M; USP_VAE_CHUNK overrides
int vae_chunk(int S) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("USP_VAE_CHUNK");
        env = e ? atoi(e) : 0;
    }
    if (env > 0) return env;
    return S <= 32 ? 16 : 4;   // 16: the 64^2-level GEMMs fill 6.9 of 7 waves instead of 3.5 of 4 (72.7 vs 75.6 ms per 64)
}

// Decoder.forward for a chunk of B latents
int decode_chunk(usp_vae* h, const float* z, float* img, int B, int S, cudaStream_t s) {
    int rc;
    const std::string d = "decoder.";
    const long long n_in = static_cast<long long>(B) * S * S;
    // z / scale -> post_quant_conv -> fp16 NHWC [B, S, S, 4]
    vae_in_kernel<<<static_cast<unsigned>((n_in + VT - 1) / VT), VT, 0, s>>>(z, W(h, "post_quant_conv.weight").d32,
                                                                               W(h, "post_quant_conv.bias").d32, h->a16, B, S,
                                                                               1.0f / h->scale, h->P);
    VTRY(h, cudaGetLastError());
    const bool split = h->P == 3;
    Act x = {h->f0, split ? nullptr : h->h0}, y = {h->f1, split ? nullptr : h->h1};
    int C = CH * MULT[3], H = S;
    if ((rc = conv(h, d + "conv_in", h->a16, B, H, H, 1, nullptr, x.f, x.h, EPI_BIAS_F32, s))) return rc;
    if ((rc = res_block(h, d + "mid.block_1", x, y, B, H, C, C, s))) return rc;
    std::swap(x, y);
    if ((rc = attn_block(h, d + "mid.attn_1", x, y, B, H, C, s))) return rc;
    std::swap(x, y);
    if ((rc = res_block(h, d + "mid.block_2", x, y, B, H, C, C, s))) return rc;
    std::swap(x, y);
    for (int lvl = 3; lvl >= 0; --lvl) {
        const int Co = CH * MULT[lvl];
        for (int blk = 0; blk <= NRES; ++blk) {
            const std::string p = d + "up." + std::to_string(lvl) + ".block." + std::to_string(blk);
            if ((rc = res_block(h, p, x, y, B, H, C, Co, s))) return rc;
            std::swap(x, y);
            C = Co;
        }
        if (lvl != 0) {
            // Upsample: nearest x2 folded into the im2col gather, then the 3x3 conv (libs/autoencoder.py:46-50)
            const __half* xop = x.h;
            if (split) {
                if ((rc = to_operand(h, x.f, h->t16, static_cast<long long>(B) * H * H, C, false, s))) return rc;
                xop = h->t16;
            }
            if ((rc = conv(h, d + "up." + std::to_string(lvl) + ".upsample.conv", xop, B, H, H, 2, nullptr, y.f, y.h,
                           EPI_BIAS_F32, s)))
                return rc;
            std::swap(x, y);
            H *= 2;
        }
    }
    if ((rc = group_norm(h, d + "norm_out", x, B, H * H, C, true, h->a16, s))) return rc;
    if ((rc = conv(h, d + "conv_out", h->a16, B, H, H, 1, nullptr, y.f, nullptr, EPI_BIAS_F32, s))) return rc;
    const long long n_out = static_cast<long long>(B) * 3 * H * H;
    vae_out_kernel<<<static_cast<unsigned>((n_out + VT - 1) / VT), VT, 0, s>>>(y.f, img, B, H * H, W(h, d + "conv_out.weight").Np);
    VTRY(h, cudaGetLastError());
    return USP_OK;
}

}  // namespace

extern "C" {

int usp_vae_create(int device, float scale_factor, usp_vae** out) {
    if (!out) return USP_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return vfail(nullptr, USP_ERR_CUDA, "no CUDA device: the decoder has no CPU fallback");
    }
    cudaDeviceProp prop;
    if (device < 0 || device >= ndev || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return vfail(nullptr, USP_ERR_INVALID, "bad device index");
    if (prop.major != 10) return vfail(nullptr, USP_ERR_UNSUPPORTED, "built for sm_100a (B200) only");
    if (!(scale_factor > 0.f)) return vfail(nullptr, USP_ERR_INVALID, "scale_factor must be positive");
    std::unique_ptr<usp_vae> h(new usp_vae());
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->scale = scale_factor;
    // the reference's state_dict order (libs/autoencoder.py:303-373, 416-418)
    const std::string d = "decoder.";
    int C = CH * MULT[3];
    add_conv(h.get(), d + "conv_in", 4, C, 3);
    add_res(h.get(), d + "mid.block_1", C, C);
    add_norm(h.get(), d + "mid.attn_1.norm", C);
    for (const char* n : {"q", "k", "v", "proj_out"}) add_conv(h.get(), d + "mid.attn_1." + n, C, C, 1);
    add_res(h.get(), d + "mid.block_2", C, C);
    // creation runs from the deepest level up, names are indexed by level
    int cin[4][3], cout[4];
    int c = C;
    for (int lvl = 3; lvl >= 0; --lvl) {
        cout[lvl] = CH * MULT[lvl];
        for (int blk = 0; blk <= NRES; ++blk) { cin[lvl][blk] = c; c = cout[lvl]; }
    }
    for (int lvl = 0; lvl < 4; ++lvl) {
        for (int blk = 0; blk <= NRES; ++blk)
            add_res(h.get(), d + "up." + std::to_string(lvl) + ".block." + std::to_string(blk), cin[lvl][blk], cout[lvl]);
        if (lvl != 0) add_conv(h.get(), d + "up." + std::to_string(lvl) + ".upsample.conv", cout[lvl], cout[lvl], 3);
    }
    add_norm(h.get(), d + "norm_out", CH);
    add_conv(h.get(), d + "conv_out", CH, 3, 3);
    add_conv(h.get(), "post_quant_conv", 4, 4, 1);
    // encoder (libs/autoencoder.py:209-272) + quant_conv
    const std::string e = "encoder.";
    add_conv(h.get(), e + "conv_in", 3, CH, 3);
    int ec = CH;
    for (int lvl = 0; lvl < 4; ++lvl) {
        const int co = CH * MULT[lvl];
        for (int blk = 0; blk < NRES; ++blk) {
            add_res(h.get(), e + "down." + std::to_string(lvl) + ".block." + std::to_string(blk), ec, co);
            ec = co;
        }
        if (lvl != 3) add_conv(h.get(), e + "down." + std::to_string(lvl) + ".downsample.conv", ec, ec, 3);
    }
    add_res(h.get(), e + "mid.block_1", ec, ec);
    add_norm(h.get(), e + "mid.attn_1.norm", ec);
    for (const char* n : {"q", "k", "v", "proj_out"}) add_conv(h.get(), e + "mid.attn_1." + n, ec, ec, 1);
    add_res(h.get(), e + "mid.block_2", ec, ec);
    add_norm(h.get(), e + "norm_out", ec);
    add_conv(h.get(), e + "conv_out", ec, 8, 3);
    add_conv(h.get(), "quant_conv", 8, 8, 1);
    *out = h.release();
    return USP_OK;
}

void usp_vae_destroy(usp_vae* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& w : h->w) {
        cudaFree(w.d32);
        cudaFree(w.d16);
        cudaFree(w.bias_pad);
    }
    cudaFree(h->slab);
    delete h;
}

const char* usp_vae_last_error(const usp_vae* h) { return h ? h->err.c_str() : g_vae_error.c_str(); }
int usp_vae_num_weights(const usp_vae* h) { return h ? static_cast<int>(h->w.size()) : 0; }
const char* usp_vae_weight_name(const usp_vae* h, int i) {
    return (h && i >= 0 && i < static_cast<int>(h->w.size())) ? h->w[i].name.c_str() : nullptr;
}

int usp_vae_set_weight(usp_vae* h, const char* name, const void* data, const int64_t* shape, int ndim) {
    if (!h || !name || !data || !shape) return USP_ERR_INVALID;
    auto it = h->idx.find(name);
    if (it == h->idx.end()) return vfail(h, USP_ERR_INVALID, std::string("unknown weight ") + name);
    VWeight& w = h->w[it->second];
    if (ndim != static_cast<int>(w.shape.size())) return vfail(h, USP_ERR_INVALID, std::string("rank mismatch for ") + name);
    for (int i = 0; i < ndim; ++i)
        if (shape[i] != w.shape[i]) return vfail(h, USP_ERR_INVALID, std::string("shape mismatch for ") + name);
    VTRY(h, cudaSetDevice(h->device));
    if (!w.d32) VTRY(h, cudaMalloc(&w.d32, w.numel * 4));
    VTRY(h, cudaMemcpy(w.d32, data, w.numel * 4, cudaMemcpyDefault));
    w.set = true;
    h->finalized = false;
    return USP_OK;
}

// The activation slab (about 0.47 GB per image of the chunk in the split mode at 256^2) stays allocated between calls;
// a caller that alternates the decoder with a memory-hungry phase can hand it back.  The next call re-allocates.
int usp_vae_release_workspace(usp_vae* h) {
    if (!h) return USP_ERR_INVALID;
    VTRY(h, cudaSetDevice(h->device));
    VTRY(h, cudaDeviceSynchronize());        // a decode enqueued on any stream may still be using it
    if (h->slab) VTRY(h, cudaFree(h->slab));
    h->slab = nullptr;
    h->ws_B = 0;
    h->ws_bytes = 0;
    return USP_OK;
}
size_t usp_vae_workspace_bytes(const usp_vae* h) { return h ? h->ws_bytes : 0; }

int usp_vae_set_precision(usp_vae* h, int mode) {
    if (!h) return USP_ERR_INVALID;
    if (mode != USP_VAE_PRECISION_FP16 && mode != USP_VAE_PRECISION_FP16X3)
        return vfail(h, USP_ERR_INVALID, "precision must be USP_VAE_PRECISION_FP16 or USP_VAE_PRECISION_FP16X3");
    const int P = mode == USP_VAE_PRECISION_FP16X3 ? 3 : 1;
    if (P != h->P) {
        h->P = P;
        h->finalized = false;     // the weights are packed per mode
        h->ws_B = 0;              // and so is the workspace
    }
    return USP_OK;
}

int usp_vae_finalize(usp_vae* h, void* stream) {
    if (!h) return USP_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VTRY(h, cudaSetDevice(h->device));
    for (auto& w : h->w)
        if (!w.set) return vfail(h, USP_ERR_STATE, "weight not set: " + w.name);
    for (size_t i = 0; i < h->w.size(); ++i) {
        VWeight& w = h->w[i];
        if (w.shape.size() != 4 || w.name == "post_quant_conv.weight" || w.name == "quant_conv.weight") continue;
        const int Co = static_cast<int>(w.shape[0]), Ci = static_cast<int>(w.shape[1]), ks = static_cast<int>(w.shape[2]);
        w.Cip = (Ci + 3) / 4 * 4;            // the RGB input is stored with a zero 4th channel
        w.Np = (Co + 127) / 128 * 128;
        const int Kp = (ks * ks * h->P * w.Cip + 63) / 64 * 64;
        if (w.d16 && Kp != w.Kp) { cudaFree(w.d16); w.d16 = nullptr; }     // the precision mode changed
        w.Kp = Kp;
        if (!w.d16) VTRY(h, cudaMalloc(&w.d16, static_cast<size_t>(w.Np) * w.Kp * 2));
        pack_conv_kernel<<<grid_for(static_cast<long long>(w.Np) * w.Kp), VT, 0, s>>>(w.d32, w.d16, Co, Ci, w.Cip, ks, w.Np, w.Kp,
                                                                                     h->P);
        VTRY(h, cudaGetLastError());
        VWeight& b = h->w[i + 1];     // the bias follows its weight
        if (!b.bias_pad) VTRY(h, cudaMalloc(&b.bias_pad, w.Np * 4));
        VTRY(h, cudaMemsetAsync(b.bias_pad, 0, w.Np * 4, s));
        VTRY(h, cudaMemcpyAsync(b.bias_pad, b.d32, Co * 4, cudaMemcpyDeviceToDevice, s));
    }
    VTRY(h, cudaStreamSynchronize(s));
    h->finalized = true;
    return USP_OK;
}

int usp_vae_decode(usp_vae* h, const float* z, float* out, int B, int S, void* stream) {
    if (!h) return USP_ERR_INVALID;
    if (!h->finalized) return vfail(h, USP_ERR_STATE, "weights not finalised");
    if (!z || !out || B < 1) return vfail(h, USP_ERR_INVALID, "null buffer or empty batch");
    if (S < 16 || S > 64 || S % 16 != 0)
        return vfail(h, USP_ERR_INVALID, "latent side must be 16, 32, 48 or 64 (the attention GEMMs need S*S % 128 == 0)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VTRY(h, cudaSetDevice(h->device));
    const int chunk = vae_chunk(S);
    int rc = ensure_workspace(h, B < chunk ? B : chunk, S);
    if (rc) return rc;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        rc = decode_chunk(h, z + static_cast<long long>(b0) * 4 * S * S, out + static_cast<long long>(b0) * 3 * 64 * S * S, nb, S, s);
        if (rc) return rc;
    }
    return USP_OK;
}

int usp_vae_encode_moments(usp_vae* h, const float* x, float* moments, int B, int R, void* stream) {
    if (!h) return USP_ERR_INVALID;
    if (!h->finalized) return vfail(h, USP_ERR_STATE, "weights not finalised");
    if (!x || !moments || B < 1) return vfail(h, USP_ERR_INVALID, "null buffer or empty batch");
    if (R < 128 || R > 512 || R % 128 != 0)
        return vfail(h, USP_ERR_INVALID, "image side must be 128, 256, 384 or 512 (latent side R / 8 in {16, 32, 48, 64})");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VTRY(h, cudaSetDevice(h->device));
    const int S = R / 8, chunk = vae_chunk(S);
    int rc = ensure_workspace(h, B < chunk ? B : chunk, S);
    if (rc) return rc;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        rc = encode_chunk(h, x + static_cast<long long>(b0) * 3 * R * R, moments + static_cast<long long>(b0) * 8 * S * S, nb, R, s);
        if (rc) return rc;
    }
    return USP_OK;
}

}  // extern "C"
