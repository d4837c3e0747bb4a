// Attention on tcgen05/TMEM for ANY sequence length: whole score rows for L <= 384, key passes of 384 with a running
// maximum beyond (the fallback behind attention3.cu, which serves U-ViT's L = 257 / 258 / 334; this kernel takes
// 336 < L: 512^2 images give 1024 patches, datasets.py:244-245).
//
// Replaces F.scaled_dot_product_attention(q, k, v) at libs/uvit.py:93-95 (and the "math" branch
// libs/uvit_t2i.py:91-107): O = softmax(Q K^T * hd^-0.5) V, no mask, no dropout, head_dim 64.
//
// One CTA per (128-query tile, sample*head).  Up to 384 keys the full score row fits in TMEM (L16 <= 384 fp32
// columns) and there is no online-softmax rescaling; longer sequences repeat the sequence below per pass of 160 keys
// (two CTAs per SM, see AttnCfg):
// the softmax warps keep a running row maximum m and sum l, multiply the O accumulator in TMEM by 2^(m_old - m_new)
// (tcgen05.ld / st) once the previous P V has completed, and the next P V accumulates on top.  K of pass j+1 is
// fetched while pass j's softmax runs; V and P buffers are reused after pass j's P V.
//   warp 8 (1 thread): TMA loads Q tile, all of K and V (3-D tensor maps, OOB rows zero-filled),
//                      issues S = Q K^T (tcgen05.mma, K-major operands) into TMEM columns [0, L16),
//                      later issues O = P V (A = P from swizzled smem, B = V as MN-major) into columns [384, 448).
//   warps 0-7 (256 threads; warps w and w+4 share TMEM lane group w, i.e. two threads per query row, each
//                      taking half of the key columns and half of the output columns): row max, exp2, row sum in fp32,
//                      write un-normalised P as 16-bit into the 128B-swizzled K-major smem layout,
//                      finally scale O by 1/sum and store heads-merged [B*L, H*64].
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;                       // query rows per CTA
constexpr int ATTN_THREADS = 288;          // 8 softmax warps + 1 control warp
constexpr int CTRL_WARP = 8;
constexpr int TILE16K = 128 * HD * 2;         // one 128-row x 64-col 16-bit tile
// Shared / tensor memory of one CTA as a function of the keys per pass.  PASS = 384 (whole rows, L <= 384): 209 KB and all
// 512 TMEM columns, one CTA per SM.  PASS = 160 (longer sequences): 105 KB and 256 columns, so TWO CTAs share an SM and
// one's softmax runs under the other's MMAs and loads - the passes of a single CTA are strictly serial.
template <int PASS>
struct AttnCfg {
    static constexpr int KV_BYTES = PASS * HD * 2;                 // K (or V) rows of one pass, 128 B each
    static constexpr int PCH = (PASS + 63) / 64;                   // 64-key column chunks of P
    static constexpr int SQ_OFF = 0;
    static constexpr int SK_OFF = SQ_OFF + TILE16K;
    static constexpr int SV_OFF = SK_OFF + KV_BYTES;
    static constexpr int SP_OFF = SV_OFF + KV_BYTES;
    static constexpr int SMEM = SP_OFF + PCH * TILE16K + 1024;
    static constexpr int O_COL = PASS;
    static constexpr int TMEM_COLS = PASS + HD <= 256 ? 256 : 512;
    static_assert(KV_BYTES % 1024 == 0 && PASS % 32 == 0 && PASS + HD <= 512, "pass size");
};

// V tile as the B operand in MN-major form: rows are keys (K dim), 64 head-dim elements contiguous per row
// (128 B, swizzled), 8-key groups 1024 B apart (SBO).  One 64-wide MN atom, so LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_v_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait32() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int PASS>
__global__ void __launch_bounds__(ATTN_THREADS, (PASS + HD <= 256) ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t bar_qk, bar_v, bar_s, bar_p, bar_o;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_max[2][QT], s_sum[2][QT];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x;
    const int bh = blockIdx.y;
    using Cfg = AttnCfg<PASS>;
    constexpr int SQ_OFF = Cfg::SQ_OFF, SK_OFF = Cfg::SK_OFF, SV_OFF = Cfg::SV_OFF, SP_OFF = Cfg::SP_OFF;
    constexpr int O_COL = Cfg::O_COL, ATTN_TMEM_COLS = Cfg::TMEM_COLS;
    const int L = a.L;
    // key passes: one pass of L keys when they fit, passes of PASS keys otherwise (the last one partial)
    const int n_pass = (L + PASS - 1) / PASS;
    // K / V arrive as n_box TMA boxes of attn_kv_box_rows(L) rows per pass (rows beyond L are zero-filled): two half-row
    // boxes up to ATTN_MAX_L keys, one box of ATTN_LONG_PASS rows beyond
    const int hrows = attn_kv_box_rows(L);
    const int n_box = L > ATTN_MAX_L ? 1 : 2;

    if (threadIdx.x == 0) {
        mbar_init(&bar_qk, 1);
        mbar_init(&bar_v, 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_p, 256);
        mbar_init(&bar_o, 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == CTRL_WARP) tmem_alloc<ATTN_TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    pdl_launch();

    if (warp == CTRL_WARP) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            const int fmt = a.opd == OPD_FP16 ? 0 : 1;
            const uint64_t qdesc = umma_desc_sw128(smem_u32(smem + SQ_OFF));
            const uint32_t idesc_o = umma_idesc(fmt, QT, HD, 0, 1);
            mbar_expect_tx(&bar_qk, TILE16K + n_box * hrows * 128);
            tma_load_3d(&tmQ, &bar_qk, smem + SQ_OFF, 0, qt * QT, bh);
            for (int c = 0; c < n_box; ++c)
                tma_load_3d(&tmK, &bar_qk, smem + SK_OFF + c * hrows * 128, 0, c * hrows, bh);
            for (int j = 0; j < n_pass; ++j) {
                const uint32_t ph = j & 1;
                const int k0 = j * PASS;
                const int Lp = (L - k0) < PASS ? (L - k0) : PASS;
                const int Lp16 = (Lp + 15) & ~15;
                // V of this pass: its buffer is free once the previous pass's P V has completed
                if (j > 0) mbar_wait(&bar_o, (j - 1) & 1);
                mbar_expect_tx(&bar_v, n_box * hrows * 128);
                for (int c = 0; c < n_box; ++c)
                    tma_load_3d(&tmV, &bar_v, smem + SV_OFF + c * hrows * 128, 0, k0 + c * hrows, bh);
                // ---- S = Q K^T (the softmax warps released the S columns when they published the previous P) ----
                mbar_wait(&bar_qk, ph);
                tc_fence_after();
                for (int n0 = 0; n0 < Lp16; n0 += 256) {
                    const int nn = (Lp16 - n0) < 256 ? (Lp16 - n0) : 256;
                    const uint32_t idesc = umma_idesc(fmt, QT, nn, 0, 0);
                    const uint64_t kdesc = umma_desc_sw128(smem_u32(smem + SK_OFF + n0 * 128));
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_f16(tmem_base + n0, qdesc + (k * 2), kdesc + (k * 2), idesc, k != 0);
                }
                umma_commit(&bar_s);
                if (j + 1 < n_pass) {
                    // K of the next pass while this pass's softmax runs
                    mbar_wait(&bar_s, ph);
                    mbar_expect_tx(&bar_qk, n_box * hrows * 128);
                    for (int c = 0; c < n_box; ++c)
                        tma_load_3d(&tmK, &bar_qk, smem + SK_OFF + c * hrows * 128, 0, k0 + PASS + c * hrows, bh);
                }
                // ---- O (+)= P V ----
                mbar_wait(&bar_v, ph);
                mbar_wait(&bar_p, ph);
                tc_fence_after();
                const int nks = Lp16 / 16;
                for (int kk = 0; kk < nks; ++kk) {
                    const uint64_t pdesc =
                        umma_desc_sw128(smem_u32(smem + SP_OFF + (kk >> 2) * TILE16K)) + ((kk & 3) * 2);
                    const uint64_t vdesc = umma_desc_v_mn(smem_u32(smem + SV_OFF + kk * 16 * 128));
                    umma_f16(tmem_base + O_COL, pdesc, vdesc, idesc_o, (kk != 0 || j != 0) ? 1 : 0);
                }
                umma_commit(&bar_o);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax / output warps (8 warps, two threads per query row) =====================
        const int lg = warp & 3;               // TMEM lane group of this warp
        const int part = warp >> 2;            // which half of the key columns / output columns
        const int row = lg * 32 + lane;        // 0..127 == TMEM lane
        const int l = qt * QT + row;
        const bool row_ok = l < L;
        // tcgen05.ld is warp-collective (.sync.aligned): skip work only when the WHOLE warp has no valid row
        const bool warp_ok = __any_sync(0xffffffffu, row_ok);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
        const float c2 = 0.125f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
        float m_run = -INFINITY;      // running row maximum (raw scores), identical in both threads of a row
        float sum = 0.f;              // this thread's share of the running row sum, in units of 2^(-m_run c2)

        for (int j = 0; j < n_pass; ++j) {
            const uint32_t ph = j & 1;
            const int k0 = j * PASS;
            const int Lp = (L - k0) < PASS ? (L - k0) : PASS;
            const int nch = (Lp + 31) / 32;
            const int c_lo = part == 0 ? 0 : (nch + 1) / 2;
            const int c_hi = part == 0 ? (nch + 1) / 2 : nch;

            mbar_wait(&bar_s, ph);
            tc_fence_after();

            float mx = -INFINITY;
            if (warp_ok) {
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_row + c * 32, r);
                    tmem_ld_wait();
                    if (c * 32 + 32 <= Lp) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) mx = fmaxf(mx, __uint_as_float(r[jj]));
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            if (c * 32 + jj < Lp) mx = fmaxf(mx, __uint_as_float(r[jj]));
                    }
                }
            }
            s_max[part][row] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float m_new = fmaxf(m_run, fmaxf(s_max[0][row], s_max[1][row]));
            const float mxs = m_new * c2;
            if (j > 0) {
                // the previous P V has to be complete before its P buffer is overwritten and O is rescaled
                mbar_wait(&bar_o, (j - 1) & 1);
                tc_fence_after();
                const float alpha = ex2_approx((m_run - m_new) * c2);
                sum *= alpha;
                if (warp_ok) {
                    uint32_t r[32];
                    tmem_ld32(t_row + O_COL + part * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) r[jj] = __float_as_uint(__uint_as_float(r[jj]) * alpha);
                    tmem_st32(t_row + O_COL + part * 32, r);
                    tmem_st_wait32();
                }
            }
            m_run = m_new;
            if (warp_ok) {
                uint8_t* prow = smem + SP_OFF + (row >> 3) * 1024 + (row & 7) * 128;
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_row + c * 32, r);
                    tmem_ld_wait();
                    float p[32];
                    if (c * 32 + 32 <= Lp) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) p[jj] = ex2_approx(fmaf(__uint_as_float(r[jj]), c2, -mxs));
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            p[jj] = (c * 32 + jj < Lp) ? ex2_approx(fmaf(__uint_as_float(r[jj]), c2, -mxs)) : 0.f;
                    }
                    // (the row sum is taken before the 16-bit rounding of P: the difference is ~2^-12/sqrt(L) relative)
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) sum += p[jj];
                    // columns [c*32, c*32+32): chunk-of-64 index, then 4 x 16-byte units with the 128B swizzle
                    uint8_t* pc = prow + (c >> 1) * TILE16K;
                    const int u0 = (c & 1) * 4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int unit = (u0 + u) ^ (row & 7);
                        uint4 w;
                        if (a.opd == OPD_FP16) {
                            w.x = Op16<OPD_FP16>::pack(p[8 * u], p[8 * u + 1]); w.y = Op16<OPD_FP16>::pack(p[8 * u + 2], p[8 * u + 3]);
                            w.z = Op16<OPD_FP16>::pack(p[8 * u + 4], p[8 * u + 5]); w.w = Op16<OPD_FP16>::pack(p[8 * u + 6], p[8 * u + 7]);
                        } else {
                            w.x = Op16<OPD_BF16>::pack(p[8 * u], p[8 * u + 1]); w.y = Op16<OPD_BF16>::pack(p[8 * u + 2], p[8 * u + 3]);
                            w.z = Op16<OPD_BF16>::pack(p[8 * u + 4], p[8 * u + 5]); w.w = Op16<OPD_BF16>::pack(p[8 * u + 6], p[8 * u + 7]);
                        }
                        if (row_ok) *reinterpret_cast<uint4*>(pc + unit * 16) = w;
                    }
                }
            }
            // S columns read, O rescaled (tcgen05.st complete), P written: hand all three to the MMA thread
            tc_fence_before();
            fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core (async proxy)
            mbar_arrive(&bar_p);
        }
        s_sum[part][row] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.0f / (s_sum[0][row] + s_sum[1][row]);

        mbar_wait(&bar_o, (n_pass - 1) & 1);
        tc_fence_after();
        if (warp_ok) {
            uint16_t* orow = reinterpret_cast<uint16_t*>(a.out16) +
                             (static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD + part * 32;
            uint32_t r[32];
            tmem_ld32(t_row + O_COL + part * 32, r);
            tmem_ld_wait();
            uint4* op = reinterpret_cast<uint4*>(orow);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]) * inv;
                uint4 u;
                if (a.opd == OPD_FP16) {
                    u.x = Op16<OPD_FP16>::pack(v[0], v[1]); u.y = Op16<OPD_FP16>::pack(v[2], v[3]);
                    u.z = Op16<OPD_FP16>::pack(v[4], v[5]); u.w = Op16<OPD_FP16>::pack(v[6], v[7]);
                } else {
                    u.x = Op16<OPD_BF16>::pack(v[0], v[1]); u.y = Op16<OPD_BF16>::pack(v[2], v[3]);
                    u.z = Op16<OPD_BF16>::pack(v[4], v[5]); u.w = Op16<OPD_BF16>::pack(v[6], v[7]);
                }
                if (row_ok) op[j] = u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CTRL_WARP) tmem_dealloc<ATTN_TMEM_COLS>(tmem_base);
}

}  // namespace

cudaError_t attention2_configure();
bool attention2_supported(const AttnArgs& a);
cudaError_t launch_attention2(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& a,
                              int num_sms, cudaStream_t s);
cudaError_t attention3_configure();
bool attention3_supported(const AttnArgs& a);
cudaError_t launch_attention3(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& a,
                              int num_sms, cudaStream_t s);

cudaError_t attention_configure() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<ATTN_MAX_L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         AttnCfg<ATTN_MAX_L>::SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_kernel<ATTN_LONG_PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnCfg<ATTN_LONG_PASS>::SMEM);
    if (e == cudaSuccess) e = attention2_configure();
    if (e == cudaSuccess) e = attention3_configure();
    if (e == cudaSuccess) done = true;
    return e;
}

// USP_ATTN_V1=1 forces the one-CTA-per-tile kernel (A/B comparison, debugging)
static bool force_v1() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("USP_ATTN_V1");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

cudaError_t launch_attention(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& a,
                             cudaStream_t s) {
    if (a.L < 1 || a.L > ATTN_LONG_MAX_L || a.D != a.H * HD) return cudaErrorInvalidValue;
    // the column re-weighting (p2p) hook lives in the persistent kernel only
    // default: two tiles in flight (attention3.cu); USP_ATTN_V3=0 / USP_ATTN_V1=1 select the older kernels
    if (!force_v1() && attention3_supported(a)) return launch_attention3(q, k, v, a, a.num_sms, s);
    if (a.vscale != nullptr && (force_v1() || !attention2_supported(a))) return cudaErrorNotSupported;
    if (!force_v1() && attention2_supported(a)) return launch_attention2(q, k, v, a, a.num_sms, s);
    dim3 grid((a.L + QT - 1) / QT, a.B * a.H);
    if (a.L > ATTN_MAX_L)
        return launch_pdl(attention_kernel<ATTN_LONG_PASS>, grid, dim3(ATTN_THREADS), AttnCfg<ATTN_LONG_PASS>::SMEM, s, q, k, v, a);
    return launch_pdl(attention_kernel<ATTN_MAX_L>, grid, dim3(ATTN_THREADS), AttnCfg<ATTN_MAX_L>::SMEM, s, q, k, v, a);
}

}  // namespace usp
