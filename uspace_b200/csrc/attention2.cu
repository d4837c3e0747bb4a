// Persistent whole-row attention (v2): one CTA per SM loops over (sample, head) items.
//
// Same math as attention.cu (softmax(Q K^T / 8) V, L <= 352, head_dim 64) with the three costs the round-1
// ncu capture showed removed:
//   * K and V of an item are TMA-loaded ONCE (two boxes of L16/2 rows, 128B swizzle) and reused by all of its query tiles,
//     and the NEXT item's K/V are prefetched into the other shared-memory buffer while this item computes
//     (v1 reloaded K/V per 128-query tile and exposed the full load latency in every CTA);
//   * P never goes through shared memory: the softmax warps write 16-bit P straight into TMEM (tcgen05.st) and
//     O = P V runs as a TMEM-A-operand UMMA (`tcgen05.mma [d], [a_tmem], b_desc`), V consumed MN-major from its
//     natural [L,64] layout;
//   * TMEM is allocated once per CTA, barriers are initialised once, the control warp issues from warp-uniform code.
//
// TMEM columns (512): S fp32 [0, L16) | O fp32 [L16, L16+64) | P words of column-part 1 [L16+64, ...).
// The P words of column-part 0 alias the S columns that the same thread has already consumed (chunk c's 16 packed
// words land in S columns [16c, 16c+16), i.e. inside chunk c/2 <= c of the same thread), so no cross-thread hazard.
//
// Per query tile: S-MMA -> (8 warps) max / exp2 / sum, P -> TMEM -> PV-MMA -> O * 1/sum -> global.
// S-MMA of tile g+1 overlaps the O read-out of tile g (disjoint TMEM columns).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int THREADS = 384;            // 2 softmax warpgroups + 1 control warpgroup (only its first warp works)
constexpr int CTRL_WARP = 8;
constexpr int QTILE_BYTES = QT * HD * 2;   // 16 KiB
constexpr int MAX_L2 = 352;
constexpr int TMEM_COLS = 512;
constexpr int TAIL_MAX = 2;             // leftover query rows handled on CUDA cores
constexpr int TAIL_THREADS = 96;        // warps 9-11

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// A operand from TMEM (M=128 lanes x 16 K-elements = 8 packed 32-bit columns), B from smem descriptor
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// V rows as the B operand in MN-major form (see attention.cu): 64 head-dim elements contiguous per key row
__device__ __forceinline__ uint64_t umma_desc_v_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <int MODE, bool EDIT>   // EDIT: p2p column re-weighting compiled in.  MODE 0: two-pass softmax (any L <= 352), 1: score-row share in registers (L <= 320),
                      // 2: registers + software-pipelined tiles (L <= 288)
__global__ void __launch_bounds__(THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t q_full[2], kv_full[2], bar_s, s_free, p_stage[6], bar_o, tail_done, tail_go;
    __shared__ float t_q[HD], t_p[MAX_L2], t_red[2][4], t_o[3][HD];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_max[2][QT], s_sum[2][QT];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = a.L;
    const int L16 = (L + 15) & ~15;
    const int hrows = L16 / 2;              // K / V arrive as two TMA boxes of L16/2 rows (multiple of 8)
    const int kv_bytes = L16 * 128;
    // 1-2 leftover query rows (L = 257 / 258) would cost a whole extra MMA tile per item in the serial
    // S-MMA -> softmax -> PV-MMA chain; they are computed by the otherwise idle warps 9-11 on CUDA cores instead,
    // from the K / V tiles already in shared memory, concurrently with the tensor-core pipeline.
    const bool tail_simt = (L >= QT) && (L % QT != 0) && (L % QT <= TAIL_MAX);
    const int n_qt = tail_simt ? L / QT : (L + QT - 1) / QT;
    const int n_items = a.B * a.H;
    const int nch = (L + 31) / 32;          // 32-column score chunks
    const int n0 = (nch + 1) / 2;           // chunks handled by column-part 0
    constexpr bool REGS = MODE >= 1;
    constexpr bool PIPE = MODE == 2;
    // PIPE keeps P in its own columns so that the next tile's S-MMA may overwrite S while P is still being consumed
    const int S_COL = 0, PP_COL = L16, O_COL = PIPE ? L16 + 16 * nch : L16, P1_COL = L16 + 64;

    // smem: Q[2] | K[2] | V[2]
    uint8_t* sQ = smem;
    uint8_t* sK = smem + 2 * QTILE_BYTES;
    uint8_t* sV = sK + 2 * kv_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&kv_full[i], 1);
        }
        mbar_init(&bar_s, 1);
        mbar_init(&s_free, 256);
        for (int i = 0; i < 6; ++i) mbar_init(&p_stage[i], 256);
        mbar_init(&bar_o, 1);
        mbar_init(&tail_done, 3);
        mbar_init(&tail_go, 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == CTRL_WARP) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    pdl_launch();

    // Register re-balancing between warpgroups (setmaxnreg): the control warpgroup needs almost nothing, the softmax
    // warpgroups keep a whole score-row share (160 fp32) in registers.  (168-72)*128 regs released == (216-168)*256 regs acquired <= 65536.
    if (warp > CTRL_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===================== tail warps (9-11): leftover query rows on CUDA cores =====================
        if (tail_simt) {
            const int tt = threadIdx.x - (CTRL_WARP + 1) * 32;   // 0..95
            const float c2 = 0.125f * 1.44269504088896340736f;
            int n = 0;
            for (int bh = blockIdx.x; bh < n_items; bh += gridDim.x, ++n) {
                const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                      ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
                const uint8_t* kbuf = sK + (n & 1) * kv_bytes;
                const uint8_t* vbuf = sV + (n & 1) * kv_bytes;
                mbar_wait(&tail_go, n & 1);
                mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
                for (int l = n_qt * QT; l < ((a.diag & 32) ? 0 : L); ++l) {   // diag 32: skip the tail rows' arithmetic
                    // query row -> fp32 in smem
                    if (tt < HD / 2) {
                        const uint32_t w = reinterpret_cast<const uint32_t*>(a.q16)[(static_cast<long long>(bh) * L + l) * (HD / 2) + tt];
                        const float2 f = a.opd == OPD_FP16 ? Op16<OPD_FP16>::unpack(w) : Op16<OPD_BF16>::unpack(w);
                        t_q[2 * tt] = f.x;
                        t_q[2 * tt + 1] = f.y;
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    // scores for keys tt, tt+96, tt+192 (log2 domain)
                    float x[3];
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int j = tt + i * TAIL_THREADS;
                        x[i] = -INFINITY;
                        if (j < L) {
                            const uint8_t* krow = kbuf + (j >> 3) * 1024 + (j & 7) * 128;
                            float acc = 0.f;
#pragma unroll 1
                            for (int u = 0; u < 8; ++u) {
                                const uint4 kk = *reinterpret_cast<const uint4*>(krow + ((u ^ (j & 7)) << 4));
                                const uint32_t kw[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = a.opd == OPD_FP16 ? Op16<OPD_FP16>::unpack(kw[e]) : Op16<OPD_BF16>::unpack(kw[e]);
                                    acc = fmaf(t_q[u * 8 + 2 * e], f.x, acc);
                                    acc = fmaf(t_q[u * 8 + 2 * e + 1], f.y, acc);
                                }
                            }
                            x[i] = acc * c2;
                            mx = fmaxf(mx, x[i]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (lane == 0) t_red[0][warp - CTRL_WARP - 1] = mx;
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    mx = fmaxf(fmaxf(t_red[0][0], t_red[0][1]), t_red[0][2]);
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int j = tt + i * TAIL_THREADS;
                        if (j < L) {
                            // P is rounded to the 16-bit operand type exactly like the tensor-core path
                            float p = ex2_approx(x[i] - mx);
                            const uint32_t w = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(p, 0.f) : Op16<OPD_BF16>::pack(p, 0.f);
                            sum += p;
                            p = (a.opd == OPD_FP16 ? Op16<OPD_FP16>::unpack(w) : Op16<OPD_BF16>::unpack(w)).x;
                            if (EDIT && cs != nullptr) p *= __ldg(cs + j);
                            t_p[j] = p;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    if (lane == 0) t_red[1][warp - CTRL_WARP - 1] = sum;
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    const float inv = 1.0f / (t_red[1][0] + t_red[1][1] + t_red[1][2]);
                    // O[d] = sum_j p_j V[j][d]: thread = (pair of d, one third of the keys)
                    const int dp = tt & 31, seg = tt >> 5;
                    float o0 = 0.f, o1 = 0.f;
                    for (int j = seg; j < L; j += 3) {
                        const uint8_t* vrow = vbuf + (j >> 3) * 1024 + (j & 7) * 128;
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(vrow + ((((dp >> 2) ^ (j & 7))) << 4) + (dp & 3) * 4);
                        const float2 f = a.opd == OPD_FP16 ? Op16<OPD_FP16>::unpack(w) : Op16<OPD_BF16>::unpack(w);
                        const float p = t_p[j];
                        o0 = fmaf(p, f.x, o0);
                        o1 = fmaf(p, f.y, o1);
                    }
                    t_o[seg][2 * dp] = o0;
                    t_o[seg][2 * dp + 1] = o1;
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    if (tt < HD / 2) {
                        const float r0 = (t_o[0][2 * tt] + t_o[1][2 * tt] + t_o[2][2 * tt]) * inv;
                        const float r1 = (t_o[0][2 * tt + 1] + t_o[1][2 * tt + 1] + t_o[2][2 * tt + 1]) * inv;
                        const uint32_t w = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(r0, r1) : Op16<OPD_BF16>::pack(r0, r1);
                        reinterpret_cast<uint32_t*>(a.out16)[((static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD) / 2 + tt] = w;
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");   // t_q / t_p / t_o are reused by the next row
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&tail_done);
            }
        }
    } else if (warp == CTRL_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===================== control warp: TMA + MMA issue (warp-uniform, one elected lane acts) ============
        const int fmt = a.opd == OPD_FP16 ? 0 : 1;
        const uint32_t idesc_o = umma_idesc(fmt, QT, HD, 0, 1);
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
        }
        __syncwarp();
        // prologue: first item's K/V and first Q tile
        if (static_cast<int>(blockIdx.x) < n_items) {
            if (elect_one()) {
                const int bh = blockIdx.x;
                mbar_expect_tx(&kv_full[0], 2 * kv_bytes);
                for (int c = 0; c < 2; ++c) {
                    tma_load_3d(&tmK, &kv_full[0], sK + c * hrows * 128, 0, c * hrows, bh);
                    tma_load_3d(&tmV, &kv_full[0], sV + c * hrows * 128, 0, c * hrows, bh);
                }
                mbar_expect_tx(&q_full[0], QTILE_BYTES);
                tma_load_3d(&tmQ, &q_full[0], sQ, 0, 0, bh);
            }
            __syncwarp();
        }
        if (PIPE) {
            // ---------- software-pipelined issue: S-MMA(g+1) goes out as soon as S(g) sits in registers ----------
            const int n_tiles = ((n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                 static_cast<int>(gridDim.x)) * n_qt;   // tiles this CTA will process
            const int nks = L16 / 16;
            auto issue_s = [&](int gg) {   // S = Q K^T for tile gg (all lanes call; one elected lane issues)
                const int nn_ = gg / n_qt;
                mbar_wait(&q_full[gg & 1], (gg >> 1) & 1);
                if (gg % n_qt == 0) mbar_wait(&kv_full[nn_ & 1], (nn_ >> 1) & 1);
                tc_fence_after();
                const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + (gg & 1) * QTILE_BYTES));
                const uint32_t kbase = smem_u32(sK + (nn_ & 1) * kv_bytes);
                if (elect_one()) {
                    for (int c0 = 0; c0 < L16; c0 += 256) {
                        const int nn = (L16 - c0) < 256 ? (L16 - c0) : 256;
                        const uint32_t idesc = umma_idesc(fmt, QT, nn, 0, 0);
                        const uint64_t kdesc = umma_desc_sw128(kbase + c0 * 128);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            umma_f16(tmem_base + S_COL + c0, qdesc + (k * 2), kdesc + (k * 2), idesc, k != 0);
                    }
                    umma_commit(&bar_s);
                }
                __syncwarp();
            };
            auto load_q = [&](int gg) {    // Q tile of this CTA's tile gg into buffer gg & 1 (elected lane only)
                const int bhq = static_cast<int>(blockIdx.x) + (gg / n_qt) * static_cast<int>(gridDim.x);
                mbar_expect_tx(&q_full[gg & 1], QTILE_BYTES);
                tma_load_3d(&tmQ, &q_full[gg & 1], sQ + (gg & 1) * QTILE_BYTES, 0, (gg % n_qt) * QT, bhq);
            };
            if (n_tiles > 0) {
                if (n_tiles > 1 && elect_one()) load_q(1);   // Q(0) was loaded by the prologue above
                __syncwarp();
                issue_s(0);
            }
            for (int g = 0; g < n_tiles; ++g) {
                const int n = g / n_qt, t = g % n_qt;
                const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                if (t == 0) {
                    // the previous item's last PV has retired: its K/V buffer may be refilled
                    if (g > 0) mbar_wait(&bar_o, (g - 1) & 1);
                    if (elect_one()) {
                        if (tail_simt) {
                            if (n > 0) mbar_wait(&tail_done, (n - 1) & 1);
                            mbar_arrive(&tail_go);
                        }
                        if (bh + static_cast<int>(gridDim.x) < n_items) {
                            const int nb = (n & 1) ^ 1;
                            const int bh2 = bh + gridDim.x;
                            mbar_expect_tx(&kv_full[nb], 2 * kv_bytes);
                            for (int c = 0; c < 2; ++c) {
                                tma_load_3d(&tmK, &kv_full[nb], sK + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh2);
                                tma_load_3d(&tmV, &kv_full[nb], sV + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh2);
                            }
                        }
                    }
                    __syncwarp();
                }
                // S(g) is in the softmax warps' registers: the S columns and Q buffer g & 1 are free again
                mbar_wait(&s_free, g & 1);
                if (g + 2 < n_tiles && elect_one()) load_q(g + 2);
                __syncwarp();
                if (g + 1 < n_tiles) issue_s(g + 1);
                // ---- O(g) = P V, stage by stage as P chunks are published ----
                const uint32_t vbase = smem_u32(sV + (n & 1) * kv_bytes);
                for (int st = 0; st < n0; ++st) {
                    mbar_wait(&p_stage[st], g & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = h == 0 ? st : n0 + st;
                            if (c >= nch) continue;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int kk = 2 * c + e;
                                if (kk >= nks) continue;
                                const uint64_t vdesc = umma_desc_v_mn(vbase + kk * 16 * 128);
                                umma_f16_ts(tmem_base + O_COL, tmem_base + PP_COL + 16 * c + 8 * e, vdesc, idesc_o,
                                            !(st == 0 && h == 0 && e == 0));
                            }
                        }
                        if (st == n0 - 1) umma_commit(&bar_o);
                    }
                    __syncwarp();
                }
            }
        } else {
        int g = 0;  // global tile counter of this CTA
        int n = 0;  // item counter of this CTA
        for (int bh = blockIdx.x; bh < n_items; bh += gridDim.x, ++n) {
            const int kb = n & 1;
            for (int t = 0; t < n_qt; ++t, ++g) {
                // previous tile's PV finished: S / P columns and (at t == 0) the other K/V buffer are free
                if (g > 0) mbar_wait(&bar_o, (g - 1) & 1);
                if (elect_one()) {
                    if (t == 0 && tail_simt) {
                        // two-way handshake with the tail warps: neither side can get a full mbarrier phase ahead
                        // (a one-way tail_done could complete two phases before it is polled and alias its parity)
                        if (n > 0) mbar_wait(&tail_done, (n - 1) & 1);   // tail warps have left the other K/V buffer
                        mbar_arrive(&tail_go);
                    }
                    if (t == 0 && bh + static_cast<int>(gridDim.x) < n_items) {   // prefetch next item's K/V
                        const int nb = kb ^ 1;
                        const int bh2 = bh + gridDim.x;
                        mbar_expect_tx(&kv_full[nb], 2 * kv_bytes);
                        for (int c = 0; c < 2; ++c) {
                            tma_load_3d(&tmK, &kv_full[nb], sK + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh2);
                            tma_load_3d(&tmV, &kv_full[nb], sV + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh2);
                        }
                    }
                    // next Q tile (this item's next tile, or the next item's first tile)
                    const bool last_t = (t + 1 == n_qt);
                    const int bhq = last_t ? bh + static_cast<int>(gridDim.x) : bh;
                    if (bhq < n_items) {
                        const int qb = (g + 1) & 1;
                        mbar_expect_tx(&q_full[qb], QTILE_BYTES);
                        tma_load_3d(&tmQ, &q_full[qb], sQ + qb * QTILE_BYTES, 0, last_t ? 0 : (t + 1) * QT, bhq);
                    }
                }
                __syncwarp();
                mbar_wait(&q_full[g & 1], (g >> 1) & 1);
                if (t == 0) mbar_wait(&kv_full[kb], (n >> 1) & 1);
                tc_fence_after();
                // ---- S = Q K^T ----
                const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + (g & 1) * QTILE_BYTES));
                const uint32_t kbase = smem_u32(sK + kb * kv_bytes);
                if (elect_one()) {
                    for (int c0 = 0; c0 < ((a.diag & 8) ? 0 : L16); c0 += 256) {
                        const int nn = (L16 - c0) < 256 ? (L16 - c0) : 256;
                        const uint32_t idesc = umma_idesc(fmt, QT, nn, 0, 0);
                        const uint64_t kdesc = umma_desc_sw128(kbase + c0 * 128);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)
                            umma_f16(tmem_base + S_COL + c0, qdesc + (k * 2), kdesc + (k * 2), idesc, k != 0);
                    }
                    umma_commit(&bar_s);
                }
                __syncwarp();
                // ---- O = P V (A = P in TMEM), issued stage by stage as the softmax warps publish P chunks:
                //      stage s = chunk s of column-part 0 and chunk n0+s of column-part 1, so the PV MMAs of the
                //      early chunks run underneath the exponentials of the later ones ----
                const uint32_t vbase = smem_u32(sV + kb * kv_bytes);
                const int nks = L16 / 16;
                for (int st = 0; st < n0; ++st) {
                    mbar_wait(&p_stage[st], g & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = h == 0 ? st : n0 + st;
                            if (c >= nch) continue;
                            const uint32_t pbase = h == 0 ? S_COL + 16 * c : P1_COL + 16 * (c - n0);
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int kk = 2 * c + e;
                                if (kk >= nks || (a.diag & 4)) continue;
                                const uint64_t vdesc = umma_desc_v_mn(vbase + kk * 16 * 128);
                                umma_f16_ts(tmem_base + O_COL, tmem_base + pbase + 8 * e, vdesc, idesc_o,
                                            !(st == 0 && h == 0 && e == 0));
                            }
                        }
                        if (st == n0 - 1) umma_commit(&bar_o);
                    }
                    __syncwarp();
                }
            }
        }
        }   // !PIPE
    } else {
        // ===================== softmax / output warps (two threads per query row) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int lg = warp & 3;
        const int part = warp >> 2;
        const int row = lg * 32 + lane;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
        const float c2 = 0.125f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
        const int c_lo = part == 0 ? 0 : n0;
        const int c_hi = part == 0 ? n0 : nch;
        if (PIPE) {
            // ---------- pipelined order: S(g) -> registers, release S, max, [finish tile g-1: O read-out], exp, P ----
            auto store_o = [&](int pbh, int pl, bool prow_ok, bool pwarp_ok, float pinv, int pg) {
                mbar_wait(&bar_o, pg & 1);
                tc_fence_after();
                if (pwarp_ok) {
                    uint16_t* op = reinterpret_cast<uint16_t*>(a.out16) +
                                   (static_cast<long long>(pbh / a.H) * L + pl) * a.D + (pbh % a.H) * HD + part * 32;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {   // 16 columns at a time: the score registers stay live
                        uint32_t r[16];
                        tmem_ld16(t_row + O_COL + part * 32 + hh * 16, r);
                        tmem_ld_wait();
                        uint32_t u[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v0 = __uint_as_float(r[2 * j]) * pinv, v1 = __uint_as_float(r[2 * j + 1]) * pinv;
                            u[j] = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(v0, v1) : Op16<OPD_BF16>::pack(v0, v1);
                        }
                        if (prow_ok) st_global_v8_b32(op + hh * 16, u);
                    }
                }
                tc_fence_before();
            };
            int g = 0;
            int p_bh = 0, p_l = 0;
            bool p_row_ok = false, p_warp_ok = false;
            float p_inv = 0.f;
            const int cnt = c_hi - c_lo;
            for (int bh = blockIdx.x; bh < n_items; bh += gridDim.x) {
                const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                      ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
                for (int t = 0; t < n_qt; ++t, ++g) {
                    const int l = t * QT + row;
                    const bool row_ok = l < L;
                    const bool warp_ok = __any_sync(0xffffffffu, row_ok);
                    uint32_t r[5][32];
                    mbar_wait(&bar_s, g & 1);
                    tc_fence_after();
                    if (warp_ok) {
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (k < cnt) tmem_ld32(t_row + S_COL + (c_lo + k) * 32, r[k]);
                        tmem_ld_wait();
                    }
                    tc_fence_before();
                    mbar_arrive(&s_free);            // the control warp may overwrite S with the next tile now
                    float mx = -INFINITY;
                    if (warp_ok) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (k < cnt) {
                                const int c = c_lo + k;
                                if (c * 32 + 32 <= L) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[k][j]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (c * 32 + j < L) mx = fmaxf(mx, __uint_as_float(r[k][j]));
                                }
                            }
                        }
                    }
                    s_max[part][row] = mx;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    mx = fmaxf(s_max[0][row], s_max[1][row]);
                    const float mxs = mx * c2;
                    // the previous tile's PV has had the whole S load / max phase to finish
                    if (g > 0) store_o(p_bh, p_l, p_row_ok, p_warp_ok, p_inv, g - 1);
                    float sum = 0.f;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (k < n0) {
                            if (warp_ok && k < cnt) {
                                const int c = c_lo + k;
                                uint32_t w[16];
                                const bool full = (c * 32 + 32 <= L);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    float p0 = ex2_approx(fmaf(__uint_as_float(r[k][2 * j]), c2, -mxs));
                                    float p1 = ex2_approx(fmaf(__uint_as_float(r[k][2 * j + 1]), c2, -mxs));
                                    if (!full) {
                                        if (c * 32 + 2 * j >= L) p0 = 0.f;
                                        if (c * 32 + 2 * j + 1 >= L) p1 = 0.f;
                                    }
                                    sum += p0 + p1;
                                    if (EDIT && cs != nullptr) {   // p2p column re-weighting, after the (unscaled) row sum
                                        p0 *= __ldg(cs + c * 32 + 2 * j);
                                        p1 *= (c * 32 + 2 * j + 1 < L) ? __ldg(cs + c * 32 + 2 * j + 1) : 0.f;
                                    }
                                    w[j] = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(p0, p1) : Op16<OPD_BF16>::pack(p0, p1);
                                }
                                tmem_st16(t_row + PP_COL + 16 * c, w);
                                tmem_st_wait();
                            }
                            tc_fence_before();
                            mbar_arrive(&p_stage[k]);
                        }
                    }
                    s_sum[part][row] = sum;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    p_inv = 1.0f / (s_sum[0][row] + s_sum[1][row]);
                    p_bh = bh; p_l = l; p_row_ok = row_ok; p_warp_ok = warp_ok;
                }
            }
            if (g > 0) store_o(p_bh, p_l, p_row_ok, p_warp_ok, p_inv, g - 1);
        } else {
        int g = 0;
        for (int bh = blockIdx.x; bh < n_items; bh += gridDim.x) {
            const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                  ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
            for (int t = 0; t < n_qt; ++t, ++g) {
                const int l = t * QT + row;
                const bool row_ok = l < L;
                const bool warp_ok = __any_sync(0xffffffffu, row_ok);   // tcgen05.ld/st are warp-collective

                mbar_wait(&bar_s, g & 1);
                tc_fence_after();
                float mx = -INFINITY;
                float sum = 0.f;
                if (REGS) {
                    // this thread's whole share of the score row (<= 5 chunks of 32) stays in registers across the
                    // max exchange: one TMEM round trip instead of two passes of per-chunk load/wait
                    uint32_t r[5][32];
                    const int cnt = c_hi - c_lo;
                    if (warp_ok && !(a.diag & 16)) {
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (k < cnt) tmem_ld32(t_row + S_COL + (c_lo + k) * 32, r[k]);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (k < cnt) {
                                const int c = c_lo + k;
                                if (c * 32 + 32 <= L) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[k][j]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (c * 32 + j < L) mx = fmaxf(mx, __uint_as_float(r[k][j]));
                                }
                            }
                        }
                    }
                    s_max[part][row] = mx;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    mx = fmaxf(s_max[0][row], s_max[1][row]);
                    const float mxs = mx * c2;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (k < n0) {
                            if (warp_ok && k < cnt && !(a.diag & 2)) {
                                const int c = c_lo + k;
                                uint32_t w[16];
                                const bool full = (c * 32 + 32 <= L);
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    float p0 = fmaf(__uint_as_float(r[k][2 * j]), c2, -mxs);
                                    float p1 = fmaf(__uint_as_float(r[k][2 * j + 1]), c2, -mxs);
                                    if (!(a.diag & 3)) {
                                        p0 = ex2_approx(p0);
                                        p1 = ex2_approx(p1);
                                    }
                                    if (!full) {
                                        if (c * 32 + 2 * j >= L) p0 = 0.f;
                                        if (c * 32 + 2 * j + 1 >= L) p1 = 0.f;
                                    }
                                    sum += p0 + p1;
                                    if (EDIT && cs != nullptr) {
                                        p0 *= __ldg(cs + c * 32 + 2 * j);
                                        p1 *= (c * 32 + 2 * j + 1 < L) ? __ldg(cs + c * 32 + 2 * j + 1) : 0.f;
                                    }
                                    w[j] = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(p0, p1) : Op16<OPD_BF16>::pack(p0, p1);
                                }
                                const uint32_t pcol = c < n0 ? S_COL + 16 * c : P1_COL + 16 * (c - n0);
                                tmem_st16(t_row + pcol, w);
                                tmem_st_wait();
                            }
                            tc_fence_before();
                            mbar_arrive(&p_stage[k]);   // every thread publishes every stage (with or without work)
                        }
                    }
                } else {
                if (warp_ok) {
                    for (int c = c_lo; c < c_hi; ++c) {
                        uint32_t r[32];
                        tmem_ld32(t_row + S_COL + c * 32, r);
                        tmem_ld_wait();
                        if (c * 32 + 32 <= L) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j < L) mx = fmaxf(mx, __uint_as_float(r[j]));
                        }
                    }
                }
                s_max[part][row] = mx;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                mx = fmaxf(s_max[0][row], s_max[1][row]);
                const float mxs = mx * c2;
                for (int k = 0; k < n0; ++k) {
                    const int c = c_lo + k;
                    if (warp_ok && c < c_hi) {
                        uint32_t r[32];
                        tmem_ld32(t_row + S_COL + c * 32, r);
                        tmem_ld_wait();
                        float p[32];
                        if (c * 32 + 32 <= L) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) p[j] = ex2_approx(fmaf(__uint_as_float(r[j]), c2, -mxs));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                p[j] = (c * 32 + j < L) ? ex2_approx(fmaf(__uint_as_float(r[j]), c2, -mxs)) : 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum += p[j];
                        if (EDIT && cs != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j < L) p[j] *= __ldg(cs + c * 32 + j);
                        }
                        uint32_t w[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            w[j] = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(p[2 * j], p[2 * j + 1])
                                                     : Op16<OPD_BF16>::pack(p[2 * j], p[2 * j + 1]);
                        const uint32_t pcol = c < n0 ? S_COL + 16 * c : P1_COL + 16 * (c - n0);
                        tmem_st16(t_row + pcol, w);
                        tmem_st_wait();
                    }
                    tc_fence_before();
                    mbar_arrive(&p_stage[k]);
                }
                }
                s_sum[part][row] = sum;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float inv = 1.0f / (s_sum[0][row] + s_sum[1][row]);

                mbar_wait(&bar_o, g & 1);
                tc_fence_after();
                if (warp_ok && !(a.diag & 16)) {
                    uint32_t r[32];
                    tmem_ld32(t_row + O_COL + part * 32, r);
                    tmem_ld_wait();
                    uint16_t* op = reinterpret_cast<uint16_t*>(a.out16) +
                                   (static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD + part * 32;
                    uint32_t u[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v0 = __uint_as_float(r[2 * j]) * inv, v1 = __uint_as_float(r[2 * j + 1]) * inv;
                        u[j] = a.opd == OPD_FP16 ? Op16<OPD_FP16>::pack(v0, v1) : Op16<OPD_BF16>::pack(v0, v1);
                    }
                    if (row_ok) {
                        st_global_v8_b32(op, u);
                        st_global_v8_b32(op + 16, u + 8);
                    }
                }
                // the O read-out above must retire before the next tile's PV-MMA overwrites O: that MMA is only
                // issued after this thread's next p_stage arrival, which follows in program order
                tc_fence_before();
            }
        }
        }   // !PIPE
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CTRL_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int smem_bytes_for(int L) {
    const int L16 = (L + 15) & ~15;
    return 2 * QTILE_BYTES + 4 * L16 * 128 + 1024;
}

}  // namespace

cudaError_t attention2_configure() {
    const int bytes = smem_bytes_for(MAX_L2);
    cudaError_t e;
#define USP_ATTN_CFG(M, E)                                                                                     \
    if ((e = cudaFuncSetAttribute(attention2_kernel<M, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != \
        cudaSuccess)                                                                                           \
        return e;
    USP_ATTN_CFG(0, false) USP_ATTN_CFG(1, false) USP_ATTN_CFG(2, false)
    USP_ATTN_CFG(0, true) USP_ATTN_CFG(1, true) USP_ATTN_CFG(2, true)
#undef USP_ATTN_CFG
    return cudaSuccess;
}

bool attention2_supported(const AttnArgs& a) {
    if (a.L > MAX_L2) return false;
    const int L16 = (a.L + 15) & ~15;
    const int nch = (a.L + 31) / 32;
    return L16 + 64 + 16 * (nch - (nch + 1) / 2) <= TMEM_COLS;
}

cudaError_t launch_attention2(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& a,
                              int num_sms, cudaStream_t s) {
    const int items = a.B * a.H;
    const int grid = items < num_sms ? items : num_sms;
    static int diag = -1;
    if (diag < 0) {
        const char* e = getenv("USP_ATTN_DIAG");
        diag = e ? atoi(e) : 0;
    }
    AttnArgs a2 = a;
    a2.diag = diag;
    const int nch = (a.L + 31) / 32;
    const int L16 = (a.L + 15) & ~15;
    static int max_mode = -1;   // USP_ATTN_MODE caps the variant (A/B comparison, debugging)
    if (max_mode < 0) {
        const char* e = getenv("USP_ATTN_MODE");
        max_mode = e ? atoi(e) : 2;
    }
    int mode = 0;
    if ((nch + 1) / 2 <= 5) mode = 1;                              // score-row share fits in registers (L <= 320)
    if (mode == 1 && L16 + 16 * nch + 64 <= TMEM_COLS) mode = 2;   // room for a private P region (L <= 288)
    if (mode > max_mode) mode = max_mode;
    // the p2p column re-weighting lives in its own instantiations: in the plain ones the hook costs registers in the
    // softmax loop (ptxas spilled 448 B in MODE 2 and the kernel ran 3x slower)
    const bool edit = a.vscale != nullptr;
#define USP_ATTN_LAUNCH(M, E) \
    return launch_pdl(attention2_kernel<M, E>, dim3(grid), dim3(THREADS), smem_bytes_for(a.L), s, q, k, v, a2)
    if (mode == 2) { if (edit) USP_ATTN_LAUNCH(2, true); USP_ATTN_LAUNCH(2, false); }
    if (mode == 1) { if (edit) USP_ATTN_LAUNCH(1, true); USP_ATTN_LAUNCH(1, false); }
    if (edit) USP_ATTN_LAUNCH(0, true);
    USP_ATTN_LAUNCH(0, false);
#undef USP_ATTN_LAUNCH
}

}  // namespace usp
