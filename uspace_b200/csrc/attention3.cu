// Persistent two-tile attention (v3): one CTA per SM, TWO 128-query tiles in flight, each owned by its own softmax
// warpgroup (one thread per query row), ping-ponging on the MUFU pipe while the tensor core serves the other tile.
//
// Replaces F.scaled_dot_product_attention (libs/uvit.py:95) / the math branch (libs/uvit_t2i.py:91-107), head_dim 64.
//
// Why: v2 (attention2.cu) keeps one tile in flight per CTA, so S-MMA -> softmax -> PV-MMA of a tile run one after the
// other (ncu r01g: tensor pipe 14 %, 3.9 us of barrier hand-offs per item).  A whole score row of L = 257 keys (272
// fp32 TMEM columns) cannot be held twice in the 512 TMEM columns, so the keys are split into blocks of <= 144
// (L = 257: 144 + 128, L = 334: 3 x 112) with a flash-style running maximum, and each tile gets half of TMEM:
//
//   slot s (columns [256 s, 256 s + 256)):  S block fp32 [0, 144) - the 16-bit P words of the same block are written
//   over its first 72 columns once the scores sit in registers - and O fp32 [192, 256).
//
// Per tile and key block: S = Q K_blk^T (tcgen05, smem x smem) -> the row's thread loads its 144 scores into
// registers, max / exp2 / sum, packs P into TMEM -> O += P V_blk (tcgen05, A operand from TMEM, V MN-major from its
// natural [L,64] layout) and right behind it, in issue order on the tensor pipe, S of the next block.  The running
// maximum is only replaced (and O rescaled in TMEM) when a block's maximum exceeds the reference by more than 2^8
// (the exponentials then stay <= 256, exact in the 16-bit operand's range; the row sum uses the same reference) -
// mathematically identical to the exact-maximum form.
//
// K and V of an item are TMA-loaded once (double-buffered across items), Q tiles rotate through 3 buffers (two in
// flight + one prefetched).  The 1-2 leftover query rows of L = 257 / 258 are computed by the otherwise idle warps
// 9-11 on CUDA cores from the resident K/V tiles (as in v2).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int THREADS = 384;            // 2 softmax warpgroups + control warp + 3 tail-row warps
constexpr int CTRL_WARP = 8;
constexpr int QTILE_BYTES = QT * HD * 2;   // 16 KiB
constexpr int NQ = 3;                   // Q tile buffers
constexpr int KBMAX = 144;              // keys per score block = fp32 score registers per softmax thread
constexpr int NCH = KBMAX / 16;
constexpr int MAX_L3 = 336;
constexpr int TMEM_COLS = 512;
constexpr int SLOT_COLS = 256;
constexpr int O_OFF = 192;
constexpr int TAIL_MAX = 2;             // leftover query rows handled on CUDA cores
constexpr int TAIL_THREADS = 96;        // warps 9-11
constexpr float RESCALE_LOG2 = 8.0f;    // lazy running-maximum update threshold (log2 domain)

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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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

// V rows as the B operand in MN-major form: 64 head-dim elements contiguous per key row (128B swizzle)
__device__ __forceinline__ uint64_t umma_desc_v_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// key blocks: nkb blocks of `kl` keys (multiple of 16, <= KBMAX), the last one shorter
__host__ __device__ inline int blocks_nkb(int L16) { return (L16 + KBMAX - 1) / KBMAX; }
__host__ __device__ inline int blocks_len(int L16) {
    const int nkb = blocks_nkb(L16);
    return (((L16 + nkb - 1) / nkb) + 15) & ~15;
}

template <int OPD, bool EDIT>
__global__ void __launch_bounds__(THREADS, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t q_full[NQ], kv_full[2], kv_free[2], bar_s[2], p_bar[2], bar_o[2], o_free[2];
    __shared__ float t_q[HD], t_p[MAX_L3 + 16], t_red[2][4], t_o[3][HD];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int L = a.L;
    const int L16 = (L + 15) & ~15;
    const int hrows = L16 / 2;              // K / V arrive as two TMA boxes of L16/2 rows (multiple of 8)
    const int kv_bytes = L16 * 128;
    const bool tail_simt = (L >= QT) && (L % QT != 0) && (L % QT <= TAIL_MAX);
    const int n_qt = tail_simt ? L / QT : (L + QT - 1) / QT;
    const int n_items = a.B * a.H;
    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int n_tiles = my_items * n_qt;
    const int nkb = blocks_nkb(L16);
    const int kl = blocks_len(L16);

    // smem: Q[3] | K[2] | V[2]
    uint8_t* sQ = smem;
    uint8_t* sK = smem + NQ * QTILE_BYTES;
    uint8_t* sV = sK + 2 * kv_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NQ; ++i) mbar_init(&q_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_free[i], n_qt + (tail_simt ? 1 : 0));
            mbar_init(&bar_s[i], 1);
            mbar_init(&p_bar[i], QT);
            mbar_init(&bar_o[i], 1);
            mbar_init(&o_free[i], QT);
        }
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

    if (warp > CTRL_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===================== tail warps (9-11): leftover query rows on CUDA cores =====================
        if (tail_simt) {
            const int tt = threadIdx.x - (CTRL_WARP + 1) * 32;   // 0..95
            const float c2 = 0.125f * 1.44269504088896340736f;
            for (int n = 0; n < my_items; ++n) {
                const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                      ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
                const uint8_t* kbuf = sK + (n & 1) * kv_bytes;
                const uint8_t* vbuf = sV + (n & 1) * kv_bytes;
                mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
                for (int l = n_qt * QT; l < ((a.diag & 32) ? 0 : L); ++l) {   // diag 32: skip the tail rows' arithmetic
                    if (tt < HD / 2) {
                        const uint32_t w = reinterpret_cast<const uint32_t*>(a.q16)[(static_cast<long long>(bh) * L + l) * (HD / 2) + tt];
                        const float2 f = Op16<OPD>::unpack(w);
                        t_q[2 * tt] = f.x;
                        t_q[2 * tt + 1] = f.y;
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    // scores for keys tt, tt+96, ... (log2 domain)
                    float x[4];
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
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
                                    const float2 f = Op16<OPD>::unpack(kw[e]);
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
                    for (int i = 0; i < 4; ++i) {
                        const int j = tt + i * TAIL_THREADS;
                        if (j < L) {
                            // P is rounded to the 16-bit operand type exactly like the tensor-core path
                            float p = ex2_approx(x[i] - mx);
                            const uint32_t w = Op16<OPD>::pack(p, 0.f);
                            sum += p;
                            p = Op16<OPD>::unpack(w).x;
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
                        const float2 f = Op16<OPD>::unpack(w);
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
                        reinterpret_cast<uint32_t*>(a.out16)[((static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD) / 2 + tt] =
                            Op16<OPD>::pack(r0, r1);
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");   // t_q / t_p / t_o are reused by the next row
                }
                asm volatile("bar.sync 2, 96;" ::: "memory");       // every tail thread has left this K/V buffer
                if (tt == 0) mbar_arrive(&kv_free[n & 1]);
            }
        }
    } else if (warp == CTRL_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===================== control warp: TMA + MMA issue (warp-uniform, one elected lane acts) ============
        constexpr int fmt = OPD == OPD_FP16 ? 0 : 1;
        const uint32_t idesc_o = umma_idesc(fmt, QT, HD, 0, 1);
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
        }
        __syncwarp();
        int next_kv = 0;                 // next item (of this CTA) whose K/V load has not been issued
        auto issue_kv = [&](int m) {     // K/V of this CTA's item m into buffer m & 1 (all lanes call)
            if (m >= 2) mbar_wait(&kv_free[m & 1], ((m >> 1) - 1) & 1);   // item m-2 (MMAs + tail rows) has left it
            if (elect_one()) {
                const int nb = m & 1;
                const int bh = static_cast<int>(blockIdx.x) + m * static_cast<int>(gridDim.x);
                mbar_expect_tx(&kv_full[nb], 2 * kv_bytes);
                for (int c = 0; c < 2; ++c) {
                    tma_load_3d(&tmK, &kv_full[nb], sK + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh);
                    tma_load_3d(&tmV, &kv_full[nb], sV + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh);
                }
            }
            __syncwarp();
        };
        auto load_q = [&](int g) {       // Q tile g of this CTA into buffer g % 3 (all lanes call)
            if (elect_one()) {
                const int n = g / n_qt, t = g - n * n_qt;
                const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                const int qb = g % NQ;
                mbar_expect_tx(&q_full[qb], QTILE_BYTES);
                tma_load_3d(&tmQ, &q_full[qb], sQ + qb * QTILE_BYTES, 0, t * QT, bh);
            }
            __syncwarp();
        };
        // S(g, kb) = Q_g K_blk^T into slot g & 1, committed on bar_s[slot]
        auto issue_s = [&](int g, int kb) {
            const int n = g / n_qt;
            const int s = g & 1;
            if (kb == 0) {
                while (next_kv <= n) issue_kv(next_kv++);
                mbar_wait(&q_full[g % NQ], (g / NQ) & 1);
                mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
                tc_fence_after();
            }
            const int key0 = kb * kl;
            const int len = (L16 - key0) < kl ? (L16 - key0) : kl;
            const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + (g % NQ) * QTILE_BYTES));
            const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + (n & 1) * kv_bytes) + key0 * 128);
            const uint32_t idesc = umma_idesc(fmt, QT, len, 0, 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_f16(tmem_base + s * SLOT_COLS, qdesc + (k * 2), kdesc + (k * 2), idesc, k != 0);
                umma_commit(&bar_s[s]);
            }
            __syncwarp();
        };
        if (n_tiles > 0) {
            issue_kv(next_kv++);
            for (int g = 0; g < NQ && g < n_tiles; ++g) load_q(g);
            if (my_items > 1) issue_kv(next_kv++);
            issue_s(0, 0);
            if (n_tiles > 1) issue_s(1, 0);
        }
        int sg[2] = {0, 1};        // tile in flight per slot
        int skb[2] = {0, 0};       // its key block whose P is awaited
        int sph[2] = {0, 0};       // completions of p_bar[slot] consumed so far
        int stl[2] = {0, 0};       // tiles finished per slot
        int live = (n_tiles > 0 ? 1 : 0) + (n_tiles > 1 ? 1 : 0);
        uint32_t idle = 0;
        while (live > 0) {
            bool progressed = false;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (sg[s] >= n_tiles) continue;
                if (!mbar_try_wait(&p_bar[s], sph[s] & 1)) continue;
                progressed = true;
                ++sph[s];
                tc_fence_after();
                const int g = sg[s], kb = skb[s];
                const int n = g / n_qt;
                // the previous tile of this slot has been read out of O
                if (kb == 0 && stl[s] > 0) {
                    mbar_wait(&o_free[s], (stl[s] - 1) & 1);
                    tc_fence_after();
                }
                const int key0 = kb * kl;
                const int len = (L16 - key0) < kl ? (L16 - key0) : kl;
                const uint32_t vbase = smem_u32(sV + (n & 1) * kv_bytes) + key0 * 128;
                const uint32_t d_o = tmem_base + s * SLOT_COLS + O_OFF;
                const uint32_t a_p = tmem_base + s * SLOT_COLS;
                if (elect_one()) {
                    for (int kk = 0; kk < len / 16; ++kk)
                        umma_f16_ts(d_o, a_p + 8 * kk, umma_desc_v_mn(vbase + kk * 16 * 128), idesc_o,
                                    !(kb == 0 && kk == 0));
                }
                __syncwarp();
                if (kb + 1 < nkb) {
                    // next key block of the same tile: behind the PV MMAs on the tensor pipe (in issue order), so the
                    // P words it overwrites have been consumed
                    issue_s(g, kb + 1);
                    skb[s] = kb + 1;
                } else {
                    if (elect_one()) {
                        umma_commit(&bar_o[s]);
                        umma_commit(&kv_free[n & 1]);
                    }
                    __syncwarp();
                    // S(g, last) has been consumed (its P is published): Q buffer g % 3 is free for tile g + 3
                    if (g + NQ < n_tiles) load_q(g + NQ);
                    ++stl[s];
                    sg[s] = g + 2;
                    skb[s] = 0;
                    if (sg[s] < n_tiles) issue_s(sg[s], 0);
                    else --live;
                }
            }
            if (next_kv < my_items && (next_kv < 2 || mbar_try_wait(&kv_free[next_kv & 1], ((next_kv >> 1) - 1) & 1))) {
                issue_kv(next_kv++);
                progressed = true;
            }
            if (progressed) idle = 0;
            else if (++idle > (1u << 24)) __trap();
        }
    } else {
        // ===================== softmax / output warps: one thread per query row, warpgroup = slot =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int s = warp >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t t_s = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + s * SLOT_COLS;
        const uint32_t t_o = t_s + O_OFF;
        const float c2 = 0.125f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
        int ks = 0;       // S blocks consumed by this slot
        int jt = 0;       // tiles finished by this slot
        for (int g = s; g < n_tiles; g += 2, ++jt) {
            const int n = g / n_qt, t = g - n * n_qt;
            const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
            const int l = t * QT + row;
            const bool row_ok = l < L;
            const bool warp_ok = t * QT + lg * 32 < L;           // tcgen05.ld/st are warp-collective
            const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                  ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
            float m_ref = 0.f;      // reference maximum (raw score units) the exponentials are taken against
            float sum = 0.f;
            for (int kb = 0; kb < nkb; ++kb, ++ks) {
                const int key0 = kb * kl;
                const int len = (L16 - key0) < kl ? (L16 - key0) : kl;
                const int vcnt = (L - key0) < len ? (L - key0) : len;   // valid keys of this block (>= 1)
                mbar_wait(&bar_s[s], ks & 1);
                tc_fence_after();
                if (warp_ok) {
                    uint32_t r[KBMAX];
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (c * 16 < len) tmem_ld16(t_s + c * 16, &r[c * 16]);
                    tmem_ld_wait();
                    if (vcnt < len) {       // padded keys of the last block: -inf (max ignores them, exp2 gives 0)
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            if (c * 16 < len && c * 16 + 16 > vcnt) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (c * 16 + j >= vcnt) r[c * 16 + j] = 0xff800000u;
                            }
                        }
                    }
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * 16 < len) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                m0 = fmaxf(m0, __uint_as_float(r[c * 16 + j]));
                                m1 = fmaxf(m1, __uint_as_float(r[c * 16 + j + 1]));
                                m2 = fmaxf(m2, __uint_as_float(r[c * 16 + j + 2]));
                                m3 = fmaxf(m3, __uint_as_float(r[c * 16 + j + 3]));
                            }
                        }
                    }
                    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    if (kb == 0) {
                        m_ref = mx;
                    } else if (__any_sync(0xffffffffu, (mx - m_ref) * c2 > RESCALE_LOG2)) {
                        // rare: move the reference and rescale what has been accumulated (O of the blocks before is
                        // complete: this block's S was issued behind their PV MMAs and has been waited for)
                        const float m_new = fmaxf(m_ref, mx);
                        const float alpha = ex2_approx((m_ref - m_new) * c2);
                        m_ref = m_new;
                        sum *= alpha;
#pragma unroll
                        for (int hh = 0; hh < 4; ++hh) {
                            uint32_t o[16];
                            tmem_ld16(t_o + hh * 16, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                            tmem_st16(t_o + hh * 16, o);
                        }
                    }
                    const float mxs = m_ref * c2;
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * 16 < len) {
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float p0 = fmaf(__uint_as_float(r[c * 16 + 2 * j]), c2, -mxs);
                                float p1 = fmaf(__uint_as_float(r[c * 16 + 2 * j + 1]), c2, -mxs);
                                if (!(a.diag & 1)) {
                                    p0 = ex2_approx(p0);
                                    p1 = ex2_approx(p1);
                                }
                                s0 += p0;
                                s1 += p1;
                                if (EDIT && cs != nullptr) {   // p2p column re-weighting, after the (unscaled) row sum
                                    const int key = key0 + c * 16 + 2 * j;
                                    p0 *= (key < L) ? __ldg(cs + key) : 0.f;
                                    p1 *= (key + 1 < L) ? __ldg(cs + key + 1) : 0.f;
                                }
                                w[j] = Op16<OPD>::pack(p0, p1);
                            }
                            tmem_st8(t_s + 8 * c, w);
                        }
                    }
                    sum += s0 + s1;
                    tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(&p_bar[s]);
            }
            // ---- read-out: O / sum -> out16 ----
            mbar_wait(&bar_o[s], jt & 1);
            tc_fence_after();
            if (warp_ok) {
                uint32_t o[2][32];
                tmem_ld32(t_o, o[0]);
                tmem_ld32(t_o + 32, o[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[s]);
                const float inv = 1.0f / sum;
                uint16_t* op = reinterpret_cast<uint16_t*>(a.out16) +
                               (static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        u[j] = Op16<OPD>::pack(__uint_as_float(o[hh][2 * j]) * inv, __uint_as_float(o[hh][2 * j + 1]) * inv);
                    if (row_ok) {
                        st_global_v8_b32(op + hh * 32, u);
                        st_global_v8_b32(op + hh * 32 + 16, u + 8);
                    }
                }
            } else {
                tc_fence_before();
                mbar_arrive(&o_free[s]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CTRL_WARP) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int smem_bytes_for(int L) {
    const int L16 = (L + 15) & ~15;
    return NQ * QTILE_BYTES + 4 * L16 * 128 + 1024;
}

}  // namespace

cudaError_t attention3_configure() {
    const int bytes = smem_bytes_for(MAX_L3);
    cudaError_t e;
#define USP_ATTN3_CFG(O, E)                                                                                      \
    if ((e = cudaFuncSetAttribute(attention3_kernel<O, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != \
        cudaSuccess)                                                                                             \
        return e;
    USP_ATTN3_CFG(OPD_FP16, false) USP_ATTN3_CFG(OPD_FP16, true)
    USP_ATTN3_CFG(OPD_BF16, false) USP_ATTN3_CFG(OPD_BF16, true)
#undef USP_ATTN3_CFG
    return cudaSuccess;
}

bool attention3_supported(const AttnArgs& a) {
    static int off = -1;     // USP_ATTN_V3=0 falls back to the one-tile-in-flight kernel (A/B comparison)
    if (off < 0) {
        const char* e = getenv("USP_ATTN_V3");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    return !off && a.L >= 1 && a.L <= MAX_L3;
}

cudaError_t launch_attention3(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v, const AttnArgs& a,
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
    const bool edit = a.vscale != nullptr;
    const int smem = smem_bytes_for(a.L);
#define USP_ATTN3_LAUNCH(O, E) \
    return launch_pdl(attention3_kernel<O, E>, dim3(grid), dim3(THREADS), smem, s, q, k, v, a2)
    if (a.opd == OPD_FP16) { if (edit) USP_ATTN3_LAUNCH(OPD_FP16, true); USP_ATTN3_LAUNCH(OPD_FP16, false); }
    if (edit) USP_ATTN3_LAUNCH(OPD_BF16, true);
    USP_ATTN3_LAUNCH(OPD_BF16, false);
#undef USP_ATTN3_LAUNCH
}

}  // namespace usp
