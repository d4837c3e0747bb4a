// Persistent two-tile attention (v3): one CTA per SM, TWO 128-query tiles in flight, each owned by its own softmax
// warpgroup (one thread per query row), ping-ponging on the MUFU pipe while the tensor core serves the other tile.
//
// Replaces F.scaled_dot_product_attention (libs/uvit.py:95) / the math branch (libs/uvit_t2i.py:91-107), head_dim 64.
//
// Why: v2 (attention2.cu) keeps one tile in flight per CTA, so S-MMA -> softmax -> PV-MMA of a tile run one after the
// other (ncu r01g: tensor pipe 14 %, 3.9 us of barrier hand-offs per item).  A whole score row of L = 257 keys (272
// fp32 TMEM columns) cannot be held twice in the 512 TMEM columns, so the keys are split into blocks of <= 96
// (L = 257: 96 + 96 + 80, L = 334: 4 x 96 / 48) with a flash-style running maximum, and each tile gets half of TMEM:
//
//   slot s (columns [256 s, 256 s + 256)):  score blocks fp32 in a 192-column window, O fp32 in [192, 256).
//   Block q of the slot (counted across its tiles) lives at window offset 48 (q % 3), 96 columns wide; the 16-bit P
//   words of the block are written over the FIRST half of its own columns once the scores sit in registers.
//   So S(q+1) = [off + 48, off + 144) never touches P(q) = [off, off + 48): the control warp issues S(q+1) as soon
//   as the softmax threads have LOADED S(q) (and PV(q-1), whose P words S(q+1) does overwrite, has been issued ahead
//   of it on the in-order tensor pipe) - the next score block is computed while the exponentials of this one run,
//   and no MMA latency sits on a warpgroup's critical path except at the read-out.
//   (First version of this file: one 144-column block in place, S(q+1) only after PV(q): the softmax warps spent
//   22 % of their samples waiting for S and 7 % for O, 64.8 us at (64,16,257) - no better than v2's 66.5.)
//
// Per tile and key block: S = Q K_blk^T (tcgen05, smem x smem) -> the row's thread loads its 96 scores into
// registers, max / exp2 / sum, packs P into TMEM -> O += P V_blk (tcgen05, A operand from TMEM, V MN-major from its
// natural [L,64] layout).  The running maximum is only replaced (and O rescaled in TMEM) when a block's maximum
// exceeds the reference by more than 2^8 (the exponentials then stay <= 256, inside the 16-bit operand's range; the
// row sum uses the same reference) - mathematically identical to the exact-maximum form.
// The exponential sections of the two warpgroups strictly alternate (named barriers, FA3-style ping-pong): one
// warpgroup owns the MUFU pipe while the other one loads its next scores and takes the row maxima.
//
// K and V of an item are TMA-loaded once by a loader warp (double-buffered across items), Q tiles rotate through 3
// buffers (two in flight + one prefetched).  The 1-2 leftover query rows of L = 257 / 258 are computed by four
// otherwise idle warps (12-15) on CUDA cores from the resident K/V tiles; as a masked third MMA tile they cost a whole
// tile's chain of hand-offs per item (66 us against 45 us at L = 256), on CUDA cores 6 us.
//
// What was tried on the way (event traces with USP_ATTN_TRACE, ncu r02c-r02e; all at (64,16,L=256/257), isolated):
//   one 144-key block in place, S(q+1) behind PV(q), polling control warp ........ 64.8 us (v2: 66.5)
//   + staggered 96-key windows, MMA warp per slot, compile-time block bodies ....... 47 us at L = 256
//   + exponential sections strictly alternating (named barriers) ................... 44.6 us (per-scheduler or
//     per-warpgroup MUFU mutexes instead: 46.6 / 51.2 us - the alternation, not the exclusion, is what helps)
//   two threads per row (16 softmax warps, row maximum exchanged through smem) ..... 58 us - slower, not kept
//   read-out by four dedicated warps (12-15) instead of the softmax warpgroups ..... 41.6 us at L = 256, 128 us at
//     L = 334 (from 45.9 / 136) but 60 us at L = 257 (from 54): the two slots' read-outs serialise behind the leftover
//     rows; as a hybrid (dedicated warps only when there are no leftover rows) 44.9 / 55.9 / 134.4 - not kept
//   640 threads (own warpgroups for read-out and leftover rows) .................... 51.5 us: setmaxnreg only
//     redistributes the CTA's LAUNCH allocation (640 x 96 = 61440 registers; asking for more in total hangs), which
//     leaves the softmax threads 144 registers and spills
//   one mbarrier arrival per warp (lane 0 behind __syncwarp) instead of per thread .. no change (48.8 vs 48.4 us)
//   handing the exponential turn over 1-4 chunks before the end of a section ........ no change (49.3-50.2 vs 49.6 us)
//   a quarter / half of the exponentials on the FMA pipe (Cody-Waite + degree-3 polynomial, 9 instructions each;
//     parity green) ................................................................... 54.3 -> 57.7 / 64.8 us at L = 257:
//     the kernel is issue- and latency-bound (45 % issue slots, 35 % MUFU pipe in ncu), not MUFU-bound
//   THREE tiles in flight (48-key blocks, 3 x 160 TMEM columns, 640 threads, turns in a ring of three) ..... 57.5 us at
//     L = 256, 60.3 us at L = 257 against 46.8 / 54.0 in the same run: twice the hand-offs per tile cost more than the
//     third warpgroup's slack buys (and 104 registers per softmax thread spill)
// Every phase of a warpgroup's chain is latency-bound (TMEM load of 96 columns ~400 cycles = 128 B/clk/SM, maxima
// 250, exponentials 950, hand-offs ~500 per block, read-out ~1700 per tile): 12k cycles per item against a MUFU floor
// of 4.4k.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int THREADS = 512;            // 2 softmax warpgroups | 2 MMA-issue warps + loader warp | 4 tail-row warps
constexpr int MMA_WARP0 = 8;            // warps 8, 9: MMA issue (and Q loads) of slot 0, 1
constexpr int LOAD_WARP = 10;           // warp 10: K/V loads (warp 11 idle)
constexpr int TAIL_WARP0 = 12;          // warps 12-15: leftover query rows on CUDA cores
constexpr int QTILE_BYTES = QT * HD * 2;   // 16 KiB
constexpr int NQ = 3;                   // Q tile buffers
constexpr int KBMAX = 96;               // keys per score block = fp32 score registers per softmax thread
constexpr int KSTEP = KBMAX / 2;        // window offset step between consecutive blocks
constexpr int MAX_L3 = 336;
constexpr int TMEM_COLS = 512;
constexpr int SLOT_COLS = 256;
constexpr int O_OFF = 192;
constexpr int TAIL_MAX = 2;             // leftover query rows handled on CUDA cores (more run as a masked extra tile)
constexpr int TAIL_THREADS = 128;       // warps 12-15
constexpr int TAIL_KEYS = 3;            // keys per tail thread (L <= 258 whenever there are tail rows: MAX_L3 = 336)
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

// event timeline of CTA 0 (build with -DUSP_ATTN_TRACE_BUILD, run with USP_ATTN_TRACE=<file>): role r = 0 / 1 softmax
// thread 0 of slot 0 / 1, 2 / 3 MMA warp of slot 0 / 1.  Compiled out by default: the checks sit on the hand-off paths.
constexpr int TRACE_N = 2048;
#ifdef USP_ATTN_TRACE_BUILD
#define USP_TR(role, tag)                                                                              \
    do {                                                                                               \
        if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_i < TRACE_N)                      \
            a.trace[(role) * TRACE_N + tr_i++] = (static_cast<unsigned long long>(tag) << 48) |        \
                                                 (static_cast<unsigned long long>(clock64()) & 0xffffffffffffull); \
    } while (0)
#else
#define USP_TR(role, tag) do { } while (0)
#endif

constexpr int STAGE_BYTES = 8 * 4096;
constexpr int MAX_DYN_SMEM = 232448 - 3072;     // opt-in limit minus static shared memory
__host__ __device__ inline bool stage_fits(int L16) {
    return NQ * QTILE_BYTES + 4 * L16 * 128 + 1024 + STAGE_BYTES <= MAX_DYN_SMEM;
}

// key blocks: nkb blocks of `kl` keys (multiple of 16, <= KBMAX), the last one shorter
__host__ __device__ inline int blocks_nkb(int L16) { return (L16 + KBMAX - 1) / KBMAX; }
__host__ __device__ inline int blocks_len(int L16) {
    const int nkb = blocks_nkb(L16);
    return (((L16 + nkb - 1) / nkb) + 15) & ~15;
}

// One score block of one query row (the thread's TMEM lane): NC chunks of 16 keys, all loop bounds compile-time so that
// the whole block is ONE basic block - ptxas then software-pipelines the FFMA / MUFU / FADD / pack streams across
// chunks.  (With a run-time chunk count every chunk was its own basic block: FFMAs, then 16 MUFUs, then the adds -
// the exponential section ran at 14.6 cycles per key against the MUFU pipe's 8; event trace r02c.)
//   loads S -> registers, releases the S window (s_free), row maximum, lazy reference update (+ O rescale),
//   p = 2^(s c2 - m c2), row sum, 16-bit P words back into the first half of the window.
template <int OPD, bool EDIT, int NC>
__device__ __forceinline__ void softmax_block(uint32_t t_s, uint32_t t_o, int vcnt, bool first, float& m_ref, float& sum,
                                              const float* cs, int key0, int L, uint64_t* s_free_bar, uint64_t* pv_bar,
                                              uint32_t pv_parity, int lock_slot) {
    const float c2 = 0.125f * 1.44269504088896340736f;  // hd^-0.5 * log2(e)
    uint32_t r[NC * 16];
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld16(t_s + c * 16, &r[c * 16]);
    tmem_ld_wait();
    // the scores sit in registers: the MMA warp may issue the next block's S-MMA over this window
    tc_fence_before();
    mbar_arrive(s_free_bar);
    if (vcnt < NC * 16) {       // padded keys (only ever in the last chunk): -inf - the maximum ignores them, 2^-inf = 0
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if ((NC - 1) * 16 + j >= vcnt) r[(NC - 1) * 16 + j] = 0xff800000u;
    }
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < NC * 16; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(r[i]));
        m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
    }
    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (first) {
        m_ref = mx;
    } else if (__any_sync(0xffffffffu, (mx - m_ref) * c2 > RESCALE_LOG2)) {
        // rare: move the reference and rescale what has been accumulated.  PV of the previous block (same tile) must
        // have retired: pv_bar completes once per PV block of this slot and is at most one completion behind here
        // (PV of the block before that was issued ahead of this block's S-MMA)
        mbar_wait(pv_bar, pv_parity);
        tc_fence_after();
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
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    // exponents x = s c2 - m c2 in place, BEFORE the turn is taken: ptxas otherwise hoists these 16 NC FFMAs to the head
    // of the exponential section, where the MUFU pipe idles behind them while the other warpgroup waits for the turn
    // (the empty asm pins the values to this point)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[c * 16 + j] = __float_as_uint(fmaf(__uint_as_float(r[c * 16 + j]), c2, -mxs));
        asm volatile("" : "+r"(r[c * 16]), "+r"(r[c * 16 + 1]), "+r"(r[c * 16 + 2]), "+r"(r[c * 16 + 3]), "+r"(r[c * 16 + 4]),
                          "+r"(r[c * 16 + 5]), "+r"(r[c * 16 + 6]), "+r"(r[c * 16 + 7]), "+r"(r[c * 16 + 8]), "+r"(r[c * 16 + 9]),
                          "+r"(r[c * 16 + 10]), "+r"(r[c * 16 + 11]), "+r"(r[c * 16 + 12]), "+r"(r[c * 16 + 13]),
                          "+r"(r[c * 16 + 14]), "+r"(r[c * 16 + 15]));
    }
    // ping-pong: the exponential sections of the two warpgroups alternate (named barriers 4 / 5)
    if (lock_slot == 0) asm volatile("bar.sync 4, 256;" ::: "memory");
    else if (lock_slot == 1) asm volatile("bar.sync 5, 256;" ::: "memory");
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const float p0 = ex2_approx(__uint_as_float(r[c * 16 + 2 * j]));
            const float p1 = ex2_approx(__uint_as_float(r[c * 16 + 2 * j + 1]));
            const float p2 = ex2_approx(__uint_as_float(r[c * 16 + 2 * j + 2]));
            const float p3 = ex2_approx(__uint_as_float(r[c * 16 + 2 * j + 3]));
            s0 += p0;
            s1 += p1;
            s2 += p2;
            s3 += p3;
            if (EDIT && cs != nullptr) {   // p2p column re-weighting, after the (unscaled) row sum
                const int key = key0 + c * 16 + 2 * j;
                w[j] = Op16<OPD>::pack(p0 * ((key < L) ? __ldg(cs + key) : 0.f),
                                       p1 * ((key + 1 < L) ? __ldg(cs + key + 1) : 0.f));
                w[j + 1] = Op16<OPD>::pack(p2 * ((key + 2 < L) ? __ldg(cs + key + 2) : 0.f),
                                           p3 * ((key + 3 < L) ? __ldg(cs + key + 3) : 0.f));
            } else {
                w[j] = Op16<OPD>::pack(p0, p1);
                w[j + 1] = Op16<OPD>::pack(p2, p3);
            }
        }
        tmem_st8(t_s + 8 * c, w);
    }
    if (lock_slot == 0) asm volatile("bar.arrive 5, 256;" ::: "memory");
    else if (lock_slot == 1) asm volatile("bar.arrive 4, 256;" ::: "memory");
    sum += (s0 + s1) + (s2 + s3);
    tmem_st_wait();
}

template <int OPD, bool EDIT>
__global__ void __launch_bounds__(THREADS, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t q_full[NQ], kv_full[2], kv_free[2], bar_s[2], s_free[2], p_bar[2][2], bar_pv[2],
        bar_o[2], o_free[2];
    __shared__ float t_q[HD], t_p[MAX_L3 + 16], t_red[2][4], t_o[4][HD];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int tr_i = 0;
    (void)tr_i;
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
    // read-out staging (4 KiB per softmax warp) when it fits: rows leave as full 128-byte lines instead of 32-byte
    // pieces of 32 different rows per store instruction
    const bool staged = stage_fits(L16);
    uint8_t* sStage = sV + 2 * kv_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NQ; ++i) mbar_init(&q_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_free[i], n_qt + (tail_simt ? 1 : 0));   // one commit per tile (its last PV) + the tail warps
            mbar_init(&bar_s[i], 1);
            mbar_init(&s_free[i], QT);
            mbar_init(&p_bar[i][0], QT);      // P(b) arrives on p_bar[slot][b & 1]: a warp that runs one block ahead of its
            mbar_init(&p_bar[i][1], QT);      // warpgroup (idle warps of a partial tile do) must not land in block b's phase
            mbar_init(&bar_pv[i], 1);
            mbar_init(&bar_o[i], 1);
            mbar_init(&o_free[i], QT);
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == MMA_WARP0) tmem_alloc<TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    pdl_launch();

    if (warp >= TAIL_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        // ============ warps 12-15: the leftover query rows (L = 257 / 258) on CUDA cores, from the resident K/V tiles ======
        // (as a masked third MMA tile they cost a whole tile's chain of hand-offs per item: 66 us against 45 at L = 256)
        const int tt = threadIdx.x - TAIL_WARP0 * 32;   // 0..127
        const float c2 = 0.125f * 1.44269504088896340736f;
        for (int n = 0; tail_simt && n < my_items; ++n) {
            const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
            {
                const float* cs = (EDIT && a.vscale != nullptr && (a.st == nullptr || a.st->attn_on != 0))
                                      ? a.vscale + static_cast<long long>(bh / a.H) * L : nullptr;   // p2p column weights
                const uint8_t* kbuf = sK + (n & 1) * kv_bytes;
                const uint8_t* vbuf = sV + (n & 1) * kv_bytes;
                mbar_wait_idle(&kv_full[n & 1], (n >> 1) & 1);     // (a whole item away: back off between polls)
                for (int l = n_qt * QT; l < ((a.diag & 32) ? 0 : L); ++l) {   // diag 32: skip the tail rows' arithmetic
                    if (tt < HD / 2) {
                        const uint32_t w = reinterpret_cast<const uint32_t*>(a.q16)[(static_cast<long long>(bh) * L + l) * (HD / 2) + tt];
                        const float2 f = Op16<OPD>::unpack(w);
                        t_q[2 * tt] = f.x;
                        t_q[2 * tt + 1] = f.y;
                    }
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    // scores for keys tt, tt+64, ... (log2 domain)
                    float x[TAIL_KEYS];
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < TAIL_KEYS; ++i) {
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
                    if (lane == 0) t_red[0][warp - TAIL_WARP0] = mx;
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    mx = fmaxf(fmaxf(t_red[0][0], t_red[0][1]), fmaxf(t_red[0][2], t_red[0][3]));
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < TAIL_KEYS; ++i) {
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
                    if (lane == 0) t_red[1][warp - TAIL_WARP0] = sum;
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    const float inv = 1.0f / ((t_red[1][0] + t_red[1][1]) + (t_red[1][2] + t_red[1][3]));
                    // O[d] = sum_j p_j V[j][d]: thread = (pair of d, every fourth key)
                    const int dp = tt & 31, seg = tt >> 5;
                    float o0 = 0.f, o1 = 0.f;
                    for (int j = seg; j < L; j += 4) {
                        const uint8_t* vrow = vbuf + (j >> 3) * 1024 + (j & 7) * 128;
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(vrow + ((((dp >> 2) ^ (j & 7))) << 4) + (dp & 3) * 4);
                        const float2 f = Op16<OPD>::unpack(w);
                        const float p = t_p[j];
                        o0 = fmaf(p, f.x, o0);
                        o1 = fmaf(p, f.y, o1);
                    }
                    t_o[seg][2 * dp] = o0;
                    t_o[seg][2 * dp + 1] = o1;
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (tt < HD / 2) {
                        const float r0 = ((t_o[0][2 * tt] + t_o[1][2 * tt]) + (t_o[2][2 * tt] + t_o[3][2 * tt])) * inv;
                        const float r1 = ((t_o[0][2 * tt + 1] + t_o[1][2 * tt + 1]) + (t_o[2][2 * tt + 1] + t_o[3][2 * tt + 1])) * inv;
                        reinterpret_cast<uint32_t*>(a.out16)[((static_cast<long long>(bh / a.H) * L + l) * a.D + (bh % a.H) * HD) / 2 + tt] =
                            Op16<OPD>::pack(r0, r1);
                    }
                    asm volatile("bar.sync 2, 128;" ::: "memory");   // t_q / t_p / t_o are reused by the next row
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");       // every tail thread has left this K/V buffer
            }
            if (tt == 0) mbar_arrive(&kv_free[n & 1]);     // (the barrier above: every tail thread has left the buffer)
        }
    } else if (warp >= LOAD_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // ============ warp 10: K/V loads (one lane) ============
        if (warp == LOAD_WARP && lane == 0) {
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            for (int m = 0; m < my_items; ++m) {
                const int nb = m & 1;
                const int bh = static_cast<int>(blockIdx.x) + m * static_cast<int>(gridDim.x);
                // buffer m & 1 was last used by item m - 2: its MMAs have retired and its tail rows are done
                if (m >= 2) mbar_wait_idle(&kv_free[nb], ((m >> 1) - 1) & 1);
                mbar_expect_tx(&kv_full[nb], 2 * kv_bytes);
                for (int c = 0; c < 2; ++c) {
                    tma_load_3d(&tmK, &kv_full[nb], sK + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh);
                    tma_load_3d(&tmV, &kv_full[nb], sV + nb * kv_bytes + c * hrows * 128, 0, c * hrows, bh);
                }
            }
        }
        __syncwarp();
    } else if (warp >= MMA_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // ============ warps 8 / 9: MMA issue for slot 0 / 1 (+ the Q tile loads); warp-uniform, one elected lane acts =====
        // Per slot the softmax warpgroup's events come in a fixed order - S(0) loaded, P(0) published, S(1) loaded, ... -
        // so the issuer is a straight sequence of blocking waits (no polling): S(b+1) goes out the moment S(b) sits in
        // registers, PV(b) the moment P(b) is published.  Everything an issue needs (descriptors, TMEM addresses) is
        // computed BEFORE the wait it follows: this warp shares its scheduler with two busy softmax warps and every
        // dependent instruction between wake-up and tcgen05.mma costs ~10 cycles of a warpgroup's critical path
        // (ncu r02c: 1300 cycles of issue code per block with the descriptors built after the wait).
        constexpr int fmt = OPD == OPD_FP16 ? 0 : 1;
        const int s = warp - MMA_WARP0;
        const uint32_t idesc_o = umma_idesc(fmt, QT, HD, 0, 1);
        const uint32_t t_slot = tmem_base + s * SLOT_COLS;
        const uint32_t d_o = t_slot + O_OFF;
        auto load_q = [&](int g, int n, int t) {       // Q tile g = (item n, tile t) of this CTA into buffer g % 3
            if (elect_one()) {
                const int bh = static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
                const int qb = g % NQ;
                mbar_expect_tx(&q_full[qb], QTILE_BYTES);
                tma_load_3d(&tmQ, &q_full[qb], sQ + qb * QTILE_BYTES, 0, t * QT, bh);
            }
            __syncwarp();
        };
        auto issue_s = [&](uint32_t d_s, uint64_t qdesc, uint64_t kdesc, uint32_t idesc) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16(d_s, qdesc + (k * 2), kdesc + (k * 2), idesc, k != 0);
                umma_commit(&bar_s[s]);
            }
            __syncwarp();
        };
        if (lane == 0) tma_prefetch_desc(&tmQ);
        __syncwarp();
        if (s == 0) {
            int n = 0, t = 0;
            for (int g = 0; g < NQ && g < n_tiles; ++g) {
                load_q(g, n, t);
                if (++t == n_qt) { t = 0; ++n; }
            }
        }
        const int len_last = L16 - (nkb - 1) * kl;                    // keys of a tile's last block
        const uint32_t idesc_full = umma_idesc(fmt, QT, nkb > 1 ? kl : len_last, 0, 0);
        const uint32_t idesc_last = umma_idesc(fmt, QT, len_last, 0, 0);
        const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK));       // + (item buffer) kv_bytes / 16 + key0 * 8
        const uint64_t vdesc0 = umma_desc_v_mn(smem_u32(sV));
        const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ));       // + (g % 3) * QTILE_BYTES / 16
        // running position of this slot: tile g = s + 2 j = (item n, tile t), block b = j nkb + kb, window offset woff
        int g = s, n = s / n_qt, t = s - n * n_qt, gq = s % NQ;
        int b = 0, woff = 0;
        if (g < n_tiles) {
            mbar_wait(&q_full[gq], (g / NQ) & 1);
            mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
            tc_fence_after();
            issue_s(t_slot, qdesc0 + gq * (QTILE_BYTES / 16), kdesc0 + (n & 1) * (kv_bytes / 16), idesc_full);
        }
        for (int j = 0; g < n_tiles; ++j) {
            const uint64_t qdesc = qdesc0 + gq * (QTILE_BYTES / 16);
            const uint64_t kdesc_item = kdesc0 + (n & 1) * (kv_bytes / 16);
            const uint64_t vdesc_item = vdesc0 + (n & 1) * (kv_bytes / 16);
            // the slot's next tile g + 2 = (item n2, tile t2), Q buffer gq2
            const int g2 = g + 2;
            int n2 = n, t2 = t + 2;
            while (t2 >= n_qt) { t2 -= n_qt; ++n2; }
            const int gq2 = gq + 2 >= NQ ? gq + 2 - NQ : gq + 2;
            bool next_issued = false;
            for (int kb = 0; kb < nkb; ++kb, ++b) {
                const int woff_next = woff == 2 * KSTEP ? 0 : woff + KSTEP;
                const bool last = kb == nkb - 1;
                // -- prepared ahead of the waits --
                const uint64_t kdesc_next = kdesc_item + (kb + 1) * kl * 8;
                const uint32_t idesc_next = (kb + 2 == nkb) ? idesc_last : idesc_full;
                const uint64_t vdesc = vdesc_item + kb * kl * 8;
                const int nkk = (last ? len_last : kl) >> 4;
                const uint32_t a_p = t_slot + woff;
                int n3 = n, t3 = t + NQ;                 // tile g + 3 (takes over this tile's Q buffer)
                while (t3 >= n_qt) { t3 -= n_qt; ++n3; }
                // ---- S(b) sits in the softmax threads' registers ----
                mbar_wait(&s_free[s], b & 1);
                tc_fence_after();
                USP_TR(2 + s, 10);
                if (!last) {
                    issue_s(t_slot + woff_next, qdesc, kdesc_next, idesc_next);
                } else {
                    // The next tile's first block goes out right here as well (its window is the rotation's next one: it
                    // overlaps neither P(b) nor anything unread) - provided its Q tile and K/V have landed; a blocking wait
                    // at this point would hold back PV(b).  Otherwise it is issued behind PV(b) below.
                    if (g2 < n_tiles &&
                        __all_sync(0xffffffffu, mbar_test_wait(&q_full[gq2], (g2 / NQ) & 1) &&
                                                    mbar_test_wait(&kv_full[n2 & 1], (n2 >> 1) & 1))) {
                        tc_fence_after();
                        issue_s(t_slot + woff_next, qdesc0 + gq2 * (QTILE_BYTES / 16), kdesc0 + (n2 & 1) * (kv_bytes / 16),
                                idesc_full);
                        next_issued = true;
                    }
                    if (g + NQ < n_tiles) load_q(g + NQ, n3, t3);
                }
                USP_TR(2 + s, 11);
                // ---- P(b) is published: O += P V ----
                mbar_wait(&p_bar[s][b & 1], (b >> 1) & 1);
                tc_fence_after();
                USP_TR(2 + s, 12);
                if (kb == 0 && j > 0) {     // the slot's previous tile has been read out of O
                    mbar_wait(&o_free[s], (j - 1) & 1);
                    tc_fence_after();
                }
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < KBMAX / 16; ++kk)
                        if (kk < nkk) umma_f16_ts(d_o, a_p + 8 * kk, vdesc + kk * 128, idesc_o, (kb | kk) != 0);
                    umma_commit(&bar_pv[s]);
                    if (last) {
                        umma_commit(&bar_o[s]);
                        umma_commit(&kv_free[n & 1]);
                    }
                }
                __syncwarp();
                USP_TR(2 + s, 13);
                woff = woff_next;
            }
            // next tile of this slot
            g = g2; n = n2; t = t2; gq = gq2;
            if (g < n_tiles && !next_issued) {
                mbar_wait(&q_full[gq], (g / NQ) & 1);
                mbar_wait(&kv_full[n & 1], (n >> 1) & 1);
                tc_fence_after();
                issue_s(t_slot + woff, qdesc0 + gq * (QTILE_BYTES / 16), kdesc0 + (n & 1) * (kv_bytes / 16), idesc_full);
            }
        }
    } else {
        // ===================== softmax / output warps: one thread per query row, warpgroup = slot =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int s = warp >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t t_slot = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + s * SLOT_COLS;
        const uint32_t t_o = t_slot + O_OFF;
        int ks = 0;       // score blocks consumed by this slot
        int jt = 0;       // tiles finished by this slot
        // Ping-pong: the exponential sections of the two warpgroups strictly alternate (named barriers 4 / 5, FA3
        // style), so that one warpgroup owns the MUFU pipe while the other one reads its next S block and takes the
        // row maxima.  Left alone the two tiles drift into phase and compute at half MUFU rate each.
        // Both warpgroups run the same number of sections (the one with fewer tiles pads with empty ones).
        const int my_secs = ((n_tiles + 1 - s) / 2) * nkb;
        const int all_secs = ((n_tiles + 1) / 2) * nkb;
        const bool lock = (a.diag & 64) == 0;   // diag 64: no ping-pong lock around the exponential sections (A/B comparison)
        if (lock && s == 1) asm volatile("bar.arrive 4, 256;" ::: "memory");   // warpgroup 0 goes first
        // Read-out of a finished tile: O / sum -> out16.  It is DEFERRED to right after the next tile's first block has
        // been published: by then PV of the tile's last block has long retired (no wait), and the ~1.5k cycles of
        // TMEM loads / scaling / stores run while the other warpgroup holds its exponential turn instead of stalling the
        // alternation at every tile end.  The MMA warp holds PV(next tile, block 0) back until o_free.
        auto read_out = [&](int p_jt, int p_bh, int p_t, bool p_row_ok, bool p_warp_ok, float p_sum) {
            const int l0 = p_t * QT + lg * 32;
            if (lg == 0) USP_TR(s, 6);
            mbar_wait(&bar_o[s], p_jt & 1);
            tc_fence_after();
            if (lg == 0) USP_TR(s, 7);
            if (p_warp_ok) {
                uint32_t o[2][32];
                tmem_ld32(t_o, o[0]);
                tmem_ld32(t_o + 32, o[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&o_free[s]);
                const float inv = 1.0f / p_sum;
                uint16_t* op = reinterpret_cast<uint16_t*>(a.out16) +
                               (static_cast<long long>(p_bh / a.H) * L + l0 + lane) * a.D + (p_bh % a.H) * HD;
                uint8_t* stg = sStage + warp * 4096;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        u[j] = Op16<OPD>::pack(__uint_as_float(o[hh][2 * j]) * inv, __uint_as_float(o[hh][2 * j + 1]) * inv);
                    if (staged) {
                        // own row (128 B) into the warp's staging tile, 16-byte units XOR-swizzled by (row & 7)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + q) ^ (lane & 7)) << 4)) =
                                make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
                    } else if (p_row_ok) {
                        st_global_v8_b32(op + hh * 32, u);
                        st_global_v8_b32(op + hh * 32 + 16, u + 8);
                    }
                }
                if (staged) {
                    __syncwarp();
                    // 8 lanes per row: every store instruction writes 4 complete 128-byte rows
                    uint8_t* ob = reinterpret_cast<uint8_t*>(a.out16) +
                                  ((static_cast<long long>(p_bh / a.H) * L + l0) * a.D + (p_bh % a.H) * HD) * 2 + (lane & 7) * 16;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + (lane >> 3);
                        const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                        if (l0 + rr < L) *reinterpret_cast<uint4*>(ob + static_cast<long long>(rr) * a.D * 2) = v;
                    }
                    __syncwarp();
                }
                if (lg == 0) USP_TR(s, 8);
            } else {
                tc_fence_before();
                mbar_arrive(&o_free[s]);
            }
        };
        bool have_prev = false;
        int p_jt = 0, p_bh = 0, p_t = 0;
        bool p_row_ok = false, p_warp_ok = false;
        float p_sum = 1.f;
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
                const uint32_t t_s = t_slot + KSTEP * (ks % 3);
                if (lg == 0) USP_TR(s, 0);
                if (a.diag & 128) {     // experiment: non-suspending poll
                    while (!mbar_test_wait(&bar_s[s], ks & 1)) {}
                } else {
                    mbar_wait(&bar_s[s], ks & 1);
                }
                tc_fence_after();
                if (lg == 0) USP_TR(s, 1);
                if (warp_ok) {
#define USP_SM_BLOCK(NC)                                                                                        \
    softmax_block<OPD, EDIT, NC>(t_s, t_o, vcnt, kb == 0, m_ref, sum, cs, key0, L, &s_free[s], &bar_pv[s], (ks - 1) & 1, \
                                 lock ? s : -1)
                    switch (len >> 4) {
                        case 6: USP_SM_BLOCK(6); break;
                        case 5: USP_SM_BLOCK(5); break;
                        case 4: USP_SM_BLOCK(4); break;
                        case 3: USP_SM_BLOCK(3); break;
                        case 2: USP_SM_BLOCK(2); break;
                        default: USP_SM_BLOCK(1); break;
                    }
#undef USP_SM_BLOCK
                    if (lg == 0) USP_TR(s, 5);
                } else {
                    tc_fence_before();
                    mbar_arrive(&s_free[s]);
                    if (lock) {
                        if (s == 0) asm volatile("bar.sync 4, 256;\n\tbar.arrive 5, 256;" ::: "memory");
                        else asm volatile("bar.sync 5, 256;\n\tbar.arrive 4, 256;" ::: "memory");
                    }
                }
                tc_fence_before();
                mbar_arrive(&p_bar[s][ks & 1]);
                if (kb == 0 && have_prev) {
                    read_out(p_jt, p_bh, p_t, p_row_ok, p_warp_ok, p_sum);
                    have_prev = false;
                }
            }
            have_prev = true;
            p_jt = jt; p_bh = bh; p_t = t; p_row_ok = row_ok; p_warp_ok = warp_ok; p_sum = sum;
        }
        if (have_prev) read_out(p_jt, p_bh, p_t, p_row_ok, p_warp_ok, p_sum);
        for (int i = my_secs; lock && i < all_secs; ++i) {     // empty sections: keep the alternation going for the other warpgroup
            if (s == 0) asm volatile("bar.sync 4, 256;\n\tbar.arrive 5, 256;" ::: "memory");
            else asm volatile("bar.sync 5, 256;\n\tbar.arrive 4, 256;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int smem_bytes_for(int L) {
    const int L16 = (L + 15) & ~15;
    return NQ * QTILE_BYTES + 4 * L16 * 128 + 1024 + (stage_fits(L16) ? STAGE_BYTES : 0);
}

}  // namespace

cudaError_t attention3_configure() {
    const int bytes = MAX_DYN_SMEM;
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
    a2.trace = nullptr;
    static const char* trace_path = getenv("USP_ATTN_TRACE");
    static unsigned long long* trace_buf = nullptr;
    static int trace_calls = 0;
    if (trace_path != nullptr) {      // debugging only: the timeline of the 4th call is dumped (synchronises!)
        if (trace_buf == nullptr) cudaMalloc(&trace_buf, 4 * TRACE_N * sizeof(unsigned long long));
        if (++trace_calls == 4) {
            cudaMemsetAsync(trace_buf, 0, 4 * TRACE_N * sizeof(unsigned long long), s);
            a2.trace = trace_buf;
        }
    }
    const bool edit = a.vscale != nullptr;
    const int smem = smem_bytes_for(a.L);
    cudaError_t err;
#define USP_ATTN3_LAUNCH(O, E) \
    err = launch_pdl(attention3_kernel<O, E>, dim3(grid), dim3(THREADS), smem, s, q, k, v, a2)
    if (a.opd == OPD_FP16) { if (edit) USP_ATTN3_LAUNCH(OPD_FP16, true); else USP_ATTN3_LAUNCH(OPD_FP16, false); }
    else { if (edit) USP_ATTN3_LAUNCH(OPD_BF16, true); else USP_ATTN3_LAUNCH(OPD_BF16, false); }
#undef USP_ATTN3_LAUNCH
    if (a2.trace != nullptr && err == cudaSuccess) {
        static unsigned long long host[4 * TRACE_N];
        cudaStreamSynchronize(s);
        cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "w")) {
            for (int r = 0; r < 4; ++r)
                for (int i = 0; i < TRACE_N && host[r * TRACE_N + i] != 0; ++i)
                    fprintf(f, "%d %d %llu\n", r, static_cast<int>(host[r * TRACE_N + i] >> 48),
                            host[r * TRACE_N + i] & 0xffffffffffffull);
            fclose(f);
        }
    }
    return err;
}

}  // namespace usp
