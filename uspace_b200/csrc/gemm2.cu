// 2-CTA (cta_group::2) tcgen05 GEMM: a CTA pair on one TPC computes a 256 x 256 output tile.
//
// Why pairs: with one CTA per 128x256 tile every SM stages 48 KiB of operands per 64-wide K block; in pair mode
// each CTA loads its own 128 rows of A and only HALF of the W tile (128 of the 256 rows) and the UMMA reads both
// halves, so per-SM operand traffic (L2 -> smem and smem -> tensor core) drops to 32 KiB per K block.
//
// Roles per CTA (352 threads):
//   warp 0    : TMA producer for this CTA's A rows and W half; completion lands on the LEADER's full barrier
//   warp 1    : TMEM allocator (both CTAs); in the leader one thread issues tcgen05.mma.cta_group::2 and
//               multicasts its commits to both CTAs' empty / tmem_full barriers
//   warps 2-9 : epilogue, two warps per TMEM lane group (each takes 128 of the 256 accumulator columns)
//   warp 10   : residual epilogue only: TMA-prefetches the residual tile in chunks (128 rows x 32 fp32, 128B swizzle)
//               three chunks ahead per column half, so the epilogue never waits on DRAM latency (ncu r01: the
//               direct-load epilogue spent 60 % of its samples in long-scoreboard stalls on the residual and ran
//               `proj` at 28 % tensor activity).  A chunk is released back to the loader as soon as its rows are in
//               registers; results leave through full 128-byte per-row global stores.
// 16-bit results (qkv, fc1) leave through per-row 32-byte global stores.  Staging them in shared memory for TMA
// stores was built and measured (round 1): no gain - the cost of the stores is their share of the SM <-> L2 traffic
// (see DESIGN.md section 4), not LSU issue, and the staging buffers cost a ring stage.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int BM = GEMM_BM;       // rows per CTA (256 per pair)
constexpr int BK = GEMM_BK;
constexpr int BN = 256;           // columns per pair; each CTA stages BN/2 rows of W
constexpr int THREADS = 352;
constexpr int EPI_WARPS = 8;
constexpr int LOADER_WARP = 10;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = (BN / 2) * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KiB per CTA per stage
constexpr int TMEM_COLS = 2 * BN;                // double-buffered accumulator
constexpr int CHUNK_BYTES = BM * 32 * 4;         // 128 rows x 32 fp32 = 16 KiB staging chunk

// LONGK (K >= 2048: fc2): the main loop per tile is long, so the residual prefetch needs little depth and the
// operand ring keeps 6 stages; short K (proj) trades two ring stages for a 3-deep residual prefetch per column half.
template <int EPI, bool LONGK>
struct Cfg2 {
    static constexpr bool STAGED = (EPI == EPI_BIAS_RESID);   // residual chunks are TMA-prefetched into smem
    static constexpr int NBUF = LONGK ? 1 : 3;                // staging chunks in flight per column half
    static constexpr int STAGES = STAGED ? (LONGK ? 6 : 4) : 6;
    static constexpr int EPI_BYTES = STAGED ? 2 * NBUF * CHUNK_BYTES : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024;
};


// NP = CTA pairs per cluster.  NP == 2 (cluster of 4): the two pairs compute horizontally adjacent tiles (same 256
// rows of A, different 256 columns of W) and every A box is fetched ONCE and multicast to the two CTAs that need it
// (pair p issues the A load of k-blocks with kb % 2 == p).  The GEMMs run at the chip's L2 -> SM delivery cap
// (~6300 B/clk: 1 MB of operands per 256x256x1024 tile is 12 TB/s at 1530 TFLOP/s), so operand bytes per flop,
// not tensor-pipe rate, set their speed; sharing A removes a quarter of them.
template <int EPI, bool LONGK, int NP>
__global__ void __cluster_dims__(2 * NP, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmR, const GemmArgs g) {
    using Cfg = Cfg2<EPI, LONGK>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool STAGED = Cfg::STAGED;
    constexpr int NBUF = Cfg::NBUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ebuf = smem + STAGES * STAGE_BYTES;             // [2 halves][NBUF][16 KiB]

    __shared__ __align__(8) uint64_t full_bar[STAGES];       // used in the leader only
    __shared__ __align__(8) uint64_t empty_bar[STAGES];      // one per CTA, arrived by the leader's multicast commit
    __shared__ __align__(8) uint64_t tmem_full_bar[2];       // one per CTA, multicast commit
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader only: 2 CTAs x 8 epilogue warps
    __shared__ __align__(8) uint64_t rfull_bar[2][3];     // staging chunk holds the residual (or is writable)
    __shared__ __align__(8) uint64_t rfree_bar[2][3];     // staging chunk's TMA store has finished reading smem
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();     // 0 .. 2*NP-1
    const uint32_t rank = crank & 1;              // position inside the pair
    const int pair = static_cast<int>(crank >> 1);
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / (2 * NP);
    const int n_clusters = gridDim.x / (2 * NP);

    // work items: NP horizontally adjacent tiles per cluster step; this pair takes column tile (item % nn) * NP + pair
    const int n_tiles_n = g.N / BN / NP;
    const int n_tiles_m = (g.M + 2 * BM - 1) / (2 * BM);
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int nkb = g.K / BK;
    const int nkb0 = g.K0 / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        if (STAGED) tma_prefetch_desc(&tmR);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], NP);     // one commit per pair that reads (or shares the A box of) this slot
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * EPI_WARPS);
            for (int j = 0; j < NBUF; ++j) {
                mbar_init(&rfull_bar[i][j], 1);
                mbar_init(&rfree_bar[i][j], 4);   // one arrive per epilogue warp of that column half
            }
        }
        fence_barrier_init();
    }
    __syncwarp();
    // (allocated by the warp that initialised the barriers: compute-sanitizer's racecheck attributes the allocator's
    //  shared-memory write to neighbouring barrier bytes and reported 1154 "hazards" against warp 0's mbarrier.init
    //  when another warp allocated - profiles/r02h_racecheck_gemm.log)
    if (warp == 0) tmem_alloc_cg2<TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    cluster_sync_all();   // peer barriers initialised, both TMEM allocations done
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();     // everything above overlapped the previous kernel's tail
    pdl_launch();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
                const int n0 = ((tile % n_tiles_n) * NP + pair) * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    // the pair leader's barrier: this CTA's own address with the peer bit cleared
                    const uint32_t lead_full = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
                    if (g.diag == 1 && (tile != cluster_id || kb >= STAGES)) {
                        // diagnostic: no operand traffic at all, the MMAs re-read whatever the ring holds
                        if (leader && elect_one()) mbar_arrive(&full_bar[stage]);
                    } else if (elect_one()) {
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        if (g.conv_C > 0) {      // implicit GEMM (launched with NP == 1 only)
                            const int cpt = g.conv_C / BK, tap = kb / cpt;
                            tma_load_im2col_cg2(&tmA0, lead_full, sa, (kb - tap * cpt) * BK,
                                                (m0 % g.conv_W) * g.conv_stride - g.conv_pad,
                                                ((m0 / g.conv_W) % g.conv_H) * g.conv_stride - g.conv_pad,
                                                m0 / (g.conv_W * g.conv_H),
                                                static_cast<uint16_t>(tap % 3), static_cast<uint16_t>(tap / 3));
                        } else if (NP == 1) {
                            if (kb < nkb0)
                                tma_load_2d_cg2(&tmA0, lead_full, sa, kb * BK, m0);
                            else
                                tma_load_2d_cg2(&tmA1, lead_full, sa, (kb - nkb0) * BK, m0);
                        } else if ((kb & 1) == pair) {
                            // this CTA's 128 rows of A are also the rows of the CTA at the same position in the
                            // other pair: one L2 read, two destinations
                            const uint16_t mask = static_cast<uint16_t>(0x5u << rank);
                            if (kb < nkb0)
                                tma_load_2d_cg2_mc(&tmA0, lead_full, sa, kb * BK, m0, mask);
                            else
                                tma_load_2d_cg2_mc(&tmA1, lead_full, sa, (kb - nkb0) * BK, m0, mask);
                        }
                        tma_load_2d_cg2(&tmB, lead_full, sb, kb * BK, n0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = umma_idesc(g.opd == OPD_FP16 ? 0 : 1, 2 * BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int t = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++t) {
                const int as = t & 1;
                const uint32_t aphase = (t >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_cg2(d_tmem, adesc + (k * 2), bdesc + (k * 2), idesc, (kb | k) != 0);
                        // frees the slot in every CTA of the cluster (NP == 2: the other pair's producers multicast
                        // A boxes into this pair's slots, so they count this pair's commit too)
                        umma_commit_cg2(&empty_bar[stage], static_cast<uint16_t>((1u << (2 * NP)) - 1));
                        if (kb == nkb - 1)   // accumulators ready (both CTAs of this pair)
                            umma_commit_cg2(&tmem_full_bar[as], static_cast<uint16_t>(0x3u << (2 * pair)));
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == LOADER_WARP) {
        // ===================== residual prefetcher (fp32 epilogues) =====================
        if (STAGED) {
            int i = 0;  // chunk counter per half (both halves advance together)
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
                const int n0 = ((tile % n_tiles_n) * NP + pair) * BN;
                for (int c = 0; c < 4; ++c, ++i) {
                    const int b = i % NBUF;
                    const uint32_t ph = (i / NBUF) & 1;
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait_idle(&rfree_bar[half][b], ph ^ 1);
                        if (elect_one()) {
                            if (g.diag & 2) {   // diagnostic: no residual traffic
                                mbar_arrive(&rfull_bar[half][b]);
                            } else {
                                mbar_expect_tx(&rfull_bar[half][b], CHUNK_BYTES);
                                tma_load_2d(&tmR, &rfull_bar[half][b], ebuf + (half * NBUF + b) * CHUNK_BYTES,
                                            n0 + (2 * c + half) * 32, m0);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int half = (warp - 2) >> 2;        // which 128 accumulator columns
        constexpr int NCH = BN / 2 / 32;         // 4 chunks of 32 columns per warp
        const int rloc = lg * 32 + lane;         // row inside the CTA's 128-row slab
        int t = 0;
        int i = 0;                               // staging chunk counter (matches the loader's)
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
            const int n0 = ((tile % n_tiles_n) * NP + pair) * BN;   // the two column halves take interleaved chunks, so that at
                                                      // any moment the CTA touches 256 B contiguous per output row
            EpiRow row = epi_row(g, EPI, m0 + rloc);

            mbar_wait_idle(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * BN;
            if (!STAGED) {
                // two TMEM loads in flight per wait: halves the number of exposed TMEM round trips
#pragma unroll 1
                for (int c = 0; c < NCH; c += 2) {
                    const int col = (c + half) * 64;       // 64-column pieces: half 0 -> 0,128; half 1 -> 64,192
                    uint32_t r0[32], r1[32];
                    tmem_ld32(t_row + col, r0);
                    tmem_ld32(t_row + col + 32, r1);
                    tmem_ld_wait();
                    epi_chunk<EPI>(g, row, n0 + col, r0, nullptr);
                    epi_chunk<EPI>(g, row, n0 + col + 32, r1, nullptr);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int col = (2 * c + half) * 32;   // interleaved 32-column chunks (matches the loader)
                    uint32_t r[32];
                    tmem_ld32(t_row + col, r);
                    {
                    const int b = i % NBUF;
                    const uint32_t ph = (i / NBUF) & 1;
                    ++i;
                    uint8_t* buf = ebuf + (half * NBUF + b) * CHUNK_BYTES + rloc * 128;
                    mbar_wait(&rfull_bar[half][b], ph);
                    // residual row out of the prefetched chunk: 128-byte rows, 16-byte units XOR-swizzled by
                    // (row & 7) -> conflict-free 128-bit reads; the chunk is handed back to the loader right away
                    float4 x4[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x4[u] = *reinterpret_cast<const float4*>(buf + ((u ^ (rloc & 7)) << 4));
                    fence_proxy_async();   // generic-proxy reads above vs. the async-proxy (TMA) refill of this chunk
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&rfree_bar[half][b]);
                    tmem_ld_wait();
                    epi_chunk<EPI>(g, row, n0 + col, r, x4);
                    }
                }
            }
            // release this accumulator stage to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], crank & ~1u);
            if (EPI == EPI_BIAS_RESID || EPI == EPI_BIAS_F32) epi_store_stats(g, row, (n0 / 128) + half);
        }
    }

    // no CTA may exit (or free TMEM) while its peer can still touch its shared memory / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
}

// resident clusters of 4 CTAs (a GPC with an odd number of TPCs leaves one idle: 33 on B200, i.e. 132 of 148 SMs)
template <int EPI, bool LONGK>
int max_clusters4() {
    static int cached = -1;
    if (cached >= 0) return cached;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = Cfg2<EPI, LONGK>::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<EPI, LONGK, 2>, &cfg) != cudaSuccess) n = 0;
    cached = n;
    return n;
}

// USP_GEMM_CL4: 0 = never, 1 = whenever the shape allows, 2 (default) = only N <= 1024 (proj / fc2 / skip_linear: their
// 260 tiles fill 3.94 of 4 waves on 33 clusters of 4 instead of 3.51 of 4 on 74 pairs; on the wide GEMMs the 16 idle
// SMs cost more than the shared A saves)
inline int gemm_cluster4_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("USP_GEMM_CL4");
        mode = e ? atoi(e) : 2;
    }
    return mode;
}

template <int EPI, bool LONGK>
cudaError_t launch2k(const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    const int n_tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * (a.N / BN);
    GemmArgs a2 = a;
    const int cl4 = gemm_cluster4_mode();
    if (cl4 && a.conv_C == 0 && (cl4 == 1 || a.N <= 1024) && (a.N / BN) % 2 == 0 && n_tiles >= num_sms) {
        int c4 = max_clusters4<EPI, LONGK>();
        if (c4 > 0) {
            if (n_tiles / 2 < c4) c4 = n_tiles / 2;
            return launch_pdl(gemm2_kernel<EPI, LONGK, 2>, dim3(4 * c4), dim3(THREADS), Cfg2<EPI, LONGK>::SMEM_BYTES, s,
                              maps.a0, maps.a1, maps.b, maps.r32, a2);
        }
    }
    int clusters = num_sms / 2;
    if (n_tiles < clusters) clusters = n_tiles;
    return launch_pdl(gemm2_kernel<EPI, LONGK, 1>, dim3(2 * clusters), dim3(THREADS), Cfg2<EPI, LONGK>::SMEM_BYTES, s,
                      maps.a0, maps.a1, maps.b, maps.r32, a2);
}

template <int EPI>
cudaError_t launch2(const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    if (EPI == EPI_BIAS_RESID && a.K >= 2048) return launch2k<EPI, true>(maps, a, num_sms, s);
    return launch2k<EPI, false>(maps, a, num_sms, s);
}

template <int EPI>
cudaError_t configure2() {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg2<EPI, false>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm2_kernel<EPI, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg2<EPI, false>::SMEM_BYTES);
    if (e != cudaSuccess || EPI != EPI_BIAS_RESID) return e;
    e = cudaFuncSetAttribute(gemm2_kernel<EPI, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg2<EPI, true>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm2_kernel<EPI, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg2<EPI, true>::SMEM_BYTES);
}

}  // namespace

cudaError_t gemm2_configure() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = configure2<EPI_QKV>()) != cudaSuccess) return e;
    if ((e = configure2<EPI_BIAS_GELU>()) != cudaSuccess) return e;
    if ((e = configure2<EPI_BIAS_RESID>()) != cudaSuccess) return e;
    if ((e = configure2<EPI_BIAS_F32>()) != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}

// the residual epilogue needs the fp32 residual tensor map (GemmMaps::r32, has_f32)
bool gemm2_supported(int epi, const GemmMaps& maps, const GemmArgs& a) {
    if (a.N % BN != 0 || a.M <= BM) return false;
    if (epi == EPI_BIAS_RESID) return maps.has_f32 && a.resid != nullptr && a.out32 != nullptr;
    return true;
}

cudaError_t launch_gemm2(int epi, const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    switch (epi) {
        case EPI_QKV: return launch2<EPI_QKV>(maps, a, num_sms, s);
        case EPI_BIAS_GELU: return launch2<EPI_BIAS_GELU>(maps, a, num_sms, s);
        case EPI_BIAS_RESID: return launch2<EPI_BIAS_RESID>(maps, a, num_sms, s);
        case EPI_BIAS_F32: return launch2<EPI_BIAS_F32>(maps, a, num_sms, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace usp
