// 2-CTA (cta_group::2) tcgen05 GEMM: a CTA pair on one TPC computes a 256 x 256 output tile.
//
// Why: with one CTA per 128x256 tile every SM pulls 48 KiB of operands per 64-wide K block from L2
// (85 FLOP/B) and the kernel saturates L2->SM bandwidth at ~60 % tensor-pipe activity (ncu, round 1).
// In pair mode each CTA loads its own 128 rows of A and only HALF of the W tile (128 of the 256 rows);
// the UMMA reads both halves, so the per-SM operand traffic drops to 32 KiB per K block (128 FLOP/B).
//
// Roles per CTA (320 threads):
//   warp 0    : TMA producer for this CTA's A rows and W half; completion lands on the LEADER's full barrier
//   warp 1    : TMEM allocator (both CTAs); in the leader one thread issues tcgen05.mma.cta_group::2 and
//               multicasts its commits to both CTAs' empty / tmem_full barriers
//   warps 2-9 : epilogue, two warps per TMEM lane group (each takes 128 of the 256 accumulator columns),
//               residual rows prefetched one chunk ahead; releases the accumulator to the leader's MMA warp
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int BM = GEMM_BM;       // rows per CTA (256 per pair)
constexpr int BK = GEMM_BK;
constexpr int BN = 256;           // columns per pair; each CTA stages BN/2 rows of W
constexpr int THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = (BN / 2) * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KiB per CTA per stage
constexpr int STAGES = 6;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr int TMEM_COLS = 2 * BN;                // double-buffered accumulator

__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t full_bar[STAGES];       // used in the leader only
    __shared__ __align__(8) uint64_t empty_bar[STAGES];      // one per CTA, arrived by the leader's multicast commit
    __shared__ __align__(8) uint64_t tmem_full_bar[2];       // one per CTA, multicast commit
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader only: 2 CTAs x 8 epilogue warps
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;

    const int n_tiles_n = g.N / BN;
    const int n_tiles_m = (g.M + 2 * BM - 1) / (2 * BM);
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int nkb = g.K / BK;
    const int nkb0 = g.K0 / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * EPI_WARPS);
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc_cg2<TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    cluster_sync_all();   // peer barriers initialised, both TMEM allocations done
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
                const int n0 = (tile % n_tiles_n) * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    const uint32_t lead_full = mapa_rank(smem_u32(&full_bar[stage]), 0);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    if (kb < nkb0)
                        tma_load_2d_cg2(&tmA0, lead_full, sa, kb * BK, m0);
                    else
                        tma_load_2d_cg2(&tmA1, lead_full, sa, (kb - nkb0) * BK, m0);
                    tma_load_2d_cg2(&tmB, lead_full, sb, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
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
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16_cg2(d_tmem, adesc + (k * 2), bdesc + (k * 2), idesc, (kb | k) != 0);
                    umma_commit_cg2(&empty_bar[stage], 0b11);   // frees the slot in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_cg2(&tmem_full_bar[as], 0b11);      // accumulators ready in both CTAs
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int half = (warp - 2) >> 2;        // which 128 accumulator columns
        constexpr int NCH = BN / 2 / 32;         // 4 chunks of 32 columns per warp
        int t = 0;
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
            const int n0 = (tile % n_tiles_n) * BN + half * (BN / 2);
            const EpiRow row = epi_row(g, EPI, m0 + lg * 32 + lane);

            float4 rb[2][8];
            if (EPI == EPI_BIAS_RESID) epi_load_resid(g, row, n0, rb[0]);   // overlaps the wait below

            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * BN + half * (BN / 2);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t r[32];
                tmem_ld32(t_row + c * 32, r);
                if (EPI == EPI_BIAS_RESID && c + 1 < NCH) epi_load_resid(g, row, n0 + (c + 1) * 32, rb[(c + 1) & 1]);
                tmem_ld_wait();
                epi_chunk<EPI>(g, row, n0 + c * 32, r, rb[c & 1]);
            }
            // release this accumulator stage to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
        }
    }

    // no CTA may exit (or free TMEM) while its peer can still touch its shared memory / barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
}

template <int EPI>
cudaError_t launch2(const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    const int n_tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * (a.N / BN);
    int clusters = num_sms / 2;
    if (n_tiles < clusters) clusters = n_tiles;
    gemm2_kernel<EPI><<<2 * clusters, THREADS, SMEM_BYTES, s>>>(maps.a0, maps.a1, maps.b, a);
    return cudaGetLastError();
}

template <int EPI>
cudaError_t configure2() {
    return cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
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

bool gemm2_supported(const GemmArgs& a) { return a.N % BN == 0 && a.M > BM; }

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
