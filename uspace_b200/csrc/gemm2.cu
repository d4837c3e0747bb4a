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
//   warp 10   : fp32 epilogues only: TMA-prefetches the residual tile chunks (128 rows x 32 fp32, 128B swizzle)
//               three chunks ahead, so the epilogue never waits on DRAM latency (ncu r01: the direct-load epilogue
//               spent 60 % of its samples in long-scoreboard stalls on the residual and ran `proj` at 28 % tensor
//               activity).  The epilogue adds accumulator + bias in place in that smem chunk and one thread TMA-stores
//               it (coalesced 128-byte rows, OOB rows clipped by the tensor map).
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
constexpr int NBUF = 3;                          // staging chunks in flight per column half

template <int EPI>
struct Cfg2 {
    static constexpr bool STAGED = (EPI == EPI_BIAS_RESID || EPI == EPI_BIAS_F32);
    static constexpr int STAGES = STAGED ? 4 : 6;
    static constexpr int EPI_BYTES = STAGED ? 2 * NBUF * CHUNK_BYTES : 0;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024;
};

__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmR,
             const __grid_constant__ CUtensorMap tmO, const GemmArgs g) {
    using Cfg = Cfg2<EPI>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool STAGED = Cfg::STAGED;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ebuf = smem + STAGES * STAGE_BYTES;             // [2 halves][NBUF][16 KiB]

    __shared__ __align__(8) uint64_t full_bar[STAGES];       // used in the leader only
    __shared__ __align__(8) uint64_t empty_bar[STAGES];      // one per CTA, arrived by the leader's multicast commit
    __shared__ __align__(8) uint64_t tmem_full_bar[2];       // one per CTA, multicast commit
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader only: 2 CTAs x 8 epilogue warps
    __shared__ __align__(8) uint64_t rfull_bar[2][NBUF];     // staging chunk holds the residual (or is writable)
    __shared__ __align__(8) uint64_t rfree_bar[2][NBUF];     // staging chunk's TMA store has finished reading smem
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
        if (STAGED) {
            tma_prefetch_desc(&tmR);
            tma_prefetch_desc(&tmO);
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 2 * EPI_WARPS);
            for (int j = 0; j < NBUF; ++j) {
                mbar_init(&rfull_bar[i][j], 1);
                mbar_init(&rfree_bar[i][j], 1);
            }
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
        // ===================== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====
        {
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
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        if (kb < nkb0)
                            tma_load_2d_cg2(&tmA0, lead_full, sa, kb * BK, m0);
                        else
                            tma_load_2d_cg2(&tmA1, lead_full, sa, (kb - nkb0) * BK, m0);
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
                        umma_commit_cg2(&empty_bar[stage], 0b11);   // frees the slot in both CTAs
                        if (kb == nkb - 1) umma_commit_cg2(&tmem_full_bar[as], 0b11);  // accumulators ready (both CTAs)
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
                const int n0 = (tile % n_tiles_n) * BN;
                for (int c = 0; c < 4; ++c, ++i) {
                    const int b = i % NBUF;
                    const uint32_t ph = (i / NBUF) & 1;
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait(&rfree_bar[half][b], ph ^ 1);
                        if (elect_one()) {
                            if (EPI == EPI_BIAS_RESID) {
                                mbar_expect_tx(&rfull_bar[half][b], CHUNK_BYTES);
                                tma_load_2d(&tmR, &rfull_bar[half][b], ebuf + (half * NBUF + b) * CHUNK_BYTES,
                                            n0 + half * (BN / 2) + c * 32, m0);
                            } else {
                                mbar_arrive(&rfull_bar[half][b]);
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
        const bool storer = (lg == 0 && lane == 0);
        const int rloc = lg * 32 + lane;         // row inside the CTA's 128-row slab
        int t = 0;
        int i = 0;                               // staging chunk counter (matches the loader's)
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
            const int n0 = (tile % n_tiles_n) * BN + half * (BN / 2);
            const EpiRow row = epi_row(g, EPI, m0 + rloc);

            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * BN + half * (BN / 2);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t r[32];
                tmem_ld32(t_row + c * 32, r);
                if (!STAGED) {
                    tmem_ld_wait();
                    epi_chunk<EPI>(g, row, n0 + c * 32, r, nullptr);
                } else {
                    const int b = i % NBUF;
                    const uint32_t ph = (i / NBUF) & 1;
                    ++i;
                    uint8_t* buf = ebuf + (half * NBUF + b) * CHUNK_BYTES + rloc * 128;
                    mbar_wait(&rfull_bar[half][b], ph);
                    tmem_ld_wait();
                    const int n = n0 + c * 32;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (g.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    }
                    // 128-byte rows, 16-byte units XOR-swizzled by (row & 7): conflict-free 128-bit accesses
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float4* p4 = reinterpret_cast<float4*>(buf + ((u ^ (rloc & 7)) << 4));
                        float4 o = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                        if (EPI == EPI_BIAS_RESID) {
                            const float4 x4 = *p4;
                            o.x += x4.x; o.y += x4.y; o.z += x4.z; o.w += x4.w;
                            v[4 * u] = o.x; v[4 * u + 1] = o.y; v[4 * u + 2] = o.z; v[4 * u + 3] = o.w;
                        }
                        *p4 = o;
                    }
                    if (g.out16 != nullptr && row.ok) {   // 16-bit copy of the new residual stream (next skip_linear operand)
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(g.out16) +
                                                             static_cast<long long>(row.m) * g.N + n);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 q;
                            q.x = pack16(g.opd, v[8 * j], v[8 * j + 1]);
                            q.y = pack16(g.opd, v[8 * j + 2], v[8 * j + 3]);
                            q.z = pack16(g.opd, v[8 * j + 4], v[8 * j + 5]);
                            q.w = pack16(g.opd, v[8 * j + 6], v[8 * j + 7]);
                            op[j] = q;
                        }
                    }
                    fence_proxy_async();                               // smem writes -> visible to the TMA store
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
                    if (storer) {
                        tma_store_2d(&tmO, ebuf + (half * NBUF + b) * CHUNK_BYTES, n, m0);
                        bulk_commit();
                        bulk_wait_read<1>();                           // the previous chunk's store has drained its smem
                        if (i >= 2) mbar_arrive(&rfree_bar[half][(i - 2) % NBUF]);
                    }
                }
            }
            // release this accumulator stage to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
        }
        if (STAGED && storer) bulk_wait<0>();   // all output tiles are in global memory before the CTA exits
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
    gemm2_kernel<EPI><<<2 * clusters, THREADS, Cfg2<EPI>::SMEM_BYTES, s>>>(maps.a0, maps.a1, maps.b, maps.r32,
                                                                          maps.o32, a);
    return cudaGetLastError();
}

template <int EPI>
cudaError_t configure2() {
    return cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                Cfg2<EPI>::SMEM_BYTES);
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

// fp32-output epilogues need the fp32 tensor maps (GemmMaps::has_f32) and write fp32 through them
bool gemm2_supported(int epi, const GemmMaps& maps, const GemmArgs& a) {
    if (a.N % BN != 0 || a.M <= BM) return false;
    if (epi == EPI_BIAS_RESID || epi == EPI_BIAS_F32) return maps.has_f32 && a.out32 != nullptr;
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
