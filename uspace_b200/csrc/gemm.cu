// tcgen05 GEMM for the U-ViT linears: C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue).
//
// Replaces, on the reference path, every nn.Linear inside Block._forward (libs/uvit.py:157-162):
//   attn.qkv  (libs/uvit.py:81,89)     -> EPI_QKV        (head-major Q/K/V, no rearrange/.float() copies)
//   attn.proj (libs/uvit.py:116,160)   -> EPI_BIAS_RESID (bias + residual add into the fp32 stream)
//   mlp.fc1+GELU (libs/timm.py:107-108)-> EPI_BIAS_GELU
//   mlp.fc2   (libs/timm.py:110)       -> EPI_BIAS_RESID
//   skip_linear(cat[x,skip]) (libs/uvit.py:158-159) -> EPI_BIAS_F32 with a two-source K loop (no concat)
//
// Structure: persistent, warp-specialised CTA per SM.
//   warp 0    : TMA producer (A and W tiles, 128B-swizzled, 64-wide K blocks) through a STAGES-deep ring
//   warp 1    : allocates TMEM, single thread issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warps 2-5 : epilogue; tcgen05.ld the fp32 accumulator (double-buffered in TMEM so the next tile's
//               main loop overlaps this tile's epilogue), fuse bias/GELU/residual, store.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace usp {

namespace {

constexpr int BM = GEMM_BM;
constexpr int BK = GEMM_BK;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB

template <int BN>
struct GemmCfg {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
            const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tiles_n = g.N / BN;
    const int n_tiles_m = (g.M + BM - 1) / BM;
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
            mbar_init(&tmem_empty_bar[i], 4);
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();
    pdl_launch();

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * BM;
                const int n0 = (tile % n_tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + A_TILE_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        if (g.conv_C > 0) {
                            const int cpt = g.conv_C / BK, tap = kb / cpt;
                            tma_load_im2col(&tmA0, &full_bar[stage], sa, (kb - tap * cpt) * BK,
                                            (m0 % g.conv_W) * g.conv_stride - g.conv_pad,
                                            ((m0 / g.conv_W) % g.conv_H) * g.conv_stride - g.conv_pad,
                                            m0 / (g.conv_W * g.conv_H),
                                            static_cast<uint16_t>(tap % 3), static_cast<uint16_t>(tap / 3));
                        } else if (kb < nkb0)
                            tma_load_2d(&tmA0, &full_bar[stage], sa, kb * BK, m0);
                        else
                            tma_load_2d(&tmA1, &full_bar[stage], sa, (kb - nkb0) * BK, m0);
                        tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n0);
                        if (BN == 256) tma_load_2d(&tmB, &full_bar[stage], sb + 128 * BK * 2, kb * BK, n0 + 128);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const uint32_t idesc = umma_idesc(g.opd == OPD_FP16 ? 0 : 1, BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int t = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
                const int as = t & 1;
                const uint32_t aphase = (t >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + A_TILE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 elements (32 B) along K inside the 128B swizzle atom
                            umma_f16(d_tmem, adesc + (k * 2), bdesc + (k * 2), idesc, (kb | k) != 0);
                        }
                        umma_commit(&empty_bar[stage]);  // smem slot is free once these MMAs retire
                        if (kb == nkb - 1) umma_commit(&tmem_full_bar[as]);  // accumulator ready for the epilogue
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int lg = warp & 3;  // TMEM lane group this warp may access
        int t = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
            const int as = t & 1;
            const uint32_t aphase = (t >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * BM;
            const int n0 = (tile % n_tiles_n) * BN;
            const int m = m0 + lg * 32 + lane;

            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();

            EpiRow row = epi_row(g, EPI, m);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                float4 rb[8];
                if (EPI == EPI_BIAS_RESID) epi_load_resid(g, row, n0 + c * 32, rb);
                tmem_ld32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * BN + c * 32, r);
                tmem_ld_wait();
                epi_chunk<EPI>(g, row, n0 + c * 32, r, rb);
            }
            // release this accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN, int EPI>
cudaError_t launch_one(const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    using Cfg = GemmCfg<BN>;
    const int n_tiles = ((a.M + BM - 1) / BM) * (a.N / BN);
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    return launch_pdl(gemm_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, maps.a0, maps.a1, maps.b,
                      a);
}

template <int BN>
cudaError_t launch_bn(int epi, const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    switch (epi) {
        case EPI_QKV: return launch_one<BN, EPI_QKV>(maps, a, num_sms, s);
        case EPI_BIAS_GELU: return launch_one<BN, EPI_BIAS_GELU>(maps, a, num_sms, s);
        case EPI_BIAS_RESID: return launch_one<BN, EPI_BIAS_RESID>(maps, a, num_sms, s);
        case EPI_BIAS_F32: return launch_one<BN, EPI_BIAS_F32>(maps, a, num_sms, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

template <int BN, int EPI>
cudaError_t configure_one() {
    return cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GemmCfg<BN>::SMEM_BYTES);
}

// opt every instantiation into its dynamic shared memory size (done once, outside any graph capture)
cudaError_t gemm2_configure();

cudaError_t gemm_configure() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = configure_one<256, EPI_QKV>()) != cudaSuccess) return e;
    if ((e = configure_one<256, EPI_BIAS_GELU>()) != cudaSuccess) return e;
    if ((e = configure_one<256, EPI_BIAS_RESID>()) != cudaSuccess) return e;
    if ((e = configure_one<256, EPI_BIAS_F32>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_QKV>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_BIAS_GELU>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_BIAS_RESID>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_BIAS_F32>()) != cudaSuccess) return e;
    if ((e = gemm2_configure()) != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}

int gemm_block_n(int N) { return (N % 256 == 0) ? 256 : 128; }
int gemm_weight_box_rows() { return 128; }

cudaError_t gemm2_configure();
bool gemm2_supported(int epi, const GemmMaps& maps, const GemmArgs& a);
cudaError_t launch_gemm2(int epi, const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s);

// USP_GEMM_1CTA=1 forces the single-CTA kernel (A/B comparison, debugging)
static bool force_1cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("USP_GEMM_1CTA");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

cudaError_t launch_gemm(int epi, const GemmMaps& maps, const GemmArgs& a, int num_sms, cudaStream_t s) {
    if (a.M <= 0 || a.N % 128 != 0 || a.K % BK != 0 || a.K0 % BK != 0 || a.K0 > a.K) return cudaErrorInvalidValue;
    // the folded-LayerNorm epilogues are implemented by the pair kernel only
    if ((a.ln_stats != nullptr || a.stats_out != nullptr) && (force_1cta() || !gemm2_supported(epi, maps, a)))
        return cudaErrorNotSupported;
    if (!force_1cta() && gemm2_supported(epi, maps, a)) {
        static int diag = -1;
        if (diag < 0) {
            const char* e = getenv("USP_GEMM_DIAG");
            diag = e ? atoi(e) : 0;
        }
        GemmArgs a2 = a;
        a2.diag = diag;
        return launch_gemm2(epi, maps, a2, num_sms, s);
    }
    if (gemm_block_n(a.N) == 256) return launch_bn<256>(epi, maps, a, num_sms, s);
    return launch_bn<128>(epi, maps, a, num_sms, s);
}

}  // namespace usp
