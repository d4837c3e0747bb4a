// Internal (C++) launcher interface between api.cu and the kernel translation units.
// Nothing here is part of the public ABI (see include/uspace_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace usp {

// tensor-core operand type (both run at the same tcgen05 kind::f16 rate)
enum OperandDtype : int { OPD_BF16 = 0, OPD_FP16 = 1 };

// ---- GEMM: C[M,N] = A[M,K] * W[N,K]^T with fused epilogues ------------------------------------
enum GemmEpilogue : int {
    EPI_QKV = 0,        // 16-bit head-major scatter: out16[which][b*H+h][l][d]      (libs/uvit.py:89-94)
    EPI_BIAS_GELU = 1,  // out16[m,n] = gelu_erf(acc + bias[n])                       (libs/timm.py:107-108)
    EPI_BIAS_RESID = 2, // out32[m,n] = resid[m,n] + acc + bias[n]; optional 16-bit copy (libs/uvit.py:160-161)
    EPI_BIAS_F32 = 3,   // out32[m,n] = acc + bias[n]; optional 16-bit copy            (libs/uvit.py:158-159)
};

struct GemmArgs {
    int M, N, K;         // K = total reduction length
    int K0;              // columns [0,K0) come from A0, [K0,K) from A1 (K0 == K: single source)
    int opd;             // OperandDtype
    const float* bias;   // [N] or nullptr
    const float* resid;  // [M,N] fp32 or nullptr
    float* out32;        // [M,N] fp32 or nullptr
    void* out16;         // [M,N] 16-bit (or QKV base) or nullptr
    int L, H;            // EPI_QKV: tokens per sample, heads (head_dim = 64)
    long long qkv_stride;  // EPI_QKV: elements between the q, k and v planes
    int diag;              // diagnostics (env USP_GEMM_DIAG, results invalid): 1 = no operand TMA loads after the first
                           // ring fill, 2 = no residual loads, 4 / 16 = no fp32 / 16-bit result stores, 8 = no GELU
    // ---- LayerNorm folded into the GEMMs (usp_config.fuse_layernorm) ----
    // consumer side (qkv / fc1): A is the UN-normalised 16-bit stream x, W is pre-scaled by gamma, and
    //   out[m,n] = rstd_m * (acc[m,n] - mean_m * ln_c[n]) + ln_d[n]
    // with ln_c[n] = sum_k W'[n,k], ln_d[n] = sum_k beta_k W[n,k] + bias[n]   (libs/uvit.py:160-161 algebraically)
    const float* ln_stats; // [M, ln_np, 2] per-row partial (sum, sum of squares) written by the producer, or nullptr
    const float* ln_c;     // [N]
    const float* ln_d;     // [N]
    int ln_np;             // partials per row
    float ln_inv_d;        // 1 / embed_dim
    int* ln_flag;          // status word (usp_nonfinite): bit 2 is raised when a row's |mean| exceeds 4 standard deviations,
                           // i.e. the 16-bit rounding of the un-normalised operand starts to cost precision
    // ---- implicit GEMM for 3x3 / stride 1 / pad 1 convolutions (csrc/vae.cu) ----
    // conv_C > 0: A is the fp16 NHWC activation [*, conv_H, conv_W, conv_C] behind an im2col tensor map (GemmMaps::a0);
    // row m = output pixel, K block kb = channels [(kb % (C/64)) * 64, +64) of filter tap kb / (C/64)
    int conv_C, conv_H, conv_W;   // conv_H / conv_W: OUTPUT height / width
    int conv_stride, conv_pad;    // 1 / 1 (same-size convolution) or 2 / 0 (Downsample: zero pad only at the far edges)
    // producer side (skip / proj / fc2): partial row statistics of the fp32 values it writes, one slot per
    // (row, 128-column group): [M, N/128, 2]
    float* stats_out;
};

struct GemmMaps {
    CUtensorMap a0, a1, b;
    CUtensorMap r32, o32;   // fp32 [M,N] residual-in / output maps (box 32 cols x 128 rows, 128B swizzle)
    bool has_f32 = false;
};

// Box sizes the GEMM expects in its tensor maps.
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
int gemm_block_n(int N);  // 256 or 128
int gemm_weight_box_rows();  // rows of the TMA box the W[N,K] tensor map must use (128)
cudaError_t launch_gemm(int epi, const GemmMaps& maps, const GemmArgs& args, int num_sms, cudaStream_t s);

// ---- attention: softmax(Q K^T / sqrt(64)) V per (sample, head) ---------------------------------
struct AttnArgs {
    int B, H, L;       // head_dim fixed at 64
    int D;             // H*64 (row pitch of out16)
    int opd;
    int num_sms;       // persistent grid size
    void* out16;       // [B*L, D] 16-bit, heads merged "(H hd)"
    const float* vscale;  // optional post-softmax column re-weighting [B, L] (p2p_rescale, tools/utils_t2i.py:196-224):
                          // O = sum_j (p_j * m_j) v_j / sum_j p_j  (no re-normalisation), or nullptr
    const struct StepState* st;  // sampling: apply vscale only while st->attn_on != 0; nullptr: always apply
    int diag;             // diagnostics (env USP_ATTN_DIAG, results invalid): 1 = no MUFU, 2 = no softmax arithmetic,
                          // 4 / 8 = no PV / S MMAs (non-pipelined variants), 16 = no TMEM reads, 32 = no tail-row arithmetic
    const void* q16;      // raw pointer to Q [B*H, L, 64] (the SIMT tail-row path reads its query rows directly)
    unsigned long long* trace;   // debugging (env USP_ATTN_TRACE=<file>): per-role event timeline of CTA 0, else nullptr
};
constexpr int ATTN_MAX_L = 384;          // keys per pass of attention_kernel (whole rows up to here)
constexpr int ATTN_LONG_MAX_L = 16384;   // sequence length limit (key passes of ATTN_LONG_PASS beyond ATTN_MAX_L)
constexpr int ATTN_LONG_PASS = 160;      // keys per pass of the long-sequence kernel (two CTAs per SM)
// tensor maps: q = [planes, L, 64] with 128-row boxes; k, v = same tensors with attn_kv_box_rows(L)-row boxes
__host__ __device__ inline int attn_kv_box_rows(int L) { return L > ATTN_MAX_L ? ATTN_LONG_PASS : ((L + 15) & ~15) / 2; }
cudaError_t launch_attention(const CUtensorMap& q, const CUtensorMap& k, const CUtensorMap& v,
                             const AttnArgs& args, cudaStream_t s);

// ---- small fused kernels ---------------------------------------------------------------------
struct StepState {       // lives in device memory; advanced by step_kernel inside the captured graph
    int next;            // index of the next grid interval
    int cur;             // interval being integrated
    float t;             // time fed to the velocity field for the current stage
    float dt;            // signed step (t_{i+1} - t_i)
    float edit;          // write_scale if the edit is active at `t`, else 0
    float write_scale;
    int didx;            // row of the edit table that belongs to `t`
    int attn_on;         // attention re-weighting active at `t` (float(f"{t:.2f}") <= t_edit)
    float attn_t_edit;   // t_edit of the attention edit, for stages between grid points (negative: no attention edit)
};

struct EmbedArgs {
    const float* x;        // [B,C,S,S] latent
    const float* tvec;     // [B] per-sample time (forward) ...
    const StepState* st;   // ... or the shared ODE time (sampling); exactly one is non-null
    const long long* y;    // [B] labels or nullptr
    const float* ctxemb;   // [B*n_ctx, D] fp32 context tokens or nullptr
    const float* w;        // patch_embed.proj.weight [D, C*p*p]
    const float* bias;     // [D]
    const float* pos;      // [L, D]
    const float* label;    // label_emb.weight [num_classes, D] or nullptr
    const float* freqs;    // [D/2]
    const float* ttok;     // mlp_time_embed: the time token(s) [1 or B, D] computed by launch_time_mlp, else nullptr
    const float* delta;    // head edit table [nsteps+1, C*S*S] or nullptr
    const float* sscale;   // optional per-sample write_scale [B] (scale sweep); nullptr: st->edit for every sample
    float* trace;          // optional "read" dump at edit_loc head: trace[st->didx][B,C,S,S] = the latent as the net sees it
    float hook_scale;      // plain forward (st == nullptr): write_scale applied with row 0 of `delta` (usp_forward_hook)
    float* out32;          // [B*L, D]
    void* out16;           // optional un-normalised 16-bit copy [B*L, D] (folded-LayerNorm path)
    float* stats;          // optional [B*L, 8, 2] partial row statistics (one slot per warp of the block)
    int opd;
    int B, C, S, p, D, L, n_ctx, has_label;
};
cudaError_t launch_embed(const EmbedArgs& a, cudaStream_t s);

// mlp_time_embed (libs/uvit.py:215-223, :320): ttok[n] = W2 silu(W1 temb(t_n) + b1) + b2 for n < rows, where
// temb is the sinusoidal timestep embedding and t_n = st->t (rows == 1, sampling) or tvec[n] (forward).
// `sincos` [rows, D] and `hidden` [rows, 4D] are scratch.  fp32 throughout (B x 8 D^2 MACs: not a hot kernel).
struct TimeMlpArgs {
    const float* tvec;
    const StepState* st;
    const float* freqs;    // [D/2]
    const float* w1;       // [4D, D]
    const float* b1;       // [4D]
    const float* w2;       // [D, 4D]
    const float* b2;       // [D]
    float* sincos;
    float* hidden;
    float* ttok;           // [rows, D]
    int rows, D;
};
cudaError_t launch_time_mlp(const TimeMlpArgs& a, cudaStream_t s);

cudaError_t launch_layernorm(const float* x, const float* g, const float* b, void* out16, int M, int D, int opd,
                             cudaStream_t s);

struct HeadArgs {
    const float* x32;     // [B*L, D]
    const float* ng;      // norm.weight
    const float* nb;      // norm.bias
    const float* w;       // decoder_pred.weight [P, D]
    const float* bias;    // [P]
    float* pf;            // [B, n_patches, P]
    int B, L, D, extras, P;
};
cudaError_t launch_head(const HeadArgs& a, cudaStream_t s);

struct FinalArgs {
    const float* pf;       // [B, n_patches, P] with P ordered (p1, p2, C)
    const float* cw;       // final_layer.weight [C,C,3,3] or nullptr (conv=False)
    const float* cb;       // [C]
    const float* delta;    // tail edit table [nsteps+1, C*S*S] or nullptr
    const float* sscale;   // optional per-sample write_scale [B] (scale sweep)
    float* trace;          // optional "read" dump at edit_loc tail: trace[st->didx][B,C,S,S] = the velocity
    const StepState* st;   // nullptr for a plain forward
    float hook_scale;      // plain forward (st == nullptr): write_scale applied with row 0 of `delta` (usp_forward_hook)
    const float* base;     // ODE: state the update starts from
    const float* aux;      // a stored derivative combination (Heun stage 2: k1), or nullptr
    float* vstore;         // running combination 1: vstore = vs_a * v + vs_b * vstore (Heun stage 1: k1), or nullptr
    float* acc2;           // running combination 2: acc2 = a2_a * v + a2_b * acc2 (rk4: k1 + 3 k2 + 3 k3), or nullptr
    int* nonfinite;        // sticky flag: set to 1 when a velocity is inf / NaN (fp16 operand overflow upstream)
    float* out;            // forward: v;  ODE: base + dt*(m1*v + m2*aux), aux read BEFORE the combinations are updated
    float m1, m2;
    float vs_a, vs_b, a2_a, a2_b;
    int B, C, S, p;
};
cudaError_t launch_final(const FinalArgs& a, cudaStream_t s);

cudaError_t launch_convert16(const float* in, void* out16, long long n, int opd, cudaStream_t s);
// Fold a LayerNorm into the Linear that consumes it: w16[n,k] = round16(W[n,k] * gamma[k]),
// c[n] = sum_k w16[n,k], d[n] = sum_k beta[k] * W[n,k] (+ bias[n])
cudaError_t launch_fold_ln(const float* W, const float* gamma, const float* beta, const float* bias, void* w16,
                           float* c, float* d, int N, int K, int opd, cudaStream_t s, int centre = 0);
// stage 0: start interval `next` (t = grid[next]) and advance; stage 1: a stage at the interval's end (t = grid[cur+1]);
// stage 2: a stage inside the interval, t = grid[cur] + frac * dt - no grid row, so no write edit; the attention edit
// follows its "%.2f" rule against st->attn_t_edit
cudaError_t launch_step(StepState* st, const float* grid, const unsigned char* mask, const unsigned char* amask,
                        int stage, cudaStream_t s, float frac = 0.f);

// GEMM on raw device pointers (api.cu): nullptr on success, else a message
// conv_C > 0: a16 is an fp16 NHWC activation [M / (conv_H * conv_W), conv_H, conv_W, conv_C] and the product is its
// 3x3 / stride 1 / pad 1 convolution with W[N, 9 * conv_C] (K ordered (ky, kx, c)), loaded through a TMA im2col map
// conv_stride == 2: the stride-2 Downsample convolution (input [.., 2 conv_H, 2 conv_W, conv_C], zero padding (0, 1, 0, 1))
const char* gemm_raw(int epilogue, const void* a16, const void* w16, const float* bias, const float* resid,
                     float* out32, void* out16, int M, int N, int K, int operand_dtype, int num_sms, cudaStream_t s,
                     int conv_C = 0, int conv_H = 0, int conv_W = 0, int conv_stride = 1);

// ---- adaptive Dormand-Prince 5(4) (csrc/ode.cu) -------------------------------------------------
// torchdiffeq's RKAdaptiveStepsizeODESolver restated with the controller on the device: time-like quantities are
// fp64, the state and the stage times handed to the velocity field fp32 (its mixed-precision convention).
constexpr int RK_STAGES = 7;          // k[0..6]; the last k of an accepted step is k[0] of the next
constexpr int RK_SOLUTION = 100;      // pseudo-stage of launch_rk_stage: ytmp = y0 + dt * sum c_sol[j] k[j] (non-FSAL methods)
enum RkMethod : int { RK_DOPRI5 = 0, RK_BOSH3 = 1, RK_ADAPTIVE_HEUN = 2 };
int rk_stages(int method);            // velocity evaluations per attempted step: 6 / 3 / 1
bool rk_fsal(int method);             // the last stage's state is the step's solution
constexpr int RK_MAX_PARTIALS = 4096;
constexpr int RK_DIGITS = 128;        // rows of the "%.2f"-indexed edit masks (0.00 .. 1.27)
struct RkState {
    double s0;          // solver time at the start of the current step (ascending; model time = sign * s)
    double dt;          // step the next attempt will try
    double s_end;
    double s0_prev;     // the step the last attempt covered (read by rk_commit)
    double dt_prev;
    double rtol, atol;
    float sign;
    float write_scale;
    float h0;           // starting-step candidate (between the two initial evaluations)
    float d0, d1;
    float ratio;        // error ratio of the last attempt
    int accept;         // decision of the last attempt
    int done;           // the last accepted step reached s_end: the result has been written
    int n_accept, n_reject, nfe;
    int n_rows;         // rows of the edit table
};
struct RkArgs {
    RkState* rs;
    StepState* st;
    float* y0;                 // [n] state at the start of the step (advanced by rk_commit)
    float* k;                  // [RK_STAGES][n]
    float* ytmp;               // [n] stage state (y1 after the last stage)
    float* out;                // [n] result
    double* partials;          // [2][RK_MAX_PARTIALS]
    const unsigned char* emask;   // [RK_DIGITS] write-edit active for digit i
    const unsigned char* amask;   // [RK_DIGITS] attention edit active for digit i
    float* eval_times;            // "read" mode: model time of the i-th velocity evaluation (row i of the trace), or nullptr
    int* eval_count;              // ... and the number of evaluations so far
    int eval_cap;                 // rows of the trace
    long long n;
    int method;                // RkMethod
};
// stage 1..6: ytmp = y0 + dt * sum_j beta[stage][j] k_j and the stage time; stage 0: time of the very first
// evaluation; stage -1: the probe point of the starting-step search (ytmp = y0 + h0 * k0)
cudaError_t launch_rk_stage(const RkArgs& a, int stage, cudaStream_t s);
// what: 0 = norms of y0/scale and k0/scale -> h0;  1 = norm of (k1 - k0)/scale -> first dt;
//       2 = error ratio of the attempted step -> accept / reject, next dt, counters
cudaError_t launch_rk_control(const RkArgs& a, int what, cudaStream_t s);
// accepted: (y0, k0) <- (y1, k6), or the dense-output polynomial at s_end into `out` when the step reached it
cudaError_t launch_rk_commit(const RkArgs& a, cudaStream_t s);

}  // namespace usp
