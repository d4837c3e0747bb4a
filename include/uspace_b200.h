/* uspace_b200 — C ABI of the Blackwell (sm_100a) U-ViT flow-matching sampler.
 *
 * This is the drop-in boundary for the one hot path of dongzhuoyao/uspace: the velocity-field call
 *     self.net(x, t, y, **kwargs)                 flow_matching.py:34
 *     self.net(x, t, context=context, **kwargs)   flow_matching_t2i.py:31
 * and the fixed-grid ODE loops around it
 *     CNF.decode / CNF.encode                     flow_matching.py:102-151, flow_matching_t2i.py:105-146
 * The reference has no FFI of its own (it is pure Python); a maintainer binds these symbols with
 * ctypes from libs/uvit.py / libs/uvit_t2i.py / flow_matching*.py — see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross the boundary;
 *   - every call returns 0 (USP_OK) or a negative usp_status; usp_last_error() gives the message;
 *   - all tensors are fp32, contiguous, NCHW for latents; pointers are DEVICE pointers unless the
 *     function name ends in _host;
 *   - work is enqueued on the caller's stream (void* == cudaStream_t, may be NULL) and the call
 *     returns without synchronising, except the *_host entry points which synchronise before returning;
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef USPACE_B200_H
#define USPACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct usp_handle usp_handle;

typedef enum usp_status {
    USP_OK = 0,
    USP_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    USP_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed       */
    USP_ERR_STATE = -3,        /* weights missing or not finalised          */
    USP_ERR_UNSUPPORTED = -4,  /* configuration outside the built path      */
    USP_ERR_NONFINITE = -5     /* a velocity evaluation produced inf / NaN  */
} usp_status;

/* Mirrors the constructor keywords of libs/uvit.py:183-202 and libs/uvit_t2i.py:193-211. */
typedef struct usp_config {
    int32_t img_size;        /* latent side, 32 for the 256^2 models      */
    int32_t patch_size;      /* 2                                         */
    int32_t in_chans;        /* 4                                         */
    int32_t embed_dim;       /* D: multiple of 128 in {256..1536}         */
    int32_t depth;           /* depth//2 in-blocks + mid + depth//2 out   */
    int32_t num_heads;       /* D / 64 (head_dim is fixed at 64)          */
    int32_t mlp_hidden;      /* int(D * mlp_ratio)                        */
    int32_t num_classes;     /* <= 0: no label token (libs/uvit.py:225)   */
    int32_t clip_dim;        /* t2i: 768; 0 for the uncond / class model  */
    int32_t num_clip_token;  /* t2i: 77;  0 for the uncond / class model  */
    int32_t qkv_bias;        /* 0/1                                       */
    int32_t conv;            /* final 3x3 conv present                    */
    int32_t skip;            /* out-blocks carry skip_linear              */
    int32_t operand_dtype;   /* tensor-core operand type: 0 bf16, 1 fp16  */
    int32_t fuse_layernorm;  /* 1: fold norm1/norm2 into the qkv / fc1 GEMMs (no LayerNorm kernels) */
    int32_t mlp_time_embed;  /* 1: time token = Linear(4D->D)(SiLU(Linear(D->4D)(sinusoid))) (libs/uvit.py:215-223);
                                weights "time_embed.0.{weight,bias}", "time_embed.2.{weight,bias}"                  */
} usp_config;

/* Post-softmax attention column re-weighting ("p2p_rescale": tools/utils_t2i.py:196-224,265-296, called from
 * libs/uvit_t2i.py:104): attn[:, :, :, j] *= colscale[b, j] after the softmax, without re-normalisation. */
typedef struct usp_attn_edit {
    const float* colscale;   /* [B, L] fp32, device or host; 1.0 leaves a key column untouched              */
    uint64_t block_mask;     /* bit i set: apply in the i-th executed block (in, mid, out order); ~0 = "all" */
    float t_edit;            /* sampling only: active while float(f"{t:.2f}") <= t_edit                      */
} usp_attn_edit;

/* torchdiffeq fixed-grid methods: "euler", "heun2"-style Heun (2 NFE), "midpoint" (2 NFE), "rk4" = the 3/8 rule of
 * rk4_alt_step_func (4 NFE). The write edit is keyed by grid point, so under midpoint / rk4 it is applied only at
 * stages that sit on a grid point (the reference would look for a delta file of the in-between time). */
enum { USP_METHOD_EULER = 0, USP_METHOD_HEUN = 1, USP_METHOD_MIDPOINT = 2, USP_METHOD_RK4 = 3 };
enum { USP_EDIT_NONE = 0, USP_EDIT_HEAD = 1, USP_EDIT_TAIL = 2 };

/* Replaces UViT.__init__ (libs/uvit.py:182-291): allocates parameter storage on `device`. */
int usp_create(const usp_config* cfg, int device, usp_handle** out);
void usp_destroy(usp_handle* h);
const char* usp_last_error(const usp_handle* h);  /* h may be NULL: message of the last failed usp_create */

/* Replaces load_state_dict (dissect_lfm.py:70-72): `name` is the reference state_dict key, `data` fp32,
 * host or device memory. Unknown names or wrong shapes return USP_ERR_INVALID. */
int usp_set_weight(usp_handle* h, const char* name, const void* data, const int64_t* shape, int ndim);
/* Packs the GEMM weights to the 16-bit operand type. Must follow the last usp_set_weight. */
int usp_finalize_weights(usp_handle* h, void* stream);
/* Number of state_dict entries the configuration expects, and the i-th name (for loaders / tests). */
int usp_num_weights(const usp_handle* h);
const char* usp_weight_name(const usp_handle* h, int i);

/* Replaces UViT.forward (libs/uvit.py:306-351, libs/uvit_t2i.py:308-342), inference only.
 *   x [B,C,S,S], t [B], context [B,n_ctx,clip_dim] or NULL, y int64 [B] or NULL  ->  out [B,C,S,S] */
int usp_forward(usp_handle* h, const float* x, const float* t, const float* context, const int64_t* y,
                float* out, int B, void* stream);

/* usp_forward with the attention edit applied unconditionally in the masked blocks (edit may be NULL). */
int usp_forward_edit(usp_handle* h, const float* x, const float* t, const float* context, const int64_t* y,
                     float* out, int B, const usp_attn_edit* edit, void* stream);

/* usp_forward_edit plus the semantic-direction hook of ONE velocity evaluation, as the reference applies it inside
 * UViT.forward (libs/uvit.py:313-314 "head": on the latent before patch_embed; :349-350 "tail": on the velocity),
 * dissect_helper_uvit (libs/dissection.py:115-186):
 *   delta != NULL      "write_attr" / "write_pca":  x + delta * write_scale, delta [C,S,S] = the row the hook loads from
 *                      delta_{t:.2f}.npy (the CALLER applies should_edit() and passes NULL when it is false);
 *   read_out != NULL   "read": the activation at edit_loc [B,C,S,S] (what the hook np.save()s).
 * delta / read_out are device pointers; edit_loc USP_EDIT_NONE with both NULL is usp_forward_edit. */
int usp_forward_hook(usp_handle* h, const float* x, const float* t, const float* context, const int64_t* y,
                     float* out, int B, int edit_loc, const float* delta, float write_scale, float* read_out,
                     const usp_attn_edit* edit, void* stream);

/* Replaces odeint(func, z, [t0,t1], method, options=dict(step_size)) as called by CNF.decode (t0=0,t1=1) and
 * CNF.encode (t0=1,t1=0) (flow_matching.py:118-125,140-147): torchdiffeq fixed-grid semantics, one CUDA graph
 * per step replayed on `stream`.  z is updated in place.
 *   delta_table: optional [n_grid, C,S,S] rows indexed by grid point (the delta_{t:.2f}.npy files of
 *   libs/dissection.py:141-157 pre-gathered by the caller), applied as x + delta*write_scale at edit_loc when
 *   "0.00" != f"{t:.2f}" <= t_edit (libs/dissection.py:21-26). */
int usp_sample(usp_handle* h, float* z, const float* context, const int64_t* y, int B, float t0, float t1,
               float step_size, int method, const float* delta_table, float write_scale, float t_edit,
               int edit_loc, void* stream);
/* usp_sample plus the attention edit of dissect_lfm_t2i.py's "p2p" mode (attn may be NULL). */
int usp_sample_edit(usp_handle* h, float* z, const float* context, const int64_t* y, int B, float t0, float t1,
                    float step_size, int method, const float* delta_table, float write_scale, float t_edit,
                    int edit_loc, const usp_attn_edit* attn, void* stream);
/* Semantic-direction sweep: the loop `for write_scale in write_scales: sample_fn(input_z, write_scale=...)` of
 * sample_for_hspace_vis (tools/utils_vis.py:189-201) as ONE batch of B * n_scales samples - copy s of latent b is
 * integrated with write_scale = write_scales[s] (host array) and lands in out[b][s] ("(b s)" order of the
 * reference's image grid). Samples are independent, so every row is bit-identical to a separate usp_sample call.
 * z: [B, C, S, S] (not modified), out: [B, n_scales, C, S, S]; context / y are per input latent ([B, ...]). */
int usp_sample_sweep(usp_handle* h, const float* z, float* out, const float* context, const int64_t* y, int B,
                     const float* write_scales, int n_scales, float t0, float t1, float step_size, int method,
                     const float* delta_table, float t_edit, int edit_loc, void* stream);
/* The "read" mode of the hook (libs/dissection.py:126-136): integrate like usp_sample and keep, for every velocity
 * evaluation, the activation at edit_loc - USP_EDIT_HEAD: the latent handed to the network; USP_EDIT_TAIL: the
 * predicted velocity. trace: [grid points, B, C, S, S] (device or host), row i <-> the evaluation at grid[i], i.e.
 * the array the reference saves as f"{batch_id}_{grid[i]:.2f}.npy"; rows never evaluated (the last grid point under
 * Euler) are zero. A later evaluation at the same grid point overwrites an earlier one, like the reference's files. */
int usp_sample_read(usp_handle* h, float* z, const float* context, const int64_t* y, int B, float t0, float t1,
                    float step_size, int method, int edit_loc, float* trace, void* stream);
/* Adaptive Dormand-Prince 5(4): replaces odeint(func, z, [t0, t1], method="dopri5", rtol=, atol=)[-1] at
 * flow_matching.py:79-84 (default sampling), :50-57 (solver="adaptive") and :172-179 (adaptive tail of "fixadp");
 * flow_matching_t2i.py likewise. Step acceptance, the next step size and the dense-output evaluation at t1 run on
 * the device; the host reads one small state record per attempted step, so this call SYNCHRONISES `stream`.
 * The rms error norm spans the whole batch (torchdiffeq's default), i.e. results depend on how a batch is split.
 * delta_digits: [n_rows, C, S, S] rows keyed by the "%.2f" digit of the evaluation time, row i <-> delta_{i/100:.2f}.npy
 * (n_rows <= 128; NULL iff edit_loc == USP_EDIT_NONE). */
typedef struct usp_adaptive_stats {
    int32_t n_accept, n_reject, nfe;
    float last_ratio;   /* error ratio of the last attempted step */
    double last_dt;     /* size of the last accepted step */
} usp_adaptive_stats;
/* method: torchdiffeq's "dopri5" (5(4), 6 evaluations per step), "bosh3" (3(2), 3) or "adaptive_heun" (2(1), 1). */
enum { USP_METHOD_DOPRI5 = 4, USP_METHOD_BOSH3 = 5, USP_METHOD_ADAPTIVE_HEUN = 6 };
int usp_sample_adaptive(usp_handle* h, float* z, const float* context, const int64_t* y, int B, float t0, float t1,
                        int method, double rtol, double atol, const float* delta_digits, int n_rows, float write_scale,
                        float t_edit, int edit_loc, const usp_attn_edit* attn, int max_steps,
                        usp_adaptive_stats* stats, void* stream);
/* dissect_name="read" (libs/dissection.py:126-136) under an adaptive solver: the same integration, and every velocity
 * evaluation - the two of the starting-step search and rejected attempts included, like the reference's hook inside
 * the net - dumps the activation at edit_loc (head: the latent the net sees, tail: the velocity) into its own row
 * trace[i] ([trace_cap][B,C,S,S] fp32, host or device) and its model time into times[i]; *n_evals receives the number
 * of evaluations (USP_ERR_STATE if it exceeds trace_cap; the first trace_cap rows are still returned). The caller
 * writes {batch_id}_{times[i]:.2f}.npy in evaluation order, so that a later evaluation at the same digit overwrites the
 * file as in the reference. */
int usp_sample_adaptive_read(usp_handle* h, float* z, const float* context, const int64_t* y, int B, float t0, float t1,
                             int method, double rtol, double atol, int edit_loc, float* trace, float* times, int trace_cap,
                             int* n_evals, int max_steps, usp_adaptive_stats* stats, void* stream);
/* ---- latent -> image decoder -------------------------------------------------------------------------------------
 * Replaces FrozenAutoencoderKL.decode (libs/autoencoder.py:446-450: z / scale_factor -> post_quant_conv -> Decoder),
 * the call dissect_lfm.py:86-98 / train_lfm.py make on every batch of sampled latents. Only the configuration of
 * libs/autoencoder.py::get_model is built (ch 128, ch_mult 1-2-4-4, 2 res blocks, z_channels 4). Weight names are the
 * reference state_dict keys "decoder.*" and "post_quant_conv.*" (fp32, host or device).
 *   z [B, 4, S, S] (S in {16, 32, 48, 64}; 32 for the 256^2 models)  ->  out [B, 3, 8S, 8S], device pointers. */
typedef struct usp_vae usp_vae;
int usp_vae_create(int device, float scale_factor, usp_vae** out);
void usp_vae_destroy(usp_vae* h);
const char* usp_vae_last_error(const usp_vae* h);   /* h may be NULL: message of the last failed usp_vae_create */
int usp_vae_num_weights(const usp_vae* h);
const char* usp_vae_weight_name(const usp_vae* h, int i);
int usp_vae_set_weight(usp_vae* h, const char* name, const void* data, const int64_t* shape, int ndim);
/* Operand precision of every GEMM in the autoencoder; call before usp_vae_finalize (changing it un-finalises).
 * FP16X3 (default): operands split into hi + lo fp16 parts, three products folded into one GEMM over 3x the K
 * (~1e-5 against the fp64 oracle); FP16: one product (3x less tensor work, ~2e-3: the precision of the TF32
 * convolutions the reference runs by default on a GPU). */
#define USP_VAE_PRECISION_FP16 0
#define USP_VAE_PRECISION_FP16X3 1
int usp_vae_set_precision(usp_vae* h, int mode);
int usp_vae_finalize(usp_vae* h, void* stream);      /* packs the convolution weights; synchronises */
/* The activation workspace is sized by the first decode / encode (chunks of <= 16 images) and kept between calls:
 * its current size, and a way to hand it back (synchronises the device; the next call allocates again). */
size_t usp_vae_workspace_bytes(const usp_vae* h);
int usp_vae_release_workspace(usp_vae* h);
int usp_vae_decode(usp_vae* h, const float* z, float* out, int B, int S, void* stream);
/* Replaces FrozenAutoencoderKL.encode_moments (libs/autoencoder.py:426-429: Encoder.forward :275-300 + quant_conv):
 *   x [B, 3, R, R] images in [-1, 1] (R in {128, 256, 384, 512})  ->  moments [B, 8, R/8, R/8] = (mean, logvar);
 * the caller draws the latent sample (libs/autoencoder.py:431-437). Weight names "encoder.*", "quant_conv.*". */
int usp_vae_encode_moments(usp_vae* h, const float* x, float* moments, int B, int R, void* stream);

/* Same with HOST buffers: copies z (and context / y / delta_table) host->device, samples, copies z back,
 * and synchronises. This is the end-to-end call bench.py times. */
int usp_sample_host(usp_handle* h, float* z_host, const float* context_host, const int64_t* y_host, int B, float t0,
                    float t1, float step_size, int method, const float* delta_table_host, float write_scale,
                    float t_edit, int edit_loc);
/* fp16 tensor-core operands saturate at 65504: a checkpoint whose activations exceed that yields inf / NaN velocities.
 * Every velocity evaluation checks its output on the device; this reads and clears the sticky flag (*flag = 1 if any
 * evaluation since the last read was non-finite). SYNCHRONISES `stream`. usp_sample_host checks it itself and returns
 * USP_ERR_NONFINITE; bf16 operands (usp_config.operand_dtype = 0) have the fp32 exponent range.
 * *flag is a bit set: 1 = a non-finite velocity; 2 = (fuse_layernorm) a token whose mean exceeded 4 standard deviations
 * reached a folded LayerNorm - its 16-bit un-normalised operand loses log2(|mean| / std) bits there, use
 * fuse_layernorm = 0 for such a checkpoint. */
int usp_nonfinite(usp_handle* h, int* flag, void* stream);
/* Number of grid points torchdiffeq builds for (t0, t1, step_size): ceil(|t1-t0|/step + 1). */
int usp_grid_size(float t0, float t1, float step_size);
/* Writes the grid itself (host memory, fp32) into out[0..cap); returns the number of points or 0 on error. */
int usp_time_grid(float t0, float t1, float step_size, float* out, int cap);

/* Introspection used by bench.py / tests. */
size_t usp_workspace_bytes(const usp_handle* h, int B);
int usp_kernels_per_forward(const usp_handle* h);   /* kernels launched by one velocity evaluation */
double usp_flops_per_forward(const usp_handle* h);  /* algorithmic FLOPs per image per forward */
int usp_last_forward_ms(usp_handle* h, float* ms);  /* device time of the most recent usp_forward/usp_sample */

/* One eager velocity evaluation with a CUDA event between consecutive launches; returns per-class device time.
 * Classes: 0 embed, 1 layernorm, 2 gemm_qkv, 3 attention, 4 gemm_proj, 5 gemm_fc1, 6 gemm_fc2, 7 gemm_skip,
 * 8 head+final, 9 context_embed.  class_ms / class_launches have USP_NUM_KERNEL_CLASSES entries. Synchronises. */
#define USP_NUM_KERNEL_CLASSES 10
int usp_profile_forward(usp_handle* h, const float* x, const float* t, const float* context, const int64_t* y,
                        float* out, int B, float* class_ms, int* class_launches, void* stream);
/* The same over `warmup` untimed + `reps` timed evaluations enqueued back to back (no host synchronisation in between,
 * so the GPU holds the clock / power state of a long run); class_ms / class_launches are per evaluation (mean). */
int usp_profile_forward_n(usp_handle* h, const float* x, const float* t, const float* context, const int64_t* y,
                          float* out, int B, int warmup, int reps, float* class_ms, int* class_launches, void* stream);

/* Kernel-level entry points (parity tests and micro-benchmarks call the kernels through the ABI).
 * All pointers are device pointers; a16/w16/q/k/v/out16 hold 16-bit operands of type `operand_dtype`. */
int usp_op_convert16(const float* in, void* out16, int64_t n, int operand_dtype, void* stream);
int usp_op_gemm(int epilogue, const void* a16, const void* a16_second, const void* w16, const float* bias,
                const float* resid, float* out32, void* out16, int M, int N, int K, int K0, int L, int H,
                int operand_dtype, void* stream);
int usp_op_attention(const void* q16, const void* k16, const void* v16, void* out16, int B, int H, int L,
                     int operand_dtype, void* stream);
int usp_op_layernorm(const float* x, const float* gamma, const float* beta, void* out16, int M, int D,
                     int operand_dtype, void* stream);

/* Token embedding alone (libs/uvit.py:175-179,316-327 without label/context tokens): out [B, 1+n_patch, D]. */
int usp_op_patch_embed(const float* x, const float* t, const float* w, const float* bias, const float* pos,
                       float* out32, int B, int C, int S, int p, int D, void* stream);
/* unpatchify (p1,p2,C) + optional 3x3 conv (libs/uvit.py:56-63,346-347): pf [B, n_patch, p*p*C] -> out [B,C,S,S];
 * conv_w / conv_b may be NULL (conv=False). */
int usp_op_unpatchify_conv(const float* pf, const float* conv_w, const float* conv_b, float* out, int B, int C,
                           int S, int p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USPACE_B200_H */
