/*
 * ldm_b200.h -- C ABI of the B200-native latent-DDPM sampling path.
 *
 * The reference (ynyeh0221/Oxford-102-Flower-GAN-VAE-latent-diffusion) has no
 * FFI layer: its boundary for this path is the PyTorch module API of
 * v2/model_train_test.py.  Each entry point below names the reference method it
 * stands behind; the Python mirror of those classes (package directory
 * oxford-102-flower-gan-vae-latent-diffusion_b200/) binds these symbols with
 * ctypes and nothing else.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer named *_dev is a device pointer on the context's device;
 *     pointers named *_host are host pointers (pinned memory recommended);
 *   - all floating point tensors are fp32, row-major, contiguous; all index
 *     tensors are int64 (the reference's dtypes);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     calls are asynchronous on it unless stated otherwise;
 *   - return value: 0 success; < 0 argument / state validation failure;
 *     > 0 a cudaError_t / CUresult.  ldm_last_error() holds the message of the
 *     last failure on the calling thread.  Nothing throws or exits.
 *   - a context is bound to one device and is not thread-safe; distinct
 *     contexts are independent.  Multi-GPU = one process and one context per GPU.
 */
#ifndef LDM_B200_H_
#define LDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDM_ABI_VERSION 1

#if defined(__GNUC__)
#define LDM_API __attribute__((visibility("default")))
#else
#define LDM_API
#endif

#define LDM_MAX_STAGES 8

/* arithmetic of the dense contractions */
#define LDM_PRECISION_FP32 0 /* fp32 CUDA-core FMA, strict mode (eps within 1e-3 of the reference) */
#define LDM_PRECISION_BF16 2 /* bf16 operands, fp32 accumulation on tcgen05 tensor cores (2e-2) */

typedef struct ldm_ctx ldm_ctx;

/* ConditionalUNet parameters (v2:501-533), device pointers to the tensors of the
 * module's state_dict, unchanged layout.  n_stages = len(hidden_dims) - 1. */
typedef struct ldm_unet_weights {
  int32_t latent_dim;                 /* v2:502 latent_dim (256) */
  int32_t time_dim;                   /* v2:503 time_emb_dim (256) */
  int32_t num_classes;                /* v2:503 (102) */
  int32_t n_stages;                   /* 4 */
  int32_t hidden[LDM_MAX_STAGES + 1]; /* v2:502 hidden_dims (256,512,1024,512,256) */
  int32_t n_t;                        /* rows of `sinusoid` (= n_steps) */
  /* sinusoid(t) = [sin(t f) | cos(t f)] for t = 0..n_t-1, (n_t, time_dim), built by the host
   * mirror with the reference's own torch expression (v2:410-414) so it is bit-identical */
  const float* sinusoid;
  const float* residual_weight;       /* () v2:533 */
  const float *time_lin1_w, *time_lin1_b, *time_lin2_w, *time_lin2_b;      /* v2:405-407 */
  const float* class_embedding;       /* (num_classes, time_dim) v2:424 */
  const float *class_lin1_w, *class_lin1_b, *class_lin2_w, *class_lin2_b;  /* v2:425-427 */
  const float *latent_proj_w, *latent_proj_b;                              /* v2:509 */
  const float* time_proj_w[LDM_MAX_STAGES];   /* time_projections[i] v2:510-512 (i < n_stages used) */
  const float* time_proj_b[LDM_MAX_STAGES];
  const float* attn_in_proj_w[LDM_MAX_STAGES]; /* (3d, d) v2:513-516; rows [2d,3d) = V */
  const float* attn_in_proj_b[LDM_MAX_STAGES];
  const float* attn_out_w[LDM_MAX_STAGES];
  const float* attn_out_b[LDM_MAX_STAGES];
  const float* block_lin_w[LDM_MAX_STAGES];    /* layers[i][0][0] v2:519 */
  const float* block_lin_b[LDM_MAX_STAGES];
  const float* block_ln_w[LDM_MAX_STAGES];     /* layers[i][0][1] v2:520 */
  const float* block_ln_b[LDM_MAX_STAGES];
  const float* stage_ln_w[LDM_MAX_STAGES];     /* layers[i][1] v2:524 */
  const float* stage_ln_b[LDM_MAX_STAGES];
  const float* down_w[LDM_MAX_STAGES];         /* layers[i][2] v2:525 */
  const float* down_b[LDM_MAX_STAGES];
  const float *final_time_w, *final_time_b, *final_class_w, *final_class_b; /* v2:528-529 */
  const float *final_norm_w, *final_norm_b;                                /* v2:530 */
  const float *final_w, *final_b;                                          /* v2:531 */
} ldm_unet_weights;

/* ResidualBlock parameters (v2:159-168) */
typedef struct ldm_resblock_weights {
  const float *conv1_w, *conv1_b, *ln1_w, *ln1_b; /* (C,C,3,3),(C),(C),(C) */
  const float *conv2_w, *conv2_b, *ln2_w, *ln2_b;
  const float *ca_w0, *ca_w2;                     /* (C/8,C,1,1), (C,C/8,1,1) v2:58-61 */
  const float* sa_w;                              /* (1,2,7,7) v2:72 */
} ldm_resblock_weights;

/* Decoder parameters (v2:242-278); index 0 -> res3/up3 (512ch), 1 -> res2/up2, 2 -> res1/up1 */
typedef struct ldm_decoder_weights {
  int32_t latent_dim; /* 256 */
  const float *fc0_w, *fc0_b, *fc1_w, *fc1_b; /* Linear(256,512), LayerNorm(512) */
  const float *fc3_w, *fc3_b, *fc4_w, *fc4_b; /* Linear(512,32768), LayerNorm(32768) */
  ldm_resblock_weights res[3];
  const float* up_w[3];  /* ConvTranspose2d weight (Cin, Cin/2, 4, 4) */
  const float* up_b[3];
  const float* up_gn_w[3];
  const float* up_gn_b[3];
  const float *fin0_w, *fin0_b, *fin_gn_w, *fin_gn_b; /* Conv2d(64,32,3), GroupNorm(8,32) */
  const float *fin3_w, *fin3_b;                       /* Conv2d(32,3,3) */
} ldm_decoder_weights;

/* v3 ConditionalUNet parameters (v3/model_train_test.py:769-803): flower + colour condition through
 * MultiConditionEmbedding (v3:739-749) and per-stage cond_projections; nn.MultiheadAttention is used in full (its
 * input (B,1,d) makes the B samples of a call attend to each other, v3:832-835); no final residual (v3:853). */
typedef struct ldm_unet3_weights {
  int32_t latent_dim, time_dim, num_classes, num_colors, n_stages;
  int32_t hidden[LDM_MAX_STAGES + 1];
  int32_t n_t;
  const float* sinusoid;                         /* (n_t, time_dim), as in ldm_unet_weights */
  const float *time_lin1_w, *time_lin1_b, *time_lin2_w, *time_lin2_b;
  const float *flower_emb, *color_emb;           /* (num_classes, time_dim), (num_colors, time_dim) */
  const float *cond_fc_w, *cond_fc_b;            /* Linear(2 time_dim, time_dim) */
  const float *latent_proj_w, *latent_proj_b;
  const float* time_proj_w[LDM_MAX_STAGES];
  const float* time_proj_b[LDM_MAX_STAGES];
  const float* cond_proj_w[LDM_MAX_STAGES];      /* cond_projections[i] v3:781 */
  const float* cond_proj_b[LDM_MAX_STAGES];
  const float* attn_in_proj_w[LDM_MAX_STAGES];   /* (3d, d): Q, K, V all used */
  const float* attn_in_proj_b[LDM_MAX_STAGES];
  const float* attn_out_w[LDM_MAX_STAGES];
  const float* attn_out_b[LDM_MAX_STAGES];
  const float* block_lin_w[LDM_MAX_STAGES];
  const float* block_lin_b[LDM_MAX_STAGES];
  const float* block_ln_w[LDM_MAX_STAGES];
  const float* block_ln_b[LDM_MAX_STAGES];
  const float* stage_ln_w[LDM_MAX_STAGES];
  const float* stage_ln_b[LDM_MAX_STAGES];
  const float* down_w[LDM_MAX_STAGES];
  const float* down_b[LDM_MAX_STAGES];
  const float *final_time_w, *final_time_b, *final_class_w, *final_class_b; /* v3:798-799 */
  const float *final_norm_w, *final_norm_b;
  const float *final_w, *final_b;
} ldm_unet3_weights;

/* v4 / v5 pixel-space SimpleUNet (v4:37-97), raw fp32 device pointers in the reference's state_dict layout
 * (Conv2d weight (Cout, Cin, k, k); ConvTranspose2d weight (Cin, Cout, 4, 4)).  res_ratio: the scalar parameter of v5
 * (v5:54, out + res_ratio * x_input, v5:144) or NULL for v4. */
typedef struct ldm_pix_conv { const float *w, *b; } ldm_pix_conv;
typedef struct ldm_pix_weights {
  int32_t in_channels;      /* 3 */
  int32_t base_channels;    /* 64 (a multiple of 64) */
  int32_t time_emb_dim;     /* 128 */
  int32_t n_t;              /* rows of the per-timestep table built at pack time (= n_steps of the sampler) */
  const float* res_ratio;
  const float *time_embed0_w, *time_embed0_b;   /* Linear(1, temb)     v4:42 */
  const float *time_embed2_w, *time_embed2_b;   /* Linear(temb, temb)  v4:44 */
  const float *time_fc_w[3], *time_fc_b[3];     /* time_fc1..3         v4:47-49 */
  ldm_pix_conv conv1[2], down1, conv2[2], down2, conv3[2], bottleneck[2], up1, conv4[2], up2, conv5[2], out_conv;
} ldm_pix_weights;

/* The conv U-Net blocks defined (and never instantiated) by the v2 script: UNetResidualBlock v2:462-486 and
 * UNetAttentionBlock v2:434-459; raw fp32 device pointers in the modules' state_dict layout. */
typedef struct ldm_ublock_res_weights {
  int32_t in_channels, out_channels, d_time;     /* channels: multiples of 64 */
  const float *norm1_w, *norm1_b;                /* LayerNorm2d(in_channels) */
  const float *conv1_w, *conv1_b;                /* (out, in, 3, 3) */
  const float *time_w, *time_b;                  /* Linear(d_time, out) */
  const float *class_w, *class_b;                /* Linear(d_time, out) */
  const float *norm2_w, *norm2_b;
  const float *conv2_w, *conv2_b;                /* (out, out, 3, 3) */
  const float *res_w, *res_b;                    /* (out, in, 1, 1) when in != out, else NULL (nn.Identity) */
} ldm_ublock_res_weights;
typedef struct ldm_ublock_attn_weights {
  int32_t channels, num_heads;                   /* channels / num_heads in {16, 32, 64, 128} */
  const float *norm_w, *norm_b;                  /* GroupNorm(1, channels) */
  const float *qkv_w, *qkv_b;                    /* (3 channels, channels, 1, 1) */
  const float *proj_w, *proj_b;                  /* (channels, channels, 1, 1) */
} ldm_ublock_attn_weights;

LDM_API int ldm_version(void);
LDM_API const char* ldm_last_error(void);

/* One context per (device, caller). precision: LDM_PRECISION_*. */
LDM_API int ldm_ctx_create(ldm_ctx** out, int device, int precision);
LDM_API int ldm_ctx_destroy(ldm_ctx* ctx);

/* ConditionalDenoiseDiffusion.__init__ (v2:565-572): the three schedule tables, HOST pointers
 * (n_steps each).  The per-step update coefficients are derived from them in fp32 with the
 * reference's own expressions (v2:584-590). Synchronous. */
LDM_API int ldm_set_schedule(ldm_ctx* ctx, const float* beta_host, const float* alpha_host,
                     const float* alpha_bar_host, int n_steps);

/* Repack ConditionalUNet weights into kernel layouts and build the per-timestep and per-class
 * bias tables (hoists v2:537-538,541-545,554-558 out of the loop). Weights are read from the
 * caller's tensors once; the context keeps its own packed copies. Synchronises the stream. */
LDM_API int ldm_unet_pack(ldm_ctx* ctx, const ldm_unet_weights* w, void* stream);

/* Class labels of the current batch (argument c of forward / p_sample / sample); NULL = the
 * c=None branch (v2:538,543,556).  Labels are range-checked on the device; an out-of-range
 * label makes the next synchronising call fail. */
LDM_API int ldm_unet_set_classes(ldm_ctx* ctx, const int64_t* c_dev, int batch, void* stream);

/* v3: pack the multi-conditional denoiser (v3:769-803) instead of the v2 one; ldm_unet_forward / ldm_sample /
 * ldm_ddpm_step then run v3's forward(x, t, flower, color) / p_sample / sample (v3:804-893) on the conditions set
 * with ldm_unet3_set_conditions.  Attention couples the rows of a call: a batch is one reference call. */
LDM_API int ldm_unet3_pack(ldm_ctx* ctx, const ldm_unet3_weights* w, void* stream);
LDM_API int ldm_unet3_set_conditions(ldm_ctx* ctx, const int64_t* flower_dev, const int64_t* color_dev, int batch,
                             void* stream);

/* ConditionalUNet.forward(x, t, c) (v2:535-561), eval mode. t_len is 1 or batch. */
LDM_API int ldm_unet_forward(ldm_ctx* ctx, const float* x_dev, const int64_t* t_dev, int t_len,
                     float* eps_out_dev, int batch, void* stream);

/* The posterior update of p_sample alone (v2:584-592) for a given eps:
 *   x <- (x - (1-alpha_t)/sqrt(1-alpha_bar_t) * eps) / sqrt(alpha_t) [+ sqrt(beta_t) * z if t > 0]
 * z = noise_dev if non-NULL, else Philox(seed; sample_offset + row, t) generated in-kernel.
 * x, eps (and noise): (batch, dim) fp32 row-major, dim a multiple of 4. */
LDM_API int ldm_ddpm_step(ldm_ctx* ctx, float* x_inout_dev, const float* eps_dev, int t,
                  const float* noise_dev, uint64_t seed, uint64_t sample_offset, int batch, int dim,
                  void* stream);

/* Standard normals from the in-kernel Philox stream (the x_T draw of v2:595 uses step = n_steps). */
LDM_API int ldm_randn(ldm_ctx* ctx, float* out_dev, uint64_t seed, uint64_t sample_offset, int step,
              int batch, int dim, void* stream);

/* ConditionalDenoiseDiffusion.sample / repeated p_sample (v2:580-598): runs timesteps
 * t_start, t_start-1, ..., t_end (inclusive) on x in place, classes from ldm_unet_set_classes.
 * noise_dev: NULL (in-kernel Philox) or (t_start - t_end + 1, batch, latent_dim) explicit draws,
 * slab j used at t = t_start - j (the t = 0 slab is ignored).  use_graph != 0 replays the whole
 * loop as one CUDA graph (captured and cached per (batch, t_start, t_end, noise mode)). */
LDM_API int ldm_sample(ldm_ctx* ctx, float* x_inout_dev, int t_start, int t_end, const float* noise_dev,
               uint64_t seed, uint64_t sample_offset, int batch, int use_graph, void* stream);

/* Repack Decoder weights (v2:242-278). Synchronises the stream. */
LDM_API int ldm_decoder_pack(ldm_ctx* ctx, const ldm_decoder_weights* w, void* stream);

/* SimpleAutoencoder.decode(z) (v2:355-357, 280-290): (batch, latent_dim) -> (batch,3,64,64) NCHW. */
LDM_API int ldm_decode(ldm_ctx* ctx, const float* z_dev, float* img_out_dev, int batch, void* stream);

/* generate_class_samples' hot section (v2:865-869) with HOST buffers: copies `c_host`
 * (batch int64 labels) to the device, draws x_T from Philox(seed; sample_offset..), runs the
 * full n_steps chain as one graph, decodes, and copies the (batch,3,64,64) images and
 * (optionally, may be NULL) the final latents back to host memory.  Synchronous on return. */
LDM_API int ldm_generate_host(ldm_ctx* ctx, const int64_t* c_host, int batch, uint64_t seed,
                      uint64_t sample_offset, float* img_out_host, float* latents_out_host,
                      void* stream);
/* The same for the v3 multi-conditional denoiser (v3:860-893 then decode): `flower_host`, `color_host` are batch int64
 * labels each; the batch is ONE reference call (its rows are coupled through the cross-batch attention, v3:832). */
LDM_API int ldm_generate3_host(ldm_ctx* ctx, const int64_t* flower_host, const int64_t* color_host, int batch, uint64_t seed,
                       uint64_t sample_offset, float* img_out_host, float* latents_out_host, void* stream);

/* ---- v4 / v5 pixel-space diffusion (bf16 = tcgen05 kernels, fp32 = strict CUDA-core path) ------------
 * ldm_pix_pack: repack SimpleUNet (v4:37-97) for the implicit-GEMM kernels and tabulate the per-stage time terms
 * time_fc_i(time_embed(t)) for t = 0..n_t-1 (v4:103-110 hoisted out of the loop).  Synchronises the stream. */
LDM_API int ldm_pix_pack(ldm_ctx* ctx, const ldm_pix_weights* w, void* stream);

/* SimpleUNet.forward(x, t) (v4:99-135 / v5:101-146): x (batch, 3, H, W) fp32 NCHW, t (batch,) fp32 (the reference
 * feeds `t.view(B, 1).float()` to a Linear, v4:104) -> eps (batch, 3, H, W) fp32 NCHW.  H, W: multiples of 4 whose
 * three resolution levels tile into 128-pixel boxes (64 x 64 and 32 x 32 do). */
LDM_API int ldm_pix_forward(ldm_ctx* ctx, const float* x_dev, const float* t_dev, float* eps_out_dev, int batch, int H, int W,
                    void* stream);

/* DiffusionModel.p_sample repeated for t = t_start .. t_end (v4:155-175) on x (batch, 3, H, W) in place; schedule from
 * ldm_set_schedule.  noise_dev: NULL (in-kernel Philox: counter quad = element / 4 of the flattened (3, H, W) sample,
 * global sample index sample_offset + n, step t) or (t_start - t_end + 1, batch, 3, H, W) explicit draws.  use_graph
 * != 0 replays the loop as one CUDA graph. */
LDM_API int ldm_pix_sample(ldm_ctx* ctx, float* x_inout_dev, int t_start, int t_end, const float* noise_dev, uint64_t seed,
                   uint64_t sample_offset, int batch, int H, int W, int use_graph, void* stream);

/* ---- conv U-Net blocks of the v2 script (module-level operators; bf16 = tcgen05, fp32 = strict) --------
 * *_pack copies and repacks one module's weights and returns a handle that lives until ldm_ctx_destroy.
 * ldm_ublock_res_forward  = UNetResidualBlock.forward(x, t, c) v2:475-486 (eval mode): x (B, Cin, H, W), t and c
 *   (B, d_time) embedding vectors (c may be NULL) -> out (B, Cout, H, W), all fp32 NCHW / row-major device pointers.
 * ldm_ublock_attn_forward = UNetAttentionBlock.forward(x) v2:444-459: x, out (B, C, H, W). H * W: a multiple of 8, >= 16. */
LDM_API int ldm_ublock_res_pack(ldm_ctx* ctx, const ldm_ublock_res_weights* w, int* handle_out, void* stream);
LDM_API int ldm_ublock_res_forward(ldm_ctx* ctx, int handle, const float* x_dev, const float* t_dev, const float* c_dev_or_null,
                           float* out_dev, int batch, int H, int W, void* stream);
LDM_API int ldm_ublock_attn_pack(ldm_ctx* ctx, const ldm_ublock_attn_weights* w, int* handle_out, void* stream);
/* Release one packed block (device weights + workspace); synchronises the device first.  The handle may be reused by a later pack. */
LDM_API int ldm_ublock_free(ldm_ctx* ctx, int handle);
LDM_API int ldm_ublock_attn_forward(ldm_ctx* ctx, int handle, const float* x_dev, float* out_dev, int batch, int H, int W,
                            void* stream);

/* Introspection for tests and the benchmark. */
LDM_API int ldm_kernel_launch_count(ldm_ctx* ctx, uint64_t* out); /* kernels launched (graph nodes count per replay) */
LDM_API int ldm_get_info(ldm_ctx* ctx, const char* key, double* out);

/* Profiling aid for the persistent loop kernel (libraries built with LDM_CHAIN_TRACE): ldm_debug_chain_trace(ctx, step,
 * NULL, 0) arms a clock64() timeline of reverse step `step` of the next ldm_sample (cluster 0 only; step < 0 switches it
 * off); ldm_debug_chain_trace(ctx, 0, out_host, n >= 16 * 5 * 96) synchronises and copies the stamps out as
 * [cluster rank][track][stamp]: tracks = two epilogue threads, weight producer, operand producer, MMA issuer;
 * stamp = clock | phase << 52 | tag << 56, in program order (0 = unused).  tools/chain_sweep.py decodes it. */
LDM_API int ldm_debug_chain_trace(ldm_ctx* ctx, int step, long long* out_host, int n);

/* Per-launch device times of everything the context launches on `stream` (eager launches only: nothing is recorded
 * while a graph is captured or replayed).  ldm_debug_ktrace(ctx, 1, stream, ...) starts a trace (CUDA event after every
 * kernel launch); ldm_debug_ktrace(ctx, 0, NULL, names, names_cap, ms, n_cap, &n) stops it, synchronises and returns the
 * n kernel names (newline separated, in launch order) and the milliseconds from the previous mark to the mark after
 * each kernel.  This is how bench.py fills `roofline_kernels` (CUDA events, not a profiler). */
LDM_API int ldm_debug_ktrace(ldm_ctx* ctx, int start, void* stream, char* names_out, int names_cap, float* ms_out, int n_cap,
                     int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* LDM_B200_H_ */
