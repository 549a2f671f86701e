// Shared declarations of the sm_100a sampling library (not part of the public ABI).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ldm_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the ABI
// ---------------------------------------------------------------------------------------------
void ldm_set_error(const char* fmt, ...);

#define LDM_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ldm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define LDM_CHECK(cond, ...)        \
  do {                              \
    if (!(cond)) {                  \
      ldm_set_error(__VA_ARGS__);   \
      return -1;                    \
    }                               \
  } while (0)

#define LDM_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float swishf(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// epilogue description shared by the fp32 (CUDA-core) and bf16 (tcgen05) GEMM kernels.
// Output element (row, col) of  acc = A[row,:] . W[col,:]  is finished as
//   v = acc + bias[col] + tab_t[tidx(row)*ld_t + col] + tab_c[cls[row]*ld_c + col] + resid[row*ld_r + col]
//   v = act(v)
// and then either stored (fp32 and/or bf16 copies) or consumed by the fused DDPM update.
// ---------------------------------------------------------------------------------------------
enum { LDM_ACT_NONE = 0, LDM_ACT_SWISH = 1, LDM_ACT_SIGMOID = 2 };

struct Epilogue {
  const float* bias = nullptr;      // [N]
  const float* tab_t = nullptr;     // per-timestep table (n_t, ld_t)
  int ld_t = 0;
  const int64_t* t_idx = nullptr;   // device timesteps (len t_len) or null -> t_const
  int t_len = 1;
  int t_const = 0;
  int n_t = 1;                      // rows of tab_t (indices are clamped; range is validated upstream)
  const float* tab_c = nullptr;     // per-class table (n_cls, ld_c)
  int ld_c = 0;
  const int32_t* cls = nullptr;     // [M] class of each row (validated copy kept by the context)
  const float* resid = nullptr;     // fp32 residual
  int ld_r = 0;
  int act = LDM_ACT_NONE;
  float* out_f32 = nullptr;
  int ld_of = 0;
  bf16* out_bf16 = nullptr;
  int ld_ob = 0;
  // attention operands straight from the in_proj GEMM (v3:832): columns < vt_col0 go to out_bf16 ([Q | K], row-major),
  // columns >= vt_col0 are written TRANSPOSED: vt[(col - vt_col0) * ld_vt + row]  (the B operand of P . V)
  bf16* vt = nullptr;
  int vt_col0 = 0, ld_vt = 0;
  // fused DDPM posterior update (v2:584-592): v is eps_theta
  int ddpm = 0;
  float* x = nullptr;               // (M, N) fp32 state, updated in place
  float c2 = 0.f, sqrt_alpha = 1.f, sigma = 0.f;  // (1-a_t)/sqrt(1-abar_t), sqrt(a_t), sqrt(beta_t) (0 at t=0)
  const float* noise = nullptr;     // explicit (M, N) draws or null -> Philox
  const unsigned long long* rng = nullptr;  // device {seed, sample_offset}: read at run time so a captured graph replays with new seeds
  int step = 0;
  // plain `bias + fp32 store` epilogues of wide outputs (Decoder.fc[3]): the tile is staged in shared memory and written row by
  // row, 512 contiguous bytes per warp instruction, instead of 32 rows x 16 bytes (gemm_tc_kernel only; needs out_f32)
  int stage_f32 = 0;
};

// ---------------------------------------------------------------------------------------------
// packed model + context
// ---------------------------------------------------------------------------------------------
struct DenseLayer {      // y = x W^T + b, W (N, K) row-major ("K-major")
  int N = 0, K = 0;
  float* w32 = nullptr;  // fp32 copy (strict path, and source of the bf16 copy)
  bf16* w16 = nullptr;   // bf16 copy (tensor-core path)
  float* b = nullptr;    // fp32 bias [N]
  CUtensorMap map_w;     // TMA descriptor of w16 (tensor-core path)
  int bn = 0;            // N tile of the tensor-core kernel for this layer
};

struct UnetModel {
  bool packed = false;
  std::vector<void*> allocs;
  int latent = 0, tdim = 0, ncls = 0, nst = 0, n_t = 0;
  int hid[LDM_MAX_STAGES + 1] = {0};
  int dmax = 0;
  DenseLayer latent_proj;
  DenseLayer block[LDM_MAX_STAGES], ov[LDM_MAX_STAGES], down[LDM_MAX_STAGES];
  DenseLayer fin;                               // K = hid[nst] + latent : [W_f | s W_f], bias (1+s) b_f
  int variant = 2;                              // 2: v2 (L = 1 attention folded into ov); 3: v3 (attention across the batch)
  int ncolors = 1;                              // v3: ncls = num_classes * num_colors condition pairs
  DenseLayer qkv[LDM_MAX_STAGES], attn_o[LDM_MAX_STAGES];   // v3: in_proj (3d x d) and out_proj
  float *ln_a_w[LDM_MAX_STAGES], *ln_a_b[LDM_MAX_STAGES];  // layers[i][0][1]
  float *ln_b_w[LDM_MAX_STAGES], *ln_b_b[LDM_MAX_STAGES];  // layers[i][1]
  float *ln_f_w = nullptr, *ln_f_b = nullptr;
  float* tab_t[LDM_MAX_STAGES + 1];             // (n_t, hid[i]): time_projections[i](time_emb(t)); [nst] = final_time_proj
  float* tab_c[LDM_MAX_STAGES + 1];             // (ncls, hid[i]): time_projections[i](class_emb(c)); [nst] = final_class_proj
  float s_res = 0.f;                            // sigmoid(residual_weight)
};

// ---------------------------------------------------------------------------------------------
// persistent cluster kernel of the reverse chain (chain.cu): folded per-phase weights in tile order
// ---------------------------------------------------------------------------------------------
#define LDM_CHAIN_CLUSTER 16      // CTAs per cluster (non-portable size; one cluster per GPC on B200)
#define LDM_CHAIN_ROWS 32         // batch rows a cluster carries through the whole chain
#define LDM_CHAIN_MAX_K 2048      // longest reduction of a phase (2 x widest hidden layer)
#define LDM_CHAIN_MAX_PHASES (LDM_MAX_STAGES + 2)
#define LDM_CHAIN_TRACE_TRACKS 5   // ldm_debug_chain_trace: stamped threads per CTA
#define LDM_CHAIN_TRACE_LEN 96     // stamps per thread
enum { LDM_PH_STAGE = 0, LDM_PH_FINAL_LN = 1, LDM_PH_MERGED = 3 };   // MERGED: stage tiles of the NEXT forward + eps tiles of the previous one

struct ChainPhaseHost {
  int type = 0, K = 0, tiles = 0, first = 0, ks = 1, d = 0, rows = 0;
  int nst_tiles = 0, eps_kb0 = 0;   // MERGED: leading stage tiles; first k-block the eps tiles read
  int dual = 0;                     // operand = raw h2 of the previous stage; weights [W1 | W2], K = width of ONE block
  float* q = nullptr;               // dual: W2 . 1 in tile order
  float* g0b = nullptr;             // MERGED: G_0 . b_fin in tile order (the eps bias seen through the next latent_proj / block Linear)
  bf16* w = nullptr;              // (rows, K) bf16, 128-row tiles
  float *bias = nullptr, *tab_t = nullptr, *tab_c = nullptr;
  CUtensorMap map;
};

struct ChainModel {
  bool ready = false;
  int n_phases = 0;
  ChainPhaseHost ph[LDM_CHAIN_MAX_PHASES];
  double peak_bytes_per_step = 0;  // L2 -> SM bytes of the busiest CTA of a cluster per reverse step
  std::vector<void*> allocs;
};

struct ConvLayer {       // implicit GEMM: out[pix, co] = sum_{tap, ci} in[pix + off(tap), ci] * w[co][tap*Cin + ci]
  int Cin = 0, Cout = 0, taps = 0;
  float* w32 = nullptr;  // (Cout, taps*Cin)
  bf16* w16 = nullptr;
  float* b = nullptr;
  CUtensorMap map_w;
  int bn = 0;            // output-channel tile of conv_tc_kernel (0: conv_tc_pick_bn(Cout))
  CUtensorMap map_w_alt; // same weights with a bn_alt-row box: used when the bn grid would not fill the SMs twice
  int bn_alt = 0;
};

struct ResBlockModel {
  int C = 0, HW = 0;     // channels, spatial side
  ConvLayer conv1, conv2;
  ConvLayer conv1s, conv2s;   // strict mode: (hi, hi, lo) bf16 split of the weights, Cin = 3 C (tensor-core path of the fp32 context)
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float *ca_w0, *ca_w2;  // (C/8, C), (C, C/8)
  float* sa_w;           // (2, 7, 7)
  float* ca_const;       // bf16 path: sigmoid(W2 swish(W1 ln2_b)) (GAP of an instance-norm output is its beta)
};

struct DecoderModel {
  bool packed = false;
  std::vector<void*> allocs;
  int latent = 0;
  DenseLayer fc0, fc3;   // fc3 rows permuted NCHW -> NHWC
  float *fc1_w, *fc1_b, *fc4_w, *fc4_b;  // fc4 affine permuted likewise
  ResBlockModel res[3];
  ConvLayer up[3][4];    // 4 sub-pixel parities (a*2+b), each 4 taps
  float *up_b[3], *up_gn_w[3], *up_gn_b[3];
  ConvLayer fin0, fin3;
  uint32_t* fin3_frag = nullptr;   // bf16 path: mma.sync B fragments of final_conv.3 (final_w_frag_kernel)
  float *fin_gn_w, *fin_gn_b;
  ConvLayer ups[3], fin0s;   // strict mode: split weights of the stacked ConvTranspose parities and of final_conv.0 (Cin = 3 C)
  bool strict_tc = false;    // the fp32 context runs its convolutions as three-term bf16 products on the tensor cores
};

struct GraphKey {
  int batch, t_start, t_end, noise_mode;
  bool operator<(const GraphKey& o) const {
    return std::tie(batch, t_start, t_end, noise_mode) < std::tie(o.batch, o.t_start, o.t_end, o.noise_mode);
  }
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  size_t n_nodes = 0;
  const float* noise = nullptr;   // captured explicit-noise pointer (must match for a replay)
  bool has_cls = false;           // captured with / without class tables
};

// v4 / v5 pixel-space SimpleUNet (pixel.cu)
struct PixModel {
  bool packed = false;
  std::vector<void*> allocs;
  int base = 0, temb = 0, n_t = 0;
  float *te0_w, *te0_b, *te2_w, *te2_b;     // time_embed
  float* tfc_w[3]; float* tfc_b[3];         // time_fc1..3
  float* tab = nullptr;                     // (n_t, 7 base): [time_fc1 | time_fc2 | time_fc3](time_embed(t)), t = 0..n_t-1
  float *in_w = nullptr, *in_b = nullptr;   // conv1.0 as (base, 27) tap-major (ky, kx, ci)
  bf16* in_w16 = nullptr;                   // base == 64: (64, 64) bf16 [w | w | 0] for the tensor-core conv1.0 kernel
  CUtensorMap in_map;
  float *out_w = nullptr, *out_b = nullptr; // out_conv as (3, 9 base)
  float* res_ratio = nullptr;               // device scalar (v5) or null (v4)
  ConvLayer out16;                          // out_conv padded to 16 output rows, bf16 (tensor-core halo kernel; base == 64)
  ConvLayer c1b, down1, c2a, c2b, down2, c3a, c3b, b0, b2, up1, c4a, c4b, up2, c5a, c5b;   // up1 / up2: 4 stacked sub-pixel kernels
  // activation workspace (NHWC bf16) for `cap` samples of cap_h x cap_w
  int cap = 0, cap_h = 0, cap_w = 0;
  std::vector<void*> ws;
  bf16 *a1, *cat5, *d1, *a2, *cat4, *d2, *a3, *x3, *bt, *x4, *a4, *x5, *a5, *x6;
  float* tsample = nullptr;                 // (cap, 7 base) per-sample time terms of forward(x, t)
  float* x_state = nullptr;                 // (cap, 3, H, W) fp32 chain state the captured graph works on
  float* eps = nullptr;                     // (cap, 3, H, W) fp32 eps of the current step (sampler)
  std::map<std::tuple<int, int, int, int, int, int>, GraphEntry> graphs;   // (batch, H, W, t_start, t_end, noise mode)
};

// stand-alone conv U-Net blocks of the v2 script (ublock.cu): type 1 = UNetResidualBlock, 2 = UNetAttentionBlock
struct UBlock {
  int type = 0, cin = 0, cout = 0, dt = 0, heads = 0;
  std::vector<void*> allocs, ws;
  float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;   // norm1 / norm (attention block), norm2
  float *tw = nullptr, *tb = nullptr, *cw = nullptr, *cb = nullptr;       // time_emb, class_emb Linears
  ConvLayer conv1, conv2;
  DenseLayer d1, d2;              // residual 1x1 (type 1); qkv and proj 1x1 (type 2)
  float* proj_perm = nullptr;     // fp32 contexts: proj weight with (head, head_dim)-ordered input columns
  size_t ws_elems = 0;
  bf16* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bf16* qkv3 = nullptr;
  float2* coef = nullptr;
  float* post = nullptr;
};

struct ldm_ctx {
  int device = 0;
  int precision = LDM_PRECISION_FP32;
  int sm_count = 148;
  std::vector<void*> allocs;      // every device allocation owned by the context
  // schedule
  int n_steps = 0;
  std::vector<float> c2, sqrt_alpha, sigma;
  // models
  UnetModel unet;
  ChainModel chain;
  DecoderModel dec;
  PixModel pix;
  std::vector<UBlock> ublocks;
  // per-batch state
  int cap = 0;                    // rows the activation workspace holds
  int batch_cls = -1;             // batch of the last set_classes (-1: none)
  bool has_cls = false;
  int32_t* cls = nullptr;         // [cap]
  int* dev_flags = nullptr;       // [0] out-of-range label / timestep seen
  // denoiser workspace (fp32 master copies + operand copies in the GEMM operand type)
  float *h = nullptr, *u = nullptr, *h2 = nullptr;   // (cap, dmax) fp32
  void *h_op = nullptr, *n_op = nullptr, *h3_op = nullptr;  // operand-typed (fp32 or bf16)
  void* af_op[2] = {nullptr, nullptr};  // [LN_f(h) | x] operand of the final GEMM, double-buffered across steps
  float* qkv = nullptr;                 // v3: (cap, 3 dmax) fp32 projections
  void* a_op = nullptr;                 // v3: attention output, operand of out_proj
  bf16 *qk16 = nullptr, *vt16 = nullptr; // v3 tensor-core attention: [Q | K] bf16 (cap, 2 dmax) and V^T bf16 (dmax, cap)
  float* x_state = nullptr;       // (cap, latent) fp32 chain state the captured graph works on
  unsigned long long* rng_dev = nullptr;  // {seed, sample_offset}
  cudaStream_t cap_stream = nullptr;      // capture-only stream (the caller may be on the legacy stream)
  std::vector<void*> ws_allocs;   // workspace allocations (freed on regrow)
  // decoder workspace
  int dec_cap = 0;
  std::vector<void*> dec_allocs;
  void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;  // ping-pong activation buffers (operand-typed, NHWC)
  float *d_f0 = nullptr, *d_f1 = nullptr;               // fp32 scratch (raw conv outputs, fc rows)
  float *d_stats = nullptr, *d_gap = nullptr, *d_ca = nullptr, *d_map = nullptr, *d_gate = nullptr;
  float2* d_part = nullptr;       // norm_coef: per-split partial sums
  int* d_cnt = nullptr;           // norm_coef: per (sample, channel block) tickets
  float* z_tmp = nullptr;
  bf16 *d_zb = nullptr, *d_h1b = nullptr;               // bf16 operands of Decoder.fc (tensor-core path)
  bf16* d_split = nullptr;                               // strict mode: (hi, lo, hi) bf16 thirds of the fp32 input of a convolution
  // generate_host staging
  int64_t* c_stage = nullptr;     // device staging of the labels copied from the host
  float* img_stage = nullptr;     // device images before the copy back
  int host_cap = 0;
  std::vector<void*> stage_allocs;
  // graphs
  std::map<GraphKey, GraphEntry> graphs;
  // TMA descriptors over the workspace (tensor-core path), rebuilt when the workspace regrows
  std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> act_maps;
  // accounting
  unsigned long long launches = 0;
  bool capturing = false;
  // persistent chain kernel (bf16 path)
  int chain_enabled = 0;          // LDM_CHAIN (default 1) and the device can co-schedule the clusters
  int use_chain = 0;              // 1: chain.cu runs the packed denoiser; 0: one kernel per layer (gemm_tc.cu + rowwise.cu)
  int chain_max_clusters = 0;     // co-resident clusters the device offers
  bf16* opbuf[LDM_MAX_STAGES] = {nullptr};   // (cap, hid[j]): h2 of stage j, operand of the next phase (chain kernel)
  bf16* caf[2] = {nullptr, nullptr};         // (cap, 3 latent): [x~ | -c_b LN_f(h) | -c_b x] operand of the merged phase, double-buffered over steps
  float4* coef_dev = nullptr;     // [n_steps] (c2, sqrt_alpha, sigma, 0)
  float4* coef_one = nullptr;     // one entry (1, 1, 0, 0): the 'coefficients' of a plain forward() in the chain kernel
  int* chain_err = nullptr;       // [2] first barrier timeout of the chain kernel: code, block
  long long* chain_trace = nullptr;   // [CS][TRACKS][LEN] tagged clock stamps (ldm_debug_chain_trace), null = off
  int chain_trace_step = 0;
  int use_pdl = 0;
  // persistent kernel of the v3 loop (v3loop.cu): folded phase list + workspace, null when the architecture is not covered
  void* v3loop = nullptr;
  int use_v3loop = 1;             // LDM_V3LOOP=0 keeps v3 on the per-layer path
  int use_attn_tc = 1;            // v3 bf16: attention on tcgen05 (LDM_ATTN_TC=0 selects the CUDA-core kernel)
  // per-launch timing (ldm_debug_ktrace): one CUDA event after every kernel launch on kt_stream while kt_on
  bool kt_on = false;
  cudaStream_t kt_stream = nullptr;
  std::vector<std::pair<std::string, cudaEvent_t>> kt_marks;
};

// Called right after a kernel launch: with tracing on, an event on the traced stream closes the interval of that kernel
// (kernels of a stream run back to back, so mark[i] - mark[i-1] is kernel i plus whatever gap precedes it).
static inline void ldm_kmark(ldm_ctx* ctx, const char* name) {
  if (!ctx->kt_on || ctx->capturing) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx->kt_stream);
  ctx->kt_marks.emplace_back(name, e);
}

#define LDM_LAUNCHED_AS(ctx, name)    \
  do {                                \
    (ctx)->launches++;                \
    ldm_kmark((ctx), name);           \
    LDM_CUDA(cudaGetLastError());     \
  } while (0)
#define LDM_LAUNCHED(ctx)             \
  do {                                \
    (ctx)->launches++;                \
    ldm_kmark((ctx), __func__);       \
    LDM_CUDA(cudaGetLastError());     \
  } while (0)

// Launch with programmatic dependent launch (PDL): the grid may start while its predecessor in the stream drains, runs its
// prologue (barrier init, TMEM allocation, descriptor prefetch, resident weights) and blocks in griddepcontrol.wait until
// the predecessor has completed and its writes are visible.
template <typename Kern, typename... Args>
static inline cudaError_t launch_maybe_pdl(Kern kern, dim3 grid, int threads, size_t smem, cudaStream_t st, int pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

__device__ __forceinline__ void ldm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ldm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


// v3loop.cu
int v3loop_pack(ldm_ctx* ctx, cudaStream_t st);
void v3loop_free(ldm_ctx* ctx);
int v3loop_supported(ldm_ctx* ctx, int B);
int v3loop_error(ldm_ctx* ctx, int* out);
#define LDM_V3LOOP_UNAVAILABLE (-77)      // launch_v3loop: the grid cannot be co-resident; the caller takes the per-layer path
int launch_v3loop(ldm_ctx* ctx, int B, int n_iter, int t_start, int sample, const int64_t* t_idx, int t_len, float* x, float* eps_out,
                  const float* noise, cudaStream_t st);

// memory helpers (api.cu)
int ldm_alloc(ldm_ctx* ctx, std::vector<void*>& pool, void** out, size_t bytes);
template <typename T>
static inline int ldm_alloc_t(ldm_ctx* ctx, std::vector<void*>& pool, T** out, size_t count) {
  return ldm_alloc(ctx, pool, (void**)out, count * sizeof(T));
}

// ---------------------------------------------------------------------------------------------
// kernel launchers (one .cu each); all return 0 or an error code and count launches in ctx
// ---------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM  C = epi(A (M,K; lda) . W (N,K)^T)
int launch_gemm_f32(ldm_ctx* ctx, const float* A, int lda, const float* W, int M, int N, int K,
                    const Epilogue& epi, cudaStream_t st);

// implicit-GEMM convolution over NHWC fp32 activations (strict path)
struct ConvGeom {
  int B, H, W;            // input batch and spatial size
  int Cin, Cout;
  int taps;
  int dy[16], dx[16];     // input offset of each tap
  int up;                 // 1: plain conv (output H x W); 2: sub-pixel of a stride-2 transposed conv
  int pa, pb;             // sub-pixel parity (output pixel (2i+pa, 2j+pb)) when up == 2
  int nchw_out;           // 1: write (B, Cout, H, W) fp32 instead of NHWC
  int act;
  // extensions used by the pixel path (0 = the defaults of the decoder's calls):
  int stride;             // 2: H, W are the OUTPUT size and tap (dy, dx) of pixel (y, x) reads input (2y + dy, 2x + dx)
  int in_pitch, out_pitch;   // channel pitch of the NHWC buffers (0: Cin / Cout)
  int relu;               // ReLU after the bias
  const float* post;      // per-channel term added after the activation, row n * post_stride
  int post_stride;
};
int launch_conv_f32(ldm_ctx* ctx, const float* in, const float* w, const float* bias, float* out,
                    const ConvGeom& g, cudaStream_t st);

// rowwise / normalisation kernels
template <typename TOP>
int launch_stage_mid(ldm_ctx* ctx, const float* u, const float* h, const float* ga, const float* ba,
                     const float* gb, const float* bb, float* h2, TOP* n_op, int ld_op, int M, int d,
                     cudaStream_t st);
template <typename TOP>
int launch_row_ln(ldm_ctx* ctx, const float* in, int ld_in, const float* g, const float* b, int act,
                  TOP* out, int ld_out, int M, int d, cudaStream_t st);
template <typename TOP>
int launch_load_x(ldm_ctx* ctx, const float* x, TOP* dst, int ld_dst, int M, int d, cudaStream_t st);
int launch_set_classes(ldm_ctx* ctx, const int64_t* c, int32_t* out, int M, int ncls, int* flags,
                       cudaStream_t st);
int launch_set_conditions(ldm_ctx* ctx, const int64_t* f, const int64_t* k, int32_t* out, int M, int nf, int nk, int* flags,
                          cudaStream_t st);
int launch_cond_pairs(ldm_ctx* ctx, const float* fe, const float* ke, float* out, int nf, int nk, int td, cudaStream_t st);
template <typename TOP>
int launch_batch_attention(ldm_ctx* ctx, const float* qkv, TOP* out, int B, int d, int heads, cudaStream_t st);
int launch_check_t(ldm_ctx* ctx, const int64_t* t, int n, int n_t, int* flags, cudaStream_t st);
int launch_ddpm_update(ldm_ctx* ctx, float* x, const float* eps, float c2, float sqrt_alpha, float sigma,
                       const float* noise, unsigned long long seed, unsigned long long sample_offset,
                       int step, int M, int d, cudaStream_t st);
int launch_set_rng(ldm_ctx* ctx, unsigned long long* rng, unsigned long long seed, unsigned long long sample_offset,
                   cudaStream_t st);
int launch_randn(ldm_ctx* ctx, float* out, unsigned long long seed, unsigned long long sample_offset,
                 int step, int M, int d, cudaStream_t st);

// pack-time helpers (fp64 accumulation, run once)
int launch_pack_matmul_nn(ldm_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K,
                          cudaStream_t st);  // C(M,N) = A(M,K) B(K,N)
int launch_pack_matvec(ldm_ctx* ctx, const float* A, const float* x, const float* add, float* y, int M,
                       int K, cudaStream_t st);  // y = A x + add
int launch_pack_final(ldm_ctx* ctx, const float* wf, const float* bf, const float* rw, float* wcat,
                      float* bcat, float* s_out, int N, int K, cudaStream_t st);
int launch_to_bf16(ldm_ctx* ctx, const float* in, bf16* out, size_t n, cudaStream_t st);
int launch_permute_rows(ldm_ctx* ctx, const float* in, float* out, int C, int P, int K, cudaStream_t st);
int launch_pack_conv(ldm_ctx* ctx, const float* w, float* out, int Cout, int Cin, int KH, int KW,
                     cudaStream_t st);
int launch_pack_convT(ldm_ctx* ctx, const float* w, float* out, int Cin, int Cout, int pa, int pb,
                      cudaStream_t st);
int launch_ca_const(ldm_ctx* ctx, const float* beta, const float* w0, const float* w2, float* out, int C,
                    cudaStream_t st);

// decoder normalisation / gating kernels (activation type T = float or bf16, NHWC)
template <typename T>
int launch_inorm_stats(ldm_ctx* ctx, const T* x, float* stats, int B, int HW, int C, int group,
                       cudaStream_t st);  // stats[(n*G+g)*2] = mean, rstd over HW x group channels
template <typename T>
int launch_norm_apply(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta,
                      T* out, int B, int HW, int C, int group, int act, cudaStream_t st);
template <typename T>
int launch_gap_norm(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta,
                    float* gap, int B, int HW, int C, cudaStream_t st);
int launch_ca_mlp(ldm_ctx* ctx, const float* gap, const float* w0, const float* w2, float* ca, int B, int C,
                  cudaStream_t st);
template <typename T>
int launch_sa_map(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta,
                  const float* ca, int ca_stride, float* map, int B, int HW, int C, cudaStream_t st);
template <typename T>
int launch_sa_apply(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta,
                    const float* ca, int ca_stride, const float* map, const float* sa_w, const T* resid,
                    T* out, int B, int H, int C, cudaStream_t st);

// persistent chain kernel (chain.cu)
int chain_init(ldm_ctx* ctx);
int chain_pack(ldm_ctx* ctx, cudaStream_t st);
int launch_chain(ldm_ctx* ctx, int B, int n_iter, int t_start, int sample, const int64_t* t_idx, int t_len, float* x,
                 float* eps_out, const float* noise, cudaStream_t st);

// tensor-core path (gemm_tc.cu)
int tc_init(ldm_ctx* ctx);
int tc_make_weight_map(ldm_ctx* ctx, const bf16* w, int N, int K, int bn, CUtensorMap* out);
int launch_gemm_tc(ldm_ctx* ctx, const bf16* A, int lda, int M, const DenseLayer& L, const Epilogue& epi,
                   cudaStream_t st);
int attn_tc_supported(int hd);
int launch_attn_prep(ldm_ctx* ctx, const float* qkv, bf16* qk, bf16* vt, int rows, int d, int ldv, cudaStream_t st);
int launch_attn_tc(ldm_ctx* ctx, const bf16* qk, int ld_qk, int qk_cols, const bf16* vt, int ldv, int L, int batches, int heads, int hd,
                   int q_col0, int k_col0, void* out, int out_bf16, int out_pitch, int out_sh, int out_se, cudaStream_t st);
