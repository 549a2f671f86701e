// C ABI of the sampling library (include/ldm_b200.h): context, weight packing, the per-step kernel
// sequence of the denoiser (v2:535-561) with the fused posterior update (v2:580-592), whole-loop CUDA
// graph capture (v2:594-598), and the decoder pass (v2:280-290).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <math.h>
#include <type_traits>

#include "common.cuh"

int tc_error_flag(int* out);
int tc_error_reset();
int conv_tc_error_flag(int* out);
int tc_pick_bn(int M, int N);
int decoder_pack_impl(ldm_ctx* ctx, const ldm_decoder_weights* w, cudaStream_t st);
int decoder_run_impl(ldm_ctx* ctx, const float* z, float* img, int B, cudaStream_t st);

// -------------------------------------------------------------------------------------------------
// errors / memory
// -------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void ldm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ldm_alloc(ldm_ctx* ctx, std::vector<void*>& pool, void** out, size_t bytes) {
  (void)ctx;
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  LDM_CUDA(cudaMalloc(&p, bytes));
  pool.push_back(p);
  *out = p;
  return 0;
}

static void free_pool(std::vector<void*>& pool) {
  for (void* p : pool) cudaFree(p);
  pool.clear();
}

void pix_drop_graphs(ldm_ctx* ctx);
void pix_free(ldm_ctx* ctx);
void ublock_free_all(ldm_ctx* ctx);

static void drop_graphs(ldm_ctx* ctx) {
  pix_drop_graphs(ctx);
  for (auto& kv : ctx->graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
  }
  ctx->graphs.clear();
}

// copy a caller tensor into a context-owned fp32 buffer
static int own_copy(ldm_ctx* ctx, std::vector<void*>& pool, const float* src, size_t n, float** out, cudaStream_t st) {
  LDM_CHECK(src != nullptr, "null weight pointer");
  LDM_TRY(ldm_alloc_t(ctx, pool, out, n));
  LDM_CUDA(cudaMemcpyAsync(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int finish_dense(ldm_ctx* ctx, std::vector<void*>& pool, DenseLayer& L, int M_hint, cudaStream_t st) {
  if (ctx->precision != LDM_PRECISION_BF16) return 0;
  LDM_TRY(ldm_alloc_t(ctx, pool, &L.w16, (size_t)L.N * L.K));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, (size_t)L.N * L.K, st));
  L.bn = tc_pick_bn(M_hint, L.N);
  LDM_TRY(tc_make_weight_map(ctx, L.w16, L.N, L.K, L.bn, &L.map_w));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
extern "C" LDM_API int ldm_version(void) { return LDM_ABI_VERSION; }
extern "C" LDM_API const char* ldm_last_error(void) { return g_err; }

extern "C" LDM_API int ldm_ctx_create(ldm_ctx** out, int device, int precision) {
  LDM_CHECK(out != nullptr, "ldm_ctx_create: out is NULL");
  *out = nullptr;
  LDM_CHECK(precision == LDM_PRECISION_FP32 || precision == LDM_PRECISION_BF16,
            "ldm_ctx_create: precision must be LDM_PRECISION_FP32 (0) or LDM_PRECISION_BF16 (2), got %d", precision);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    ldm_set_error("ldm_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return e != cudaSuccess ? (int)e : -1;
  }
  LDM_CHECK(device >= 0 && device < ndev, "ldm_ctx_create: device %d out of range [0,%d)", device, ndev);
  LDM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LDM_CUDA(cudaGetDeviceProperties(&prop, device));
  LDM_CHECK(prop.major == 10, "ldm_ctx_create: built for sm_100a only, device %d is sm_%d%d", device, prop.major, prop.minor);
  ldm_ctx* ctx = new ldm_ctx();
  ctx->device = device;
  ctx->precision = precision;
  ctx->sm_count = prop.multiProcessorCount;
  int r = ldm_alloc_t(ctx, ctx->allocs, &ctx->dev_flags, 4);
  if (r == 0) r = ldm_alloc_t(ctx, ctx->allocs, &ctx->rng_dev, 2);
  if (r == 0) r = (int)cudaMemset(ctx->dev_flags, 0, 4 * sizeof(int));
  if (r == 0) r = (int)cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking);
  if (r == 0 && precision == LDM_PRECISION_BF16) r = tc_init(ctx);
  if (r == 0) r = ldm_alloc_t(ctx, ctx->allocs, &ctx->chain_err, 2);
  if (r == 0) r = (int)cudaMemset(ctx->chain_err, 0, 2 * sizeof(int));
  if (r == 0) r = ldm_alloc_t(ctx, ctx->allocs, &ctx->coef_one, 1);
  if (r == 0) { const float4 one = make_float4(1.f, 1.f, 0.f, 0.f); r = (int)cudaMemcpy(ctx->coef_one, &one, sizeof(one), cudaMemcpyHostToDevice); }
  {
    // The persistent cluster kernel is the denoiser of both modes: bf16 operands (2e-2), or, in the strict mode, bf16
    // (hi, lo) splits of weights and operands with the three products hi.hi + hi.lo + lo.hi on the tensor cores (1e-3).
    // LDM_CHAIN=0 selects the one-kernel-per-layer sequence (strict mode: fp32 FMA on the CUDA cores).
    const char* ch = getenv("LDM_CHAIN");
    if (r == 0 && precision != LDM_PRECISION_BF16) r = tc_init(ctx);
    ctx->chain_enabled = r == 0 ? (ch ? atoi(ch) : 1) : 0;
    if (ctx->chain_enabled && chain_init(ctx) != 0) ctx->chain_enabled = 0;   // ldm_last_error() keeps the reason
    ctx->use_chain = ctx->chain_enabled;
  }
  const char* atc = getenv("LDM_ATTN_TC");
  ctx->use_attn_tc = atc ? atoi(atc) : 1;
  const char* pdl = getenv("LDM_PDL");
  ctx->use_pdl = pdl ? atoi(pdl) : 1;      // programmatic dependent launch of the decoder / pixel-path kernels (LDM_PDL=0: plain stream order)
  if (r != 0) {
    free_pool(ctx->allocs);
    delete ctx;
    return r;
  }
  *out = ctx;
  return 0;
}

extern "C" LDM_API int ldm_ctx_destroy(ldm_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  drop_graphs(ctx);
  for (auto& m : ctx->kt_marks) cudaEventDestroy(m.second);
  ctx->kt_marks.clear();
  free_pool(ctx->allocs);
  free_pool(ctx->ws_allocs);
  free_pool(ctx->dec_allocs);
  free_pool(ctx->unet.allocs);
  free_pool(ctx->dec.allocs);
  free_pool(ctx->stage_allocs);
  free_pool(ctx->chain.allocs);
  v3loop_free(ctx);
  pix_free(ctx);
  ublock_free_all(ctx);
  if (ctx->coef_dev) cudaFree(ctx->coef_dev);
  if (ctx->chain_trace) cudaFree(ctx->chain_trace);
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  delete ctx;
  return 0;
}

extern "C" LDM_API int ldm_set_schedule(ldm_ctx* ctx, const float* beta, const float* alpha, const float* alpha_bar, int n_steps) {
  LDM_CHECK(ctx && beta && alpha && alpha_bar && n_steps > 0, "ldm_set_schedule: bad arguments");
  ctx->n_steps = n_steps;
  ctx->c2.resize(n_steps);
  ctx->sqrt_alpha.resize(n_steps);
  ctx->sigma.resize(n_steps);
  for (int t = 0; t < n_steps; ++t) {
    // fp32, operation for operation as v2:584-590: (1 - alpha_t) / sqrt(1 - alpha_bar_t), sqrt(alpha_t), sqrt(beta_t)
    volatile float one_minus_a = 1.0f - alpha[t];
    volatile float one_minus_ab = 1.0f - alpha_bar[t];
    volatile float den = sqrtf(one_minus_ab);
    ctx->c2[t] = one_minus_a / den;
    ctx->sqrt_alpha[t] = sqrtf(alpha[t]);
    ctx->sigma[t] = t > 0 ? sqrtf(beta[t]) : 0.0f;   // v2:588: no noise at t = 0
  }
  drop_graphs(ctx);
  {  // device copy for the persistent chain kernel: (c2, sqrt_alpha, sigma, 0) per timestep
    LDM_CUDA(cudaSetDevice(ctx->device));
    LDM_CUDA(cudaDeviceSynchronize());
    if (ctx->coef_dev) { cudaFree(ctx->coef_dev); ctx->coef_dev = nullptr; }
    std::vector<float4> h(n_steps);
    for (int t = 0; t < n_steps; ++t) h[t] = make_float4(ctx->c2[t], ctx->sqrt_alpha[t], ctx->sigma[t], 0.f);
    LDM_CUDA(cudaMalloc((void**)&ctx->coef_dev, sizeof(float4) * n_steps));
    LDM_CUDA(cudaMemcpy(ctx->coef_dev, h.data(), sizeof(float4) * n_steps, cudaMemcpyHostToDevice));
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------
// denoiser packing
// -------------------------------------------------------------------------------------------------
static int dense_from(ldm_ctx* ctx, std::vector<void*>& pool, DenseLayer& L, const float* w, const float* b, int N, int K,
                      cudaStream_t st) {
  L.N = N;
  L.K = K;
  LDM_TRY(own_copy(ctx, pool, w, (size_t)N * K, &L.w32, st));
  LDM_TRY(own_copy(ctx, pool, b, (size_t)N, &L.b, st));
  return 0;
}

extern "C" LDM_API int ldm_unet_pack(ldm_ctx* ctx, const ldm_unet_weights* w, void* stream) {
  LDM_CHECK(ctx && w, "ldm_unet_pack: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  UnetModel& U = ctx->unet;
  const int nst = w->n_stages;
  LDM_CHECK(nst >= 1 && nst <= LDM_MAX_STAGES, "ldm_unet_pack: n_stages %d out of range", nst);
  LDM_CHECK(w->latent_dim % 64 == 0 && w->time_dim % 64 == 0 && w->latent_dim > 0 && w->time_dim > 0,
            "ldm_unet_pack: latent_dim and time_emb_dim must be multiples of 64");
  for (int i = 0; i <= nst; ++i)
    LDM_CHECK(w->hidden[i] % 128 == 0 && w->hidden[i] >= 128 && w->hidden[i] <= 1024,
              "ldm_unet_pack: hidden_dims[%d] = %d must be a multiple of 128 in [128, 1024]", i, w->hidden[i]);
  LDM_CHECK(w->hidden[nst] == w->latent_dim, "ldm_unet_pack: hidden_dims[-1] must equal latent_dim (final is applied to x, v2:561)");
  LDM_CHECK(w->n_t >= 1 && w->num_classes >= 1 && w->sinusoid, "ldm_unet_pack: sinusoid table missing");
  cudaDeviceSynchronize();
  drop_graphs(ctx);
  v3loop_free(ctx);
  free_pool(U.allocs);
  U = UnetModel();
  U.latent = w->latent_dim; U.tdim = w->time_dim; U.ncls = w->num_classes; U.nst = nst; U.n_t = w->n_t;
  for (int i = 0; i <= nst; ++i) { U.hid[i] = w->hidden[i]; U.dmax = U.hid[i] > U.dmax ? U.hid[i] : U.dmax; }
  auto& P = U.allocs;
  const int td = U.tdim;
  const int M_hint = 256;

  // --- embeddings hoisted out of the loop: TE = time_emb(t) for every t, CE = class_emb(c) for every class
  float *sinus, *te1, *te, *ce0, *ce1, *ce;
  std::vector<void*> tmp;
  LDM_TRY(own_copy(ctx, tmp, w->sinusoid, (size_t)U.n_t * td, &sinus, st));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &te1, (size_t)U.n_t * 2 * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &te, (size_t)U.n_t * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &ce1, (size_t)U.ncls * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &ce, (size_t)U.ncls * td));
  ce0 = const_cast<float*>(w->class_embedding);
  LDM_CHECK(w->time_lin1_w && w->time_lin1_b && w->time_lin2_w && w->time_lin2_b && w->class_embedding && w->class_lin1_w &&
            w->class_lin1_b && w->class_lin2_w && w->class_lin2_b, "ldm_unet_pack: embedding weights missing");
  {
    Epilogue e; e.bias = w->time_lin1_b; e.act = LDM_ACT_SWISH; e.out_f32 = te1; e.ld_of = 2 * td;
    LDM_TRY(launch_gemm_f32(ctx, sinus, td, w->time_lin1_w, U.n_t, 2 * td, td, e, st));        // v2:418 lin1 + act
    Epilogue e2; e2.bias = w->time_lin2_b; e2.out_f32 = te; e2.ld_of = td;
    LDM_TRY(launch_gemm_f32(ctx, te1, 2 * td, w->time_lin2_w, U.n_t, td, 2 * td, e2, st));     // v2:418 lin2
    Epilogue e3; e3.bias = w->class_lin1_b; e3.act = LDM_ACT_SWISH; e3.out_f32 = ce1; e3.ld_of = td;
    LDM_TRY(launch_gemm_f32(ctx, ce0, td, w->class_lin1_w, U.ncls, td, td, e3, st));           // v2:430-431
    Epilogue e4; e4.bias = w->class_lin2_b; e4.out_f32 = ce; e4.ld_of = td;
    LDM_TRY(launch_gemm_f32(ctx, ce1, td, w->class_lin2_w, U.ncls, td, td, e4, st));
  }
  // --- per-stage bias tables: T_i[t] = tp_i(TE[t]), C_i[c] = tp_i(CE[c]) (same Linear, bias in both: v2:541-545)
  for (int i = 0; i <= nst; ++i) {
    const int d = U.hid[i];
    const float* tw = i < nst ? w->time_proj_w[i] : w->final_time_w;
    const float* tb = i < nst ? w->time_proj_b[i] : w->final_time_b;
    const float* cw = i < nst ? w->time_proj_w[i] : w->final_class_w;
    const float* cb = i < nst ? w->time_proj_b[i] : w->final_class_b;
    LDM_CHECK(tw && tb && cw && cb, "ldm_unet_pack: projection weights of stage %d missing", i);
    LDM_TRY(ldm_alloc_t(ctx, P, &U.tab_t[i], (size_t)U.n_t * d));
    LDM_TRY(ldm_alloc_t(ctx, P, &U.tab_c[i], (size_t)U.ncls * d));
    Epilogue e; e.bias = tb; e.out_f32 = U.tab_t[i]; e.ld_of = d;
    LDM_TRY(launch_gemm_f32(ctx, te, td, tw, U.n_t, d, td, e, st));
    Epilogue e2; e2.bias = cb; e2.out_f32 = U.tab_c[i]; e2.ld_of = d;
    LDM_TRY(launch_gemm_f32(ctx, ce, td, cw, U.ncls, d, td, e2, st));
  }
  // --- dense layers
  LDM_TRY(dense_from(ctx, P, U.latent_proj, w->latent_proj_w, w->latent_proj_b, U.hid[0], U.latent, st));
  LDM_TRY(finish_dense(ctx, P, U.latent_proj, M_hint, st));
  for (int i = 0; i < nst; ++i) {
    const int d = U.hid[i], dn = U.hid[i + 1];
    LDM_CHECK(w->block_lin_w[i] && w->attn_in_proj_w[i] && w->attn_in_proj_b[i] && w->attn_out_w[i] && w->attn_out_b[i] &&
              w->down_w[i] && w->block_ln_w[i] && w->stage_ln_w[i], "ldm_unet_pack: weights of stage %d missing", i);
    LDM_TRY(dense_from(ctx, P, U.block[i], w->block_lin_w[i], w->block_lin_b[i], d, d, st));
    LDM_TRY(finish_dense(ctx, P, U.block[i], M_hint, st));
    // L = 1 attention == out_proj(V(.)): W_ov = W_out . W_v, b_ov = W_out . b_v + b_out (V = in_proj rows [2d, 3d))
    U.ov[i].N = d; U.ov[i].K = d;
    LDM_TRY(ldm_alloc_t(ctx, P, &U.ov[i].w32, (size_t)d * d));
    LDM_TRY(ldm_alloc_t(ctx, P, &U.ov[i].b, (size_t)d));
    LDM_TRY(launch_pack_matmul_nn(ctx, w->attn_out_w[i], w->attn_in_proj_w[i] + (size_t)2 * d * d, U.ov[i].w32, d, d, d, st));
    LDM_TRY(launch_pack_matvec(ctx, w->attn_out_w[i], w->attn_in_proj_b[i] + 2 * d, w->attn_out_b[i], U.ov[i].b, d, d, st));
    LDM_TRY(finish_dense(ctx, P, U.ov[i], M_hint, st));
    LDM_TRY(dense_from(ctx, P, U.down[i], w->down_w[i], w->down_b[i], dn, d, st));
    LDM_TRY(finish_dense(ctx, P, U.down[i], M_hint, st));
    LDM_TRY(own_copy(ctx, P, w->block_ln_w[i], d, &U.ln_a_w[i], st));
    LDM_TRY(own_copy(ctx, P, w->block_ln_b[i], d, &U.ln_a_b[i], st));
    LDM_TRY(own_copy(ctx, P, w->stage_ln_w[i], d, &U.ln_b_w[i], st));
    LDM_TRY(own_copy(ctx, P, w->stage_ln_b[i], d, &U.ln_b_b[i], st));
  }
  LDM_TRY(own_copy(ctx, P, w->final_norm_w, U.hid[nst], &U.ln_f_w, st));
  LDM_TRY(own_copy(ctx, P, w->final_norm_b, U.hid[nst], &U.ln_f_b, st));
  // --- final(h) + s final(x) as ONE contraction over K = [h | x]: [W_f | s W_f], bias (1+s) b_f  (v2:560-561)
  {
    LDM_CHECK(w->final_w && w->final_b && w->residual_weight, "ldm_unet_pack: final weights missing");
    const int N = U.latent, K = U.hid[nst];
    U.fin.N = N; U.fin.K = 2 * K;
    float* s_dev;
    LDM_TRY(ldm_alloc_t(ctx, P, &U.fin.w32, (size_t)N * 2 * K));
    LDM_TRY(ldm_alloc_t(ctx, P, &U.fin.b, (size_t)N));
    LDM_TRY(ldm_alloc_t(ctx, tmp, &s_dev, 1));
    LDM_TRY(launch_pack_final(ctx, w->final_w, w->final_b, w->residual_weight, U.fin.w32, U.fin.b, s_dev, N, K, st));
    LDM_TRY(finish_dense(ctx, P, U.fin, M_hint, st));
    LDM_CUDA(cudaMemcpyAsync(&U.s_res, s_dev, sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  cudaError_t e = cudaStreamSynchronize(st);
  free_pool(tmp);
  LDM_CUDA(e);
  // the persistent chain kernel covers the architectures chain_pack accepts; any other ConditionalUNet shape runs the
  // one-kernel-per-layer sequence (ldm_get_info("chain") tells which; ldm_last_error() keeps chain_pack's reason)
  ctx->use_chain = ctx->chain_enabled;
  if (ctx->use_chain && chain_pack(ctx, st) != 0) ctx->use_chain = 0;
  // the activation workspace depends on the path (operand buffers of the chain): rebuild it on the next call
  cudaDeviceSynchronize();
  ctx->act_maps.clear();
  free_pool(ctx->ws_allocs);
  ctx->cap = 0;
  ctx->batch_cls = -1;
  ctx->has_cls = false;
  U.packed = true;
  return 0;
}

// -------------------------------------------------------------------------------------------------
// workspace
// -------------------------------------------------------------------------------------------------
static int ensure_workspace(ldm_ctx* ctx, int B) {
  UnetModel& U = ctx->unet;
  LDM_CHECK(U.packed, "denoiser weights not packed (call ldm_unet_pack first)");
  if (B <= ctx->cap) return 0;
  LDM_CHECK(!ctx->capturing, "workspace growth during graph capture");
  cudaDeviceSynchronize();
  drop_graphs(ctx);
  ctx->act_maps.clear();
  free_pool(ctx->ws_allocs);
  const int cap = ceil_div(B, 128) * 128;
  const size_t op = ctx->precision == LDM_PRECISION_BF16 ? 2 : 4;
  auto& P = ctx->ws_allocs;
  const size_t nd = (size_t)cap * U.dmax, naf = (size_t)cap * (U.hid[U.nst] + U.latent);
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->h, nd));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->u, nd));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->h2, nd));
  LDM_TRY(ldm_alloc(ctx, P, &ctx->n_op, nd * op));
  LDM_TRY(ldm_alloc(ctx, P, &ctx->h3_op, nd * op));
  if (op == 2) LDM_TRY(ldm_alloc(ctx, P, &ctx->h_op, nd * op));
  else ctx->h_op = ctx->h;      // strict path: the fp32 master copy is the operand
  LDM_TRY(ldm_alloc(ctx, P, &ctx->af_op[0], naf * op));
  LDM_TRY(ldm_alloc(ctx, P, &ctx->af_op[1], naf * op));
  if (ctx->use_chain)
  {
    // + 64 rows: the last cluster of a launch writes its (up to 64-row) block unconditionally; the tensor maps stop at B.
    // Strict mode: every operand row is its (hi, lo, hi) thirds.
    const size_t capc = (size_t)cap + 64, mult = ctx->precision == LDM_PRECISION_BF16 ? 1 : 3;
    for (int j = 0; j < U.nst; ++j) LDM_TRY(ldm_alloc_t(ctx, P, &ctx->opbuf[j], capc * U.hid[j] * mult));
    for (int k = 0; k < 2; ++k) {
      LDM_TRY(ldm_alloc_t(ctx, P, &ctx->caf[k], capc * 3 * U.latent * mult));
      LDM_CUDA(cudaMemset(ctx->caf[k], 0, capc * 3 * U.latent * mult * sizeof(bf16)));
    }
  }
  if (U.variant == 3) {
    LDM_TRY(ldm_alloc_t(ctx, P, &ctx->qkv, (size_t)cap * 3 * U.dmax));
    LDM_TRY(ldm_alloc(ctx, P, &ctx->a_op, nd * op));
    LDM_CUDA(cudaMemset(ctx->a_op, 0, nd * op));
    if (op == 2) {
      LDM_TRY(ldm_alloc_t(ctx, P, &ctx->qk16, (size_t)cap * 2 * U.dmax));
      LDM_TRY(ldm_alloc_t(ctx, P, &ctx->vt16, (size_t)U.dmax * cap));
      LDM_CUDA(cudaMemset(ctx->qk16, 0, (size_t)cap * 2 * U.dmax * sizeof(bf16)));
      LDM_CUDA(cudaMemset(ctx->vt16, 0, (size_t)U.dmax * cap * sizeof(bf16)));
    }
  }
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->x_state, (size_t)cap * U.latent));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->cls, (size_t)cap));
  // zero everything once: rows beyond the batch are read by full 128-row TMA boxes
  LDM_CUDA(cudaMemset(ctx->n_op, 0, nd * op));
  LDM_CUDA(cudaMemset(ctx->h3_op, 0, nd * op));
  LDM_CUDA(cudaMemset(ctx->h_op, 0, op == 2 ? nd * op : nd * 4));
  LDM_CUDA(cudaMemset(ctx->af_op[0], 0, naf * op));
  LDM_CUDA(cudaMemset(ctx->af_op[1], 0, naf * op));
  ctx->cap = cap;
  ctx->batch_cls = -1;
  ctx->has_cls = false;
  return 0;
}

static int ensure_workspace(ldm_ctx* ctx, int B);

// -------------------------------------------------------------------------------------------------
// v3 multi-conditional denoiser (v3/model_train_test.py:739-853)
// -------------------------------------------------------------------------------------------------
extern "C" LDM_API int ldm_unet3_pack(ldm_ctx* ctx, const ldm_unet3_weights* w, void* stream) {
  LDM_CHECK(ctx && w, "ldm_unet3_pack: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  UnetModel& U = ctx->unet;
  const int nst = w->n_stages;
  LDM_CHECK(nst >= 1 && nst <= LDM_MAX_STAGES, "ldm_unet3_pack: n_stages %d out of range", nst);
  LDM_CHECK(w->latent_dim % 64 == 0 && w->time_dim % 64 == 0 && w->latent_dim > 0 && w->time_dim > 0,
            "ldm_unet3_pack: latent_dim and time_emb_dim must be multiples of 64");
  for (int i = 0; i <= nst; ++i)
    LDM_CHECK(w->hidden[i] % 128 == 0 && w->hidden[i] >= 128 && w->hidden[i] <= 1024,
              "ldm_unet3_pack: hidden_dims[%d] = %d must be a multiple of 128 in [128, 1024]", i, w->hidden[i]);
  LDM_CHECK(w->n_t >= 1 && w->num_classes >= 1 && w->num_colors >= 1 && w->sinusoid, "ldm_unet3_pack: sinusoid table missing");
  LDM_CHECK(w->time_lin1_w && w->time_lin1_b && w->time_lin2_w && w->time_lin2_b && w->flower_emb && w->color_emb && w->cond_fc_w &&
            w->cond_fc_b && w->latent_proj_w && w->latent_proj_b && w->final_time_w && w->final_time_b && w->final_class_w &&
            w->final_class_b && w->final_norm_w && w->final_norm_b && w->final_w && w->final_b, "ldm_unet3_pack: weights missing");
  cudaDeviceSynchronize();
  drop_graphs(ctx);
  free_pool(U.allocs);
  U = UnetModel();
  U.variant = 3;
  U.latent = w->latent_dim; U.tdim = w->time_dim; U.ncolors = w->num_colors; U.ncls = w->num_classes * w->num_colors;
  U.nst = nst; U.n_t = w->n_t;
  for (int i = 0; i <= nst; ++i) { U.hid[i] = w->hidden[i]; U.dmax = U.hid[i] > U.dmax ? U.hid[i] : U.dmax; }
  auto& P = U.allocs;
  const int td = U.tdim, M_hint = 256;
  // --- embeddings hoisted out of the loop: TE = time_emb(t) for every t, CE = multi_cond_emb(f, k) for every pair
  float *sinus, *te1, *te, *pairs, *ce;
  std::vector<void*> tmp;
  LDM_TRY(own_copy(ctx, tmp, w->sinusoid, (size_t)U.n_t * td, &sinus, st));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &te1, (size_t)U.n_t * 2 * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &te, (size_t)U.n_t * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &pairs, (size_t)U.ncls * 2 * td));
  LDM_TRY(ldm_alloc_t(ctx, tmp, &ce, (size_t)U.ncls * td));
  {
    Epilogue e; e.bias = w->time_lin1_b; e.act = LDM_ACT_SWISH; e.out_f32 = te1; e.ld_of = 2 * td;
    LDM_TRY(launch_gemm_f32(ctx, sinus, td, w->time_lin1_w, U.n_t, 2 * td, td, e, st));
    Epilogue e2; e2.bias = w->time_lin2_b; e2.out_f32 = te; e2.ld_of = td;
    LDM_TRY(launch_gemm_f32(ctx, te1, 2 * td, w->time_lin2_w, U.n_t, td, 2 * td, e2, st));
    LDM_TRY(launch_cond_pairs(ctx, w->flower_emb, w->color_emb, pairs, w->num_classes, w->num_colors, td, st));   // v3:746-748
    Epilogue e3; e3.bias = w->cond_fc_b; e3.out_f32 = ce; e3.ld_of = td;
    LDM_TRY(launch_gemm_f32(ctx, pairs, 2 * td, w->cond_fc_w, U.ncls, td, 2 * td, e3, st));                      // v3:749
  }
  for (int i = 0; i <= nst; ++i) {   // T_i[t] = time_projections[i](TE[t]); C_i[p] = cond_projections[i](CE[p])   (v3:818-822, 844-846)
    const int d = U.hid[i];
    const float* tw = i < nst ? w->time_proj_w[i] : w->final_time_w;
    const float* tb = i < nst ? w->time_proj_b[i] : w->final_time_b;
    const float* cw = i < nst ? w->cond_proj_w[i] : w->final_class_w;
    const float* cb = i < nst ? w->cond_proj_b[i] : w->final_class_b;
    LDM_CHECK(tw && tb && cw && cb, "ldm_unet3_pack: projection weights of stage %d missing", i);
    LDM_TRY(ldm_alloc_t(ctx, P, &U.tab_t[i], (size_t)U.n_t * d));
    LDM_TRY(ldm_alloc_t(ctx, P, &U.tab_c[i], (size_t)U.ncls * d));
    Epilogue e; e.bias = tb; e.out_f32 = U.tab_t[i]; e.ld_of = d;
    LDM_TRY(launch_gemm_f32(ctx, te, td, tw, U.n_t, d, td, e, st));
    Epilogue e2; e2.bias = cb; e2.out_f32 = U.tab_c[i]; e2.ld_of = d;
    LDM_TRY(launch_gemm_f32(ctx, ce, td, cw, U.ncls, d, td, e2, st));
  }
  LDM_TRY(dense_from(ctx, P, U.latent_proj, w->latent_proj_w, w->latent_proj_b, U.hid[0], U.latent, st));
  LDM_TRY(finish_dense(ctx, P, U.latent_proj, M_hint, st));
  for (int i = 0; i < nst; ++i) {
    const int d = U.hid[i], dn = U.hid[i + 1];
    LDM_CHECK(w->block_lin_w[i] && w->attn_in_proj_w[i] && w->attn_in_proj_b[i] && w->attn_out_w[i] && w->attn_out_b[i] &&
              w->down_w[i] && w->block_ln_w[i] && w->stage_ln_w[i], "ldm_unet3_pack: weights of stage %d missing", i);
    LDM_CHECK(d % 8 == 0 && d / 8 <= 128, "ldm_unet3_pack: head_dim %d unsupported", d / 8);
    LDM_TRY(dense_from(ctx, P, U.block[i], w->block_lin_w[i], w->block_lin_b[i], d, d, st));
    LDM_TRY(finish_dense(ctx, P, U.block[i], M_hint, st));
    LDM_TRY(dense_from(ctx, P, U.qkv[i], w->attn_in_proj_w[i], w->attn_in_proj_b[i], 3 * d, d, st));
    LDM_TRY(finish_dense(ctx, P, U.qkv[i], M_hint, st));
    LDM_TRY(dense_from(ctx, P, U.attn_o[i], w->attn_out_w[i], w->attn_out_b[i], d, d, st));
    LDM_TRY(finish_dense(ctx, P, U.attn_o[i], M_hint, st));
    LDM_TRY(dense_from(ctx, P, U.down[i], w->down_w[i], w->down_b[i], dn, d, st));
    LDM_TRY(finish_dense(ctx, P, U.down[i], M_hint, st));
    LDM_TRY(own_copy(ctx, P, w->block_ln_w[i], d, &U.ln_a_w[i], st));
    LDM_TRY(own_copy(ctx, P, w->block_ln_b[i], d, &U.ln_a_b[i], st));
    LDM_TRY(own_copy(ctx, P, w->stage_ln_w[i], d, &U.ln_b_w[i], st));
    LDM_TRY(own_copy(ctx, P, w->stage_ln_b[i], d, &U.ln_b_b[i], st));
  }
  LDM_TRY(own_copy(ctx, P, w->final_norm_w, U.hid[nst], &U.ln_f_w, st));
  LDM_TRY(own_copy(ctx, P, w->final_norm_b, U.hid[nst], &U.ln_f_b, st));
  {  // final(h) alone (`return out`, v3:853): the [W_f | s W_f] layout of the v2 path with s = sigmoid(-1e30) = 0
    const int N = U.latent, K = U.hid[nst];
    U.fin.N = N; U.fin.K = 2 * K;
    float *s_dev, *rw_dev;
    LDM_TRY(ldm_alloc_t(ctx, P, &U.fin.w32, (size_t)N * 2 * K));
    LDM_TRY(ldm_alloc_t(ctx, P, &U.fin.b, (size_t)N));
    LDM_TRY(ldm_alloc_t(ctx, tmp, &s_dev, 1));
    LDM_TRY(ldm_alloc_t(ctx, tmp, &rw_dev, 1));
    const float minus_big = -1e30f;
    LDM_CUDA(cudaMemcpyAsync(rw_dev, &minus_big, sizeof(float), cudaMemcpyHostToDevice, st));
    LDM_TRY(launch_pack_final(ctx, w->final_w, w->final_b, rw_dev, U.fin.w32, U.fin.b, s_dev, N, K, st));
    LDM_TRY(finish_dense(ctx, P, U.fin, M_hint, st));
    U.s_res = 0.f;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  free_pool(tmp);
  LDM_CUDA(e);
  ctx->use_chain = 0;   // attention couples the rows of a call: the per-layer sequence runs v3
  cudaDeviceSynchronize();
  ctx->act_maps.clear();
  free_pool(ctx->ws_allocs);
  ctx->cap = 0;
  ctx->batch_cls = -1;
  ctx->has_cls = false;
  U.packed = true;
  LDM_TRY(v3loop_pack(ctx, st));   // the whole loop as one persistent kernel where the architecture allows (v3loop.cu)
  return 0;
}

extern "C" LDM_API int ldm_unet3_set_conditions(ldm_ctx* ctx, const int64_t* flower_dev, const int64_t* color_dev, int batch,
                                        void* stream) {
  LDM_CHECK(ctx && flower_dev && color_dev && batch > 0, "ldm_unet3_set_conditions: bad arguments");
  LDM_CHECK(ctx->unet.packed && ctx->unet.variant == 3, "ldm_unet3_set_conditions: pack a v3 denoiser first (ldm_unet3_pack)");
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  ctx->batch_cls = batch;
  ctx->has_cls = true;
  const int nk = ctx->unet.ncolors;
  return launch_set_conditions(ctx, flower_dev, color_dev, ctx->cls, batch, ctx->unet.ncls / nk, nk, ctx->dev_flags, (cudaStream_t)stream);
}

extern "C" LDM_API int ldm_unet_set_classes(ldm_ctx* ctx, const int64_t* c_dev, int batch, void* stream) {
  LDM_CHECK(ctx && batch > 0, "ldm_unet_set_classes: bad arguments");
  LDM_CHECK(ctx->unet.variant != 3, "ldm_unet_set_classes: a v3 denoiser takes (flower, color): use ldm_unet3_set_conditions");
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  ctx->batch_cls = batch;
  ctx->has_cls = c_dev != nullptr;
  if (c_dev) LDM_TRY(launch_set_classes(ctx, c_dev, ctx->cls, batch, ctx->unet.ncls, ctx->dev_flags, (cudaStream_t)stream));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// one denoiser evaluation (+ optional fused posterior update)
// -------------------------------------------------------------------------------------------------
struct StepMode {
  // forward(): timesteps from the device, eps written to eps_out
  const int64_t* t_idx = nullptr;
  int t_len = 1;
  float* eps_out = nullptr;
  // p_sample(): compile-time-of-the-graph timestep, x updated in place
  int sample = 0;
  int t = 0;
  float* x = nullptr;
  const float* noise = nullptr;
};

template <typename TOP>
static int gemm(ldm_ctx* ctx, const TOP* A, int lda, int M, const DenseLayer& L, const Epilogue& e, cudaStream_t st) {
  if constexpr (std::is_same<TOP, float>::value) return launch_gemm_f32(ctx, A, lda, L.w32, M, L.N, L.K, e, st);
  else return launch_gemm_tc(ctx, A, lda, M, L, e, st);
}

// route the operand-typed copy of an epilogue's result
template <typename TOP>
static void set_op_out(Epilogue& e, TOP* p, int ld) {
  if constexpr (std::is_same<TOP, float>::value) {
    if (p != e.out_f32) {   // a distinct fp32 operand buffer (only one fp32 destination exists)
      e.out_f32 = p;
      e.ld_of = ld;
    }
  } else {
    e.out_bf16 = p;
    e.ld_ob = ld;
  }
}

template <typename TOP>
static int denoise(ldm_ctx* ctx, int B, int parity, const StepMode& md, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  const int nst = U.nst, dl = U.hid[nst], ldaf = dl + U.latent;
  TOP* af = (TOP*)ctx->af_op[parity];
  TOP* af_next = (TOP*)ctx->af_op[parity ^ 1];
  TOP *h_op = (TOP*)ctx->h_op, *n_op = (TOP*)ctx->n_op, *h3_op = (TOP*)ctx->h3_op;
  const int32_t* cls = ctx->has_cls ? ctx->cls : nullptr;
  auto tables = [&](Epilogue& e, int i) {
    e.tab_t = U.tab_t[i]; e.ld_t = U.hid[i]; e.n_t = U.n_t;
    e.t_idx = md.sample ? nullptr : md.t_idx; e.t_len = md.t_len; e.t_const = md.t;
    if (cls) { e.tab_c = U.tab_c[i]; e.ld_c = U.hid[i]; e.cls = cls; }
  };
  {  // h = latent_proj(x) + T_0[t] + C_0[c]                                   (v2:539,541-545)
    Epilogue e; e.bias = U.latent_proj.b; tables(e, 0);
    e.out_f32 = ctx->h; e.ld_of = U.hid[0];
    set_op_out<TOP>(e, h_op, U.hid[0]);
    LDM_TRY(gemm<TOP>(ctx, af + dl, ldaf, B, U.latent_proj, e, st));
  }
  for (int i = 0; i < nst; ++i) {
    const int d = U.hid[i], dn = U.hid[i + 1];
    {  // u = Linear_b(h)                                                      (v2:519)
      Epilogue e; e.bias = U.block[i].b; e.out_f32 = ctx->u; e.ld_of = d;
      LDM_TRY(gemm<TOP>(ctx, h_op, d, B, U.block[i], e, st));
    }
    // h2 = swish(LN_a(u)) + h ; n = LN_b(h2)                                  (v2:520-522,547-549)
    LDM_TRY(launch_stage_mid<TOP>(ctx, ctx->u, ctx->h, U.ln_a_w[i], U.ln_a_b[i], U.ln_b_w[i], U.ln_b_b[i], ctx->h2, n_op, d, B, d, st));
    if (U.variant == 3) {  // h3 = h2 + out_proj(softmax(Q K^T / sqrt(hd)) V) over the rows of the call   (v3:832-838)
      bool tc_attn = false;
      if constexpr (std::is_same<TOP, bf16>::value) tc_attn = attn_tc_supported(d / 8) && ctx->use_attn_tc;
      Epilogue eq; eq.bias = U.qkv[i].b;
      if (tc_attn) {   // the in_proj epilogue writes the attention operands directly: [Q | K] bf16 row-major, V transposed
        eq.out_bf16 = ctx->qk16; eq.ld_ob = 2 * d; eq.vt = ctx->vt16; eq.vt_col0 = 2 * d; eq.ld_vt = ctx->cap;
      } else {
        eq.out_f32 = ctx->qkv; eq.ld_of = 3 * d;
      }
      LDM_TRY(gemm<TOP>(ctx, n_op, d, B, U.qkv[i], eq, st));
      if (tc_attn) {   // softmax(Q K^T) V on the tensor cores (gemm_tc.cu: attn_tc_kernel)
        LDM_TRY(launch_attn_tc(ctx, ctx->qk16, 2 * d, 2 * d, ctx->vt16, ctx->cap, B, 1, 8, d / 8, 0, d, ctx->a_op, 1, d, d / 8, 1, st));
      } else {
        LDM_TRY(launch_batch_attention<TOP>(ctx, ctx->qkv, (TOP*)ctx->a_op, B, d, 8, st));
      }
      Epilogue e; e.bias = U.attn_o[i].b; e.resid = ctx->h2; e.ld_r = d;
      if constexpr (std::is_same<TOP, float>::value) { e.out_f32 = (float*)h3_op; e.ld_of = d; }
      else { e.out_bf16 = h3_op; e.ld_ob = d; }
      LDM_TRY(gemm<TOP>(ctx, (TOP*)ctx->a_op, d, B, U.attn_o[i], e, st));
    } else {  // h3 = h2 + out_proj(V(n))                                      (v2:550-552)
      Epilogue e; e.bias = U.ov[i].b; e.resid = ctx->h2; e.ld_r = d;
      if constexpr (std::is_same<TOP, float>::value) { e.out_f32 = (float*)h3_op; e.ld_of = d; }
      else { e.out_bf16 = h3_op; e.ld_ob = d; }
      LDM_TRY(gemm<TOP>(ctx, n_op, d, B, U.ov[i], e, st));
    }
    {  // h = down(h3) + T_{i+1}[t] + C_{i+1}[c]                               (v2:553 then v2:541-545 / 554-558)
      Epilogue e; e.bias = U.down[i].b; tables(e, i + 1);
      e.out_f32 = ctx->h; e.ld_of = dn;
      if (i + 1 < nst) set_op_out<TOP>(e, h_op, dn);
      LDM_TRY(gemm<TOP>(ctx, h3_op, d, B, U.down[i], e, st));
    }
  }
  // [LN_f(h) | x] -> final                                                    (v2:559-561)
  LDM_TRY(launch_row_ln<TOP>(ctx, ctx->h, dl, U.ln_f_w, U.ln_f_b, LDM_ACT_NONE, af, ldaf, B, dl, st));
  {
    Epilogue e; e.bias = U.fin.b;
    if (!md.sample) {
      e.out_f32 = md.eps_out; e.ld_of = U.latent;
    } else {
      e.ddpm = 1; e.x = md.x; e.step = md.t;
      e.c2 = ctx->c2[md.t]; e.sqrt_alpha = ctx->sqrt_alpha[md.t]; e.sigma = ctx->sigma[md.t];
      e.noise = md.noise; e.rng = ctx->rng_dev;
      // the new x also becomes the x-operand of the NEXT step (other buffer: this step's tiles still read af)
      if constexpr (std::is_same<TOP, float>::value) { e.out_f32 = af_next + dl; e.ld_of = ldaf; }
      else { e.out_bf16 = af_next + dl; e.ld_ob = ldaf; }
    }
    LDM_TRY(gemm<TOP>(ctx, af, ldaf, B, U.fin, e, st));
  }
  return 0;
}

template <typename TOP>
static int stage_x(ldm_ctx* ctx, const float* x, int B, int parity, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  const int dl = U.hid[U.nst];
  return launch_load_x<TOP>(ctx, x, (TOP*)ctx->af_op[parity] + dl, dl + U.latent, B, U.latent, st);
}

static int check_classes_match(ldm_ctx* ctx, int batch) {
  LDM_CHECK(ctx->batch_cls == batch || (ctx->batch_cls == -1 && !ctx->has_cls),
            "class labels were set for batch %d but this call has batch %d (call ldm_unet_set_classes)", ctx->batch_cls, batch);
  return 0;
}

extern "C" LDM_API int ldm_unet_forward(ldm_ctx* ctx, const float* x_dev, const int64_t* t_dev, int t_len, float* eps_out_dev,
                                int batch, void* stream) {
  LDM_CHECK(ctx && x_dev && t_dev && eps_out_dev && batch > 0, "ldm_unet_forward: bad arguments");
  LDM_CHECK(t_len == 1 || t_len == batch, "ldm_unet_forward: t must have 1 or batch (%d) entries, got %d", batch, t_len);
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  LDM_TRY(check_classes_match(ctx, batch));
  LDM_TRY(launch_check_t(ctx, t_dev, t_len, ctx->unet.n_t, ctx->dev_flags, st));
  StepMode md; md.t_idx = t_dev; md.t_len = t_len; md.eps_out = eps_out_dev;
  if (v3loop_supported(ctx, batch)) {
    const int r = launch_v3loop(ctx, batch, 1, 0, 0, t_dev, t_len, const_cast<float*>(x_dev), eps_out_dev, nullptr, st);
    if (r != LDM_V3LOOP_UNAVAILABLE) return r;
  }
  if (ctx->precision == LDM_PRECISION_BF16) {
    LDM_TRY(stage_x<bf16>(ctx, x_dev, batch, 0, st));
    if (ctx->use_chain) return launch_chain(ctx, batch, 1, 0, 0, t_dev, t_len, const_cast<float*>(x_dev), eps_out_dev, nullptr, st);
    return denoise<bf16>(ctx, batch, 0, md, st);
  }
  if (ctx->use_chain) return launch_chain(ctx, batch, 1, 0, 0, t_dev, t_len, const_cast<float*>(x_dev), eps_out_dev, nullptr, st);
  LDM_TRY(stage_x<float>(ctx, x_dev, batch, 0, st));
  return denoise<float>(ctx, batch, 0, md, st);
}

extern "C" LDM_API int ldm_ddpm_step(ldm_ctx* ctx, float* x, const float* eps, int t, const float* noise, uint64_t seed,
                             uint64_t sample_offset, int batch, int dim, void* stream) {
  LDM_CHECK(ctx && x && eps && batch > 0 && dim > 0 && dim % 4 == 0, "ldm_ddpm_step: bad arguments (batch %d, dim %d: a positive multiple of 4)", batch, dim);
  LDM_CHECK(ctx->n_steps > 0, "ldm_ddpm_step: schedule not set");
  LDM_CHECK(t >= 0 && t < ctx->n_steps, "ldm_ddpm_step: t = %d outside [0, %d)", t, ctx->n_steps);
  LDM_CUDA(cudaSetDevice(ctx->device));
  return launch_ddpm_update(ctx, x, eps, ctx->c2[t], ctx->sqrt_alpha[t], ctx->sigma[t], noise, seed, sample_offset, t,
                            batch, dim, (cudaStream_t)stream);
}

extern "C" LDM_API int ldm_randn(ldm_ctx* ctx, float* out, uint64_t seed, uint64_t sample_offset, int step, int batch, int dim,
                         void* stream) {
  LDM_CHECK(ctx && out && batch > 0 && dim > 0, "ldm_randn: bad arguments");
  LDM_CUDA(cudaSetDevice(ctx->device));
  return launch_randn(ctx, out, seed, sample_offset, step, batch, dim, (cudaStream_t)stream);
}

// -------------------------------------------------------------------------------------------------
// the sampling loop
// -------------------------------------------------------------------------------------------------
static int run_chain(ldm_ctx* ctx, int B, int t_start, int t_end, const float* noise, cudaStream_t st) {
  // x lives in ctx->x_state; its operand copy must already be staged in af_op[0]
  const size_t slab = (size_t)B * ctx->unet.latent;
  if (ctx->use_chain)   // the whole loop is one persistent kernel (both precisions)
    return launch_chain(ctx, B, t_start - t_end + 1, t_start, 1, nullptr, 1, ctx->x_state, nullptr, noise, st);
  if (v3loop_supported(ctx, B)) {   // v3: one persistent kernel for the whole loop (the rows of a call are coupled: one grid)
    const int r = launch_v3loop(ctx, B, t_start - t_end + 1, t_start, 1, nullptr, 1, ctx->x_state, nullptr, noise, st);
    if (r != LDM_V3LOOP_UNAVAILABLE) return r;
  }
  for (int t = t_start, j = 0; t >= t_end; --t, ++j) {
    StepMode md; md.sample = 1; md.t = t; md.x = ctx->x_state;
    md.noise = noise ? noise + (size_t)j * slab : nullptr;
    if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(denoise<bf16>(ctx, B, j & 1, md, st));
    else LDM_TRY(denoise<float>(ctx, B, j & 1, md, st));
  }
  return 0;
}

extern "C" LDM_API int ldm_sample(ldm_ctx* ctx, float* x_inout, int t_start, int t_end, const float* noise, uint64_t seed,
                          uint64_t sample_offset, int batch, int use_graph, void* stream) {
  LDM_CHECK(ctx && x_inout && batch > 0, "ldm_sample: bad arguments");
  LDM_CHECK(ctx->n_steps > 0, "ldm_sample: schedule not set (ldm_set_schedule)");
  LDM_CHECK(ctx->unet.packed, "ldm_sample: denoiser not packed (ldm_unet_pack)");
  LDM_CHECK(t_end >= 0 && t_start >= t_end && t_start < ctx->n_steps, "ldm_sample: need 0 <= t_end <= t_start < n_steps (%d), got %d..%d",
            ctx->n_steps, t_start, t_end);
  LDM_CHECK(ctx->n_steps <= ctx->unet.n_t, "ldm_sample: time table has %d rows but the schedule has %d steps", ctx->unet.n_t, ctx->n_steps);
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  LDM_TRY(check_classes_match(ctx, batch));
  const size_t xbytes = (size_t)batch * ctx->unet.latent * sizeof(float);
  if (x_inout != ctx->x_state) LDM_CUDA(cudaMemcpyAsync(ctx->x_state, x_inout, xbytes, cudaMemcpyDeviceToDevice, st));
  LDM_TRY(launch_set_rng(ctx, ctx->rng_dev, seed, sample_offset, st));
  if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(stage_x<bf16>(ctx, ctx->x_state, batch, 0, st));
  else LDM_TRY(stage_x<float>(ctx, ctx->x_state, batch, 0, st));

  if (!use_graph) {
    LDM_TRY(run_chain(ctx, batch, t_start, t_end, noise, st));
  } else {
    GraphKey key{batch, t_start, t_end, noise ? 1 : 0};
    auto it = ctx->graphs.find(key);
    if (it != ctx->graphs.end() && (it->second.noise != noise || it->second.has_cls != ctx->has_cls)) {
      cudaGraphExecDestroy(it->second.exec);
      cudaGraphDestroy(it->second.graph);
      ctx->graphs.erase(it);
      it = ctx->graphs.end();
    }
    if (it == ctx->graphs.end()) {
      // capture on the context's own stream (the caller may be on the legacy default stream)
      GraphEntry ge;
      const unsigned long long before = ctx->launches;
      LDM_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
      ctx->capturing = true;
      int r = run_chain(ctx, batch, t_start, t_end, noise, ctx->cap_stream);
      ctx->capturing = false;
      cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &ge.graph);
      if (r != 0) { if (ge.graph) cudaGraphDestroy(ge.graph); return r; }
      LDM_CUDA(ce);
      ge.n_nodes = (size_t)(ctx->launches - before);
      ctx->launches = before;   // captured, not yet executed
      LDM_CUDA(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
      ge.noise = noise;
      ge.has_cls = ctx->has_cls;
      it = ctx->graphs.emplace(key, ge).first;
    }
    LDM_CUDA(cudaGraphLaunch(it->second.exec, st));
    ctx->launches += it->second.n_nodes;
  }
  if (x_inout != ctx->x_state) LDM_CUDA(cudaMemcpyAsync(x_inout, ctx->x_state, xbytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// decoder
// -------------------------------------------------------------------------------------------------
extern "C" LDM_API int ldm_decoder_pack(ldm_ctx* ctx, const ldm_decoder_weights* w, void* stream) {
  LDM_CHECK(ctx && w, "ldm_decoder_pack: null argument");
  LDM_CUDA(cudaSetDevice(ctx->device));
  return decoder_pack_impl(ctx, w, (cudaStream_t)stream);
}

extern "C" LDM_API int ldm_decode(ldm_ctx* ctx, const float* z_dev, float* img_out_dev, int batch, void* stream) {
  LDM_CHECK(ctx && z_dev && img_out_dev && batch > 0, "ldm_decode: bad arguments");
  LDM_CHECK(ctx->dec.packed, "ldm_decode: decoder not packed (ldm_decoder_pack)");
  LDM_CUDA(cudaSetDevice(ctx->device));
  return decoder_run_impl(ctx, z_dev, img_out_dev, batch, (cudaStream_t)stream);
}

// -------------------------------------------------------------------------------------------------
// host-buffer entry point (v2:865-869)
// -------------------------------------------------------------------------------------------------
extern "C" LDM_API int ldm_generate_host(ldm_ctx* ctx, const int64_t* c_host, int batch, uint64_t seed, uint64_t sample_offset,
                                 float* img_out_host, float* latents_out_host, void* stream) {
  LDM_CHECK(ctx && img_out_host && batch > 0, "ldm_generate_host: bad arguments");
  LDM_CHECK(ctx->n_steps > 0 && ctx->unet.packed && ctx->dec.packed, "ldm_generate_host: schedule, denoiser and decoder must be set first");
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  if (batch > ctx->host_cap) {
    cudaDeviceSynchronize();
    free_pool(ctx->stage_allocs);
    LDM_TRY(ldm_alloc_t(ctx, ctx->stage_allocs, &ctx->c_stage, (size_t)batch));
    LDM_TRY(ldm_alloc_t(ctx, ctx->stage_allocs, &ctx->img_stage, (size_t)batch * 3 * 64 * 64));
    ctx->host_cap = batch;
  }
  if (c_host) LDM_CUDA(cudaMemcpyAsync(ctx->c_stage, c_host, (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  LDM_TRY(ldm_unet_set_classes(ctx, c_host ? ctx->c_stage : nullptr, batch, st));
  LDM_TRY(launch_randn(ctx, ctx->x_state, seed, sample_offset, ctx->n_steps, batch, ctx->unet.latent, st));   // v2:595
  LDM_TRY(ldm_sample(ctx, ctx->x_state, ctx->n_steps - 1, 0, nullptr, seed, sample_offset, batch, 1, st));
  LDM_TRY(decoder_run_impl(ctx, ctx->x_state, ctx->img_stage, batch, st));
  LDM_CUDA(cudaMemcpyAsync(img_out_host, ctx->img_stage, (size_t)batch * 3 * 64 * 64 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (latents_out_host)
    LDM_CUDA(cudaMemcpyAsync(latents_out_host, ctx->x_state, (size_t)batch * ctx->unet.latent * sizeof(float), cudaMemcpyDeviceToHost, st));
  LDM_CUDA(cudaStreamSynchronize(st));
  int flags = 0;
  LDM_CUDA(cudaMemcpy(&flags, ctx->dev_flags, sizeof(int), cudaMemcpyDeviceToHost));
  if (flags & 1) {
    cudaMemset(ctx->dev_flags, 0, sizeof(int));
    ldm_set_error("ldm_generate_host: class label out of range [0, %d)", ctx->unet.ncls);
    return -2;
  }
  return 0;
}

// v3 (v3:860-893 + decode): host (flower, color) labels in, host images out, one call.  The 2 x batch labels are staged in
// c_stage; the chain runs as one call of `batch` rows (the rows of a v3 call are coupled through the attention).
extern "C" LDM_API int ldm_generate3_host(ldm_ctx* ctx, const int64_t* flower_host, const int64_t* color_host, int batch, uint64_t seed,
                                          uint64_t sample_offset, float* img_out_host, float* latents_out_host, void* stream) {
  LDM_CHECK(ctx && flower_host && color_host && img_out_host && batch > 0, "ldm_generate3_host: bad arguments");
  LDM_CHECK(ctx->n_steps > 0 && ctx->unet.packed && ctx->unet.variant == 3 && ctx->dec.packed,
            "ldm_generate3_host: schedule, v3 denoiser and decoder must be set first");
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_workspace(ctx, batch));
  if (2 * batch > ctx->host_cap) {
    cudaDeviceSynchronize();
    free_pool(ctx->stage_allocs);
    LDM_TRY(ldm_alloc_t(ctx, ctx->stage_allocs, &ctx->c_stage, (size_t)2 * batch));
    LDM_TRY(ldm_alloc_t(ctx, ctx->stage_allocs, &ctx->img_stage, (size_t)2 * batch * 3 * 64 * 64));
    ctx->host_cap = 2 * batch;
  }
  LDM_CUDA(cudaMemcpyAsync(ctx->c_stage, flower_host, (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  LDM_CUDA(cudaMemcpyAsync(ctx->c_stage + batch, color_host, (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  LDM_TRY(ldm_unet3_set_conditions(ctx, ctx->c_stage, ctx->c_stage + batch, batch, st));
  LDM_TRY(launch_randn(ctx, ctx->x_state, seed, sample_offset, ctx->n_steps, batch, ctx->unet.latent, st));   // v3:889
  LDM_TRY(ldm_sample(ctx, ctx->x_state, ctx->n_steps - 1, 0, nullptr, seed, sample_offset, batch, 1, st));
  LDM_TRY(decoder_run_impl(ctx, ctx->x_state, ctx->img_stage, batch, st));
  LDM_CUDA(cudaMemcpyAsync(img_out_host, ctx->img_stage, (size_t)batch * 3 * 64 * 64 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (latents_out_host)
    LDM_CUDA(cudaMemcpyAsync(latents_out_host, ctx->x_state, (size_t)batch * ctx->unet.latent * sizeof(float), cudaMemcpyDeviceToHost, st));
  LDM_CUDA(cudaStreamSynchronize(st));
  int flags = 0;
  LDM_CUDA(cudaMemcpy(&flags, ctx->dev_flags, sizeof(int), cudaMemcpyDeviceToHost));
  if (flags & 1) {
    cudaMemset(ctx->dev_flags, 0, sizeof(int));
    ldm_set_error("ldm_generate3_host: flower / colour label out of range");
    return -2;
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------
// introspection
// -------------------------------------------------------------------------------------------------
extern "C" LDM_API int ldm_kernel_launch_count(ldm_ctx* ctx, uint64_t* out) {
  LDM_CHECK(ctx && out, "ldm_kernel_launch_count: bad arguments");
  *out = ctx->launches;
  return 0;
}

extern "C" LDM_API int ldm_debug_chain_trace(ldm_ctx* ctx, int step, long long* out_host, int n) {
  LDM_CHECK(ctx, "ldm_debug_chain_trace: null context");
  LDM_CUDA(cudaSetDevice(ctx->device));
  const int total = LDM_CHAIN_CLUSTER * LDM_CHAIN_TRACE_TRACKS * LDM_CHAIN_TRACE_LEN;
  if (out_host) {   // read back (and keep tracing)
    LDM_CHECK(ctx->chain_trace && n >= total, "ldm_debug_chain_trace: tracing is off or the buffer is shorter than %d", total);
    LDM_CUDA(cudaDeviceSynchronize());
    LDM_CUDA(cudaMemcpy(out_host, ctx->chain_trace, sizeof(long long) * total, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (step < 0) {   // off
    ctx->chain_trace_step = 0;
    if (ctx->chain_trace) { LDM_CUDA(cudaDeviceSynchronize()); cudaFree(ctx->chain_trace); ctx->chain_trace = nullptr; }
    drop_graphs(ctx);
    return 0;
  }
  if (!ctx->chain_trace) LDM_CUDA(cudaMalloc((void**)&ctx->chain_trace, sizeof(long long) * total));
  LDM_CUDA(cudaMemset(ctx->chain_trace, 0, sizeof(long long) * total));
  ctx->chain_trace_step = step;
  drop_graphs(ctx);
  return 0;
}

extern "C" LDM_API int ldm_debug_ktrace(ldm_ctx* ctx, int start, void* stream, char* names_out, int names_cap, float* ms_out,
                                        int n_cap, int* n_out) {
  LDM_CHECK(ctx, "ldm_debug_ktrace: null context");
  LDM_CUDA(cudaSetDevice(ctx->device));
  auto drop = [&]() {
    for (auto& m : ctx->kt_marks) cudaEventDestroy(m.second);
    ctx->kt_marks.clear();
  };
  if (start) {
    drop();
    ctx->kt_stream = (cudaStream_t)stream;
    ctx->kt_on = true;
    ldm_kmark(ctx, "start");
    return 0;
  }
  ctx->kt_on = false;
  LDM_CHECK(names_out && ms_out && n_out && names_cap > 0 && n_cap > 0, "ldm_debug_ktrace: output buffers missing");
  LDM_CHECK(!ctx->kt_marks.empty(), "ldm_debug_ktrace: no trace was started");
  LDM_CUDA(cudaEventSynchronize(ctx->kt_marks.back().second));
  int n = 0;
  size_t pos = 0;
  for (size_t i = 1; i < ctx->kt_marks.size() && n < n_cap; ++i) {
    const std::string& nm = ctx->kt_marks[i].first;
    if (pos + nm.size() + 2 > (size_t)names_cap) break;
    float ms = 0.f;
    LDM_CUDA(cudaEventElapsedTime(&ms, ctx->kt_marks[i - 1].second, ctx->kt_marks[i].second));
    memcpy(names_out + pos, nm.data(), nm.size());
    pos += nm.size();
    names_out[pos++] = '\n';
    ms_out[n++] = ms;
  }
  names_out[pos] = 0;
  *n_out = n;
  drop();
  return 0;
}

extern "C" LDM_API int ldm_get_info(ldm_ctx* ctx, const char* key, double* out) {
  LDM_CHECK(ctx && key && out, "ldm_get_info: bad arguments");
  if (!strcmp(key, "precision")) { *out = ctx->precision; return 0; }
  if (!strcmp(key, "sm_count")) { *out = ctx->sm_count; return 0; }
  if (!strcmp(key, "n_steps")) { *out = ctx->n_steps; return 0; }
  if (!strcmp(key, "graphs")) { *out = (double)ctx->graphs.size(); return 0; }
  if (!strcmp(key, "residual_gate")) { *out = ctx->unet.s_res; return 0; }
  if (!strcmp(key, "chain")) { *out = ctx->use_chain; return 0; }
  if (!strcmp(key, "chain_max_clusters")) { *out = ctx->chain_max_clusters; return 0; }
  if (!strcmp(key, "chain_peak_bytes_per_step")) { *out = ctx->chain.peak_bytes_per_step; return 0; }
  if (!strcmp(key, "launches_per_step")) {
    if (ctx->use_chain || (ctx->v3loop && ctx->use_v3loop)) { *out = 0.0; return 0; }   // the loop is one launch
    *out = ctx->unet.packed ? 3.0 + (ctx->unet.variant == 3 ? 6.0 : 4.0) * ctx->unet.nst : 0.0;   // G0, (G1, R1, G2[, A, G2b], G3) x stages, R_f, G_f
    return 0;
  }
  if (!strcmp(key, "device_flags")) {   // bit 0: class label out of range, bit 1: timestep out of range (clears the flags)
    int f = 0;
    LDM_CUDA(cudaMemcpy(&f, ctx->dev_flags, sizeof(int), cudaMemcpyDeviceToHost));
    LDM_CUDA(cudaMemset(ctx->dev_flags, 0, sizeof(int)));
    *out = f;
    return 0;
  }
  if (!strcmp(key, "tc_error")) {       // barrier timeout record of the tensor-core kernels (0 = none)
    int f = 0;
    if (ctx->precision == LDM_PRECISION_BF16) { LDM_TRY(tc_error_flag(&f)); if (f) tc_error_reset(); }
    if (!f && ctx->precision == LDM_PRECISION_BF16) LDM_TRY(conv_tc_error_flag(&f));
    if (!f) {
      int ce[2] = {0, 0};
      LDM_CUDA(cudaMemcpy(ce, ctx->chain_err, sizeof(ce), cudaMemcpyDeviceToHost));
      if (ce[0]) { f = (ce[0] << 16) | (ce[1] & 0xFFFF); LDM_CUDA(cudaMemset(ctx->chain_err, 0, sizeof(ce))); }
    }
    if (!f) LDM_TRY(v3loop_error(ctx, &f));
    *out = f;
    return 0;
  }
  ldm_set_error("ldm_get_info: unknown key '%s'", key);
  return -1;
}
