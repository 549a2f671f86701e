// VAE decoder pass (Decoder.forward v2:280-290) over NHWC activations: weight repacking and the kernel
// sequence.  Convolutions are implicit GEMMs (rows = pixels, K = taps x Cin); ConvTranspose2d(4,2,1) is
// four sub-pixel 2x2-tap convolutions; LayerNorm2d / GroupNorm / CALayer / SpatialAttention are the
// memory-bound kernels of decoder_norm.cu.
#include "common.cuh"
#include "pix_out.cuh"

int tc_pick_bn(int M, int N);
int conv_tc_pick_bn(int Cout);
int launch_conv_tc(ldm_ctx* ctx, const bf16* in, const ConvLayer& L, const float* bias, bf16* out, int B, int H, int W,
                   int up, cudaStream_t st);
int launch_norm_coef_bf16_ws(ldm_ctx* ctx, const bf16* x, const float* gamma, const float* beta, float2* coef, int B, int HW,
                             int C, int group, float2* part, int* cnt, cudaStream_t st);
int launch_coef_apply_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, bf16* out, int B, int HW, int C, int act,
                           cudaStream_t st);
int launch_sa_map_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, float* map, int B, int HW, int C,
                       cudaStream_t st);
int launch_sa_apply_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, const float* map,
                         const float* sa_w, const bf16* resid, bf16* out, float* gate, int B, int H, int C, cudaStream_t st);
int launch_conv_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                     int B, int H, int W, int relu, const float* post, int post_stride, const PixOutArgs* fin, int ddpm,
                     cudaStream_t st);
int launch_conv_tc_f32out(ldm_ctx* ctx, const bf16* in, const ConvLayer& L, const float* bias, float* out32, int B, int H, int W,
                          int up, cudaStream_t st);
int launch_split3(ldm_ctx* ctx, const float* x, bf16* out, size_t npix, int C, cudaStream_t st);
int launch_pack_conv_split(ldm_ctx* ctx, const float* w, bf16* out, int rows, int taps, int Cin, cudaStream_t st);
int launch_conv_out3(ldm_ctx* ctx, const bf16* in, const float* w, const float* bias, float* out, int B, int H, int W,
                     cudaStream_t st);
int launch_final_gn_conv3(ldm_ctx* ctx, const bf16* x, const float2* coef, const uint32_t* wf, const float* bias, float* out, int B,
                          int H, int W, cudaStream_t st);
int launch_final_w_frag(ldm_ctx* ctx, const float* w, uint32_t* wf, cudaStream_t st);
int launch_sa_map_gate_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, const float* sa_w, float* gate,
                            int B, int H, int C, cudaStream_t st, int* done);
int convt_halo_supported(int H, int W, int Cin, int Cout);
int launch_convt_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int relu, cudaStream_t st);

namespace {

constexpr int kDecChunk = 256;   // samples per pass: bounds the activation workspace (3 x 256 MiB fp32)

int own(ldm_ctx* ctx, std::vector<void*>& pool, const float* src, size_t n, float** out, cudaStream_t st) {
  LDM_CHECK(src != nullptr, "ldm_decoder_pack: null weight pointer");
  LDM_TRY(ldm_alloc_t(ctx, pool, out, n));
  LDM_CUDA(cudaMemcpyAsync(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pack_conv3(ldm_ctx* ctx, std::vector<void*>& pool, ConvLayer& L, const float* w, const float* b, int Cout, int Cin,
               cudaStream_t st) {
  LDM_CHECK(w && b, "ldm_decoder_pack: conv weights missing");
  L.Cin = Cin; L.Cout = Cout; L.taps = 9;
  LDM_TRY(ldm_alloc_t(ctx, pool, &L.w32, (size_t)Cout * 9 * Cin));
  LDM_TRY(launch_pack_conv(ctx, w, L.w32, Cout, Cin, 3, 3, st));
  LDM_TRY(own(ctx, pool, b, Cout, &L.b, st));
  return 0;
}

void conv3_geom(ConvGeom& g, int B, int H, int Cin, int Cout) {
  g = ConvGeom();
  g.B = B; g.H = H; g.W = H; g.Cin = Cin; g.Cout = Cout; g.taps = 9; g.up = 1;
  for (int t = 0; t < 9; ++t) { g.dy[t] = t / 3 - 1; g.dx[t] = t % 3 - 1; }
}

// sub-pixel (pa, pb) of ConvTranspose2d(4, 2, 1): tap j in {0,1} per axis reads input offset 0 / (-1 or +1)
void convT_geom(ConvGeom& g, int B, int H, int Cin, int Cout, int pa, int pb) {
  g = ConvGeom();
  g.B = B; g.H = H; g.W = H; g.Cin = Cin; g.Cout = Cout; g.taps = 4; g.up = 2; g.pa = pa; g.pb = pb;
  for (int t = 0; t < 4; ++t) {
    const int jy = t >> 1, jx = t & 1;
    g.dy[t] = jy == 0 ? 0 : (pa == 0 ? -1 : 1);
    g.dx[t] = jx == 0 ? 0 : (pb == 0 ? -1 : 1);
  }
}

int ensure_dec_workspace(ldm_ctx* ctx, int B) {
  if (B <= ctx->dec_cap && (!ctx->dec.strict_tc || ctx->d_split)) return 0;
  if (B < ctx->dec_cap) B = ctx->dec_cap;
  cudaDeviceSynchronize();
  for (void* p : ctx->dec_allocs) cudaFree(p);
  ctx->dec_allocs.clear();
  auto& P = ctx->dec_allocs;
  const size_t act = (size_t)B * 64 * 64 * 64;   // largest activation: up1 output (B, 64, 64, 64)
  LDM_TRY(ldm_alloc(ctx, P, &ctx->d_a, act * sizeof(float)));
  LDM_TRY(ldm_alloc(ctx, P, &ctx->d_b, act * sizeof(float)));
  LDM_TRY(ldm_alloc(ctx, P, &ctx->d_c, act * sizeof(float)));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_f0, (size_t)B * 512));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_f1, (size_t)B * 512));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_stats, (size_t)B * 512 * 2));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_gap, (size_t)B * 512));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_ca, (size_t)B * 512));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_map, (size_t)B * 1024 * 2));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_gate, (size_t)B * 1024));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_part, (size_t)B * 1024));          // norm_coef pixel-split partial sums
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_cnt, (size_t)B * 16));             // ... and their tickets (left zero by every launch)
  LDM_CUDA(cudaMemset(ctx->d_cnt, 0, (size_t)B * 16 * sizeof(int)));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_zb, (size_t)B * 256 + 64));
  LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_h1b, (size_t)B * 512));
  if (ctx->dec.strict_tc) LDM_TRY(ldm_alloc_t(ctx, P, &ctx->d_split, act * 3));   // (hi, lo, hi) thirds of the largest convolution input
  ctx->act_maps.clear();   // descriptors over the old workspace are stale
  ctx->dec_cap = B;
  return 0;
}

// ResidualBlock.forward (v2:170-178): x in `X`, result in `OUT`, scratch `Y` (raw conv outputs)
// strict-mode convolution: fp32 NHWC in -> (hi, lo, hi) bf16 thirds -> three-term product on the tensor cores -> fp32 NHWC out
int conv_strict(ldm_ctx* ctx, const float* X, const ConvLayer& Ls, const float* bias, float* Y, int B, int H, int up, cudaStream_t st) {
  LDM_TRY(launch_split3(ctx, X, ctx->d_split, (size_t)B * H * H, Ls.Cin / 3, st));
  return launch_conv_tc_f32out(ctx, ctx->d_split, Ls, bias, Y, B, H, H, up, st);
}

int res_block_f32(ldm_ctx* ctx, const ResBlockModel& R, int B, const float* X, float* Y, float* OUT, cudaStream_t st) {
  const int C = R.C, H = R.HW, P = H * H;
  const bool tc = ctx->dec.strict_tc;
  ConvGeom g;
  conv3_geom(g, B, H, C, C);
  if (tc) LDM_TRY(conv_strict(ctx, X, R.conv1s, R.conv1.b, Y, B, H, 1, st));
  else LDM_TRY(launch_conv_f32(ctx, X, R.conv1.w32, R.conv1.b, Y, g, st));                              // conv1
  LDM_TRY(launch_inorm_stats<float>(ctx, Y, ctx->d_stats, B, P, C, 1, st));                              // ln1 statistics
  LDM_TRY(launch_norm_apply<float>(ctx, Y, ctx->d_stats, R.ln1_w, R.ln1_b, OUT, B, P, C, 1, LDM_ACT_SWISH, st));  // swish(ln1(.))
  if (tc) LDM_TRY(conv_strict(ctx, OUT, R.conv2s, R.conv2.b, Y, B, H, 1, st));
  else LDM_TRY(launch_conv_f32(ctx, OUT, R.conv2.w32, R.conv2.b, Y, g, st));                             // conv2
  LDM_TRY(launch_inorm_stats<float>(ctx, Y, ctx->d_stats, B, P, C, 1, st));                              // ln2 statistics
  LDM_TRY(launch_gap_norm<float>(ctx, Y, ctx->d_stats, R.ln2_w, R.ln2_b, ctx->d_gap, B, P, C, st));      // CALayer avg_pool
  LDM_TRY(launch_ca_mlp(ctx, ctx->d_gap, R.ca_w0, R.ca_w2, ctx->d_ca, B, C, st));                        // CALayer conv_du
  LDM_TRY(launch_sa_map<float>(ctx, Y, ctx->d_stats, R.ln2_w, R.ln2_b, ctx->d_ca, C, ctx->d_map, B, P, C, st));
  LDM_TRY(launch_sa_apply<float>(ctx, Y, ctx->d_stats, R.ln2_w, R.ln2_b, ctx->d_ca, C, ctx->d_map, R.sa_w, X, OUT, B, H, C, st));
  return 0;
}

// up block (v2:255-271): ConvTranspose2d -> GroupNorm (8 channels per group) -> Swish
int up_block_f32(ldm_ctx* ctx, const DecoderModel& D, int idx, int B, int H, int Cin, const float* X, float* Y, float* OUT,
                 cudaStream_t st) {
  const int Cout = Cin / 2;
  if (D.strict_tc) {
    LDM_TRY(conv_strict(ctx, X, D.ups[idx], D.up_b[idx], Y, B, H, 2, st));   // four sub-pixel parities, one launch
  } else {
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        ConvGeom g;
        convT_geom(g, B, H, Cin, Cout, pa, pb);
        LDM_TRY(launch_conv_f32(ctx, X, D.up[idx][pa * 2 + pb].w32, D.up_b[idx], Y, g, st));
      }
  }
  const int P = 4 * H * H;
  LDM_TRY(launch_inorm_stats<float>(ctx, Y, ctx->d_stats, B, P, Cout, 8, st));
  LDM_TRY(launch_norm_apply<float>(ctx, Y, ctx->d_stats, D.up_gn_w[idx], D.up_gn_b[idx], OUT, B, P, Cout, 8, LDM_ACT_SWISH, st));
  return 0;
}

int decode_chunk_f32(ldm_ctx* ctx, const float* z, float* img, int B, cudaStream_t st) {
  DecoderModel& D = ctx->dec;
  float *A = (float*)ctx->d_a, *Bf = (float*)ctx->d_b, *C = (float*)ctx->d_c;
  {  // Decoder.fc (v2:246-253); fc.3 rows were permuted so the result is already NHWC (B, 8, 8, 512)
    Epilogue e; e.bias = D.fc0.b; e.out_f32 = ctx->d_f0; e.ld_of = 512;
    LDM_TRY(launch_gemm_f32(ctx, z, D.latent, D.fc0.w32, B, 512, D.latent, e, st));
    LDM_TRY(launch_row_ln<float>(ctx, ctx->d_f0, 512, D.fc1_w, D.fc1_b, LDM_ACT_SWISH, ctx->d_f1, 512, B, 512, st));
    Epilogue e2; e2.bias = D.fc3.b; e2.out_f32 = Bf; e2.ld_of = 32768;
    LDM_TRY(launch_gemm_f32(ctx, ctx->d_f1, 512, D.fc3.w32, B, 32768, 512, e2, st));
    LDM_TRY(launch_row_ln<float>(ctx, Bf, 32768, D.fc4_w, D.fc4_b, LDM_ACT_SWISH, A, 32768, B, 32768, st));
  }
  // x in A.  res -> C (scratch Bf), up -> A (scratch Bf)
  LDM_TRY(res_block_f32(ctx, D.res[0], B, A, Bf, C, st));
  LDM_TRY(up_block_f32(ctx, D, 0, B, 8, 512, C, Bf, A, st));
  LDM_TRY(res_block_f32(ctx, D.res[1], B, A, Bf, C, st));
  LDM_TRY(up_block_f32(ctx, D, 1, B, 16, 256, C, Bf, A, st));
  LDM_TRY(res_block_f32(ctx, D.res[2], B, A, Bf, C, st));
  LDM_TRY(up_block_f32(ctx, D, 2, B, 32, 128, C, Bf, A, st));
  // final_conv (v2:272-278): conv 64->32, GroupNorm(8, 32), Swish, conv 32->3, Sigmoid; output NCHW
  ConvGeom g;
  conv3_geom(g, B, 64, 64, 32);
  if (D.strict_tc) LDM_TRY(conv_strict(ctx, A, D.fin0s, D.fin0.b, Bf, B, 64, 1, st));
  else LDM_TRY(launch_conv_f32(ctx, A, D.fin0.w32, D.fin0.b, Bf, g, st));
  LDM_TRY(launch_inorm_stats<float>(ctx, Bf, ctx->d_stats, B, 4096, 32, 4, st));
  LDM_TRY(launch_norm_apply<float>(ctx, Bf, ctx->d_stats, D.fin_gn_w, D.fin_gn_b, C, B, 4096, 32, 4, LDM_ACT_SWISH, st));
  conv3_geom(g, B, 64, 32, 3);
  g.nchw_out = 1;
  g.act = LDM_ACT_SIGMOID;
  LDM_TRY(launch_conv_f32(ctx, C, D.fin3.w32, D.fin3.b, img, g, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 tensor-core pass: activations NHWC bf16, every convolution on tcgen05 (conv_tc.cu), statistics in fp32
// ---------------------------------------------------------------------------------------------------------------
int res_block_bf16(ldm_ctx* ctx, const ResBlockModel& R, int B, const bf16* X, bf16* Y, bf16* OUT, cudaStream_t st) {
  const int C = R.C, H = R.HW, P = H * H;
  float2* coef = reinterpret_cast<float2*>(ctx->d_stats);
  LDM_TRY(launch_conv_tc(ctx, X, R.conv1, R.conv1.b, Y, B, H, H, 1, st));                              // conv1
  LDM_TRY(launch_norm_coef_bf16_ws(ctx, Y, R.ln1_w, R.ln1_b, coef, B, P, C, 1, ctx->d_part, ctx->d_cnt, st));   // ln1 as (scale, shift)
  LDM_TRY(launch_coef_apply_bf16(ctx, Y, coef, OUT, B, P, C, LDM_ACT_SWISH, st));                       // swish(ln1(.))
  LDM_TRY(launch_conv_tc(ctx, OUT, R.conv2, R.conv2.b, Y, B, H, H, 1, st));                             // conv2
  LDM_TRY(launch_norm_coef_bf16_ws(ctx, Y, R.ln2_w, R.ln2_b, coef, B, P, C, 1, ctx->d_part, ctx->d_cnt, st));   // ln2 as (scale, shift)
  // CALayer (v2:64-67): the average pool of an instance-normalised map is its beta, so the channel gate is a
  // per-channel constant computed at pack time (ca_const), the same for every sample
  int gated = 0;
  LDM_TRY(launch_sa_map_gate_bf16(ctx, Y, coef, R.ca_const, R.sa_w, ctx->d_gate, B, H, C, st, &gated));      // map + 7x7 gate, one CTA per sample
  if (!gated) LDM_TRY(launch_sa_map_bf16(ctx, Y, coef, R.ca_const, ctx->d_map, B, P, C, st));
  LDM_TRY(launch_sa_apply_bf16(ctx, Y, coef, R.ca_const, gated ? nullptr : ctx->d_map, R.sa_w, X, OUT, ctx->d_gate, B, H, C, st));
  return 0;
}

int up_block_bf16(ldm_ctx* ctx, const DecoderModel& D, int idx, int B, int H, int Cin, const bf16* X, bf16* Y, bf16* OUT,
                  cudaStream_t st) {
  const int Cout = Cin / 2, P = 4 * H * H;
  if (convt_halo_supported(H, H, Cin, Cout))      // up1 (128 -> 64 at 64 x 64): resident weights, every pixel run loaded once
    LDM_TRY(launch_convt_halo(ctx, X, Cin, D.up[idx][0], D.up_b[idx], Y, Cout, B, H, H, 0, st));
  else
    LDM_TRY(launch_conv_tc(ctx, X, D.up[idx][0], D.up_b[idx], Y, B, H, H, 2, st));   // four sub-pixel parities, one launch
  float2* coef = reinterpret_cast<float2*>(ctx->d_stats);
  LDM_TRY(launch_norm_coef_bf16_ws(ctx, Y, D.up_gn_w[idx], D.up_gn_b[idx], coef, B, P, Cout, 8, ctx->d_part, ctx->d_cnt, st));
  LDM_TRY(launch_coef_apply_bf16(ctx, Y, coef, OUT, B, P, Cout, LDM_ACT_SWISH, st));
  return 0;
}

int decode_chunk_bf16(ldm_ctx* ctx, const float* z, float* img, int B, cudaStream_t st) {
  DecoderModel& D = ctx->dec;
  bf16 *A = (bf16*)ctx->d_a, *Bf = (bf16*)ctx->d_b, *C = (bf16*)ctx->d_c;
  {  // Decoder.fc (v2:246-253) on the tensor cores; LayerNorm statistics over the fp32 accumulators
    LDM_TRY(launch_load_x<bf16>(ctx, z, ctx->d_zb, D.latent, B, D.latent, st));
    Epilogue e; e.bias = D.fc0.b; e.out_f32 = ctx->d_f0; e.ld_of = 512;
    LDM_TRY(launch_gemm_tc(ctx, ctx->d_zb, D.latent, B, D.fc0, e, st));
    LDM_TRY(launch_row_ln<bf16>(ctx, ctx->d_f0, 512, D.fc1_w, D.fc1_b, LDM_ACT_SWISH, ctx->d_h1b, 512, B, 512, st));
    Epilogue e2; e2.bias = D.fc3.b; e2.out_f32 = (float*)ctx->d_b; e2.ld_of = 32768; e2.stage_f32 = 1;
    LDM_TRY(launch_gemm_tc(ctx, ctx->d_h1b, 512, B, D.fc3, e2, st));
    LDM_TRY(launch_row_ln<bf16>(ctx, (const float*)ctx->d_b, 32768, D.fc4_w, D.fc4_b, LDM_ACT_SWISH, A, 32768, B, 32768, st));
  }
  LDM_TRY(res_block_bf16(ctx, D.res[0], B, A, Bf, C, st));
  LDM_TRY(up_block_bf16(ctx, D, 0, B, 8, 512, C, Bf, A, st));
  LDM_TRY(res_block_bf16(ctx, D.res[1], B, A, Bf, C, st));
  LDM_TRY(up_block_bf16(ctx, D, 1, B, 16, 256, C, Bf, A, st));
  LDM_TRY(res_block_bf16(ctx, D.res[2], B, A, Bf, C, st));
  LDM_TRY(up_block_bf16(ctx, D, 2, B, 32, 128, C, Bf, A, st));
  // final_conv (v2:272-278)
  // 64 -> 32 at 64 x 64: the halo kernel (weights resident, every pixel tile loaded once with its halo) instead of nine
  // re-loaded taps per tile
  LDM_TRY(launch_conv_halo(ctx, A, 64, D.fin0, D.fin0.b, Bf, 32, B, 64, 64, 0, nullptr, 0, nullptr, 0, st));
  float2* coef = reinterpret_cast<float2*>(ctx->d_stats);
  LDM_TRY(launch_norm_coef_bf16_ws(ctx, Bf, D.fin_gn_w, D.fin_gn_b, coef, B, 4096, 32, 4, ctx->d_part, ctx->d_cnt, st));   // GroupNorm(8, 32)
  // GroupNorm apply + Swish + final_conv.3 + Sigmoid as ONE pass (LDM_DEC_FUSE_OUT=0: the two-kernel sequence)
  static const bool fuse_out = !(getenv("LDM_DEC_FUSE_OUT") && atoi(getenv("LDM_DEC_FUSE_OUT")) == 0);
  if (fuse_out) return launch_final_gn_conv3(ctx, Bf, coef, D.fin3_frag, D.fin3.b, img, B, 64, 64, st);
  LDM_TRY(launch_coef_apply_bf16(ctx, Bf, coef, C, B, 4096, 32, LDM_ACT_SWISH, st));
  LDM_TRY(launch_conv_out3(ctx, C, D.fin3.w32, D.fin3.b, img, B, 64, 64, st));
  return 0;
}

// bf16 copies + TMA descriptors of the decoder weights (tensor-core path)
int dense_bf16(ldm_ctx* ctx, std::vector<void*>& pool, DenseLayer& L, int M_hint, cudaStream_t st) {
  LDM_TRY(ldm_alloc_t(ctx, pool, &L.w16, (size_t)L.N * L.K));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, (size_t)L.N * L.K, st));
  L.bn = tc_pick_bn(M_hint, L.N);
  return tc_make_weight_map(ctx, L.w16, L.N, L.K, L.bn, &L.map_w);
}
int conv_bf16(ldm_ctx* ctx, std::vector<void*>& pool, ConvLayer& L, cudaStream_t st) {
  const size_t n = (size_t)L.Cout * L.taps * L.Cin;
  LDM_TRY(ldm_alloc_t(ctx, pool, &L.w16, n));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, n, st));
  LDM_TRY(tc_make_weight_map(ctx, L.w16, L.Cout, L.taps * L.Cin, conv_tc_pick_bn(L.Cout), &L.map_w));
  if (L.Cout % 256 == 0) {      // 128-row box: each CTA of a pair loads half of a 256-channel tile (conv_tc_pair_kernel)
    L.bn_alt = 128;
    LDM_TRY(tc_make_weight_map(ctx, L.w16, L.Cout, L.taps * L.Cin, 128, &L.map_w_alt));
  }
  return 0;
}

}  // namespace

int decoder_pack_impl(ldm_ctx* ctx, const ldm_decoder_weights* w, cudaStream_t st) {
  DecoderModel& D = ctx->dec;
  LDM_CHECK(w->latent_dim > 0 && w->latent_dim % 16 == 0, "ldm_decoder_pack: latent_dim must be a positive multiple of 16");
  cudaDeviceSynchronize();
  for (void* p : D.allocs) cudaFree(p);
  D = DecoderModel();
  D.latent = w->latent_dim;
  auto& P = D.allocs;
  // fc
  D.fc0.N = 512; D.fc0.K = D.latent;
  LDM_TRY(own(ctx, P, w->fc0_w, (size_t)512 * D.latent, &D.fc0.w32, st));
  LDM_TRY(own(ctx, P, w->fc0_b, 512, &D.fc0.b, st));
  LDM_TRY(own(ctx, P, w->fc1_w, 512, &D.fc1_w, st));
  LDM_TRY(own(ctx, P, w->fc1_b, 512, &D.fc1_b, st));
  // fc.3 produces view(B, 512, 8, 8) (v2:282): flat index c*64 + p.  Permute its rows (and fc.4's affine, a
  // LayerNorm over the whole row, so statistics are unchanged) to p*512 + c: the output is NHWC directly.
  LDM_CHECK(w->fc3_w && w->fc3_b && w->fc4_w && w->fc4_b, "ldm_decoder_pack: fc weights missing");
  D.fc3.N = 32768; D.fc3.K = 512;
  LDM_TRY(ldm_alloc_t(ctx, P, &D.fc3.w32, (size_t)32768 * 512));
  LDM_TRY(ldm_alloc_t(ctx, P, &D.fc3.b, 32768));
  LDM_TRY(ldm_alloc_t(ctx, P, &D.fc4_w, 32768));
  LDM_TRY(ldm_alloc_t(ctx, P, &D.fc4_b, 32768));
  LDM_TRY(launch_permute_rows(ctx, w->fc3_w, D.fc3.w32, 512, 64, 512, st));
  LDM_TRY(launch_permute_rows(ctx, w->fc3_b, D.fc3.b, 512, 64, 1, st));
  LDM_TRY(launch_permute_rows(ctx, w->fc4_w, D.fc4_w, 512, 64, 1, st));
  LDM_TRY(launch_permute_rows(ctx, w->fc4_b, D.fc4_b, 512, 64, 1, st));
  const int chans[3] = {512, 256, 128}, sides[3] = {8, 16, 32};
  for (int i = 0; i < 3; ++i) {
    const int C = chans[i];
    const ldm_resblock_weights& r = w->res[i];
    ResBlockModel& R = D.res[i];
    R.C = C; R.HW = sides[i];
    LDM_TRY(pack_conv3(ctx, P, R.conv1, r.conv1_w, r.conv1_b, C, C, st));
    LDM_TRY(pack_conv3(ctx, P, R.conv2, r.conv2_w, r.conv2_b, C, C, st));
    LDM_TRY(own(ctx, P, r.ln1_w, C, &R.ln1_w, st));
    LDM_TRY(own(ctx, P, r.ln1_b, C, &R.ln1_b, st));
    LDM_TRY(own(ctx, P, r.ln2_w, C, &R.ln2_w, st));
    LDM_TRY(own(ctx, P, r.ln2_b, C, &R.ln2_b, st));
    LDM_TRY(own(ctx, P, r.ca_w0, (size_t)(C / 8) * C, &R.ca_w0, st));
    LDM_TRY(own(ctx, P, r.ca_w2, (size_t)C * (C / 8), &R.ca_w2, st));
    LDM_TRY(own(ctx, P, r.sa_w, 98, &R.sa_w, st));
    LDM_TRY(ldm_alloc_t(ctx, P, &R.ca_const, (size_t)C));
    LDM_TRY(launch_ca_const(ctx, R.ln2_b, R.ca_w0, R.ca_w2, R.ca_const, C, st));
    // ConvTranspose2d(C, C/2, 4, 2, 1) -> four 2x2-tap sub-pixel kernels
    LDM_CHECK(w->up_w[i] && w->up_b[i] && w->up_gn_w[i] && w->up_gn_b[i], "ldm_decoder_pack: up-block weights missing");
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        ConvLayer& L = D.up[i][pa * 2 + pb];
        L.Cin = C; L.Cout = C / 2; L.taps = 4;
        LDM_TRY(ldm_alloc_t(ctx, P, &L.w32, (size_t)(C / 2) * 4 * C));
        LDM_TRY(launch_pack_convT(ctx, w->up_w[i], L.w32, C, C / 2, pa, pb, st));
      }
    LDM_TRY(own(ctx, P, w->up_b[i], C / 2, &D.up_b[i], st));
    LDM_TRY(own(ctx, P, w->up_gn_w[i], C / 2, &D.up_gn_w[i], st));
    LDM_TRY(own(ctx, P, w->up_gn_b[i], C / 2, &D.up_gn_b[i], st));
  }
  LDM_TRY(pack_conv3(ctx, P, D.fin0, w->fin0_w, w->fin0_b, 32, 64, st));
  LDM_TRY(own(ctx, P, w->fin_gn_w, 32, &D.fin_gn_w, st));
  LDM_TRY(own(ctx, P, w->fin_gn_b, 32, &D.fin_gn_b, st));
  LDM_TRY(pack_conv3(ctx, P, D.fin3, w->fin3_w, w->fin3_b, 3, 32, st));
  if (ctx->precision == LDM_PRECISION_BF16) {
    LDM_CHECK(D.latent % 64 == 0, "ldm_decoder_pack: the tensor-core path needs latent_dim %% 64 == 0");
    LDM_TRY(dense_bf16(ctx, P, D.fc0, 256, st));
    LDM_TRY(dense_bf16(ctx, P, D.fc3, 256, st));
    for (int i = 0; i < 3; ++i) {
      LDM_TRY(conv_bf16(ctx, P, D.res[i].conv1, st));
      LDM_TRY(conv_bf16(ctx, P, D.res[i].conv2, st));
      // the four sub-pixel kernels stacked along the output-channel axis: rows [z*Cout, (z+1)*Cout) = parity z
      ConvLayer& U0 = D.up[i][0];
      const size_t per = (size_t)U0.Cout * 4 * U0.Cin;
      LDM_TRY(ldm_alloc_t(ctx, P, &U0.w16, 4 * per));
      for (int z = 0; z < 4; ++z) LDM_TRY(launch_to_bf16(ctx, D.up[i][z].w32, U0.w16 + (size_t)z * per, per, st));
      LDM_TRY(tc_make_weight_map(ctx, U0.w16, 4 * U0.Cout, 4 * U0.Cin, conv_tc_pick_bn(U0.Cout), &U0.map_w));
      if (U0.Cout % 256 == 0) {
        U0.bn_alt = 128;
        LDM_TRY(tc_make_weight_map(ctx, U0.w16, 4 * U0.Cout, 4 * U0.Cin, 128, &U0.map_w_alt));
      }
    }
    LDM_TRY(conv_bf16(ctx, P, D.fin0, st));
    LDM_TRY(ldm_alloc_t(ctx, P, &D.fin3_frag, (size_t)36 * 32));
    LDM_TRY(launch_final_w_frag(ctx, D.fin3.w32, D.fin3_frag, st));
  }
  if (ctx->precision != LDM_PRECISION_BF16) {
    // strict mode: the 3x3 / transposed convolutions as three-term bf16 products on the tensor cores (LDM_DEC_F32=1 keeps
    // them on the CUDA cores); split weights: per tap [hi | hi | lo] of the Cin columns
    const char* e = getenv("LDM_DEC_F32");
    D.strict_tc = !(e && atoi(e));
    if (D.strict_tc) {
      auto split_layer = [&](const ConvLayer& L, ConvLayer& S) -> int {
        S = ConvLayer();
        S.Cin = 3 * L.Cin; S.Cout = L.Cout; S.taps = L.taps; S.b = L.b;
        LDM_TRY(ldm_alloc_t(ctx, P, &S.w16, (size_t)L.Cout * L.taps * S.Cin));
        LDM_TRY(launch_pack_conv_split(ctx, L.w32, S.w16, L.Cout, L.taps, L.Cin, st));
        LDM_TRY(tc_make_weight_map(ctx, S.w16, S.Cout, S.taps * S.Cin, conv_tc_pick_bn(S.Cout), &S.map_w));
        if (S.Cout % 256 == 0) {
          S.bn_alt = 128;
          LDM_TRY(tc_make_weight_map(ctx, S.w16, S.Cout, S.taps * S.Cin, 128, &S.map_w_alt));
        }
        return 0;
      };
      for (int i = 0; i < 3; ++i) {
        LDM_TRY(split_layer(D.res[i].conv1, D.res[i].conv1s));
        LDM_TRY(split_layer(D.res[i].conv2, D.res[i].conv2s));
        const ConvLayer& U0 = D.up[i][0];
        ConvLayer& S = D.ups[i];
        S = ConvLayer();
        S.Cin = 3 * U0.Cin; S.Cout = U0.Cout; S.taps = 4;
        const size_t per = (size_t)U0.Cout * 4 * S.Cin;
        LDM_TRY(ldm_alloc_t(ctx, P, &S.w16, 4 * per));
        for (int z = 0; z < 4; ++z) LDM_TRY(launch_pack_conv_split(ctx, D.up[i][z].w32, S.w16 + (size_t)z * per, U0.Cout, 4, U0.Cin, st));
        LDM_TRY(tc_make_weight_map(ctx, S.w16, 4 * S.Cout, 4 * S.Cin, conv_tc_pick_bn(S.Cout), &S.map_w));
        if (S.Cout % 256 == 0) {
          S.bn_alt = 128;
          LDM_TRY(tc_make_weight_map(ctx, S.w16, 4 * S.Cout, 4 * S.Cin, 128, &S.map_w_alt));
        }
      }
      LDM_TRY(split_layer(D.fin0, D.fin0s));
    }
  }
  LDM_CUDA(cudaStreamSynchronize(st));
  D.packed = true;
  return 0;
}

int decoder_run_impl(ldm_ctx* ctx, const float* z, float* img, int B, cudaStream_t st) {
  const int chunk = B < kDecChunk ? B : kDecChunk;
  LDM_TRY(ensure_dec_workspace(ctx, chunk));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = B - b0 < chunk ? B - b0 : chunk;
    const float* zc = z + (size_t)b0 * ctx->dec.latent;
    float* ic = img + (size_t)b0 * 3 * 64 * 64;
    if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(decode_chunk_bf16(ctx, zc, ic, nb, st));
    else LDM_TRY(decode_chunk_f32(ctx, zc, ic, nb, st));
  }
  return 0;
}
