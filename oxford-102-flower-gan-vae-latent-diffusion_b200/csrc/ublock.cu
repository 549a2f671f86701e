// The convolutional U-Net blocks the v2 script defines next to its MLP denoiser: UNetResidualBlock
// (v2:462-486) and UNetAttentionBlock (v2:434-459).  The reference never instantiates them, so there is no end-to-end
// target; they are exposed as stand-alone operators with module-level parity, composed from the kernels of the sampling
// path: implicit-GEMM 3x3 convolutions (conv_tc.cu) with the time / class embedding broadcast-add as the epilogue's
// per-sample term, LayerNorm2d / GroupNorm as (scale, shift) coefficients + one fused apply pass (decoder_norm.cu), the
// 1x1 convolutions as tcgen05 GEMMs over pixels (gemm_tc.cu) and the 4-head spatial self-attention over the H W tokens
// as attn_tc_kernel (softmax(QK^T)V on tcgen05).  fp32 contexts run the same sequence on the CUDA-core kernels (strict mode).
// I/O is the module's own NCHW fp32.
#include "common.cuh"

int launch_conv_tc_ex(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int mode, int relu, const float* post, int post_stride, cudaStream_t st);
int tc_make_weight_map(ldm_ctx* ctx, const bf16* w, int N, int K, int bn, CUtensorMap* out);
int tc_init(ldm_ctx* ctx);
int tc_pick_bn(int M, int N);
int conv_tc_pick_bn(int Cout);
int launch_norm_coef_bf16(ldm_ctx* ctx, const bf16* x, const float* gamma, const float* beta, float2* coef, int B, int HW,
                          int C, int group, cudaStream_t st);
int launch_coef_apply_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, bf16* out, int B, int HW, int C, int act,
                           cudaStream_t st);

namespace {


// (B, C, HW) fp32 -> (B, HW, C) bf16 / fp32 through a 32 x 32 shared-memory tile
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    tile[i][tx] = (c0 + i < C && p0 + tx < HW) ? in[((size_t)n * C + c0 + i) * HW + p0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (p0 + i < HW && c0 + tx < C) out[((size_t)n * HW + p0 + i) * C + c0 + tx] = from_f32<T>(tile[tx][i]);
}

// out (B, C, HW) fp32 = a (B, HW, C) [+ r16 (B, HW, C)] [+ r32 (B, C, HW) fp32]     (a, r16: bf16 or fp32)
template <typename T>
__global__ void __launch_bounds__(256)
combine_nchw_kernel(const T* __restrict__ a, const T* __restrict__ r16, const float* __restrict__ r32, float* __restrict__ out,
                    int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    float v = 0.f;
    if (p0 + i < HW && c0 + tx < C) {
      const size_t k = ((size_t)n * HW + p0 + i) * C + c0 + tx;
      v = to_f32<T>(a[k]);
      if (r16) v += to_f32<T>(r16[k]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && p0 + tx < HW) {
      const size_t k = ((size_t)n * C + c0 + i) * HW + p0 + tx;
      out[k] = tile[tx][i] + (r32 ? r32[k] : 0.f);
    }
}

// post[n, co] = swish(Wt[co, :] . t[n, :] + bt[co]) [+ swish(Wc[co, :] . c[n, :] + bc[co])]      (v2:478-482)
__global__ void __launch_bounds__(128)
ub_emb_kernel(const float* __restrict__ t, const float* __restrict__ c, const float* __restrict__ wt, const float* __restrict__ bt,
              const float* __restrict__ wc, const float* __restrict__ bc, float* __restrict__ post, int Cout, int dt) {
  const int n = blockIdx.y, co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= Cout) return;
  float a = bt[co];
  for (int k = 0; k < dt; ++k) a += wt[(size_t)co * dt + k] * t[(size_t)n * dt + k];
  float v = swishf(a);
  if (c) {
    float b = bc[co];
    for (int k = 0; k < dt; ++k) b += wc[(size_t)co * dt + k] * c[(size_t)n * dt + k];
    v += swishf(b);
  }
  post[(size_t)n * Cout + co] = v;
}

// GroupNorm(1, C) (v2:439): statistics over ALL C x HW values of a sample -> per-channel (scale, shift).  One CTA per
// sample; sums are centred on the sample's first value.
template <typename T>
__global__ void __launch_bounds__(1024)
gn1_coef_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float2* __restrict__ coef,
                int HW, int C) {
  __shared__ float red[2][32];
  __shared__ float stat[2];
  const int n = blockIdx.x;
  const T* base = x + (size_t)n * HW * C;
  const float piv = to_f32<T>(base[0]);
  const size_t total = (size_t)HW * C;
  float s1 = 0.f, s2 = 0.f;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const float d = to_f32<T>(base[i]) - piv;
    s1 += d; s2 += d * d;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = red[0][threadIdx.x], b = red[1][threadIdx.x];
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) {
      const float md = a / (float)total;
      const float var = fmaxf(b / (float)total - md * md, 0.f);
      stat[0] = piv + md;
      stat[1] = rsqrtf(var + 1e-5f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = stat[1] * gamma[c];
    coef[(size_t)n * C + c] = make_float2(sc, beta[c] - stat[0] * sc);
  }
}

// V^T for attn_tc_kernel: vt[(n * C + ch), p] = qkv[(n * HW + p), 2 C + ch]
__global__ void __launch_bounds__(256)
vt_from_qkv_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ vt, int C, int HW) {
  __shared__ bf16 tile[32][34];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (p0 + i < HW && c0 + tx < C) tile[i][tx] = qkv[((size_t)n * HW + p0 + i) * 3 * C + 2 * C + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && p0 + tx < HW) vt[((size_t)n * C + c0 + i) * HW + p0 + tx] = tile[tx][i];
}


// ---- strict fp32 path (module-level parity within 1e-3): CUDA-core kernels over NHWC fp32 -----------------------------
// LayerNorm2d coefficients: one thread per (sample, channel), two passes over the H W values of that channel
__global__ void ln2d_coef_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float2* __restrict__ coef, int HW, int C, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i / C, c = i - n * C;
  const float* base = x + (size_t)n * HW * C + c;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += base[(size_t)p * C];
  const float mean = s / (float)HW;
  float q = 0.f;
  for (int p = 0; p < HW; ++p) { const float d = base[(size_t)p * C] - mean; q += d * d; }
  const float sc = rsqrtf(q / (float)HW + 1e-5f) * gamma[c];
  coef[i] = make_float2(sc, beta[c] - mean * sc);
}

__global__ void coef_apply_f32_kernel(const float* __restrict__ x, const float2* __restrict__ coef, float* __restrict__ out, int HW, int C,
                                      int act, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const size_t n = i / ((size_t)HW * C);
  const float2 k = coef[n * C + c];
  float v = x[i] * k.x + k.y;
  if (act == LDM_ACT_SWISH) v = swishf(v);
  out[i] = v;
}

// proj weight with its input columns permuted from the reference's (head_dim, head) order to (head, head_dim)
__global__ void permute_proj_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C) return;
  const int o = i / C, col = i - o * C, hd = C / heads, h = col / hd, e = col - h * hd;
  out[i] = w[(size_t)o * C + e * heads + h];
}

int conv3_f32(ldm_ctx* ctx, const float* in, const ConvLayer& L, float* out, int B, int H, int W, const float* post, cudaStream_t st) {
  ConvGeom g = ConvGeom();
  g.B = B; g.H = H; g.W = W; g.Cin = L.Cin; g.Cout = L.Cout; g.taps = 9; g.up = 1; g.post = post; g.post_stride = L.Cout;
  for (int t = 0; t < 9; ++t) { g.dy[t] = t / 3 - 1; g.dx[t] = t % 3 - 1; }
  return launch_conv_f32(ctx, in, L.w32, L.b, out, g, st);
}

int own(ldm_ctx* ctx, std::vector<void*>& pool, const float* src, size_t n, float** out, cudaStream_t st) {
  LDM_CHECK(src != nullptr, "ldm_ublock_*_pack: null weight pointer");
  LDM_TRY(ldm_alloc_t(ctx, pool, out, n));
  LDM_CUDA(cudaMemcpyAsync(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pack_conv3(ldm_ctx* ctx, std::vector<void*>& P, ConvLayer& L, const float* w, const float* b, int Cout, int Cin, cudaStream_t st) {
  LDM_CHECK(w && b, "ldm_ublock_res_pack: convolution weights missing");
  L.Cin = Cin; L.Cout = Cout; L.taps = 9; L.bn = conv_tc_pick_bn(Cout);
  const size_t n = (size_t)Cout * 9 * Cin;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w32, n));
  LDM_TRY(launch_pack_conv(ctx, w, L.w32, Cout, Cin, 3, 3, st));
  LDM_TRY(own(ctx, P, b, Cout, &L.b, st));
  if (ctx->precision != LDM_PRECISION_BF16) return 0;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w16, n));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, n, st));
  return tc_make_weight_map(ctx, L.w16, Cout, 9 * Cin, L.bn, &L.map_w);
}

// 1x1 convolution (Cout, Cin, 1, 1) == Linear over pixels
int pack_dense(ldm_ctx* ctx, std::vector<void*>& P, DenseLayer& L, const float* w, const float* b, int N, int K, cudaStream_t st) {
  L.N = N; L.K = K;
  LDM_TRY(own(ctx, P, w, (size_t)N * K, &L.w32, st));
  LDM_TRY(own(ctx, P, b, N, &L.b, st));
  if (ctx->precision != LDM_PRECISION_BF16) return 0;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w16, (size_t)N * K));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, (size_t)N * K, st));
  L.bn = tc_pick_bn(4096, N);
  return tc_make_weight_map(ctx, L.w16, N, K, L.bn, &L.map_w);
}

int grow(ldm_ctx* ctx, UBlock& u, size_t elems) {   // activation workspace: 6 NHWC bf16 buffers of `elems`
  if (elems <= u.ws_elems) return 0;
  LDM_CUDA(cudaDeviceSynchronize());
  for (void* p : u.ws) cudaFree(p);
  u.ws.clear();
  for (int i = 0; i < 6; ++i) LDM_TRY(ldm_alloc_t(ctx, u.ws, &u.buf[i], elems));
  LDM_TRY(ldm_alloc_t(ctx, u.ws, &u.qkv3, 3 * elems));
  LDM_TRY(ldm_alloc_t(ctx, u.ws, &u.coef, elems / 16 + 4096));     // (B, C) float2: far fewer than B * HW * C / 16 entries for HW >= 16
  LDM_TRY(ldm_alloc_t(ctx, u.ws, &u.post, elems / 16 + 4096));
  u.ws_elems = elems;
  return 0;
}

int get_block(ldm_ctx* ctx, int handle, int type, UBlock** out) {
  LDM_CHECK(handle >= 0 && handle < (int)ctx->ublocks.size() && ctx->ublocks[handle].type == type, "ldm_ublock: bad handle %d", handle);
  *out = &ctx->ublocks[handle];
  return 0;
}

int check_common(ldm_ctx* ctx, int B, int H, int W) {
  LDM_CHECK(B > 0 && H > 0 && W > 0 && H * W >= 16 && (H * W) % 8 == 0, "ldm_ublock_*: need H * W >= 16 and a multiple of 8 (got %d x %d)", H, W);
  return 0;
}

}  // namespace

static void ublock_release(UBlock& u) {
  for (void* p : u.ws) cudaFree(p);
  for (void* p : u.allocs) cudaFree(p);
  u = UBlock();      // type 0: a free slot
}

void ublock_free_all(ldm_ctx* ctx) {
  for (UBlock& u : ctx->ublocks) ublock_release(u);
  ctx->ublocks.clear();
}

// a finished block goes into the first free slot (handles of released blocks are reused), or at the end
static int ublock_store(ldm_ctx* ctx, UBlock& u) {
  for (size_t i = 0; i < ctx->ublocks.size(); ++i)
    if (ctx->ublocks[i].type == 0) { ctx->ublocks[i] = u; return (int)i; }
  ctx->ublocks.push_back(u);
  return (int)ctx->ublocks.size() - 1;
}

// Release the packed weights and the workspace of one conv U-Net block; its handle may be handed out again.
extern "C" LDM_API int ldm_ublock_free(ldm_ctx* ctx, int handle) {
  LDM_CHECK(ctx, "ldm_ublock_free: null context");
  LDM_CHECK(handle >= 0 && handle < (int)ctx->ublocks.size() && ctx->ublocks[handle].type != 0, "ldm_ublock_free: bad handle %d", handle);
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_CUDA(cudaDeviceSynchronize());      // no kernel of this context may still read the block
  ublock_release(ctx->ublocks[handle]);
  return 0;
}

extern "C" LDM_API int ldm_ublock_res_pack(ldm_ctx* ctx, const ldm_ublock_res_weights* w, int* handle_out, void* stream) {
  LDM_CHECK(ctx && w && handle_out, "ldm_ublock_res_pack: null argument");
  LDM_CHECK(w->in_channels % 64 == 0 && w->out_channels % 64 == 0 && w->in_channels > 0 && w->out_channels > 0 && w->d_time > 0,
            "ldm_ublock_res_pack: channel counts must be multiples of 64 (%d -> %d)", w->in_channels, w->out_channels);
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(tc_init(ctx));
  UBlock u;
  u.type = 1; u.cin = w->in_channels; u.cout = w->out_channels; u.dt = w->d_time;
  auto& P = u.allocs;
  auto body = [&]() -> int {
  LDM_TRY(own(ctx, P, w->norm1_w, u.cin, &u.n1w, st));
  LDM_TRY(own(ctx, P, w->norm1_b, u.cin, &u.n1b, st));
  LDM_TRY(own(ctx, P, w->norm2_w, u.cout, &u.n2w, st));
  LDM_TRY(own(ctx, P, w->norm2_b, u.cout, &u.n2b, st));
  LDM_TRY(own(ctx, P, w->time_w, (size_t)u.cout * u.dt, &u.tw, st));
  LDM_TRY(own(ctx, P, w->time_b, u.cout, &u.tb, st));
  LDM_TRY(own(ctx, P, w->class_w, (size_t)u.cout * u.dt, &u.cw, st));
  LDM_TRY(own(ctx, P, w->class_b, u.cout, &u.cb, st));
  LDM_TRY(pack_conv3(ctx, P, u.conv1, w->conv1_w, w->conv1_b, u.cout, u.cin, st));
  LDM_TRY(pack_conv3(ctx, P, u.conv2, w->conv2_w, w->conv2_b, u.cout, u.cout, st));
  if (u.cin != u.cout) {
    LDM_CHECK(w->res_w && w->res_b, "ldm_ublock_res_pack: the 1x1 residual convolution is required when in_channels != out_channels");
    LDM_TRY(pack_dense(ctx, P, u.d1, w->res_w, w->res_b, u.cout, u.cin, st));
  }
  LDM_CUDA(cudaStreamSynchronize(st));
  return 0;
  };
  const int rc = body();
  if (rc != 0) { cudaStreamSynchronize(st); ublock_release(u); return rc; }   // a pack that fails midway leaves nothing behind
  *handle_out = ublock_store(ctx, u);
  return 0;
}

// UNetResidualBlock.forward(x, t, c) (v2:475-486), eval mode (Dropout = identity).  x (B, Cin, H, W), t / c (B, d_time)
// fp32 (c may be NULL) -> out (B, Cout, H, W) fp32.
extern "C" LDM_API int ldm_ublock_res_forward(ldm_ctx* ctx, int handle, const float* x, const float* t, const float* c, float* out,
                                              int B, int H, int W, void* stream) {
  LDM_CHECK(ctx && x && t && out, "ldm_ublock_res_forward: null argument");
  UBlock* up;
  LDM_TRY(get_block(ctx, handle, 1, &up));
  UBlock& u = *up;
  LDM_TRY(check_common(ctx, B, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  const int HW = H * W, cm = u.cin > u.cout ? u.cin : u.cout;
  const bool f32 = ctx->precision != LDM_PRECISION_BF16;
  LDM_TRY(grow(ctx, u, (size_t)B * HW * cm * (f32 ? 2 : 1)));
  const dim3 tg(ceil_div(HW, 32), ceil_div(u.cin, 32), B);
  const dim3 og(ceil_div(HW, 32), ceil_div(u.cout, 32), B);
  if (f32) {
    float *xh = (float*)u.buf[0], *a = (float*)u.buf[1], *h1 = (float*)u.buf[2], *a2 = (float*)u.buf[3], *h2 = (float*)u.buf[4],
          *r = (float*)u.buf[5];
    const size_t n_in = (size_t)B * HW * u.cin, n_out = (size_t)B * HW * u.cout;
    nchw_to_nhwc_kernel<float><<<tg, 256, 0, st>>>(x, xh, u.cin, HW);
    ln2d_coef_f32_kernel<<<ceil_div(B * u.cin, 128), 128, 0, st>>>(xh, u.n1w, u.n1b, u.coef, HW, u.cin, B * u.cin);
    coef_apply_f32_kernel<<<(unsigned)((n_in + 255) / 256), 256, 0, st>>>(xh, u.coef, a, HW, u.cin, LDM_ACT_SWISH, n_in);
    ub_emb_kernel<<<dim3(ceil_div(u.cout, 128), B), 128, 0, st>>>(t, c, u.tw, u.tb, u.cw, u.cb, u.post, u.cout, u.dt);
    ctx->launches += 4;
    LDM_TRY(conv3_f32(ctx, a, u.conv1, h1, B, H, W, u.post, st));
    ln2d_coef_f32_kernel<<<ceil_div(B * u.cout, 128), 128, 0, st>>>(h1, u.n2w, u.n2b, u.coef, HW, u.cout, B * u.cout);
    coef_apply_f32_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(h1, u.coef, a2, HW, u.cout, LDM_ACT_SWISH, n_out);
    ctx->launches += 2;
    LDM_TRY(conv3_f32(ctx, a2, u.conv2, h2, B, H, W, nullptr, st));
    if (u.cin == u.cout) {
      combine_nchw_kernel<float><<<og, 256, 0, st>>>(h2, nullptr, x, out, u.cout, HW);
    } else {
      Epilogue e; e.bias = u.d1.b; e.out_f32 = r; e.ld_of = u.cout;
      LDM_TRY(launch_gemm_f32(ctx, xh, u.cin, u.d1.w32, B * HW, u.cout, u.cin, e, st));
      combine_nchw_kernel<float><<<og, 256, 0, st>>>(h2, r, nullptr, out, u.cout, HW);
    }
    LDM_LAUNCHED(ctx);
    return 0;
  }
  bf16 *xh = u.buf[0], *a = u.buf[1], *h1 = u.buf[2], *a2 = u.buf[3], *h2 = u.buf[4], *r = u.buf[5];
  nchw_to_nhwc_kernel<bf16><<<tg, 256, 0, st>>>(x, xh, u.cin, HW);
  LDM_LAUNCHED(ctx);
  // h = conv1(act(norm1(x))) + act(time_emb(t)) [+ act(class_emb(c))]
  LDM_TRY(launch_norm_coef_bf16(ctx, xh, u.n1w, u.n1b, u.coef, B, HW, u.cin, 1, st));
  LDM_TRY(launch_coef_apply_bf16(ctx, xh, u.coef, a, B, HW, u.cin, LDM_ACT_SWISH, st));
  ub_emb_kernel<<<dim3(ceil_div(u.cout, 128), B), 128, 0, st>>>(t, c, u.tw, u.tb, u.cw, u.cb, u.post, u.cout, u.dt);
  LDM_LAUNCHED(ctx);
  LDM_TRY(launch_conv_tc_ex(ctx, a, u.cin, u.conv1, u.conv1.b, h1, u.cout, B, H, W, 1, 0, u.post, u.cout, st));
  // h = conv2(act(norm2(h)))
  LDM_TRY(launch_norm_coef_bf16(ctx, h1, u.n2w, u.n2b, u.coef, B, HW, u.cout, 1, st));
  LDM_TRY(launch_coef_apply_bf16(ctx, h1, u.coef, a2, B, HW, u.cout, LDM_ACT_SWISH, st));
  LDM_TRY(launch_conv_tc_ex(ctx, a2, u.cout, u.conv2, u.conv2.b, h2, u.cout, B, H, W, 1, 0, nullptr, 0, st));
  // + residual(x): identity (the fp32 input itself) or the 1x1 convolution
  if (u.cin == u.cout) {
    combine_nchw_kernel<bf16><<<og, 256, 0, st>>>(h2, nullptr, x, out, u.cout, HW);
  } else {
    Epilogue e; e.bias = u.d1.b; e.out_bf16 = r; e.ld_ob = u.cout;
    LDM_TRY(launch_gemm_tc(ctx, xh, u.cin, B * HW, u.d1, e, st));
    combine_nchw_kernel<bf16><<<og, 256, 0, st>>>(h2, r, nullptr, out, u.cout, HW);
  }
  LDM_LAUNCHED(ctx);
  return 0;
}

extern "C" LDM_API int ldm_ublock_attn_pack(ldm_ctx* ctx, const ldm_ublock_attn_weights* w, int* handle_out, void* stream) {
  LDM_CHECK(ctx && w && handle_out, "ldm_ublock_attn_pack: null argument");
  LDM_CHECK(w->channels > 0 && w->channels % 64 == 0 && w->num_heads > 0 && w->channels % w->num_heads == 0,
            "ldm_ublock_attn_pack: channels must be a multiple of 64 and of num_heads (%d, %d heads)", w->channels, w->num_heads);
  LDM_CHECK(attn_tc_supported(w->channels / w->num_heads), "ldm_ublock_attn_pack: head_dim %d unsupported (16, 32, 64, 128)",
            w->channels / w->num_heads);
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(tc_init(ctx));
  UBlock u;
  u.type = 2; u.cin = u.cout = w->channels; u.heads = w->num_heads;
  auto& P = u.allocs;
  auto body = [&]() -> int {
  LDM_TRY(own(ctx, P, w->norm_w, u.cin, &u.n1w, st));
  LDM_TRY(own(ctx, P, w->norm_b, u.cin, &u.n1b, st));
  LDM_TRY(pack_dense(ctx, P, u.d1, w->qkv_w, w->qkv_b, 3 * u.cin, u.cin, st));
  LDM_TRY(pack_dense(ctx, P, u.d2, w->proj_w, w->proj_b, u.cin, u.cin, st));
  if (ctx->precision != LDM_PRECISION_BF16) {   // strict path: standard (head, head_dim) attention output, proj columns permuted
    LDM_TRY(ldm_alloc_t(ctx, P, &u.proj_perm, (size_t)u.cin * u.cin));
    permute_proj_kernel<<<ceil_div(u.cin * u.cin, 256), 256, 0, st>>>(u.d2.w32, u.proj_perm, u.cin, u.heads);
    LDM_LAUNCHED(ctx);
  }
  LDM_CUDA(cudaStreamSynchronize(st));
  return 0;
  };
  const int rc = body();
  if (rc != 0) { cudaStreamSynchronize(st); ublock_release(u); return rc; }
  *handle_out = ublock_store(ctx, u);
  return 0;
}

// UNetAttentionBlock.forward(x) (v2:444-459): GroupNorm(1, C) -> qkv 1x1 -> 4-head attention over the H W tokens -> the
// reference's (head_dim, head) channel interleave (out.permute(0, 3, 1, 2).reshape, v2:456-457) -> proj 1x1 -> + x.
extern "C" LDM_API int ldm_ublock_attn_forward(ldm_ctx* ctx, int handle, const float* x, float* out, int B, int H, int W, void* stream) {
  LDM_CHECK(ctx && x && out, "ldm_ublock_attn_forward: null argument");
  UBlock* up;
  LDM_TRY(get_block(ctx, handle, 2, &up));
  UBlock& u = *up;
  LDM_TRY(check_common(ctx, B, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  const int HW = H * W, C = u.cin, hd = C / u.heads;
  const bool f32 = ctx->precision != LDM_PRECISION_BF16;
  LDM_TRY(grow(ctx, u, (size_t)B * HW * C * (f32 ? 2 : 1)));
  const dim3 tg(ceil_div(HW, 32), ceil_div(C, 32), B);
  if (f32) {
    float *xh = (float*)u.buf[0], *xn = (float*)u.buf[1], *qkv = (float*)u.qkv3, *att = (float*)u.buf[3], *pr = (float*)u.buf[4];
    const size_t n = (size_t)B * HW * C;
    nchw_to_nhwc_kernel<float><<<tg, 256, 0, st>>>(x, xh, C, HW);
    gn1_coef_kernel<float><<<B, 1024, 0, st>>>(xh, u.n1w, u.n1b, u.coef, HW, C);
    coef_apply_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xh, u.coef, xn, HW, C, LDM_ACT_NONE, n);
    ctx->launches += 3;
    LDM_CUDA(cudaGetLastError());
    {
      Epilogue e; e.bias = u.d1.b; e.out_f32 = qkv; e.ld_of = 3 * C;
      LDM_TRY(launch_gemm_f32(ctx, xn, C, u.d1.w32, B * HW, 3 * C, C, e, st));
    }
    for (int b = 0; b < B; ++b)      // the tokens of one sample attend to each other: one call per sample
      LDM_TRY(launch_batch_attention<float>(ctx, qkv + (size_t)b * HW * 3 * C, att + (size_t)b * HW * C, HW, C, u.heads, st));
    {
      Epilogue e; e.bias = u.d2.b; e.out_f32 = pr; e.ld_of = C;
      LDM_TRY(launch_gemm_f32(ctx, att, C, u.proj_perm, B * HW, C, C, e, st));
    }
    combine_nchw_kernel<float><<<tg, 256, 0, st>>>(pr, nullptr, x, out, C, HW);
    LDM_LAUNCHED(ctx);
    return 0;
  }
  bf16 *xh = u.buf[0], *xn = u.buf[1], *qkv = u.qkv3, *vt = u.buf[2], *att = u.buf[3], *pr = u.buf[4];
  nchw_to_nhwc_kernel<bf16><<<tg, 256, 0, st>>>(x, xh, C, HW);
  LDM_LAUNCHED(ctx);
  gn1_coef_kernel<bf16><<<B, 1024, 0, st>>>(xh, u.n1w, u.n1b, u.coef, HW, C);
  LDM_LAUNCHED(ctx);
  LDM_TRY(launch_coef_apply_bf16(ctx, xh, u.coef, xn, B, HW, C, LDM_ACT_NONE, st));
  {
    Epilogue e; e.bias = u.d1.b; e.out_bf16 = qkv; e.ld_ob = 3 * C;
    LDM_TRY(launch_gemm_tc(ctx, xn, C, B * HW, u.d1, e, st));
  }
  vt_from_qkv_kernel<<<tg, 256, 0, st>>>(qkv, vt, C, HW);
  LDM_LAUNCHED(ctx);
  // channel of (head h, dim e) in the reference's output = e * heads + h
  LDM_TRY(launch_attn_tc(ctx, qkv, 3 * C, 3 * C, vt, HW, HW, B, u.heads, hd, 0, C, att, 1, C, 1, u.heads, st));
  {
    Epilogue e; e.bias = u.d2.b; e.out_bf16 = pr; e.ld_ob = C;
    LDM_TRY(launch_gemm_tc(ctx, att, C, B * HW, u.d2, e, st));
  }
  combine_nchw_kernel<bf16><<<tg, 256, 0, st>>>(pr, nullptr, x, out, C, HW);
  LDM_LAUNCHED(ctx);
  return 0;
}
