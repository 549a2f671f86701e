// Memory-bound kernels of the VAE decoder over NHWC activations (T = float strict path, bf16 tensor path):
// LayerNorm2d / GroupNorm statistics and application (v2:144-156, v2:257,263,269,274), the CALayer
// squeeze-excite (v2:53-67) and SpatialAttention gating fused with the residual add + Swish of
// ResidualBlock.forward (v2:69-81,170-178).
#include "common.cuh"

namespace {

// thread layout of the per-(sample, channel-block) reductions: 32 channels x 8 pixel lanes
constexpr int CB = 32, PL = 8;

// stats[(n*G + g)*2 + {0,1}] = mean, 1/sqrt(var + eps) over HW pixels x cg channels (biased var, two-pass)
template <typename T>
__global__ void __launch_bounds__(CB * PL)
inorm_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, int HW, int C, int cg) {
  __shared__ float red[PL][CB];
  __shared__ float chv[CB];
  const int n = blockIdx.y, c0 = blockIdx.x * CB, tx = threadIdx.x % CB, ty = threadIdx.x / CB;
  const T* base = x + (size_t)n * HW * C + c0 + tx;
  const float cnt = (float)HW * (float)cg;
  float s = 0.f;
  for (int p = ty; p < HW; p += PL) s += to_f32<T>(base[(size_t)p * C]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) t += red[i][tx];
    chv[tx] = t;
  }
  __syncthreads();
  const int g0 = (tx / cg) * cg;
  float gs = 0.f;
  for (int i = 0; i < cg; ++i) gs += chv[g0 + i];
  const float mean = gs / cnt;
  float q = 0.f;
  for (int p = ty; p < HW; p += PL) {
    float d = to_f32<T>(base[(size_t)p * C]) - mean;
    q += d * d;
  }
  __syncthreads();
  red[ty][tx] = q;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) t += red[i][tx];
    chv[tx] = t;
  }
  __syncthreads();
  if (ty == 0 && tx % cg == 0) {
    float gq = 0.f;
    for (int i = 0; i < cg; ++i) gq += chv[tx + i];
    const int G = C / cg, g = (c0 + tx) / cg;
    stats[((size_t)n * G + g) * 2 + 0] = mean;
    stats[((size_t)n * G + g) * 2 + 1] = 1.0f / sqrtf(gq / cnt + 1e-5f);
  }
}

template <typename T>
__global__ void norm_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  T* __restrict__ out, int HW, int C, int cg, int act, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const size_t n = i / ((size_t)HW * C);
  const int G = C / cg;
  const float* s = stats + (n * G + c / cg) * 2;
  float v = (to_f32<T>(x[i]) - s[0]) * s[1] * gamma[c] + beta[c];
  if (act == LDM_ACT_SWISH) v = swishf(v);
  out[i] = from_f32<T>(v);
}

// gap[n][c] = mean over pixels of LayerNorm2d(x)  (AdaptiveAvgPool2d(1) of v2:65 applied to ln2's output)
template <typename T>
__global__ void __launch_bounds__(CB * PL)
gap_norm_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ gap, int HW, int C) {
  __shared__ float red[PL][CB];
  const int n = blockIdx.y, c = blockIdx.x * CB + threadIdx.x % CB, ty = threadIdx.x / CB;
  const T* base = x + (size_t)n * HW * C + c;
  const float mean = stats[((size_t)n * C + c) * 2], rstd = stats[((size_t)n * C + c) * 2 + 1];
  const float g = gamma[c], b = beta[c];
  float s = 0.f;
  for (int p = ty; p < HW; p += PL) s += (to_f32<T>(base[(size_t)p * C]) - mean) * rstd * g + b;
  red[ty][threadIdx.x % CB] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) t += red[i][threadIdx.x];
    gap[(size_t)n * C + c] = t / (float)HW;
  }
}

// ca[n][c] = sigmoid(W2 . swish(W0 . gap[n]))   (conv_du, v2:57-62; 1x1 convs without bias)
__global__ void __launch_bounds__(256)
ca_mlp_kernel(const float* __restrict__ gap, const float* __restrict__ w0, const float* __restrict__ w2,
              float* __restrict__ ca, int C) {
  extern __shared__ float sm[];
  float* sg = sm;        // [C]
  float* sh = sm + C;    // [C/8]
  const int n = blockIdx.x, R = C / 8;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sg[c] = gap[(size_t)n * C + c];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < R; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w0[(size_t)j * C + c] * sg[c];
    s = warp_sum(s);
    if (lane == 0) sh[j] = swishf(s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < R; ++j) s += w2[(size_t)c * R + j] * sh[j];
    ca[(size_t)n * C + c] = sigmoidf_(s);
  }
}

// map[n][p][0] = mean_c z, map[n][p][1] = max_c z with z = ca * LayerNorm2d(x)   (v2:76-78 on v2:67's output)
template <typename T>
__global__ void __launch_bounds__(256)
sa_map_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ ca, int ca_stride,
              float* __restrict__ map, int HW, int C, int npix) {
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pix >= npix) return;
  const int n = pix / HW;
  const T* row = x + (size_t)pix * C;
  const float* st = stats + (size_t)n * C * 2;
  const float* cav = ca + (size_t)n * ca_stride;
  float s = 0.f, m = -INFINITY;
  for (int c = lane; c < C; c += 32) {
    float z = cav[c] * ((to_f32<T>(row[c]) - st[2 * c]) * st[2 * c + 1] * gamma[c] + beta[c]);
    s += z;
    m = fmaxf(m, z);
  }
  s = warp_sum(s);
  m = warp_max(m);
  if (lane == 0) {
    map[(size_t)pix * 2 + 0] = s / (float)C;
    map[(size_t)pix * 2 + 1] = m;
  }
}

// out = swish(z * sigmoid(conv7x7(map)) + resid)   (v2:79-81 then v2:176-177)
template <typename T>
__global__ void __launch_bounds__(256)
sa_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ ca, int ca_stride,
                const float* __restrict__ map, const float* __restrict__ sa_w, const T* __restrict__ resid,
                T* __restrict__ out, int H, int C, int npix) {
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pix >= npix) return;
  const int HW = H * H, n = pix / HW, rem = pix - n * HW, y = rem / H, xx = rem - y * H;
  float a = 0.f;
  for (int t = lane; t < 98; t += 32) {
    const int ch = t / 49, k = t - ch * 49, ky = k / 7, kx = k - ky * 7;
    const int yy = y + ky - 3, xq = xx + kx - 3;
    if (yy >= 0 && yy < H && xq >= 0 && xq < H) a += sa_w[t] * map[((size_t)n * HW + yy * H + xq) * 2 + ch];
  }
  const float gate = sigmoidf_(warp_sum(a));
  const T* row = x + (size_t)pix * C;
  const T* rr = resid + (size_t)pix * C;
  T* orow = out + (size_t)pix * C;
  const float* st = stats + (size_t)n * C * 2;
  const float* cav = ca + (size_t)n * ca_stride;
  for (int c = lane; c < C; c += 32) {
    float z = cav[c] * ((to_f32<T>(row[c]) - st[2 * c]) * st[2 * c + 1] * gamma[c] + beta[c]);
    orow[c] = from_f32<T>(swishf(z * gate + to_f32<T>(rr[c])));
  }
}

}  // namespace

#define LDM_LAUNCHED(ctx)         \
  do {                            \
    (ctx)->launches++;            \
    LDM_CUDA(cudaGetLastError()); \
  } while (0)

template <typename T>
int launch_inorm_stats(ldm_ctx* ctx, const T* x, float* stats, int B, int HW, int C, int group, cudaStream_t st) {
  LDM_CHECK(C % CB == 0 && CB % group == 0, "inorm_stats: C %% 32 == 0 and group | 32 required (C=%d group=%d)", C, group);
  inorm_stats_kernel<T><<<dim3(C / CB, B), CB * PL, 0, st>>>(x, stats, HW, C, group);
  LDM_LAUNCHED(ctx);
  return 0;
}
template <typename T>
int launch_norm_apply(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta, T* out,
                      int B, int HW, int C, int group, int act, cudaStream_t st) {
  const size_t total = (size_t)B * HW * C;
  norm_apply_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, stats, gamma, beta, out, HW, C, group, act, total);
  LDM_LAUNCHED(ctx);
  return 0;
}
template <typename T>
int launch_gap_norm(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta, float* gap,
                    int B, int HW, int C, cudaStream_t st) {
  LDM_CHECK(C % CB == 0, "gap_norm: C %% 32 == 0 required");
  gap_norm_kernel<T><<<dim3(C / CB, B), CB * PL, 0, st>>>(x, stats, gamma, beta, gap, HW, C);
  LDM_LAUNCHED(ctx);
  return 0;
}
int launch_ca_mlp(ldm_ctx* ctx, const float* gap, const float* w0, const float* w2, float* ca, int B, int C,
                  cudaStream_t st) {
  ca_mlp_kernel<<<B, 256, (C + C / 8) * sizeof(float), st>>>(gap, w0, w2, ca, C);
  LDM_LAUNCHED(ctx);
  return 0;
}
template <typename T>
int launch_sa_map(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta, const float* ca,
                  int ca_stride, float* map, int B, int HW, int C, cudaStream_t st) {
  const int npix = B * HW;
  sa_map_kernel<T><<<ceil_div(npix, 8), 256, 0, st>>>(x, stats, gamma, beta, ca, ca_stride, map, HW, C, npix);
  LDM_LAUNCHED(ctx);
  return 0;
}
template <typename T>
int launch_sa_apply(ldm_ctx* ctx, const T* x, const float* stats, const float* gamma, const float* beta,
                    const float* ca, int ca_stride, const float* map, const float* sa_w, const T* resid, T* out, int B,
                    int H, int C, cudaStream_t st) {
  const int npix = B * H * H;
  sa_apply_kernel<T><<<ceil_div(npix, 8), 256, 0, st>>>(x, stats, gamma, beta, ca, ca_stride, map, sa_w, resid, out, H, C, npix);
  LDM_LAUNCHED(ctx);
  return 0;
}

#define INST(T)                                                                                                        \
  template int launch_inorm_stats<T>(ldm_ctx*, const T*, float*, int, int, int, int, cudaStream_t);                    \
  template int launch_norm_apply<T>(ldm_ctx*, const T*, const float*, const float*, const float*, T*, int, int, int,  \
                                    int, int, cudaStream_t);                                                           \
  template int launch_gap_norm<T>(ldm_ctx*, const T*, const float*, const float*, const float*, float*, int, int, int, \
                                  cudaStream_t);                                                                       \
  template int launch_sa_map<T>(ldm_ctx*, const T*, const float*, const float*, const float*, const float*, int,      \
                                float*, int, int, int, cudaStream_t);                                                  \
  template int launch_sa_apply<T>(ldm_ctx*, const T*, const float*, const float*, const float*, const float*, int,    \
                                  const float*, const float*, const T*, T*, int, int, int, cudaStream_t);
INST(float)
INST(bf16)
