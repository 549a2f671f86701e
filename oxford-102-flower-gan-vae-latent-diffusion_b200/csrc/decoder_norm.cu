// Memory-bound kernels of the VAE decoder over NHWC activations (T = float strict path, bf16 tensor path):
// LayerNorm2d / GroupNorm statistics and application (v2:144-156, v2:257,263,269,274), the CALayer
// squeeze-excite (v2:53-67) and SpatialAttention gating fused with the residual add + Swish of
// ResidualBlock.forward (v2:69-81,170-178).
#include <limits.h>

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

// ---------------------------------------------------------------------------------------------------------------
// bf16 fast path: 16-byte (8-channel) accesses, one pass statistics, (scale, shift) precomputed per (sample, channel)
// ---------------------------------------------------------------------------------------------------------------
// Swish of a value that is rounded to bf16 right away: the two-instruction division (2 ulp of fp32) instead of the IEEE sequence.
// The apply passes spend ~150 instructions per 16-byte load with the latter and are issue-bound before they are memory-bound.
__device__ __forceinline__ float swishf_bf16out(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// Every kernel of the bf16 path starts with ldm_pdl_wait(): launched through launch_maybe_pdl, the grid may be scheduled while
// its predecessor in the stream drains and blocks there until the predecessor's writes are visible (a no-op for a plain launch).
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
    f[2 * e] = __low2float(h2);
    f[2 * e + 1] = __high2float(h2);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h2);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// One CTA per (sample, CB-channel block, pixel split): threads = CB / 8 channel-octets x 2048 / CB pixel lanes, four 16-byte
// loads in flight per thread.  With S > 1 splits every CTA leaves its per-channel partial sums (centred on the sample's
// first pixel, so they add) in `part`; the last CTA of a (sample, block) to finish, by ticket, adds them IN SPLIT ORDER
// (deterministic) and writes the coefficients.  `cnt` must be zero on entry and is left zero.
template <int CB>
__global__ void __launch_bounds__(256)
norm_coef2_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float2* __restrict__ coef, int HW, int C, int cg, float2* __restrict__ part, int* __restrict__ cnt, int S) {
  ldm_pdl_wait();
  constexpr int OCT = CB / 8, PL = 256 / OCT;
  __shared__ float s1[PL][CB + 1], s2[PL][CB + 1];
  __shared__ int s_last;
  const int n = blockIdx.y, cb = blockIdx.x, c0 = cb * CB, sp = blockIdx.z;
  const int oct = threadIdx.x % OCT, pl = threadIdx.x / OCT;
  const bf16* base = x + (size_t)n * HW * C + c0 + oct * 8;
  float piv[8], a[8], b[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base)), piv);
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = 0.f; b[e] = 0.f; }
  const int per = HW / S, p_hi = (sp + 1) * per;
  int p = sp * per + pl;
  for (; p + 3 * PL < p_hi; p += 4 * PL) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + i * PL) * C));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = f[e] - piv[e]; a[e] += d; b[e] = fmaf(d, d, b[e]); }
    }
  }
  for (; p < p_hi; p += PL) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)p * C)), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = f[e] - piv[e]; a[e] += d; b[e] = fmaf(d, d, b[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[pl][oct * 8 + e] = a[e]; s2[pl][oct * 8 + e] = b[e]; }
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
  if (threadIdx.x < CB) {
#pragma unroll 8
    for (int i = 0; i < PL; ++i) { t1 += s1[i][threadIdx.x]; t2 += s2[i][threadIdx.x]; }
  }
  if (S > 1) {
    float2* mine = part + (((size_t)n * gridDim.x + cb) * S) * CB;
    if (threadIdx.x < CB) mine[(size_t)sp * CB + threadIdx.x] = make_float2(t1, t2);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int ticket = atomicAdd(&cnt[n * gridDim.x + cb], 1);
      s_last = ticket == S - 1;
      if (s_last) cnt[n * gridDim.x + cb] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < CB) {
      t1 = 0.f; t2 = 0.f;
      for (int i = 0; i < S; ++i) { const float2 v = __ldcg(mine + (size_t)i * CB + threadIdx.x); t1 += v.x; t2 += v.y; }
    }
  }
  __syncthreads();
  if (threadIdx.x < CB) {
    const int c = threadIdx.x;
    // per-channel (mean, M2), then merge the cg channels of the group (Chan et al.)
    const float inv = 1.0f / (float)HW;
    const float pv = __bfloat162float(x[(size_t)n * HW * C + c0 + c]);
    const float dm = t1 * inv;
    s1[0][c] = pv + dm;
    s2[0][c] = fmaxf(t2 - t1 * dm, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < CB) {
    const int c = threadIdx.x, g0 = (c / cg) * cg;
    float gm = 0.f;
    for (int i = 0; i < cg; ++i) gm += s1[0][g0 + i];
    gm /= (float)cg;
    float gq = 0.f;
    for (int i = 0; i < cg; ++i) { const float d = s1[0][g0 + i] - gm; gq += s2[0][g0 + i] + (float)HW * d * d; }
    const float rstd = rsqrtf(gq / ((float)HW * (float)cg) + 1e-5f);
    const float sc = rstd * gamma[c0 + c];
    coef[(size_t)n * C + c0 + c] = make_float2(sc, beta[c0 + c] - gm * sc);
  }
}

// out = act(x * scale + shift), 8 channels per thread
__global__ void __launch_bounds__(256)
coef_apply_bf16_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, bf16* __restrict__ out, int HWC8, int C8,
                       int act, size_t total8) {
  ldm_pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int n = (int)(i / HWC8), c8 = (int)(i % C8);
  float f[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
  const float4* cf = reinterpret_cast<const float4*>(coef + (size_t)n * C8 * 8 + c8 * 8);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 k = __ldg(cf + e);   // (scale, shift) of two channels
    float v0 = f[2 * e] * k.x + k.y, v1 = f[2 * e + 1] * k.z + k.w;
    if (act == LDM_ACT_SWISH) { v0 = swishf_bf16out(v0); v1 = swishf_bf16out(v1); }
    f[2 * e] = v0; f[2 * e + 1] = v1;
  }
  reinterpret_cast<uint4*>(out)[i] = pack8(f);
}
// The same for power-of-two channel counts with HW C / 8 a multiple of 1024 (every layer of the decoder): blockIdx.y = sample,
// a thread keeps its channel octet (256 is a multiple of C / 8), fetches the octet's coefficients once and streams FOUR
// 16-byte loads per iteration (independent, issued before the first use); 32-bit index arithmetic, no division.
constexpr int kCaUnroll = 4;
__global__ void __launch_bounds__(256)
coef_apply_bf16_stream_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, bf16* __restrict__ out, int HWC8, int C8,
                              int act, int iters) {
  ldm_pdl_wait();
  const int n = blockIdx.y, c8 = threadIdx.x & (C8 - 1);
  const uint4* xi = reinterpret_cast<const uint4*>(x) + (size_t)n * HWC8;
  uint4* xo = reinterpret_cast<uint4*>(out) + (size_t)n * HWC8;
  float sc[8], sh[8];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef + ((size_t)n * C8 + c8) * 8);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 k = __ldg(cf + e);
      sc[2 * e] = k.x; sh[2 * e] = k.y; sc[2 * e + 1] = k.z; sh[2 * e + 1] = k.w;
    }
  }
  int i = blockIdx.x * (256 * kCaUnroll * iters) + threadIdx.x;
  for (int it = 0; it < iters; ++it, i += 256 * kCaUnroll) {
    uint4 u[kCaUnroll];
#pragma unroll
    for (int j = 0; j < kCaUnroll; ++j) u[j] = __ldcs(xi + i + 256 * j);
#pragma unroll
    for (int j = 0; j < kCaUnroll; ++j) {
      float f[8];
      unpack8(u[j], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = f[e] * sc[e] + sh[e];
        if (act == LDM_ACT_SWISH) v = swishf_bf16out(v);
        f[e] = v;
      }
      xo[i + 256 * j] = pack8(f);
    }
  }
}

// SpatialAttention input (v2:76-78): per pixel, mean and max over channels of z = ca[c] * (x * scale + shift).
// A warp walks `run` consecutive pixels of ONE sample: LPP = min(32, C / 8) lanes share a pixel (NO = C / (8 LPP) channel
// octets per lane), 32 / LPP pixels side by side; the lane's (ca * scale, ca * shift) stay in registers for the whole run.
template <int NO>
__global__ void __launch_bounds__(256)
sa_map2_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, const float* __restrict__ ca,
               float* __restrict__ map, int HW, int C, int npix, int run) {
  ldm_pdl_wait();
  const int lpp = C / (8 * NO) < 32 ? C / (8 * NO) : 32, ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int first = wid * run;
  if (first >= npix) return;
  const int n = first / HW, sub = lane % lpp, pj = lane / lpp;
  float A[NO][8], Bv[NO][8];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const int c = (sub + o * lpp) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 k = __ldg(coef + (size_t)n * C + c + e);
      const float g = __ldg(ca + c + e);
      A[o][e] = g * k.x; Bv[o][e] = g * k.y;
    }
  }
  const float inv_c = 1.0f / (float)C;
  for (int i = 0; i < run; i += 2 * ppw) {       // two pixel groups per iteration: two independent 16-byte loads in flight
    uint4 u[2][NO];
    int pix[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      pix[h] = first + i + h * ppw + pj;
      const bool ok = i + h * ppw < run && pix[h] < npix;
      pix[h] = ok ? pix[h] : -1;
#pragma unroll
      for (int o = 0; o < NO; ++o)
        u[h][o] = ok ? __ldg(reinterpret_cast<const uint4*>(x + (size_t)pix[h] * C + (sub + o * lpp) * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float s = 0.f, m = -INFINITY;
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        float f[8];
        unpack8(u[h][o], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float z = fmaf(f[e], A[o][e], Bv[o][e]); s += z; m = fmaxf(m, z); }
      }
      for (int o = lpp >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      }
      if (pix[h] >= 0 && sub == 0) *reinterpret_cast<float2*>(map + (size_t)pix[h] * 2) = make_float2(s * inv_c, m);
    }
  }
}

// sa_map2 + sa_gate as ONE kernel for maps of up to 32 x 32 pixels: one CTA per sample walks the sample's pixels (the lane
// layout of sa_map2_kernel, 16 warps, two pixel groups per warp in flight), keeps the [mean, max] map in shared memory and,
// after a CTA barrier, finishes gate = sigmoid(conv7x7(map)) from it (the tap order of sa_gate_kernel: same bits).  The map
// never reaches global memory and the separate gate launch (10 - 16 us for 16 k - 262 k pixels) is gone.
template <int NO>
__global__ void __launch_bounds__(512, 2)      // two CTAs per SM: all 256 samples of a full batch in one wave
sa_map_gate_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, const float* __restrict__ ca,
                   const float* __restrict__ sa_w, float* __restrict__ gate, int H, int C) {
  __shared__ float2 smap[1024];
  __shared__ float w[98];
  const int HW = H * H, n = blockIdx.x;
  const int lpp = C / (8 * NO) < 32 ? C / (8 * NO) : 32, ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % lpp, pj = lane / lpp;
  if (threadIdx.x < 98) w[threadIdx.x] = sa_w[threadIdx.x];
  ldm_pdl_wait();
  float A[NO][8], Bv[NO][8];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const int c = (sub + o * lpp) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 k = __ldg(coef + (size_t)n * C + c + e);
      const float g = __ldg(ca + c + e);
      A[o][e] = g * k.x; Bv[o][e] = g * k.y;
    }
  }
  const float inv_c = 1.0f / (float)C;
  const bf16* xs = x + (size_t)n * HW * C;
  for (int i = warp * 2 * ppw; i < HW; i += 16 * 2 * ppw) {       // HW is a multiple of 2 ppw (checked by the launcher)
    uint4 u[2][NO];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int o = 0; o < NO; ++o)
        u[h][o] = __ldg(reinterpret_cast<const uint4*>(xs + (size_t)(i + h * ppw + pj) * C + (sub + o * lpp) * 8));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float sm = 0.f, m = -INFINITY;
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        float f[8];
        unpack8(u[h][o], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float z = fmaf(f[e], A[o][e], Bv[o][e]); sm += z; m = fmaxf(m, z); }
      }
      for (int o = lpp >> 1; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      }
      if (sub == 0) smap[i + h * ppw + pj] = make_float2(sm * inv_c, m);
    }
  }
  __syncthreads();
  for (int pix = threadIdx.x; pix < HW; pix += 512) {
    const int y = pix / H, xx = pix - y * H;
    float a = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
      const int yy = y + ky - 3;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int xq = xx + kx - 3;
        if (xq < 0 || xq >= H) continue;
        const float2 v = smap[yy * H + xq];
        a = fmaf(w[ky * 7 + kx], v.x, a);
        a = fmaf(w[49 + ky * 7 + kx], v.y, a);
      }
    }
    gate[(size_t)n * HW + pix] = sigmoidf_(a);
  }
}

// gate = sigmoid(conv7x7([mean, max] map)) (v2:79-80): one thread per pixel; the map of a sample is a few KiB (L1 / L2)
__global__ void __launch_bounds__(256)
sa_gate_kernel(const float* __restrict__ map, const float* __restrict__ sa_w, float* __restrict__ gate, int H, int npix) {
  ldm_pdl_wait();
  __shared__ float w[98];
  if (threadIdx.x < 98) w[threadIdx.x] = sa_w[threadIdx.x];
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int HW = H * H, n = pix / HW, rem = pix - n * HW, y = rem / H, xx = rem - y * H;
  const float2* mp = reinterpret_cast<const float2*>(map) + (size_t)n * HW;
  float a = 0.f;
  for (int ky = 0; ky < 7; ++ky) {
    const int yy = y + ky - 3;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const int xq = xx + kx - 3;
      if (xq < 0 || xq >= H) continue;
      const float2 v = __ldg(mp + yy * H + xq);
      a = fmaf(w[ky * 7 + kx], v.x, a);
      a = fmaf(w[49 + ky * 7 + kx], v.y, a);
    }
  }
  gate[pix] = sigmoidf_(a);
}

// out = swish(z * gate + resid)   (v2:81 then v2:176-177), same lane layout as sa_map2
template <int NO>
__global__ void __launch_bounds__(256)
sa_apply2_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, const float* __restrict__ ca,
                 const float* __restrict__ gate, const bf16* __restrict__ resid, bf16* __restrict__ out, int HW, int C, int npix,
                 int run) {
  ldm_pdl_wait();
  const int lpp = C / (8 * NO) < 32 ? C / (8 * NO) : 32, ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int first = wid * run;
  if (first >= npix) return;
  const int n = first / HW, sub = lane % lpp, pj = lane / lpp;
  float A[NO][8], Bv[NO][8];
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const int c = (sub + o * lpp) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 k = __ldg(coef + (size_t)n * C + c + e);
      const float g = __ldg(ca + c + e);
      A[o][e] = g * k.x; Bv[o][e] = g * k.y;
    }
  }
  for (int i = 0; i < run; i += 2 * ppw) {
    uint4 u[2][NO], r[2][NO];
    float gt[2];
    int pix[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      pix[h] = first + i + h * ppw + pj;
      const bool ok = i + h * ppw < run && pix[h] < npix;
      pix[h] = ok ? pix[h] : -1;
      gt[h] = ok ? __ldg(gate + pix[h]) : 0.f;
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const size_t off = (size_t)(ok ? pix[h] : 0) * C + (sub + o * lpp) * 8;
        u[h][o] = ok ? __ldg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0, 0, 0, 0);
        r[h][o] = ok ? __ldg(reinterpret_cast<const uint4*>(resid + off)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (pix[h] < 0) continue;
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        float f[8], rr[8];
        unpack8(u[h][o], f);
        unpack8(r[h][o], rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = swishf_bf16out(fmaf(fmaf(f[e], A[o][e], Bv[o][e]), gt[h], rr[e]));
        *reinterpret_cast<uint4*>(out + (size_t)pix[h] * C + (sub + o * lpp) * 8) = pack8(f);
      }
    }
  }
}

}  // namespace


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


// ---- bf16 fast path launchers
// splits of the pixel range: power of two, >= 4 pixels per thread, until two CTAs per SM exist.  Measured over the ten statistics
// passes of a 256-image decode (LDM_NORM_BLOCKS): 296 blocks 1.424 ms per decode, 592: 1.434, 1184: 1.474, 2368: 1.503, no
// splits: 1.455 - the ticket / partial-sum epilogue of a split costs more than the extra CTAs bring
static int norm_splits(int blocks, int HW, int pl, int C) {
  static const int want = getenv("LDM_NORM_BLOCKS") ? atoi(getenv("LDM_NORM_BLOCKS")) : 296;
  int s = 1;
  while (blocks * s < want && HW / (2 * s) >= 4 * pl && 2 * s * C <= 1024 && s < 16) s *= 2;
  return s;
}
static bool norm_wide() {
  static const bool on = !(getenv("LDM_NORM_WIDE") && atoi(getenv("LDM_NORM_WIDE")) == 0);
  return on;
}
// `part` (B x 1024 float2) and `cnt` (B x 16 ints, zero) enable the pixel splits; without them one CTA walks the whole sample
int launch_norm_coef_bf16_ws(ldm_ctx* ctx, const bf16* x, const float* gamma, const float* beta, float2* coef, int B, int HW,
                             int C, int group, float2* part, int* cnt, cudaStream_t st) {
  LDM_CHECK(C % 64 == 0 || C == 32, "norm_coef: C must be 32 or a multiple of 64 (C=%d)", C);
  LDM_CHECK(group == 1 || group == 2 || group == 4 || group == 8, "norm_coef: group size %d unsupported", group);
  if (C == 32) {
    const int S = part && cnt ? norm_splits(B, HW, 64, C) : 1;
    LDM_CUDA(launch_maybe_pdl(norm_coef2_kernel<32>, dim3(1, B, S), 256, 0, st, ctx->use_pdl, x, gamma, beta, coef, HW, C, group, part, cnt, S));
  } else if (C % 256 == 0 && HW <= 256 && B >= 128 && norm_wide()) {
    // small maps (8 x 8, 16 x 16) of a full batch: 256 channels per CTA - a quarter of the CTAs, each with two to four batches of
    // loads (decode of 256: 1.281 -> 1.273 ms; of 40: 0.452 -> 0.458, hence the batch condition)
    const int S = part && cnt ? norm_splits(B * (C / 256), HW, 8, C) : 1;
    LDM_CUDA(launch_maybe_pdl(norm_coef2_kernel<256>, dim3(C / 256, B, S), 256, 0, st, ctx->use_pdl, x, gamma, beta, coef, HW, C, group, part, cnt, S));
  } else {
    const int S = part && cnt && C / 64 <= 16 ? norm_splits(B * (C / 64), HW, 32, C) : 1;
    LDM_CUDA(launch_maybe_pdl(norm_coef2_kernel<64>, dim3(C / 64, B, S), 256, 0, st, ctx->use_pdl, x, gamma, beta, coef, HW, C, group, part, cnt, S));
  }
  LDM_LAUNCHED_AS(ctx, "launch_norm_coef_bf16");
  return 0;
}
int launch_norm_coef_bf16(ldm_ctx* ctx, const bf16* x, const float* gamma, const float* beta, float2* coef, int B, int HW,
                          int C, int group, cudaStream_t st) {
  return launch_norm_coef_bf16_ws(ctx, x, gamma, beta, coef, B, HW, C, group, nullptr, nullptr, st);
}
int launch_coef_apply_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, bf16* out, int B, int HW, int C, int act,
                           cudaStream_t st) {
  const size_t total8 = (size_t)B * HW * C / 8;
  const int C8 = C / 8, HWC8 = HW * C8;
  if (C8 >= 1 && C8 <= 256 && (C8 & (C8 - 1)) == 0 && HWC8 % (256 * kCaUnroll) == 0 && B <= 65535) {
    const int chunks = HWC8 / (256 * kCaUnroll);
    int iters = 1;      // enough blocks to fill the machine a few times over, long enough per thread to amortise the coefficient fetch
    while (iters < 8 && chunks % (2 * iters) == 0 && (long long)(chunks / (2 * iters)) * B >= 4ll * ctx->sm_count) iters *= 2;
    LDM_CUDA(launch_maybe_pdl(coef_apply_bf16_stream_kernel, dim3(chunks / iters, B), 256, 0, st, ctx->use_pdl, x, coef, out, HWC8, C8, act, iters));
  } else {
    LDM_CUDA(launch_maybe_pdl(coef_apply_bf16_kernel, dim3((unsigned)((total8 + 255) / 256)), 256, 0, st, ctx->use_pdl, x, coef, out, HWC8, C8, act, total8));
  }
  LDM_LAUNCHED(ctx);
  return 0;
}
// pixels a warp walks: a power of two dividing HW (one sample per warp), a multiple of 2 ppw, short enough that the grid
// holds a few thousand warps (a small map with 512 channels would otherwise run on 64 CTAs)
static int sa_run(int npix, int HW, int ppw) {
  int run = 32;
  while (run > HW || (run > 2 * ppw && npix / run < 4096)) run >>= 1;
  return run < 2 * ppw ? 2 * ppw : run;
}
// map + gate in one launch (sa_map_gate_kernel); returns 1 when it ran, 0 when the shape needs the two-kernel sequence
int launch_sa_map_gate_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, const float* sa_w, float* gate,
                            int B, int H, int C, cudaStream_t st, int* done) {
  static const bool on = !(getenv("LDM_DEC_FUSE_GATE") && atoi(getenv("LDM_DEC_FUSE_GATE")) == 0);
  const int HW = H * H, no = C > 256 ? 2 : 1, lpp = C / (8 * no) < 32 ? C / (8 * no) : 32;
  *done = 0;
  // one CTA per sample: only with enough samples to occupy the machine (measured: 1.305 -> 1.292 ms per decode at B = 256, bit-identical;
  // 0.310 -> 0.323 ms at B = 3, where three CTAs would walk the maps alone)
  if (!on || B < 96 || C % 64 != 0 || C > 512 || HW > 1024 || HW % (2 * (32 / lpp)) != 0) return 0;
  if (no == 2) LDM_CUDA(launch_maybe_pdl(sa_map_gate_kernel<2>, dim3(B), 512, 0, st, ctx->use_pdl, x, coef, ca, sa_w, gate, H, C));
  else LDM_CUDA(launch_maybe_pdl(sa_map_gate_kernel<1>, dim3(B), 512, 0, st, ctx->use_pdl, x, coef, ca, sa_w, gate, H, C));
  LDM_LAUNCHED_AS(ctx, "launch_sa_map_gate");
  *done = 1;
  return 0;
}
int launch_sa_map_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, float* map, int B, int HW, int C,
                       cudaStream_t st) {
  LDM_CHECK(C % 64 == 0 && C <= 512 && HW % 4 == 0, "sa_map: C %% 64 == 0, C <= 512 and HW %% 4 == 0 required (C=%d, HW=%d)", C, HW);
  const int npix = B * HW, no = C > 256 ? 2 : 1, lpp = C / (8 * no) < 32 ? C / (8 * no) : 32;
  const int run = sa_run(npix, HW, 32 / lpp), warps = ceil_div(npix, run);
  LDM_CHECK(HW % run == 0, "sa_map: HW (%d) must be a multiple of the warp run (%d)", HW, run);
  if (no == 2) LDM_CUDA(launch_maybe_pdl(sa_map2_kernel<2>, dim3(ceil_div(warps, 8)), 256, 0, st, ctx->use_pdl, x, coef, ca, map, HW, C, npix, run));
  else LDM_CUDA(launch_maybe_pdl(sa_map2_kernel<1>, dim3(ceil_div(warps, 8)), 256, 0, st, ctx->use_pdl, x, coef, ca, map, HW, C, npix, run));
  LDM_LAUNCHED(ctx);
  return 0;
}
// gate: (B, H, H) scratch of the spatial-attention gate
int launch_sa_apply_bf16(ldm_ctx* ctx, const bf16* x, const float2* coef, const float* ca, const float* map,
                         const float* sa_w, const bf16* resid, bf16* out, float* gate, int B, int H, int C, cudaStream_t st) {
  LDM_CHECK(C % 64 == 0 && C <= 512, "sa_apply: C %% 64 == 0 and C <= 512 required (C=%d)", C);
  const int HW = H * H, npix = B * HW, no = C > 256 ? 2 : 1, lpp = C / (8 * no) < 32 ? C / (8 * no) : 32;
  const int run = sa_run(npix, HW, 32 / lpp), warps = ceil_div(npix, run);
  LDM_CHECK(HW % run == 0, "sa_apply: HW (%d) must be a multiple of the warp run (%d)", HW, run);
  if (map) {      // null: the gate is already there (launch_sa_map_gate_bf16)
    LDM_CUDA(launch_maybe_pdl(sa_gate_kernel, dim3(ceil_div(npix, 256)), 256, 0, st, ctx->use_pdl, map, sa_w, gate, H, npix));
    LDM_LAUNCHED_AS(ctx, "launch_sa_gate");
  }
  if (no == 2) LDM_CUDA(launch_maybe_pdl(sa_apply2_kernel<2>, dim3(ceil_div(warps, 8)), 256, 0, st, ctx->use_pdl, x, coef, ca, (const float*)gate, resid, out, HW, C, npix, run));
  else LDM_CUDA(launch_maybe_pdl(sa_apply2_kernel<1>, dim3(ceil_div(warps, 8)), 256, 0, st, ctx->use_pdl, x, coef, ca, (const float*)gate, resid, out, HW, C, npix, run));
  LDM_LAUNCHED(ctx);
  return 0;
}

// fp32 NHWC (npix, C) -> bf16 (npix, 3 C) = [hi | lo | hi] per pixel: the operand of a strict-mode convolution
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, bf16* __restrict__ out, size_t total4, int C4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const size_t pix = i / C4;
  const int c4 = (int)(i - pix * C4);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat162 hi[2], lo[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const bf16 h0 = __float2bfloat16_rn(f[2 * e]), h1 = __float2bfloat16_rn(f[2 * e + 1]);
    hi[e] = __halves2bfloat162(h0, h1);
    lo[e] = __halves2bfloat162(__float2bfloat16_rn(f[2 * e] - __bfloat162float(h0)), __float2bfloat16_rn(f[2 * e + 1] - __bfloat162float(h1)));
  }
  const uint2 H = make_uint2(*reinterpret_cast<uint32_t*>(&hi[0]), *reinterpret_cast<uint32_t*>(&hi[1]));
  const uint2 Lo = make_uint2(*reinterpret_cast<uint32_t*>(&lo[0]), *reinterpret_cast<uint32_t*>(&lo[1]));
  bf16* o = out + pix * (size_t)(12 * C4) + (size_t)c4 * 4;
  *reinterpret_cast<uint2*>(o) = H;
  *reinterpret_cast<uint2*>(o + 4 * C4) = Lo;
  *reinterpret_cast<uint2*>(o + 8 * C4) = H;
}
int launch_split3(ldm_ctx* ctx, const float* x, bf16* out, size_t npix, int C, cudaStream_t st) {
  LDM_CHECK(C % 4 == 0, "split3: C %% 4 == 0 required");
  const size_t total4 = npix * (size_t)(C / 4);
  split3_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(x, out, total4, C / 4);
  LDM_LAUNCHED(ctx);
  return 0;
}
// (rows, taps * Cin) fp32 -> (rows, taps * 3 Cin) bf16: per tap [hi | hi | lo]
__global__ void pack_conv_split_kernel(const float* __restrict__ w, bf16* __restrict__ out, size_t total, int Cin) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t rt = i / Cin;            // (row, tap)
  const int ci = (int)(i - rt * Cin);
  const float v = w[i];
  const bf16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  bf16* o = out + rt * (size_t)(3 * Cin) + ci;
  o[0] = hi; o[Cin] = hi; o[2 * Cin] = lo;
}
int launch_pack_conv_split(ldm_ctx* ctx, const float* w, bf16* out, int rows, int taps, int Cin, cudaStream_t st) {
  const size_t total = (size_t)rows * taps * Cin;
  pack_conv_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, out, total, Cin);
  LDM_LAUNCHED(ctx);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// final_conv[1..4] as ONE pass (v2:274-278): GroupNorm(8, 32) apply + Swish -> Conv2d(32, 3, 3, padding 1) -> Sigmoid,
// raw conv output NHWC bf16 (B, H, W, 32) in, NCHW fp32 image out.  The separate apply pass re-wrote and re-read the
// 67 MB tensor, and the one-thread-per-pixel convolution behind it sat on the fp32 FMA rate (2 592 FMA per pixel = 73 us
// per 256 images on 148 SMs).  Here a CTA normalises a 16 x 32 pixel tile with its halo WHILE loading it (the zero
// padding applies to the activated tensor: out-of-image slots stay 0), keeps it in shared memory channel-octet major
// (an 8-pixel x 8-channel ldmatrix tile is 128 contiguous bytes for every tap shift) and runs the convolution as
// warp-level mma.sync m16n8k16 (rows = 16 neighbouring pixels, k = 16 channels of one tap).  N = 3 is no tcgen05 shape
// (at N = 16 its rate is set by the 4 KB pixel operand per MMA, no better than this).  The weights stay at fp32 precision as
// a bf16 (hi, lo) pair that shares ONE MMA: columns 0-2 of the n = 8 tile are the hi parts of the three output channels,
// columns 3-5 the lo parts, and the epilogue adds column c + 3 to column c.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kF3TW = 32, kF3TH = 16, kF3HW = kF3TW + 2, kF3HH = kF3TH + 2, kF3HP = kF3HW * kF3HH;
constexpr int kF3Frag = 9 * 2 * 2;     // B fragments: [tap][k half][b0 | b1], one 32-bit word per lane each
__global__ void __launch_bounds__(256, 3)
final_gn_conv3_kernel(const bf16* __restrict__ x, const float2* __restrict__ coef, const uint32_t* __restrict__ wf /* [36][32] */,
                      const float* __restrict__ bias, float* __restrict__ out, int H, int W) {
  __shared__ uint4 tile[4][kF3HP];                // activated halo tile, [channel octet][slot]
  const int n = blockIdx.z, y0 = blockIdx.y * kF3TH, x0 = blockIdx.x * kF3TW, HW = H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- weight fragments (final_w_frag_kernel, pack time) straight into registers: 36 independent coalesced loads
  uint32_t bfr[kF3Frag];
#pragma unroll
  for (int f = 0; f < kF3Frag; ++f) bfr[f] = __ldg(wf + f * 32 + lane);
  ldm_pdl_wait();      // the fragments are pack-time constants: fetched before the predecessor's results are awaited
  // ---- halo tile: GroupNorm apply + Swish on the way in (the arithmetic of coef_apply_bf16_stream_kernel)
  {
    const int j = threadIdx.x & 3;      // channel octet of this thread (256 is a multiple of 4)
    float sc[8], sh[8];
    const float4* cf = reinterpret_cast<const float4*>(coef + (size_t)n * 32 + j * 8);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 k = __ldg(cf + e);
      sc[2 * e] = k.x; sh[2 * e] = k.y; sc[2 * e + 1] = k.z; sh[2 * e + 1] = k.w;
    }
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)n * HW * 32);
    // 4 * 612 = 2448 sixteen-byte words: ten per thread (the last one partial), five loads in flight at a time
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint4 u[5];
      int hpv[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int i = threadIdx.x + (half * 5 + q) * 256;
        const int hp = i >> 2, hy = hp / kF3HW, hx = hp - hy * kF3HW;
        const int yy = y0 + hy - 1, xx = x0 + hx - 1;
        const bool in = i < 4 * kF3HP && yy >= 0 && yy < H && xx >= 0 && xx < W;
        hpv[q] = i < 4 * kF3HP ? (in ? hp : -1 - hp) : INT_MIN;
        u[q] = in ? __ldcs(src + ((size_t)yy * W + xx) * 4 + j) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        if (hpv[q] == INT_MIN) continue;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (hpv[q] >= 0) {
          float f[8];
          unpack8(u[q], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = swishf_bf16out(f[e] * sc[e] + sh[e]);
          v = pack8(f);
        }
        tile[j][hpv[q] >= 0 ? hpv[q] : -1 - hpv[q]] = v;
      }
    }
  }
  __syncthreads();
  // ---- warp `warp` finishes tile rows 2 warp, 2 warp + 1: four m16 tiles (row, x half)
  float acc[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
  const int lm = lane >> 3, lr = lane & 7;     // ldmatrix: this lane supplies row lr of 8 x 8 matrix lm
  const uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(&tile[0][0]);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int slot0 = (2 * warp + (m >> 1) + tap / 3) * kF3HW + (m & 1) * 16 + tap % 3;
        const uint32_t addr = tile_a + (uint32_t)(((2 * kc + (lm >> 1)) * kF3HP + slot0 + (lm & 1) * 8 + lr) * 16);
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(acc[m][0]), "+f"(acc[m][1]), "+f"(acc[m][2]), "+f"(acc[m][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfr[(tap * 2 + kc) * 2]), "r"(bfr[(tap * 2 + kc) * 2 + 1]));
      }
    }
  }
  // ---- accumulator fragment: (row g, columns 2t, 2t + 1) and (row g + 8, ...); channel c = column c (hi) + column c + 3 (lo).
  //      Lane t of a quad ends up with channel t of its two pixels (t = 3: idle).
  const int g = lane >> 2, t = lane & 3, qb = lane & ~3;
  const float bs = t < 3 ? __ldg(bias + t) : 0.f;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int y = y0 + 2 * warp + (m >> 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float ce = acc[m][2 * h], co = acc[m][2 * h + 1];
      const float c0 = __shfl_sync(0xffffffffu, ce, qb), c1 = __shfl_sync(0xffffffffu, co, qb);
      const float c2 = __shfl_sync(0xffffffffu, ce, qb + 1), c3 = __shfl_sync(0xffffffffu, co, qb + 1);
      const float c4 = __shfl_sync(0xffffffffu, ce, qb + 2), c5 = __shfl_sync(0xffffffffu, co, qb + 2);
      const float v = t == 0 ? c0 + c3 : (t == 1 ? c1 + c4 : c2 + c5);
      const int xq = x0 + (m & 1) * 16 + g + 8 * h;
      if (t < 3 && y < H && xq < W) out[(size_t)n * 3 * HW + (size_t)t * HW + (size_t)y * W + xq] = sigmoidf_(v + bs);
    }
  }
}
// B fragments of the 3 x (9 * 32) fp32 weights for final_gn_conv3_kernel: [tap][k half][b0 | b1][lane]; lane (g = lane / 4,
// t = lane % 4) holds column g, rows k = 2t, 2t + 1 (b0) and k = 2t + 8, 2t + 9 (b1): g < 3 the bf16 hi part of output
// channel g, 3 <= g < 6 the lo part (w - hi) of channel g - 3, zero above
__global__ void final_w_frag_kernel(const float* __restrict__ w, uint32_t* __restrict__ wf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kF3Frag * 32) return;
  const int l = i & 31, f = i >> 5, reg = f & 1, kc = (f >> 1) & 1, tap = f >> 2;
  const int g = l >> 2, t = l & 3;
  uint32_t v = 0;
  if (g < 6) {
    const float* wp = w + (size_t)(g % 3) * 288 + tap * 32 + kc * 16 + reg * 8 + 2 * t;
    const float w0 = wp[0], w1 = wp[1];
    bf16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
    if (g >= 3) { h0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)); h1 = __float2bfloat16_rn(w1 - __bfloat162float(h1)); }
    __nv_bfloat162 h2 = __halves2bfloat162(h0, h1);
    v = *reinterpret_cast<uint32_t*>(&h2);
  }
  wf[i] = v;
}
int launch_final_w_frag(ldm_ctx* ctx, const float* w, uint32_t* wf, cudaStream_t st) {
  final_w_frag_kernel<<<ceil_div(kF3Frag * 32, 256), 256, 0, st>>>(w, wf);
  LDM_LAUNCHED(ctx);
  return 0;
}
int launch_final_gn_conv3(ldm_ctx* ctx, const bf16* x, const float2* coef, const uint32_t* wf, const float* bias, float* out, int B,
                          int H, int W, cudaStream_t st) {
  LDM_CHECK(((uintptr_t)x & 15) == 0 && wf != nullptr, "final_gn_conv3: input must be 16-byte aligned, weights packed");
  LDM_CUDA(launch_maybe_pdl(final_gn_conv3_kernel, dim3(ceil_div(W, kF3TW), ceil_div(H, kF3TH), B), 256, 0, st, ctx->use_pdl, x, coef, wf, bias, out, H, W));
  LDM_LAUNCHED_AS(ctx, "final_gn_conv3");
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
