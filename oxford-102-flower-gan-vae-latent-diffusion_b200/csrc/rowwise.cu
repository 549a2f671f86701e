// Memory-bound row kernels of the denoiser step: LayerNorm + Swish + residual (v2:546-549,559),
// operand staging, label validation, the stand-alone DDPM update (v2:584-592) and Philox draws.
// One warp owns one row: the row lives in registers, statistics are warp-shuffle reductions,
// global accesses are 128-bit and coalesced.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kWarpsPerCta = 4;

template <typename TOP> struct Pack4;
template <> struct Pack4<float> {
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
};
template <> struct Pack4<bf16> {
  static __device__ __forceinline__ void store(bf16* p, float a, float b, float c, float d) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(p) = pk;
  }
};

// mean and 1/sqrt(var + eps) of a row held as NV4 float4 per lane (biased variance, two-pass)
template <int NV4>
__device__ __forceinline__ void row_stats(const float4 (&v)[NV4], int d, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + 1e-5f);
}

// h2 = swish(LN_a(u)) + h ; n = LN_b(h2)           (v2:546-549)
template <int NV4, typename TOP>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
stage_mid_kernel(const float* __restrict__ u, const float* __restrict__ h, const float* __restrict__ ga,
                 const float* __restrict__ ba, const float* __restrict__ gb, const float* __restrict__ bb,
                 float* __restrict__ h2, TOP* __restrict__ n_op, int ld_op, int M, int d) {
  ldm_pdl_launch_dependents();      // PDL (no-ops without the launch attribute): let the next kernel set itself up ...
  ldm_pdl_wait();                   // ... and wait for the previous one's results
  const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float4 v[NV4];
  const float4* up = reinterpret_cast<const float4*>(u + (size_t)row * d);
#pragma unroll
  for (int j = 0; j < NV4; ++j) v[j] = up[j * 32 + lane];
  float mean, rstd;
  row_stats<NV4>(v, d, mean, rstd);
  const float4* hp = reinterpret_cast<const float4*>(h + (size_t)row * d);
  float4* h2p = reinterpret_cast<float4*>(h2 + (size_t)row * d);
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int q = j * 32 + lane;
    float4 g = reinterpret_cast<const float4*>(ga)[q], b = reinterpret_cast<const float4*>(ba)[q], r = hp[q];
    v[j].x = swishf((v[j].x - mean) * rstd * g.x + b.x) + r.x;
    v[j].y = swishf((v[j].y - mean) * rstd * g.y + b.y) + r.y;
    v[j].z = swishf((v[j].z - mean) * rstd * g.z + b.z) + r.z;
    v[j].w = swishf((v[j].w - mean) * rstd * g.w + b.w) + r.w;
    h2p[q] = v[j];
  }
  row_stats<NV4>(v, d, mean, rstd);
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int q = j * 32 + lane;
    float4 g = reinterpret_cast<const float4*>(gb)[q], b = reinterpret_cast<const float4*>(bb)[q];
    Pack4<TOP>::store(n_op + (size_t)row * ld_op + q * 4, (v[j].x - mean) * rstd * g.x + b.x,
                      (v[j].y - mean) * rstd * g.y + b.y, (v[j].z - mean) * rstd * g.z + b.z,
                      (v[j].w - mean) * rstd * g.w + b.w);
  }
}

// out = act(LN(in))   (v2:559 final_norm; also Decoder.fc LayerNorm(512), v2:247)
template <int NV4, typename TOP>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
row_ln_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ g_, const float* __restrict__ b_,
              int act, TOP* __restrict__ out, int ld_out, int M, int d) {
  ldm_pdl_launch_dependents();
  ldm_pdl_wait();
  const int row = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float4 v[NV4];
  const float4* ip = reinterpret_cast<const float4*>(in + (size_t)row * ld_in);
#pragma unroll
  for (int j = 0; j < NV4; ++j) v[j] = ip[j * 32 + lane];
  float mean, rstd;
  row_stats<NV4>(v, d, mean, rstd);
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    const int q = j * 32 + lane;
    float4 g = reinterpret_cast<const float4*>(g_)[q], b = reinterpret_cast<const float4*>(b_)[q];
    float o0 = (v[j].x - mean) * rstd * g.x + b.x, o1 = (v[j].y - mean) * rstd * g.y + b.y;
    float o2 = (v[j].z - mean) * rstd * g.z + b.z, o3 = (v[j].w - mean) * rstd * g.w + b.w;
    if (act == LDM_ACT_SWISH) { o0 = swishf(o0); o1 = swishf(o1); o2 = swishf(o2); o3 = swishf(o3); }
    Pack4<TOP>::store(out + (size_t)row * ld_out + q * 4, o0, o1, o2, o3);
  }
}

// LayerNorm over a long row (Decoder.fc LayerNorm(32768), v2:251): one CTA per row, three passes
// over an L2-resident row (sum; centred sum of squares; write).
template <typename TOP>
__global__ void __launch_bounds__(256)
row_ln_big_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ g_, const float* __restrict__ b_,
                  int act, TOP* __restrict__ out, int ld_out, int d) {
  __shared__ float red[8];
  __shared__ float bc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float4* ip = reinterpret_cast<const float4*>(in + (size_t)row * ld_in);
  const int n4 = d >> 2;
  float s = 0.f;
  ldm_pdl_wait();
  for (int q = tid; q < n4; q += 256) { float4 v = ip[q]; s += (v.x + v.y) + (v.z + v.w); }
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = t / (float)d; }
  __syncthreads();
  const float mean = bc;
  float qq = 0.f;
  for (int q = tid; q < n4; q += 256) {
    float4 v = ip[q];
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
    qq += (a * a + b * b) + (c * c + e * e);
  }
  qq = warp_sum(qq);
  __syncthreads();
  if (lane == 0) red[wid] = qq;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = 1.0f / sqrtf(t / (float)d + 1e-5f); }
  __syncthreads();
  const float rstd = bc;
  for (int q = tid; q < n4; q += 256) {
    float4 v = ip[q], g = reinterpret_cast<const float4*>(g_)[q], b = reinterpret_cast<const float4*>(b_)[q];
    float o0 = (v.x - mean) * rstd * g.x + b.x, o1 = (v.y - mean) * rstd * g.y + b.y;
    float o2 = (v.z - mean) * rstd * g.z + b.z, o3 = (v.w - mean) * rstd * g.w + b.w;
    if (act == LDM_ACT_SWISH) { o0 = swishf(o0); o1 = swishf(o1); o2 = swishf(o2); o3 = swishf(o3); }
    Pack4<TOP>::store(out + (size_t)row * ld_out + q * 4, o0, o1, o2, o3);
  }
}

template <typename TOP>
__global__ void load_x_kernel(const float* __restrict__ x, TOP* __restrict__ dst, int ld_dst, int M, int d4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * d4) return;
  const int row = i / d4, q = i - row * d4;
  float4 v = reinterpret_cast<const float4*>(x)[i];
  Pack4<TOP>::store(dst + (size_t)row * ld_dst + q * 4, v.x, v.y, v.z, v.w);
}

__global__ void set_classes_kernel(const int64_t* __restrict__ c, int32_t* __restrict__ out, int M, int ncls,
                                   int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  long long v = c[i];
  if (v < 0 || v >= ncls) { atomicOr(flags, 1); v = v < 0 ? 0 : ncls - 1; }
  out[i] = (int32_t)v;
}

// v3: condition pair index f * nk + k of every row (range-checked like nn.Embedding would)
__global__ void set_conditions_kernel(const int64_t* __restrict__ f, const int64_t* __restrict__ k, int32_t* __restrict__ out,
                                      int M, int nf, int nk, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  long long a = f[i], b = k[i];
  if (a < 0 || a >= nf) { atomicOr(flags, 1); a = a < 0 ? 0 : nf - 1; }
  if (b < 0 || b >= nk) { atomicOr(flags, 1); b = b < 0 ? 0 : nk - 1; }
  out[i] = (int32_t)(a * nk + b);
}
// v3: out[(f * nk + k)] = [flower_emb[f] | color_emb[k]]   (torch.cat of v3:748)
__global__ void cond_pairs_kernel(const float* __restrict__ fe, const float* __restrict__ ke, float* __restrict__ out, int nf,
                                  int nk, int td) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nf * nk * 2 * td) return;
  const int e = (int)(i % (2 * td));
  const int p = (int)(i / (2 * td)), a = p / nk, b = p - a * nk;
  out[i] = e < td ? fe[(size_t)a * td + e] : ke[(size_t)b * td + (e - td)];
}

// v3 attention ACROSS the batch (v3:832-835, nn.MultiheadAttention on (L = B, N = 1, E = d), eval mode):
//   out[i, h*hd:(h+1)*hd] = sum_j softmax_j(q_i . k_j / sqrt(hd)) v_j     per head h.
// qkv: (B, 3d) fp32 = [Q | K | V].  One warp per (query row, head); 32 keys per tile, lane j scores key j, online
// softmax in fp32, lanes own output dims e = lane, lane + 32, ...
template <typename TOP, int HD>
__global__ void __launch_bounds__(256)
batch_attention_kernel(const float* __restrict__ qkv, TOP* __restrict__ out, int B, int d) {
  // 8 warps x 2 query rows per block; K / V tiles of 32 keys staged in shared memory (row pitch HD + 1: lane j reads
  // key j without bank conflicts); compile-time head_dim so that the dot products unroll
  constexpr int R = 2, NC = (HD + 31) / 32;
  __shared__ float ks[32][HD + 1];
  __shared__ float vs[32][HD + 1];
  __shared__ float qs[8 * R][HD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, row0 = blockIdx.x * (8 * R) + w * R;
  const float scale = rsqrtf((float)HD);
#pragma unroll
  for (int r = 0; r < R; ++r)
    for (int e = lane; e < HD; e += 32)
      qs[w * R + r][e] = row0 + r < B ? qkv[(size_t)(row0 + r) * 3 * d + head * HD + e] * scale : 0.f;   // q * hd^-0.5 as torch does
  float m[R], l[R], o[R][NC];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) o[r][c] = 0.f;
  }
  for (int j0 = 0; j0 < B; j0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * HD; idx += 256) {
      const int j = idx / HD, e = idx - j * HD;
      const bool in = j0 + j < B;
      ks[j][e] = in ? qkv[(size_t)(j0 + j) * 3 * d + d + head * HD + e] : 0.f;
      vs[j][e] = in ? qkv[(size_t)(j0 + j) * 3 * d + 2 * d + head * HD + e] : 0.f;
    }
    __syncthreads();
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll 8
    for (int e = 0; e < HD; ++e) {
      const float kv = ks[lane][e];
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] += qs[w * R + r][e] * kv;
    }
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (j0 + lane >= B) s[r] = -INFINITY;
      const float mn = fmaxf(m[r], warp_max(s[r]));
      p[r] = __expf(s[r] - mn);
      const float corr = __expf(m[r] - mn);
      l[r] = l[r] * corr + warp_sum(p[r]);
      m[r] = mn;
#pragma unroll
      for (int c = 0; c < NC; ++c) o[r][c] *= corr;
    }
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      float pj[R];
#pragma unroll
      for (int r = 0; r < R; ++r) pj[r] = __shfl_sync(0xffffffffu, p[r], j);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float vv = (c * 32 + lane < HD) ? vs[j][c * 32 + lane] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) o[r][c] += pj[r] * vv;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= B) continue;
    const float inv = 1.0f / l[r];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c * 32 + lane < HD) out[(size_t)(row0 + r) * d + head * HD + c * 32 + lane] = from_f32<TOP>(o[r][c] * inv);
  }
}

__global__ void check_t_kernel(const int64_t* __restrict__ t, int n, int n_t, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = t[i];
  if (v < 0 || v >= n_t) atomicOr(flags, 2);
}

__global__ void ddpm_update_kernel(float* __restrict__ x, const float* __restrict__ eps, float c2, float sqrt_alpha,
                                   float sigma, const float* __restrict__ noise, unsigned long long seed,
                                   unsigned long long sample_offset, int step, int M, int d4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * d4) return;
  const int row = i / d4, q = i - row * d4;
  float4 xv = reinterpret_cast<float4*>(x)[i];
  float4 e = reinterpret_cast<const float4*>(eps)[i];
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sigma > 0.0f) {
    if (noise) z = reinterpret_cast<const float4*>(noise)[i];
    else z = philox_normal4(seed, sample_offset + (unsigned long long)row, (uint32_t)step, (uint32_t)q);
  }
  xv.x = ddpm_update_one(xv.x, e.x, c2, sqrt_alpha, sigma, z.x);
  xv.y = ddpm_update_one(xv.y, e.y, c2, sqrt_alpha, sigma, z.y);
  xv.z = ddpm_update_one(xv.z, e.z, c2, sqrt_alpha, sigma, z.z);
  xv.w = ddpm_update_one(xv.w, e.w, c2, sqrt_alpha, sigma, z.w);
  reinterpret_cast<float4*>(x)[i] = xv;
}

__global__ void randn_kernel(float* __restrict__ out, unsigned long long seed, unsigned long long sample_offset,
                             int step, int M, int d4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * d4) return;
  const int row = i / d4, q = i - row * d4;
  reinterpret_cast<float4*>(out)[i] =
      philox_normal4(seed, sample_offset + (unsigned long long)row, (uint32_t)step, (uint32_t)q);
}

__global__ void set_rng_kernel(unsigned long long* rng, unsigned long long seed, unsigned long long off) {
  rng[0] = seed;
  rng[1] = off;
}

template <typename F>
int dispatch_nv4(int d, F&& f) {
  switch (d / 128) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
    case 8: return f(std::integral_constant<int, 8>());
  }
  return -1;
}

}  // namespace


template <typename TOP>
int launch_stage_mid(ldm_ctx* ctx, const float* u, const float* h, const float* ga, const float* ba,
                     const float* gb, const float* bb, float* h2, TOP* n_op, int ld_op, int M, int d,
                     cudaStream_t st) {
  LDM_CHECK(d % 128 == 0 && d >= 128 && d <= 1024, "stage_mid: hidden dim %d must be a multiple of 128 in [128,1024]", d);
  dim3 grid(ceil_div(M, kWarpsPerCta)), block(kWarpsPerCta * 32);
  dispatch_nv4(d, [&](auto nv) {
    (void)launch_maybe_pdl(stage_mid_kernel<decltype(nv)::value, TOP>, grid, kWarpsPerCta * 32, 0, st, ctx->use_pdl, u, h, ga, ba, gb, bb, h2,
                           n_op, ld_op, M, d);      // a launch error is picked up by LDM_LAUNCHED below
    return 0;
  });
  LDM_LAUNCHED(ctx);
  return 0;
}
template int launch_stage_mid<float>(ldm_ctx*, const float*, const float*, const float*, const float*, const float*,
                                     const float*, float*, float*, int, int, int, cudaStream_t);
template int launch_stage_mid<bf16>(ldm_ctx*, const float*, const float*, const float*, const float*, const float*,
                                    const float*, float*, bf16*, int, int, int, cudaStream_t);

// row_ln_big_kernel with 1024 threads: the three passes over the L2-resident row (sum; centred sum of squares; apply) were 32
// dependent iterations of 256 threads each; here a pass is d / 4096 iterations with four independent 16-byte loads in flight
// per thread (32 registers: two CTAs per SM, a batch of 256 rows is one wave).  d % 16384 == 0.
template <typename TOP>
__global__ void __launch_bounds__(1024, 2)
row_ln_wide_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ g_, const float* __restrict__ b_,
                   int act, TOP* __restrict__ out, int ld_out, int d) {
  __shared__ float red[32];
  __shared__ float bc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float4* ip = reinterpret_cast<const float4*>(in + (size_t)row * ld_in);
  const int n4 = d >> 2;
  ldm_pdl_wait();
  float s = 0.f;
  for (int q = tid; q < n4; q += 4096) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ip[q + 1024 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (wid == 0) { const float t = warp_sum(red[lane]); if (lane == 0) bc = t / (float)d; }
  __syncthreads();
  const float mean = bc;
  float qq = 0.f;
  for (int q = tid; q < n4; q += 4096) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ip[q + 1024 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
      qq += (a * a + b * b) + (c * c + e * e);
    }
  }
  qq = warp_sum(qq);
  __syncthreads();
  if (lane == 0) red[wid] = qq;
  __syncthreads();
  if (wid == 0) { const float t = warp_sum(red[lane]); if (lane == 0) bc = 1.0f / sqrtf(t / (float)d + 1e-5f); }
  __syncthreads();
  const float rstd = bc;
  for (int q = tid; q < n4; q += 2048) {
    float4 v[2], g[2], b[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      v[j] = ip[q + 1024 * j];
      g[j] = __ldg(reinterpret_cast<const float4*>(g_) + q + 1024 * j);
      b[j] = __ldg(reinterpret_cast<const float4*>(b_) + q + 1024 * j);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float o0 = (v[j].x - mean) * rstd * g[j].x + b[j].x, o1 = (v[j].y - mean) * rstd * g[j].y + b[j].y;
      float o2 = (v[j].z - mean) * rstd * g[j].z + b[j].z, o3 = (v[j].w - mean) * rstd * g[j].w + b[j].w;
      if (act == LDM_ACT_SWISH) { o0 = swishf(o0); o1 = swishf(o1); o2 = swishf(o2); o3 = swishf(o3); }
      Pack4<TOP>::store(out + (size_t)row * ld_out + (size_t)(q + 1024 * j) * 4, o0, o1, o2, o3);
    }
  }
}

template <typename TOP>
int launch_row_ln(ldm_ctx* ctx, const float* in, int ld_in, const float* g, const float* b, int act, TOP* out,
                  int ld_out, int M, int d, cudaStream_t st) {
  LDM_CHECK(d % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0, "row_ln: dims must be multiples of 4");
  if (d % 128 == 0 && d <= 1024) {
    dim3 grid(ceil_div(M, kWarpsPerCta)), block(kWarpsPerCta * 32);
    dispatch_nv4(d, [&](auto nv) {
      (void)launch_maybe_pdl(row_ln_kernel<decltype(nv)::value, TOP>, grid, kWarpsPerCta * 32, 0, st, ctx->use_pdl, in, ld_in, g, b, act, out,
                             ld_out, M, d);
      return 0;
    });
  } else {
    static const bool reg_rows = !(getenv("LDM_ROW_LN_WIDE") && atoi(getenv("LDM_ROW_LN_WIDE")) == 0);
    if (reg_rows && d % 16384 == 0)
      LDM_CUDA(launch_maybe_pdl(row_ln_wide_kernel<TOP>, dim3(M), 1024, 0, st, ctx->use_pdl, in, ld_in, g, b, act, out, ld_out, d));
    else
      LDM_CUDA(launch_maybe_pdl(row_ln_big_kernel<TOP>, dim3(M), 256, 0, st, ctx->use_pdl, in, ld_in, g, b, act, out, ld_out, d));
  }
  LDM_LAUNCHED(ctx);
  return 0;
}
template int launch_row_ln<float>(ldm_ctx*, const float*, int, const float*, const float*, int, float*, int, int, int,
                                  cudaStream_t);
template int launch_row_ln<bf16>(ldm_ctx*, const float*, int, const float*, const float*, int, bf16*, int, int, int,
                                 cudaStream_t);

template <typename TOP>
int launch_load_x(ldm_ctx* ctx, const float* x, TOP* dst, int ld_dst, int M, int d, cudaStream_t st) {
  LDM_CHECK(d % 4 == 0, "load_x: dim must be a multiple of 4");
  const int n = M * (d / 4);
  load_x_kernel<TOP><<<ceil_div(n, 256), 256, 0, st>>>(x, dst, ld_dst, M, d / 4);
  LDM_LAUNCHED(ctx);
  return 0;
}
template int launch_load_x<float>(ldm_ctx*, const float*, float*, int, int, int, cudaStream_t);
template int launch_load_x<bf16>(ldm_ctx*, const float*, bf16*, int, int, int, cudaStream_t);

int launch_set_classes(ldm_ctx* ctx, const int64_t* c, int32_t* out, int M, int ncls, int* flags, cudaStream_t st) {
  set_classes_kernel<<<ceil_div(M, 256), 256, 0, st>>>(c, out, M, ncls, flags);
  LDM_LAUNCHED(ctx);
  return 0;
}

int launch_set_conditions(ldm_ctx* ctx, const int64_t* f, const int64_t* k, int32_t* out, int M, int nf, int nk, int* flags,
                          cudaStream_t st) {
  set_conditions_kernel<<<ceil_div(M, 256), 256, 0, st>>>(f, k, out, M, nf, nk, flags);
  LDM_LAUNCHED(ctx);
  return 0;
}
int launch_cond_pairs(ldm_ctx* ctx, const float* fe, const float* ke, float* out, int nf, int nk, int td, cudaStream_t st) {
  const size_t n = (size_t)nf * nk * 2 * td;
  cond_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fe, ke, out, nf, nk, td);
  LDM_LAUNCHED(ctx);
  return 0;
}
template <typename TOP>
int launch_batch_attention(ldm_ctx* ctx, const float* qkv, TOP* out, int B, int d, int heads, cudaStream_t st) {
  const int hd = d / heads;
  LDM_CHECK(d % heads == 0 && (hd == 16 || hd == 32 || hd == 48 || hd == 64 || hd == 96 || hd == 128),
            "batch_attention: head_dim %d unsupported (d=%d)", hd, d);
  const dim3 grid(ceil_div(B, 16), heads);
  switch (hd) {
    case 16: batch_attention_kernel<TOP, 16><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
    case 32: batch_attention_kernel<TOP, 32><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
    case 48: batch_attention_kernel<TOP, 48><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
    case 64: batch_attention_kernel<TOP, 64><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
    case 96: batch_attention_kernel<TOP, 96><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
    default: batch_attention_kernel<TOP, 128><<<grid, 256, 0, st>>>(qkv, out, B, d); break;
  }
  LDM_LAUNCHED(ctx);
  return 0;
}
template int launch_batch_attention<float>(ldm_ctx*, const float*, float*, int, int, int, cudaStream_t);
template int launch_batch_attention<bf16>(ldm_ctx*, const float*, bf16*, int, int, int, cudaStream_t);

int launch_check_t(ldm_ctx* ctx, const int64_t* t, int n, int n_t, int* flags, cudaStream_t st) {
  check_t_kernel<<<ceil_div(n, 256), 256, 0, st>>>(t, n, n_t, flags);
  LDM_LAUNCHED(ctx);
  return 0;
}

int launch_ddpm_update(ldm_ctx* ctx, float* x, const float* eps, float c2, float sqrt_alpha, float sigma,
                       const float* noise, unsigned long long seed, unsigned long long sample_offset, int step,
                       int M, int d, cudaStream_t st) {
  LDM_CHECK(d % 4 == 0, "ddpm_update: dim must be a multiple of 4");
  const int n = M * (d / 4);
  ddpm_update_kernel<<<ceil_div(n, 256), 256, 0, st>>>(x, eps, c2, sqrt_alpha, sigma, noise, seed, sample_offset,
                                                       step, M, d / 4);
  LDM_LAUNCHED(ctx);
  return 0;
}

int launch_randn(ldm_ctx* ctx, float* out, unsigned long long seed, unsigned long long sample_offset, int step,
                 int M, int d, cudaStream_t st) {
  LDM_CHECK(d % 4 == 0, "randn: dim must be a multiple of 4");
  const int n = M * (d / 4);
  randn_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, seed, sample_offset, step, M, d / 4);
  LDM_LAUNCHED(ctx);
  return 0;
}

int launch_set_rng(ldm_ctx* ctx, unsigned long long* rng, unsigned long long seed, unsigned long long sample_offset,
                   cudaStream_t st) {
  set_rng_kernel<<<1, 1, 0, st>>>(rng, seed, sample_offset);
  LDM_LAUNCHED(ctx);
  return 0;
}
