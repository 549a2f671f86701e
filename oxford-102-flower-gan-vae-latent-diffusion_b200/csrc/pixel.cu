// v4 / v5 pixel-space diffusion: SimpleUNet.forward (v4:99-135, v5:101-146) and DiffusionModel.p_sample /
// sample (v4:155-175) over NHWC bf16 activations.
//
//   x (fp32 NCHW state) --pix_conv_in--> conv1.0+ReLU --conv_tc--> ... 15 implicit-GEMM layers ... --pix_conv_out--> eps
//
// * every Conv2d(3x3), the two Conv2d(4, stride 2) and the two ConvTranspose2d(4, 2, 1) with 64..512 channels run on the
//   tensor cores (conv_tc.cu: TMA boxes as the im2col, tcgen05 / TMEM); bias, ReLU and the per-stage time term
//   (x_i = conv_i(x) + t_emb_i, added AFTER the ReLU, v4:114,118,122) are its epilogue;
// * torch.cat([x5, x2]) / torch.cat([x6, x1]) (v4:127,131) cost nothing: the producers write channel slices of one
//   concat buffer (output pitch) and the stride-2 convolutions read their slice back through a pitched TMA map;
// * the time path (Linear(1,128) on the RAW timestep -> ReLU -> Linear -> time_fc1..3, v4:103-110) does not depend on x:
//   it is a (n_t, 7 base) table built at pack time for the sampler, one tiny kernel for forward(x, t) with per-sample t;
// * conv1.0 (3 input channels, K = 27) and out_conv (3 output channels) are no tensor-core shapes: CUDA cores, fused with
//   the NCHW fp32 <-> NHWC bf16 layout change and, in the sampler, with the posterior update and its Philox noise, so
//   eps never goes to memory.
#include "common.cuh"
#include "philox.cuh"
#include "pix_out.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>

int convt_halo_supported(int H, int W, int Cin, int Cout);
int launch_convt_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int relu, cudaStream_t st);
int launch_conv_tc_ex(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int mode, int relu, const float* post, int post_stride, cudaStream_t st);
int tc_make_weight_map(ldm_ctx* ctx, const bf16* w, int N, int K, int bn, CUtensorMap* out);
int tc_init(ldm_ctx* ctx);
int conv_halo_supported(int H, int W, int Cin, int Cout);
int conv_halo_stream_supported(int H, int W, int Cin, int Cout);
int launch_conv_halo_stream(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const CUtensorMap& map_w, const float* bias,
                            bf16* out, int out_pitch, int B, int H, int W, int relu, const float* post, int post_stride, cudaStream_t st);
int launch_conv_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                     int B, int H, int W, int relu, const float* post, int post_stride, const PixOutArgs* fin, int ddpm,
                     cudaStream_t st);

namespace {

#define HW_MULT4(H, W) ((((H) * (W)) & 3) == 0)

// ------------------------------------------------------------------------------------------------------------------
// time terms: out[row] = [time_fc1 | time_fc2 | time_fc3](Linear2(relu(Linear1(t))))     (v4:103-110)
// t == nullptr: t = row index (the sampler's table).  One block per row.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pix_time_terms_kernel(const float* __restrict__ t, const float* __restrict__ w0, const float* __restrict__ b0,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ wf1,
                      const float* __restrict__ bf1, const float* __restrict__ wf2, const float* __restrict__ bf2,
                      const float* __restrict__ wf3, const float* __restrict__ bf3, float* __restrict__ out, int temb, int base) {
  extern __shared__ float sm[];
  float* h = sm;            // [temb]
  float* te = sm + temb;    // [temb]
  const int row = blockIdx.x;
  const float tv = t ? t[row] : (float)row;
  for (int i = threadIdx.x; i < temb; i += blockDim.x) h[i] = fmaxf(w0[i] * tv + b0[i], 0.f);
  __syncthreads();
  for (int i = threadIdx.x; i < temb; i += blockDim.x) {
    float acc = 0.f;
    const float* wr = w2 + (size_t)i * temb;
    for (int j = 0; j < temb; ++j) acc += wr[j] * h[j];
    te[i] = acc + b2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 7 * base; c += blockDim.x) {
    const float* wr;
    float bias;
    if (c < base) { wr = wf1 + (size_t)c * temb; bias = bf1[c]; }
    else if (c < 3 * base) { wr = wf2 + (size_t)(c - base) * temb; bias = bf2[c - base]; }
    else { wr = wf3 + (size_t)(c - 3 * base) * temb; bias = bf3[c - 3 * base]; }
    float acc = 0.f;
    for (int j = 0; j < temb; ++j) acc += wr[j] * te[j];
    out[(size_t)row * 7 * base + c] = acc + bias;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// conv1.0 + ReLU (v4:55-56): x (B, 3, H, W) fp32 NCHW -> (B, H, W, C) bf16 NHWC.  One thread per (4 pixels of a row, 8
// output channels): the 3 x 6 x 3 input window sits in registers and every weight fetched from shared memory feeds four
// pixels (the kernel is bound by shared-memory wavefronts otherwise).  Weights are k-major in shared memory, and the 8
// channels of group g are stored as two 4-float runs at g * 4 and C / 2 + g * 4 so that the 8 lanes of a quarter warp
// read one contiguous 128-byte line per 16-byte load (no bank conflicts).  W % 4 == 0.
// ------------------------------------------------------------------------------------------------------------------
template <typename TOUT>
__global__ void __launch_bounds__(256)
pix_conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /* (C, 27) */, const float* __restrict__ b,
                   TOUT* __restrict__ out, int H, int W, int C, int total_quads) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;             // [27][C] permuted
  float* bs = sm + C * 27;    // [C]
  const int half = C >> 1;
  for (int i = threadIdx.x; i < C * 27; i += blockDim.x) {
    const int k = i / C, pos = i - k * C;
    const int hi = pos >= half, r = pos - hi * half, g = r >> 2, j = (r & 3) + 4 * hi;     // channel g * 8 + j
    ws[i] = __ldg(w + (g * 8 + j) * 27 + k);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = b[i];
  __syncthreads();
  const int groups = C >> 3;
  const long long total = (long long)total_quads * groups;
#pragma unroll 1
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
  const int quad = (int)(gid / groups), g = (int)(gid - (long long)quad * groups);
  const int qpr = W >> 2, HW = H * W;
  const int row = quad / qpr, x0 = (quad - row * qpr) << 2;
  const int n = row / H, y = row - n * H;
  float in[3][6][3];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
#pragma unroll
    for (int cx = 0; cx < 6; ++cx) {
      const int xx = x0 + cx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) in[ky][cx][ci] = ok ? __ldg(x + ((size_t)n * 3 + ci) * HW + (size_t)yy * W + xx) : 0.f;
    }
  }
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = bs[g * 8 + j];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* wr = ws + ((ky * 3 + kx) * 3 + ci) * C;
        const float4 w0 = *reinterpret_cast<const float4*>(wr + g * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + half + g * 4);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float v = in[ky][p + kx][ci];
          acc[p][0] += v * w0.x; acc[p][1] += v * w0.y; acc[p][2] += v * w0.z; acc[p][3] += v * w0.w;
          acc[p][4] += v * w1.x; acc[p][5] += v * w1.y; acc[p][6] += v * w1.z; acc[p][7] += v * w1.w;
        }
      }
  const size_t p0 = (size_t)n * HW + (size_t)y * W + x0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if constexpr (sizeof(TOUT) == 4) {      // strict fp32 path
      float4* d4 = reinterpret_cast<float4*>(out + (p0 + p) * C + g * 8);
      d4[0] = make_float4(fmaxf(acc[p][0], 0.f), fmaxf(acc[p][1], 0.f), fmaxf(acc[p][2], 0.f), fmaxf(acc[p][3], 0.f));
      d4[1] = make_float4(fmaxf(acc[p][4], 0.f), fmaxf(acc[p][5], 0.f), fmaxf(acc[p][6], 0.f), fmaxf(acc[p][7], 0.f));
    } else {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(acc[p][2 * j], 0.f), fmaxf(acc[p][2 * j + 1], 0.f));
        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
      }
      reinterpret_cast<uint4*>(out + (p0 + p) * C)[g] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// conv1.0 + ReLU on the tensor cores (C == 64): the im2col row of a pixel is only 27 values, so the 128 builder threads
// write it straight into shared memory in the 128-byte-swizzled K-major operand layout - as bf16 HIGH and LOW parts
// (x = hi + lo, |lo| <= 2^-9 |x|), K = 27 + 27 (+ 10 zero columns) = 64, against the bf16 weights repeated for both
// parts - and ONE UMMA 128 x 64 x 64 per 128 pixels does the arithmetic the CUDA-core kernel needs 1700 FMAs per pixel
// for.  The input therefore keeps ~17 bits of mantissa; the weights are bf16 like every other layer's.
//   warp 0: weight TMA, TMEM, MMA issue        warps 1-4: build the A rows, then the epilogue (ReLU, bf16 NHWC store)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160)
pix_conv_in_tc_kernel(const __grid_constant__ CUtensorMap map_w, const float* __restrict__ x, const float* __restrict__ bias,
                      bf16* __restrict__ out, int H, int W, int total_pix) {
  __shared__ __align__(1024) uint8_t a_s[128 * 128];
  __shared__ __align__(1024) uint8_t w_s[64 * 128];
  __shared__ __align__(8) uint64_t bar_w, bar_d;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&map_w);
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_d, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(&bar_w, 64 * 128);
    tc::tma_load_2d(w_s, &map_w, &bar_w, 0, 0);
  }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  if (threadIdx.x >= 32 && threadIdx.x < 96) bias_s[threadIdx.x - 32] = bias[threadIdx.x - 32];
  ldm_pdl_wait();
  if (warp >= 1) {
    const int q = warp & 3, r = q * 32 + lane;          // TMEM lane = tile row this thread owns
    const int p = blockIdx.x * 128 + r;
    const int HW = H * W;
    uint32_t hi[14], lo[14];                           // 27 values as bf16 pairs (the 28th is zero)
    float v[28];
    v[27] = 0.f;
    if (p < total_pix) {
      const int n = p / HW, rem = p - n * HW, y = rem / W, xx = rem - y * W;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
        const bool ok = yy >= 0 && yy < H && xc >= 0 && xc < W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) v[tap * 3 + ci] = ok ? __ldg(x + ((size_t)n * 3 + ci) * HW + (size_t)yy * W + xc) : 0.f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 27; ++k) v[k] = 0.f;
    }
    // K layout: [hi_0..hi_26 | lo_0..lo_26 | 0 x 10]; element k sits at byte 2k of the 128-byte row
    __nv_bfloat16 e[64];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      e[k] = __float2bfloat16_rn(v[k]);
      e[27 + k] = __float2bfloat16_rn(v[k] - __bfloat162float(e[k]));
    }
#pragma unroll
    for (int k = 54; k < 64; ++k) e[k] = __float2bfloat16_rn(0.f);
    (void)hi; (void)lo;
    uint8_t* row = a_s + r * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h2 = __halves2bfloat162(e[c * 8 + 2 * j], e[c * 8 + 2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    tc::fence_proxy_async();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, 64);
      if (tc::mbar_wait(&bar_w, 0, 41)) {
        tc::fence_after_sync();
        const uint64_t da = tc::make_desc_sw128(tc::smem_u32(a_s)), dw = tc::make_desc_sw128(tc::smem_u32(w_s));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem_base, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
      }
      tc::umma_commit(&bar_d);
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const int p = blockIdx.x * 128 + r;
    tc::mbar_wait(&bar_d, 0, 42);
    tc::fence_after_sync();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)p * 64);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float acc[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
      if (p < total_pix) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(acc[2 * j] + bias_s[c0 + 2 * j], 0.f), fmaxf(acc[2 * j + 1] + bias_s[c0 + 2 * j + 1], 0.f));
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        dst[c0 / 8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[c0 / 8 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tmem_base);
}

// (C, 27) fp32 -> (64, 64) bf16: [w | w | 0] for the [hi | lo | 0] K layout of pix_conv_in_tc_kernel
__global__ void pix_pack_in_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64) return;
  const int co = i >> 6, k = i & 63;
  out[i] = __float2bfloat16_rn(k < 27 ? w[co * 27 + k] : (k < 54 ? w[co * 27 + k - 27] : 0.f));
}

// The posterior update of p_sample alone (v4:159-168) on the flattened (B, 3 H W) state: four consecutive elements are
// one Philox counter (oracle/philox.py); seed / first sample index are read from device memory so that a captured
// graph replays with new seeds.
__global__ void __launch_bounds__(256)
pix_ddpm_kernel(float* __restrict__ x, const float* __restrict__ eps, float c2, float sqrt_alpha, float sigma,
                const float* __restrict__ noise, const unsigned long long* __restrict__ rng, int step, int total4, int d4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int row = i / d4, q = i - row * d4;
  float4 xv = reinterpret_cast<float4*>(x)[i];
  const float4 e = reinterpret_cast<const float4*>(eps)[i];
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sigma > 0.0f) {
    if (noise) z = reinterpret_cast<const float4*>(noise)[i];
    else z = philox_normal4(rng[0], rng[1] + (unsigned long long)row, (uint32_t)step, (uint32_t)q);
  }
  xv.x = ddpm_update_one(xv.x, e.x, c2, sqrt_alpha, sigma, z.x);
  xv.y = ddpm_update_one(xv.y, e.y, c2, sqrt_alpha, sigma, z.y);
  xv.z = ddpm_update_one(xv.z, e.z, c2, sqrt_alpha, sigma, z.z);
  xv.w = ddpm_update_one(xv.w, e.w, c2, sqrt_alpha, sigma, z.w);
  reinterpret_cast<float4*>(x)[i] = xv;
}

// ------------------------------------------------------------------------------------------------------------------
// out_conv (v4:96,134) [+ res_ratio * x_input, v5:144], one thread per pixel over NHWC bf16, three fp32 outputs.
// DDPM == 0: eps (B, 3, H, W) fp32 NCHW is stored.   DDPM == 1: the posterior update of p_sample (v4:159-168) is applied
// to the state x in place with explicit or in-kernel Philox noise; eps is never stored.
// ------------------------------------------------------------------------------------------------------------------
// C == 64, W % 32 == 0: one warp per 32-pixel row segment, lane = channel pair (2 lane, 2 lane + 1).  The lane's 54
// weights stay in registers for the whole kernel, a pixel is ONE coalesced 128-byte load per tap row (three loads per
// column, each column feeds its three neighbouring outputs), and the 32 x 3 per-lane partial sums are reduced across the
// warp with a transposing shuffle tree (31 shuffles per 32 values), leaving pixel x0 + lane on lane `lane`.
template <int DDPM>
__global__ void __launch_bounds__(128)
pix_conv_out64_kernel(const PixOutArgs a) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (seg >= (a.total_pix >> 5)) return;
  const int H = a.H, W = a.W, HW = H * W, spr = W >> 5;
  const int row = seg / spr, x0 = (seg - row * spr) << 5;
  const int n = row / H, y = row - n * H;
  float w[3][9][2];
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(a.w + o * 576 + tap * 64) + lane);
      w[o][tap][0] = t.x; w[o][tap][1] = t.y;
    }
  float acc[3][32];
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[o][i] = 0.f;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(a.in) + (size_t)n * HW * 32 + lane;
#pragma unroll
  for (int j = 0; j < 34; ++j) {          // input column x0 - 1 + j feeds outputs j - 2, j - 1, j (kx = 2, 1, 0)
    const int xx = x0 - 1 + j;
    float v[3][2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      const bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H;
      const uint32_t u = ok ? __ldg(base + ((size_t)yy * W + xx) * 32) : 0u;
      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&u);
      v[ky][0] = __low2float(h2); v[ky][1] = __high2float(h2);
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int p = j - kx;
      if (p < 0 || p >= 32) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int o = 0; o < 3; ++o)
          acc[o][p] += v[ky][0] * w[o][ky * 3 + kx][0] + v[ky][1] * w[o][ky * 3 + kx][1];
    }
  }
  float eps[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = upper ? acc[o][i] : acc[o][i + off];
        const float keep = upper ? acc[o][i + off] : acc[o][i];
        acc[o][i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    eps[o] = acc[o][0] + __ldg(a.bias + o);
  }
  const int rem = y * W + x0 + lane;
  float z[3] = {0.f, 0.f, 0.f};
  if (DDPM && a.sigma > 0.f) {
    if (a.noise) {
#pragma unroll
      for (int c = 0; c < 3; ++c) z[c] = a.noise[(size_t)n * 3 * HW + (size_t)c * HW + rem];
    } else {
      // the 4 lanes of a pixel quad share the Philox counter of each channel: lane (q, c) draws channel c for all four
      const int c_own = (lane & 3) < 3 ? (lane & 3) : 0;
      const uint32_t e = (uint32_t)(c_own * HW + (rem & ~3));
      const float4 z4 = philox_normal4(a.rng[0], a.rng[1] + (unsigned long long)n, (uint32_t)a.step, e >> 2);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int src = (lane & ~3) + c;
        const float t0 = __shfl_sync(0xffffffffu, z4.x, src), t1 = __shfl_sync(0xffffffffu, z4.y, src);
        const float t2 = __shfl_sync(0xffffffffu, z4.z, src), t3 = __shfl_sync(0xffffffffu, z4.w, src);
        const int jj = lane & 3;
        z[c] = jj == 0 ? t0 : (jj == 1 ? t1 : (jj == 2 ? t2 : t3));
      }
    }
  }
  pix_finish<DDPM>(a, eps, n, rem, HW, z);
}

// generic shape: one thread per pixel, weights in shared memory
template <int DDPM>
__global__ void __launch_bounds__(128)
pix_conv_out_kernel(const PixOutArgs a) {
  extern __shared__ float ws[];   // [3 * 9 * C]
  const int C = a.C, K = 9 * C;
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) ws[i] = a.w[i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.total_pix) return;
  const int H = a.H, W = a.W, HW = H * W, n = p / HW, rem = p - n * HW, y = rem / W, x = rem - y * W;
  float acc0 = a.bias[0], acc1 = a.bias[1], acc2 = a.bias[2];
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const uint4* src = reinterpret_cast<const uint4*>(a.in + ((size_t)n * HW + (size_t)yy * W + xx) * C);
    const float* w0 = ws + tap * C;
    for (int j = 0; j < C / 8; ++j) {
      const uint4 u = __ldg(src + j);
      const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&wd[e]);
        const float f0 = __low2float(h2), f1 = __high2float(h2);
        const int k = j * 8 + e * 2;
        acc0 += f0 * w0[k] + f1 * w0[k + 1];
        acc1 += f0 * w0[K + k] + f1 * w0[K + k + 1];
        acc2 += f0 * w0[2 * K + k] + f1 * w0[2 * K + k + 1];
      }
    }
  }
  float eps[3] = {acc0, acc1, acc2};
  float z[3] = {0.f, 0.f, 0.f};
  if (DDPM && a.sigma > 0.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (a.noise) z[c] = a.noise[(size_t)n * 3 * HW + (size_t)c * HW + rem];
      else {
        const uint32_t e = (uint32_t)(c * HW + rem);
        const float4 z4 = philox_normal4(a.rng[0], a.rng[1] + (unsigned long long)n, (uint32_t)a.step, e >> 2);
        const int j = e & 3;
        z[c] = j == 0 ? z4.x : (j == 1 ? z4.y : (j == 2 ? z4.z : z4.w));
      }
    }
  }
  pix_finish<DDPM>(a, eps, n, rem, HW, z);
}

// ------------------------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------------------------
int own(ldm_ctx* ctx, std::vector<void*>& pool, const float* src, size_t n, float** out, cudaStream_t st) {
  LDM_CHECK(src != nullptr, "ldm_pix_pack: null weight pointer");
  LDM_TRY(ldm_alloc_t(ctx, pool, out, n));
  LDM_CUDA(cudaMemcpyAsync(*out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int pick_bn(int Cout) { return Cout >= 256 ? 256 : Cout; }

// Conv2d (Cout, Cin, k, k) -> (Cout, k*k*Cin) tap-major bf16 + TMA map
int pack_conv(ldm_ctx* ctx, std::vector<void*>& P, ConvLayer& L, const ldm_pix_conv& c, int Cout, int Cin, int k, cudaStream_t st) {
  LDM_CHECK(c.w && c.b, "ldm_pix_pack: convolution weights missing");
  LDM_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "ldm_pix_pack: channel counts must be multiples of 64 (Cin=%d, Cout=%d)", Cin, Cout);
  L.Cin = Cin; L.Cout = Cout; L.taps = k * k; L.bn = pick_bn(Cout);
  const size_t n = (size_t)Cout * k * k * Cin;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w32, n));
  LDM_TRY(launch_pack_conv(ctx, c.w, L.w32, Cout, Cin, k, k, st));
  LDM_TRY(own(ctx, P, c.b, Cout, &L.b, st));
  if (ctx->precision != LDM_PRECISION_BF16) return 0;      // strict path: fp32 weights only
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w16, n));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, n, st));
  if (L.bn > 128) {
    L.bn_alt = 128;
    LDM_TRY(tc_make_weight_map(ctx, L.w16, Cout, L.taps * Cin, L.bn_alt, &L.map_w_alt));
  }
  return tc_make_weight_map(ctx, L.w16, Cout, L.taps * Cin, L.bn, &L.map_w);
}

// ConvTranspose2d (Cin, Cout, 4, 4) -> four sub-pixel 2x2-tap kernels stacked along the output-channel axis
int pack_convT(ldm_ctx* ctx, std::vector<void*>& P, ConvLayer& L, const ldm_pix_conv& c, int Cin, int Cout, cudaStream_t st) {
  LDM_CHECK(c.w && c.b, "ldm_pix_pack: transposed-convolution weights missing");
  LDM_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "ldm_pix_pack: channel counts must be multiples of 64 (Cin=%d, Cout=%d)", Cin, Cout);
  L.Cin = Cin; L.Cout = Cout; L.taps = 4; L.bn = pick_bn(Cout);
  const size_t per = (size_t)Cout * 4 * Cin;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w32, 4 * per));
  for (int z = 0; z < 4; ++z) LDM_TRY(launch_pack_convT(ctx, c.w, L.w32 + (size_t)z * per, Cin, Cout, z >> 1, z & 1, st));
  LDM_TRY(own(ctx, P, c.b, Cout, &L.b, st));
  if (ctx->precision != LDM_PRECISION_BF16) return 0;
  LDM_TRY(ldm_alloc_t(ctx, P, &L.w16, 4 * per));
  LDM_TRY(launch_to_bf16(ctx, L.w32, L.w16, 4 * per, st));
  return tc_make_weight_map(ctx, L.w16, 4 * Cout, 4 * Cin, L.bn, &L.map_w);
}

int launch_time_terms(ldm_ctx* ctx, const float* t, float* out, int rows, cudaStream_t st) {
  PixModel& M = ctx->pix;
  pix_time_terms_kernel<<<rows, 128, 2 * M.temb * sizeof(float), st>>>(t, M.te0_w, M.te0_b, M.te2_w, M.te2_b, M.tfc_w[0], M.tfc_b[0],
                                                                         M.tfc_w[1], M.tfc_b[1], M.tfc_w[2], M.tfc_b[2], out, M.temb, M.base);
  LDM_LAUNCHED_AS(ctx, "pix_time_terms");
  return 0;
}

int ensure_pix_workspace(ldm_ctx* ctx, int B, int H, int W) {
  PixModel& M = ctx->pix;
  if (B <= M.cap && H == M.cap_h && W == M.cap_w) return 0;
  LDM_CUDA(cudaDeviceSynchronize());
  for (auto& kv : M.graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
  }
  M.graphs.clear();
  for (void* p : M.ws) cudaFree(p);
  M.ws.clear();
  M.cap = 0;
  const size_t es = ctx->precision == LDM_PRECISION_BF16 ? 1 : 2;      // strict path: the same buffers hold fp32 activations
  const size_t c = M.base * es, p1 = (size_t)B * H * W, p2 = p1 / 4, p3 = p1 / 16;
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.a1, p1 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.cat5, p1 * 2 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.d1, p2 * 2 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.a2, p2 * 2 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.cat4, p2 * 4 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.d2, p3 * 4 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.a3, p3 * 4 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.x3, p3 * 4 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.bt, p3 * 8 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.x4, p3 * 4 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.a4, p2 * 2 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.x5, p2 * 2 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.a5, p1 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.x6, p1 * c));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.tsample, (size_t)B * 7 * M.base));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.x_state, p1 * 3));
  LDM_TRY(ldm_alloc_t(ctx, M.ws, &M.eps, p1 * 3));
  M.cap = B; M.cap_h = H; M.cap_w = W;
  return 0;
}

int use_halo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_PIX_HALO");
    // bit 0: resident-weight halo kernel (Cin = 64).  bit 1: streamed-weight halo kernel (Cin > 64) - parity-green but
    // slower than conv_tc_kernel (measured at B = 64: conv2.x 36 vs 30 us, conv4.0 64 vs 47, conv5.0 112 vs 79): one
    // persistent CTA per SM keeps only ~5 weight tiles (40-80 KB) in flight against ~1 us of L2 latency, two conv_tc
    // CTAs per SM keep ~200 KB.  Off by default.
    v = e ? atoi(e) : 1;
  }
  return v;
}

// 3x3 convolution + bias + ReLU [+ time term]: the halo kernel where it applies (Cin = 64 on a large image), else conv_tc
int conv3(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, bf16* out, int out_pitch, int B, int H, int W,
          const float* post, int post_stride, cudaStream_t st) {
  if (use_halo() && conv_halo_supported(H, W, L.Cin, L.Cout))
    return launch_conv_halo(ctx, in, in_pitch, L, L.b, out, out_pitch, B, H, W, 1, post, post_stride, nullptr, 0, st);
  if ((use_halo() & 2) && L.bn == L.Cout && conv_halo_stream_supported(H, W, L.Cin, L.Cout))
    return launch_conv_halo_stream(ctx, in, in_pitch, L, L.map_w, L.b, out, out_pitch, B, H, W, 1, post, post_stride, st);
  return launch_conv_tc_ex(ctx, in, in_pitch, L, L.b, out, out_pitch, B, H, W, 1, 1, post, post_stride, st);
}

// ------------------------------------------------------------------------------------------------------------------
// strict fp32 path (eps within 1e-3 of the reference): the same layer sequence on the CUDA-core implicit-GEMM kernel
// (gemm_f32.cu: conv_f32_kernel), NHWC fp32 activations in the same workspace buffers
// ------------------------------------------------------------------------------------------------------------------
__global__ void pix_axpy_kernel(float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __fadd_rn(y[i], __fmul_rn(__ldg(a), x[i]));
}

int conv_f32(ldm_ctx* ctx, const float* in, int in_pitch, const ConvLayer& L, const float* w, float* out, int out_pitch, int B, int H, int W,
             int mode /* 1: 3x3, 0: 4x4 stride 2 (H, W = input size), 2: transposed, parity pa/pb */, int pa, int pb, int relu,
             const float* post, int post_stride, cudaStream_t st) {
  ConvGeom g = ConvGeom();
  g.B = B; g.Cin = L.Cin; g.Cout = L.Cout; g.in_pitch = in_pitch; g.out_pitch = out_pitch; g.relu = relu; g.post = post; g.post_stride = post_stride;
  g.up = 1;
  if (mode == 1) {
    g.H = H; g.W = W; g.taps = 9;
    for (int t = 0; t < 9; ++t) { g.dy[t] = t / 3 - 1; g.dx[t] = t % 3 - 1; }
  } else if (mode == 0) {
    g.H = H / 2; g.W = W / 2; g.taps = 16; g.stride = 2;
    for (int t = 0; t < 16; ++t) { g.dy[t] = t / 4 - 1; g.dx[t] = t % 4 - 1; }
  } else {
    g.H = H; g.W = W; g.taps = 4; g.up = 2; g.pa = pa; g.pb = pb;
    for (int t = 0; t < 4; ++t) {
      g.dy[t] = (t >> 1) == 0 ? 0 : (pa == 0 ? -1 : 1);
      g.dx[t] = (t & 1) == 0 ? 0 : (pb == 0 ? -1 : 1);
    }
  }
  return launch_conv_f32(ctx, in, w, L.b, out, g, st);
}

int run_forward_f32(ldm_ctx* ctx, const float* x, const float* terms, int tstride, int B, int H, int W, PixOutArgs fin, int ddpm,
                    cudaStream_t st) {
  PixModel& M = ctx->pix;
  const int c = M.base, H2 = H / 2, H4 = H / 4, W2 = W / 2, W4 = W / 4, P1 = B * H * W;
  float *a1 = (float*)M.a1, *cat5 = (float*)M.cat5, *d1 = (float*)M.d1, *a2 = (float*)M.a2, *cat4 = (float*)M.cat4, *d2 = (float*)M.d2,
        *a3 = (float*)M.a3, *x3 = (float*)M.x3, *bt = (float*)M.bt, *x4 = (float*)M.x4, *a4 = (float*)M.a4, *x5 = (float*)M.x5,
        *a5 = (float*)M.a5, *x6 = (float*)M.x6;
  {
    const long long items = (long long)(P1 / 4) * (c / 8);
    const long long want = (items + 255) / 256;
    const unsigned grid = (unsigned)(want < 2ll * ctx->sm_count ? want : 2ll * ctx->sm_count);
    pix_conv_in_kernel<float><<<grid, 256, (size_t)c * 28 * sizeof(float), st>>>(x, M.in_w, M.in_b, a1, H, W, c, P1 / 4);
    LDM_LAUNCHED_AS(ctx, "pix_conv_in");
  }
  const float *t1 = terms, *t2 = terms + c, *t3 = terms + 3 * c;
  auto c3 = [&](const float* in, int ip, const ConvLayer& L, float* out, int op, int h, int w, const float* post) {
    return conv_f32(ctx, in, ip, L, L.w32, out, op, B, h, w, 1, 0, 0, 1, post, post ? tstride : 0, st);
  };
  auto up = [&](const float* in, int ip, const ConvLayer& L, float* out, int op, int h, int w) {
    const size_t per = (size_t)L.Cout * 4 * L.Cin;
    for (int z = 0; z < 4; ++z) LDM_TRY(conv_f32(ctx, in, ip, L, L.w32 + (size_t)z * per, out, op, B, h, w, 2, z >> 1, z & 1, 0, nullptr, 0, st));
    return 0;
  };
  LDM_TRY(c3(a1, c, M.c1b, cat5 + c, 2 * c, H, W, t1));
  LDM_TRY(conv_f32(ctx, cat5 + c, 2 * c, M.down1, M.down1.w32, d1, 2 * c, B, H, W, 0, 0, 0, 0, nullptr, 0, st));
  LDM_TRY(c3(d1, 2 * c, M.c2a, a2, 2 * c, H2, W2, nullptr));
  LDM_TRY(c3(a2, 2 * c, M.c2b, cat4 + 2 * c, 4 * c, H2, W2, t2));
  LDM_TRY(conv_f32(ctx, cat4 + 2 * c, 4 * c, M.down2, M.down2.w32, d2, 4 * c, B, H2, W2, 0, 0, 0, 0, nullptr, 0, st));
  LDM_TRY(c3(d2, 4 * c, M.c3a, a3, 4 * c, H4, W4, nullptr));
  LDM_TRY(c3(a3, 4 * c, M.c3b, x3, 4 * c, H4, W4, t3));
  LDM_TRY(c3(x3, 4 * c, M.b0, bt, 8 * c, H4, W4, nullptr));
  LDM_TRY(c3(bt, 8 * c, M.b2, x4, 4 * c, H4, W4, nullptr));
  LDM_TRY(up(x4, 4 * c, M.up1, cat4, 4 * c, H4, W4));
  LDM_TRY(c3(cat4, 4 * c, M.c4a, a4, 2 * c, H2, W2, nullptr));
  LDM_TRY(c3(a4, 2 * c, M.c4b, x5, 2 * c, H2, W2, nullptr));
  LDM_TRY(up(x5, 2 * c, M.up2, cat5, 2 * c, H2, W2));
  LDM_TRY(c3(cat5, 2 * c, M.c5a, a5, c, H, W, nullptr));
  LDM_TRY(c3(a5, c, M.c5b, x6, c, H, W, nullptr));
  // out_conv -> eps (B, 3, H, W) [+ res_ratio * x]; the sampler then applies the posterior update
  float* eps = ddpm ? M.eps : fin.out;
  {
    ConvGeom g = ConvGeom();
    g.B = B; g.H = H; g.W = W; g.Cin = c; g.Cout = 3; g.taps = 9; g.up = 1; g.nchw_out = 1;
    for (int t = 0; t < 9; ++t) { g.dy[t] = t / 3 - 1; g.dx[t] = t % 3 - 1; }
    LDM_TRY(launch_conv_f32(ctx, x6, M.out_w, M.out_b, eps, g, st));
  }
  const size_t n3 = (size_t)P1 * 3;
  if (M.res_ratio) {
    pix_axpy_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(eps, M.res_ratio, x, n3);
    LDM_LAUNCHED_AS(ctx, "pix_axpy");
  }
  if (ddpm) {
    const int total4 = (int)(n3 / 4);
    pix_ddpm_kernel<<<ceil_div(total4, 256), 256, 0, st>>>(fin.out, eps, fin.c2, fin.sqrt_alpha, fin.sigma, fin.noise, fin.rng, fin.step,
                                                           total4, 3 * H * W / 4);
    LDM_LAUNCHED_AS(ctx, "pix_ddpm");
  }
  return 0;
}

// One forward over the workspace.  terms: (rows, 7 base) time terms, row stride tstride (0: one row for the batch).
// fin: what out_conv does with eps.
int run_forward(ldm_ctx* ctx, const float* x, const float* terms, int tstride, int B, int H, int W, PixOutArgs fin, int ddpm,
                cudaStream_t st) {
  if (ctx->precision != LDM_PRECISION_BF16) return run_forward_f32(ctx, x, terms, tstride, B, H, W, fin, ddpm, st);
  PixModel& M = ctx->pix;
  const int c = M.base, H2 = H / 2, H4 = H / 4, W2 = W / 2, W4 = W / 4;
  const int P1 = B * H * W;
  static int in_tc = -1;
  if (in_tc < 0) {
    const char* e = getenv("LDM_PIX_IN_TC");
    in_tc = e ? atoi(e) : 1;
  }
  if (in_tc && c == 64 && M.in_w16) {
    pix_conv_in_tc_kernel<<<ceil_div(P1, 128), 160, 0, st>>>(M.in_map, x, M.in_b, M.a1, H, W, P1);
  } else
  {   // persistent blocks (two per SM): the weights are staged in shared memory once per block
    const long long items = (long long)(P1 / 4) * (c / 8);
    const long long want = (items + 255) / 256;
    const unsigned grid = (unsigned)(want < 2ll * ctx->sm_count ? want : 2ll * ctx->sm_count);
    pix_conv_in_kernel<bf16><<<grid, 256, (size_t)c * 28 * sizeof(float), st>>>(x, M.in_w, M.in_b, M.a1, H, W, c, P1 / 4);
  }
  LDM_LAUNCHED_AS(ctx, "pix_conv_in");
  const float *t1 = terms, *t2 = terms + c, *t3 = terms + 3 * c;
  // encoder (v4:113-122); x1 -> cat5[:, c:2c], x2 -> cat4[:, 2c:4c]
  LDM_TRY(conv3(ctx, M.a1, c, M.c1b, M.cat5 + c, 2 * c, B, H, W, t1, tstride, st));
  LDM_TRY(launch_conv_tc_ex(ctx, M.cat5 + c, 2 * c, M.down1, M.down1.b, M.d1, 2 * c, B, H, W, 0, 0, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.d1, 2 * c, M.c2a, M.a2, 2 * c, B, H2, W2, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.a2, 2 * c, M.c2b, M.cat4 + 2 * c, 4 * c, B, H2, W2, t2, tstride, st));
  LDM_TRY(launch_conv_tc_ex(ctx, M.cat4 + 2 * c, 4 * c, M.down2, M.down2.b, M.d2, 4 * c, B, H2, W2, 0, 0, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.d2, 4 * c, M.c3a, M.a3, 4 * c, B, H4, W4, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.a3, 4 * c, M.c3b, M.x3, 4 * c, B, H4, W4, t3, tstride, st));
  // bottleneck (v4:124)
  LDM_TRY(conv3(ctx, M.x3, 4 * c, M.b0, M.bt, 8 * c, B, H4, W4, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.bt, 8 * c, M.b2, M.x4, 4 * c, B, H4, W4, nullptr, 0, st));
  // decoder (v4:126-133)
  LDM_TRY(launch_conv_tc_ex(ctx, M.x4, 4 * c, M.up1, M.up1.b, M.cat4, 4 * c, B, H4, W4, 2, 0, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.cat4, 4 * c, M.c4a, M.a4, 2 * c, B, H2, W2, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.a4, 2 * c, M.c4b, M.x5, 2 * c, B, H2, W2, nullptr, 0, st));
  if (convt_halo_supported(H2, W2, 2 * c, c) && M.up2.bn == 64)      // 128 -> 64: resident weights, every pixel run loaded once
    LDM_TRY(launch_convt_halo(ctx, M.x5, 2 * c, M.up2, M.up2.b, M.cat5, 2 * c, B, H2, W2, 0, st));
  else
    LDM_TRY(launch_conv_tc_ex(ctx, M.x5, 2 * c, M.up2, M.up2.b, M.cat5, 2 * c, B, H2, W2, 2, 0, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.cat5, 2 * c, M.c5a, M.a5, c, B, H, W, nullptr, 0, st));
  LDM_TRY(conv3(ctx, M.a5, c, M.c5b, M.x6, c, B, H, W, nullptr, 0, st));
  fin.in = M.x6; fin.w = M.out_w; fin.bias = M.out_b; fin.res_ratio = M.res_ratio; fin.x_in = x;
  fin.H = H; fin.W = W; fin.C = c; fin.total_pix = P1;
  if (use_halo() && M.out16.w16 && conv_halo_supported(H, W, c, 16)) {
    // out_conv on the tensor cores; in the sampler eps goes through a (B, 3, H, W) fp32 buffer to a full-occupancy
    // update kernel (the four epilogue warps of a persistent tensor-core CTA are too few for the Philox arithmetic:
    // fused, the kernel took 57 us per step at B = 64 against 25 us with a plain store)
    PixOutArgs f2 = fin;
    if (ddpm) f2.out = M.eps;
    LDM_TRY(launch_conv_halo(ctx, M.x6, c, M.out16, M.out_b, nullptr, 0, B, H, W, 0, nullptr, 0, &f2, 0, st));
    if (ddpm) {
      const int total4 = P1 * 3 / 4;
      pix_ddpm_kernel<<<ceil_div(total4, 256), 256, 0, st>>>(fin.out, M.eps, fin.c2, fin.sqrt_alpha, fin.sigma, fin.noise, fin.rng,
                                                             fin.step, total4, 3 * H * W / 4);
      LDM_LAUNCHED_AS(ctx, "pix_ddpm");
    }
    return 0;
  }
  if (c == 64 && W % 32 == 0 && HW_MULT4(H, W)) {
    const int blocks = ceil_div(P1 / 32, 4);
    if (ddpm) pix_conv_out64_kernel<1><<<blocks, 128, 0, st>>>(fin);
    else pix_conv_out64_kernel<0><<<blocks, 128, 0, st>>>(fin);
  } else {
    const size_t smem = (size_t)27 * c * sizeof(float);
    if (ddpm) pix_conv_out_kernel<1><<<ceil_div(P1, 128), 128, smem, st>>>(fin);
    else pix_conv_out_kernel<0><<<ceil_div(P1, 128), 128, smem, st>>>(fin);
  }
  LDM_LAUNCHED_AS(ctx, "pix_conv_out");
  return 0;
}

int check_shape(int B, int H, int W) {
  LDM_CHECK(B > 0 && H >= 8 && W >= 8 && H % 4 == 0 && W % 4 == 0, "pixel path: need batch > 0 and H, W multiples of 4 (got %d x %d x %d)", B, H, W);
  LDM_CHECK((long long)B * H * W * 8 < (1ll << 31), "pixel path: batch %d at %dx%d exceeds the 32-bit pixel index range", B, H, W);
  return 0;
}

int run_loop(ldm_ctx* ctx, int B, int H, int W, int t_start, int t_end, const float* noise, cudaStream_t st) {
  PixModel& M = ctx->pix;
  const size_t slab = (size_t)B * 3 * H * W;
  for (int t = t_start, j = 0; t >= t_end; --t, ++j) {
    PixOutArgs f = {};
    f.out = M.x_state;
    f.noise = noise ? noise + (size_t)j * slab : nullptr;
    f.rng = ctx->rng_dev;
    f.c2 = ctx->c2[t]; f.sqrt_alpha = ctx->sqrt_alpha[t]; f.sigma = ctx->sigma[t]; f.step = t;
    LDM_TRY(run_forward(ctx, M.x_state, M.tab + (size_t)t * 7 * M.base, 0, B, H, W, f, 1, st));
  }
  return 0;
}

}  // namespace

void pix_drop_graphs(ldm_ctx* ctx) {
  for (auto& kv : ctx->pix.graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
  }
  ctx->pix.graphs.clear();
}

void pix_free(ldm_ctx* ctx) {
  pix_drop_graphs(ctx);
  for (void* p : ctx->pix.ws) cudaFree(p);
  for (void* p : ctx->pix.allocs) cudaFree(p);
  ctx->pix.ws.clear();
  ctx->pix.allocs.clear();
}

extern "C" LDM_API int ldm_pix_pack(ldm_ctx* ctx, const ldm_pix_weights* w, void* stream) {
  LDM_CHECK(ctx && w, "ldm_pix_pack: null argument");
  LDM_CHECK(w->in_channels == 3, "ldm_pix_pack: in_channels must be 3 (got %d)", w->in_channels);
  LDM_CHECK(w->base_channels >= 64 && w->base_channels % 64 == 0 && w->base_channels <= 256,
            "ldm_pix_pack: base_channels must be 64, 128, 192 or 256 (got %d)", w->base_channels);
  LDM_CHECK(w->time_emb_dim > 0 && w->time_emb_dim <= 4096 && w->n_t > 0, "ldm_pix_pack: bad time_emb_dim / n_t");
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  if (ctx->precision == LDM_PRECISION_BF16) LDM_TRY(tc_init(ctx));
  LDM_CUDA(cudaDeviceSynchronize());
  pix_free(ctx);
  ctx->pix = PixModel();
  PixModel& M = ctx->pix;
  auto& P = M.allocs;
  const int c = w->base_channels, te = w->time_emb_dim;
  M.base = c; M.temb = te; M.n_t = w->n_t;
  LDM_TRY(own(ctx, P, w->time_embed0_w, te, &M.te0_w, st));
  LDM_TRY(own(ctx, P, w->time_embed0_b, te, &M.te0_b, st));
  LDM_TRY(own(ctx, P, w->time_embed2_w, (size_t)te * te, &M.te2_w, st));
  LDM_TRY(own(ctx, P, w->time_embed2_b, te, &M.te2_b, st));
  const int tc[3] = {c, 2 * c, 4 * c};
  for (int i = 0; i < 3; ++i) {
    LDM_TRY(own(ctx, P, w->time_fc_w[i], (size_t)tc[i] * te, &M.tfc_w[i], st));
    LDM_TRY(own(ctx, P, w->time_fc_b[i], tc[i], &M.tfc_b[i], st));
  }
  if (w->res_ratio) LDM_TRY(own(ctx, P, w->res_ratio, 1, &M.res_ratio, st));
  // conv1.0 and out_conv: tap-major fp32 for the CUDA-core kernels
  LDM_CHECK(w->conv1[0].w && w->conv1[0].b && w->out_conv.w && w->out_conv.b, "ldm_pix_pack: conv1.0 / out_conv weights missing");
  LDM_TRY(ldm_alloc_t(ctx, P, &M.in_w, (size_t)c * 27));
  LDM_TRY(launch_pack_conv(ctx, w->conv1[0].w, M.in_w, c, 3, 3, 3, st));
  LDM_TRY(own(ctx, P, w->conv1[0].b, c, &M.in_b, st));
  if (c == 64 && ctx->precision == LDM_PRECISION_BF16) {
    LDM_TRY(ldm_alloc_t(ctx, P, &M.in_w16, (size_t)64 * 64));
    pix_pack_in_tc_kernel<<<16, 256, 0, st>>>(M.in_w, M.in_w16);
    LDM_LAUNCHED_AS(ctx, "pix_pack_in_tc");
    LDM_TRY(tc_make_weight_map(ctx, M.in_w16, 64, 64, 64, &M.in_map));
  }
  LDM_TRY(ldm_alloc_t(ctx, P, &M.out_w, (size_t)3 * 9 * c));
  LDM_TRY(launch_pack_conv(ctx, w->out_conv.w, M.out_w, 3, c, 3, 3, st));
  LDM_TRY(own(ctx, P, w->out_conv.b, 3, &M.out_b, st));
  if (c == 64 && ctx->precision == LDM_PRECISION_BF16) {   // out_conv on the tensor cores: N padded 3 -> 16 with zero rows (conv_halo_kernel, BN = 16)
    M.out16.Cin = c; M.out16.Cout = 16; M.out16.taps = 9; M.out16.bn = 16;
    LDM_TRY(ldm_alloc_t(ctx, P, &M.out16.w16, (size_t)16 * 9 * c));
    LDM_CUDA(cudaMemsetAsync(M.out16.w16, 0, (size_t)16 * 9 * c * sizeof(bf16), st));
    LDM_TRY(launch_to_bf16(ctx, M.out_w, M.out16.w16, (size_t)3 * 9 * c, st));
    LDM_TRY(tc_make_weight_map(ctx, M.out16.w16, 16, 9 * c, 16, &M.out16.map_w));
  }
  LDM_TRY(pack_conv(ctx, P, M.c1b, w->conv1[1], c, c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.down1, w->down1, 2 * c, c, 4, st));
  LDM_TRY(pack_conv(ctx, P, M.c2a, w->conv2[0], 2 * c, 2 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.c2b, w->conv2[1], 2 * c, 2 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.down2, w->down2, 4 * c, 2 * c, 4, st));
  LDM_TRY(pack_conv(ctx, P, M.c3a, w->conv3[0], 4 * c, 4 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.c3b, w->conv3[1], 4 * c, 4 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.b0, w->bottleneck[0], 8 * c, 4 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.b2, w->bottleneck[1], 4 * c, 8 * c, 3, st));
  LDM_TRY(pack_convT(ctx, P, M.up1, w->up1, 4 * c, 2 * c, st));
  LDM_TRY(pack_conv(ctx, P, M.c4a, w->conv4[0], 2 * c, 4 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.c4b, w->conv4[1], 2 * c, 2 * c, 3, st));
  LDM_TRY(pack_convT(ctx, P, M.up2, w->up2, 2 * c, c, st));
  LDM_TRY(pack_conv(ctx, P, M.c5a, w->conv5[0], c, 2 * c, 3, st));
  LDM_TRY(pack_conv(ctx, P, M.c5b, w->conv5[1], c, c, 3, st));
  // the sampler's time terms for t = 0 .. n_t-1
  LDM_TRY(ldm_alloc_t(ctx, P, &M.tab, (size_t)M.n_t * 7 * c));
  LDM_TRY(launch_time_terms(ctx, nullptr, M.tab, M.n_t, st));
  LDM_CUDA(cudaStreamSynchronize(st));
  M.packed = true;
  return 0;
}

extern "C" LDM_API int ldm_pix_forward(ldm_ctx* ctx, const float* x_dev, const float* t_dev, float* eps_out_dev, int batch, int H, int W,
                                       void* stream) {
  LDM_CHECK(ctx && x_dev && t_dev && eps_out_dev, "ldm_pix_forward: null argument");
  LDM_CHECK(ctx->pix.packed, "ldm_pix_forward: model not packed (ldm_pix_pack)");
  LDM_TRY(check_shape(batch, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_pix_workspace(ctx, batch, H, W));
  PixModel& M = ctx->pix;
  LDM_TRY(launch_time_terms(ctx, t_dev, M.tsample, batch, st));
  PixOutArgs f = {};
  f.out = eps_out_dev;
  return run_forward(ctx, x_dev, M.tsample, 7 * M.base, batch, H, W, f, 0, st);
}

extern "C" LDM_API int ldm_pix_sample(ldm_ctx* ctx, float* x_inout, int t_start, int t_end, const float* noise, uint64_t seed,
                                      uint64_t sample_offset, int batch, int H, int W, int use_graph, void* stream) {
  LDM_CHECK(ctx && x_inout, "ldm_pix_sample: null argument");
  LDM_CHECK(ctx->pix.packed, "ldm_pix_sample: model not packed (ldm_pix_pack)");
  LDM_CHECK(ctx->n_steps > 0, "ldm_pix_sample: schedule not set (ldm_set_schedule)");
  LDM_CHECK(t_end >= 0 && t_start >= t_end && t_start < ctx->n_steps, "ldm_pix_sample: need 0 <= t_end <= t_start < n_steps (%d), got %d..%d",
            ctx->n_steps, t_start, t_end);
  LDM_CHECK(ctx->n_steps <= ctx->pix.n_t, "ldm_pix_sample: time table has %d rows but the schedule has %d steps", ctx->pix.n_t, ctx->n_steps);
  LDM_TRY(check_shape(batch, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  LDM_CUDA(cudaSetDevice(ctx->device));
  LDM_TRY(ensure_pix_workspace(ctx, batch, H, W));
  PixModel& M = ctx->pix;
  const size_t xbytes = (size_t)batch * 3 * H * W * sizeof(float);
  LDM_CUDA(cudaMemcpyAsync(M.x_state, x_inout, xbytes, cudaMemcpyDeviceToDevice, st));
  LDM_TRY(launch_set_rng(ctx, ctx->rng_dev, seed, sample_offset, st));
  if (!use_graph) {
    LDM_TRY(run_loop(ctx, batch, H, W, t_start, t_end, noise, st));
  } else {
    const auto key = std::make_tuple(batch, H, W, t_start, t_end, noise ? 1 : 0);
    auto it = M.graphs.find(key);
    if (it != M.graphs.end() && it->second.noise != noise) {
      cudaGraphExecDestroy(it->second.exec);
      cudaGraphDestroy(it->second.graph);
      M.graphs.erase(it);
      it = M.graphs.end();
    }
    if (it == M.graphs.end()) {
      GraphEntry ge;
      const unsigned long long before = ctx->launches;
      LDM_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal));
      ctx->capturing = true;
      int r = run_loop(ctx, batch, H, W, t_start, t_end, noise, ctx->cap_stream);
      ctx->capturing = false;
      cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &ge.graph);
      if (r != 0) { if (ge.graph) cudaGraphDestroy(ge.graph); return r; }
      LDM_CUDA(ce);
      ge.n_nodes = (size_t)(ctx->launches - before);
      ctx->launches = before;
      LDM_CUDA(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
      ge.noise = noise;
      it = M.graphs.emplace(key, ge).first;
    }
    LDM_CUDA(cudaGraphLaunch(it->second.exec, st));
    ctx->launches += it->second.n_nodes;
  }
  LDM_CUDA(cudaMemcpyAsync(x_inout, M.x_state, xbytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}
