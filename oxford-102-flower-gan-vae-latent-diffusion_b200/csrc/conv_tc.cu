// Implicit-GEMM convolution on the sm_100a tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
//   out[pix, co] = bias[co] + sum_{tap, ci} in[pix + off(tap), ci] * w[co][tap * Cin + ci]
//
// covers the 3x3 convolutions of ResidualBlock (v2:163,165) and final_conv[0] (v2:273) and, as four sub-pixel
// 2x2-tap convolutions (one per output parity, blockIdx.z), ConvTranspose2d(4, 2, 1) of up3/up2/up1 (v2:256,262,268).
//
// Activations are NHWC bf16.  A CTA owns 128 consecutive output pixels (a bn x bh x bw box of the (N, H, W) pixel
// grid) and BN output channels.  There is no im2col buffer: the A operand of tap (dy, dx) and channel block c0 is
// ONE 4-D TMA box (64 channels, bw, bh, bn) of the input at coordinates (c0, dx, y0 + dy, n0); rows and columns that
// fall outside the image are zero-filled by the TMA unit, which is exactly the zero padding of the convolution.  The
// box lands in shared memory as 128 rows of 128 bytes with the 128-byte swizzle, i.e. the K-major UMMA operand
// layout.  Weights (Cout, taps * Cin) are a plain 2-D TMA box (64, BN).
//
//   warp 0 : TMA producer (mbarrier ring)        warp 1 : TMEM allocator + tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2-5 : epilogue (tcgen05.ld, + bias, bf16 pack, 16-byte stores; sub-pixel scatter for the transposed conv)
#include "common.cuh"
#include "tc_ptx.cuh"

int tc_init(ldm_ctx* ctx);

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvTcArgs {
  int H, W;         // input spatial size
  int Cin, Cout;    // Cout: output channels of one parity
  int taps;         // 9 (3x3), 4 (sub-pixel of the transposed conv) or 16 (4x4 stride-2 conv)
  int up;           // 1: 3x3; 2: sub-pixel of ConvTranspose2d(4,2,1); 0: Conv2d(4, stride 2, pad 1) - H, W are then the OUTPUT size
  int total_pix;    // B * H * W pixels of the grid the CTAs tile
  int stages;
  int in_pitch;     // channel pitch of the input buffer (>= Cin: the input may be a channel slice of a concat buffer)
  int out_pitch;    // channel pitch of the output buffer
  int relu;         // ReLU after the bias
  const float* post;   // optional per-channel term added AFTER the activation (v4:114: x1 = conv1(x) + t_emb1), row n * post_stride
  int post_stride;     // 0: one row for the whole batch
  const float* bias;
  bf16* out;        // (B, up*H, up*W, out_pitch), already offset to the first output channel
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ConvTcArgs a) {
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kABytes = BM * BK * 2, kWBytes = BN * BK * 2, kStageBytes = kABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, nc0 = blockIdx.y * BN, z = blockIdx.z;
  const int pa = z >> 1, pb = z & 1;
  const int HW = a.H * a.W;
  const int n0 = m0 / HW, y0 = (m0 - n0 * HW) / a.W;
  const int cblocks = a.Cin / BK;
  const int nkb = a.taps * cblocks, S = a.stages;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_a);
    tc::prefetch_tmap(&map_w);
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<kTmemCols>(&tmem_slot);
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_s[i] = a.bias ? a.bias[nc0 + i] : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        if (!tc::mbar_wait(&empty_bar[s], ph ^ 1u, 1)) break;
        const int tap = kb / cblocks, cb = kb - tap * cblocks;
        int dy, dx;
        if (a.up == 0) {
          // Conv2d(4, stride 2, pad 1): input row 2Y + ky - 1 = 2 (Y + dy) + p with ky -> (dy, p) = (-1,1), (0,0), (0,1), (1,0);
          // the 5-D map splits rows and columns into (index / 2, parity), so a tap is again ONE box of 128 output pixels
          const int ky = tap >> 2, kx = tap & 3;
          dy = ky == 0 ? -1 : (ky == 3 ? 1 : 0);
          dx = kx == 0 ? -1 : (kx == 3 ? 1 : 0);
          const int py = (ky == 0 || ky == 2) ? 1 : 0, px = (kx == 0 || kx == 2) ? 1 : 0;
          uint8_t* sa = smem + (size_t)s * kStageBytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          tc::tma_load_5d(sa, &map_a, &full_bar[s], px * a.in_pitch + cb * BK, dx, py, y0 + dy, n0);
          tc::tma_load_2d(sa + kABytes, &map_w, &full_bar[s], tap * a.Cin + cb * BK, nc0);
          continue;
        } else if (a.up == 1) {
          dy = tap / 3 - 1;
          dx = tap - (tap / 3) * 3 - 1;
        } else {   // sub-pixel (pa, pb) of ConvTranspose2d(4, 2, 1): tap 0 reads offset 0, tap 1 reads -1 (parity 0) or +1
          dy = (tap >> 1) == 0 ? 0 : (pa == 0 ? -1 : 1);
          dx = (tap & 1) == 0 ? 0 : (pb == 0 ? -1 : 1);
        }
        uint8_t* sa = smem + (size_t)s * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        tc::tma_load_4d(sa, &map_a, &full_bar[s], cb * BK, dx, y0 + dy, n0);
        tc::tma_load_2d(sa + kABytes, &map_w, &full_bar[s], tap * a.Cin + cb * BK, z * a.Cout + nc0);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
      bool ok = true;
      for (int kb = 0; kb < nkb && ok; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        ok = tc::mbar_wait(&full_bar[s], ph, 2);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t da = tc::make_desc_sw128(a_addr), dw = tc::make_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          tc::umma_bf16(tmem_base, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int p = m0 + q * 32 + lane;
    const bool valid = p < a.total_pix;
    size_t op = (size_t)p;
    if (a.up == 2) {
      const int n = p / HW, rem = p - n * HW, y = rem / a.W, x = rem - y * a.W;
      op = ((size_t)n * (2 * a.H) + (size_t)(2 * y + pa)) * (size_t)(2 * a.W) + (size_t)(2 * x + pb);
    }
    bf16* dst = a.out + op * (size_t)a.out_pitch + nc0;
    const float* prow = a.post ? a.post + (size_t)(valid ? p / HW : 0) * a.post_stride + nc0 : nullptr;
    tc::mbar_wait(&tmem_full_bar, 0, 3);
    tc::fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] += bias_s[c0 + j];
          if (a.relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (prow) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(prow + c0) + j);
            v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
        d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

// final_conv[3] + Sigmoid (v2:277-278): Conv2d(32, 3, 3, padding 1) over NHWC bf16 -> NCHW fp32.  N = 3 output
// channels is no tensor-core shape: one thread per pixel on the CUDA cores, weights in shared memory.
__global__ void __launch_bounds__(256)
conv_out3_kernel(const bf16* __restrict__ in, const float* __restrict__ w /* (3, 9*32) */, const float* __restrict__ bias,
                 float* __restrict__ out, int H, int W, int total_pix) {
  __shared__ float ws[3 * 288];
  for (int i = threadIdx.x; i < 3 * 288; i += 256) ws[i] = w[i];
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= total_pix) return;
  const int HW = H * W, n = p / HW, rem = p - n * HW, y = rem / W, x = rem - y * W;
  float acc0 = bias[0], acc1 = bias[1], acc2 = bias[2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const uint4* src = reinterpret_cast<const uint4*>(in + ((size_t)n * HW + (size_t)yy * W + xx) * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = __ldg(src + j);
      const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&wd[e]);
        const float f0 = __low2float(h2), f1 = __high2float(h2);
        const int ci = j * 8 + e * 2, k = tap * 32 + ci;
        acc0 += f0 * ws[k] + f1 * ws[k + 1];
        acc1 += f0 * ws[288 + k] + f1 * ws[288 + k + 1];
        acc2 += f0 * ws[576 + k] + f1 * ws[576 + k + 1];
      }
    }
  }
  float* o = out + (size_t)n * 3 * HW + rem;
  o[0] = sigmoidf_(acc0);
  o[HW] = sigmoidf_(acc1);
  o[2 * HW] = sigmoidf_(acc2);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode4 = nullptr;
bool g_attr_set = false;

int conv_init(ldm_ctx* ctx) {
  LDM_TRY(tc_init(ctx));
  if (!g_encode4) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    LDM_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not found in the driver");
    g_encode4 = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (!g_attr_set) {
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    g_attr_set = true;
  }
  return 0;
}

// NHWC bf16 activation (B, H, W, C of pitch P) as a 4-D tensor (C, W, H, B); box = 64 channels x 128 pixels
int make_act_map(const bf16* base, int B, int H, int W, int C, int P, CUtensorMap* out) {
  LDM_CHECK(W > 0 && H > 0 && W <= 128 && 128 % W == 0, "conv_tc: spatial size %dx%d does not tile into 128-pixel boxes", H, W);
  int bw = W, bh = 128 / W, bn = 1;
  if (bh > H) { bh = H; bn = 128 / (H * W); }
  LDM_CHECK(bw * bh * bn == 128 && H % bh == 0, "conv_tc: spatial size %dx%d does not tile into 128-pixel boxes", H, W);
  LDM_CHECK(((uintptr_t)base & 15) == 0 && C % 64 == 0 && P % 8 == 0 && P >= C,
            "conv_tc: activation must be 16-byte aligned with C %% 64 == 0 (C=%d, pitch %d)", C, P);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode4(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ldm_set_error("cuTensorMapEncodeTiled (4-D activation) failed: CUresult %d (B=%d H=%d W=%d C=%d)", (int)r, B, H, W, C);
    return (int)r;
  }
  return 0;
}

// Input of Conv2d(4, stride 2, pad 1): NHWC bf16 (B, H, W, C of pitch P) as the 5-D tensor
// (px * P + c, W/2, py, H/2, B) - rows and columns split into (index / 2, parity); box = 64 channels x 128 OUTPUT pixels
int make_act_map_down(const bf16* base, int B, int H, int W, int C, int P, CUtensorMap* out) {
  LDM_CHECK(H % 2 == 0 && W % 2 == 0, "conv_tc: stride-2 convolution needs even H, W (%dx%d)", H, W);
  const int Ho = H / 2, Wo = W / 2;
  LDM_CHECK(Wo > 0 && Wo <= 128 && 128 % Wo == 0, "conv_tc: output size %dx%d does not tile into 128-pixel boxes", Ho, Wo);
  int bw = Wo, bh = 128 / Wo, bn = 1;
  if (bh > Ho) { bh = Ho; bn = 128 / (Ho * Wo); }
  LDM_CHECK(bw * bh * bn == 128 && Ho % bh == 0, "conv_tc: output size %dx%d does not tile into 128-pixel boxes", Ho, Wo);
  LDM_CHECK(((uintptr_t)base & 15) == 0 && C % 64 == 0 && P % 8 == 0 && P >= C,
            "conv_tc: activation must be 16-byte aligned with C %% 64 == 0 (C=%d, pitch %d)", C, P);
  cuuint64_t dims[5] = {(cuuint64_t)(P + C), (cuuint64_t)Wo, 2, (cuuint64_t)Ho, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)2 * P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)2 * W * P * 2, (cuuint64_t)H * W * P * 2};
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode4(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<bf16*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ldm_set_error("cuTensorMapEncodeTiled (5-D stride-2 activation) failed: CUresult %d (B=%d H=%d W=%d C=%d P=%d)", (int)r, B, H, W, C, P);
    return (int)r;
  }
  return 0;
}

template <int BN>
int launch_bn(ldm_ctx* ctx, const CUtensorMap& ma, const ConvLayer& L, const ConvTcArgs& a0, int nz, cudaStream_t st) {
  ConvTcArgs a = a0;
  const int nkb = a.taps * (a.Cin / BK);
  const size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)BN * BK * 2;
  int stages = nkb < kMaxStages ? nkb : kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes > 100 * 1024) --stages;   // two CTAs per SM: one finishes while the other loads
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  dim3 grid(ceil_div(a.total_pix, BM), a.Cout / BN, nz);
  conv_tc_kernel<BN><<<grid, kThreads, smem, st>>>(ma, L.map_w, a);
  ctx->launches++;
  LDM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int conv_tc_pick_bn(int Cout) { return Cout >= 256 ? 256 : Cout; }

// in: (B, H, W, Cin of pitch in_pitch) bf16; out: (B, up*H, up*W, Cout of pitch out_pitch) bf16 (mode 0: (B, H/2, W/2, .)).
// L.w16 / L.map_w: (nz * Cout, taps * Cin), box (64, bn).  mode = ConvTcArgs::up.
int launch_conv_tc_ex(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int mode, int relu, const float* post, int post_stride, cudaStream_t st) {
  LDM_TRY(conv_init(ctx));
  LDM_CHECK(L.w16 != nullptr, "conv_tc: layer not packed for the tensor-core path");
  LDM_CHECK(L.Cin % BK == 0, "conv_tc: Cin %% 64 == 0 required (Cin=%d)", L.Cin);
  LDM_CHECK((mode == 1 && L.taps == 9) || (mode == 2 && L.taps == 4) || (mode == 0 && L.taps == 16), "conv_tc: unsupported taps/mode combination");
  LDM_CHECK(out_pitch % 8 == 0 && ((uintptr_t)out & 15) == 0, "conv_tc: output must be 16-byte aligned (pitch %d)", out_pitch);
  CUtensorMap ma;
  ConvTcArgs a;
  if (mode == 0) {
    LDM_TRY(make_act_map_down(in, B, H, W, L.Cin, in_pitch, &ma));
    a.H = H / 2; a.W = W / 2;
  } else {
    LDM_TRY(make_act_map(in, B, H, W, L.Cin, in_pitch, &ma));
    a.H = H; a.W = W;
  }
  a.Cin = L.Cin; a.Cout = L.Cout; a.taps = L.taps; a.up = mode; a.total_pix = B * a.H * a.W; a.stages = 0;
  a.in_pitch = in_pitch; a.out_pitch = out_pitch; a.relu = relu; a.post = post; a.post_stride = post_stride;
  a.bias = bias; a.out = out;
  const int nz = mode == 2 ? 4 : 1;
  switch (L.bn ? L.bn : conv_tc_pick_bn(L.Cout)) {
    case 32: return launch_bn<32>(ctx, ma, L, a, nz, st);
    case 64: return launch_bn<64>(ctx, ma, L, a, nz, st);
    case 128: return launch_bn<128>(ctx, ma, L, a, nz, st);
    case 256: return launch_bn<256>(ctx, ma, L, a, nz, st);
  }
  ldm_set_error("conv_tc: unsupported Cout %d", L.Cout);
  return -1;
}

int launch_conv_tc(ldm_ctx* ctx, const bf16* in, const ConvLayer& L, const float* bias, bf16* out, int B, int H, int W,
                   int up, cudaStream_t st) {
  return launch_conv_tc_ex(ctx, in, L.Cin, L, bias, out, L.Cout, B, H, W, up, 0, nullptr, 0, st);
}

int launch_conv_out3(ldm_ctx* ctx, const bf16* in, const float* w, const float* bias, float* out, int B, int H, int W,
                     cudaStream_t st) {
  const int total = B * H * W;
  conv_out3_kernel<<<ceil_div(total, 256), 256, 0, st>>>(in, w, bias, out, H, W, total);
  ctx->launches++;
  LDM_CUDA(cudaGetLastError());
  return 0;
}

// barrier-timeout record of this translation unit's kernels (read-and-clear)
int conv_tc_error_flag(int* out) {
  LDM_CUDA(cudaMemcpyFromSymbol(out, g_tc_error, sizeof(int)));
  if (*out) {
    int z = 0;
    LDM_CUDA(cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int)));
  }
  return 0;
}
