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
#include "pix_out.cuh"

#include <cstdlib>

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
  float* out32;     // strict mode: fp32 output instead (same indexing, pitch in floats); `out` is then unused
};

// MT = pixel tiles per CTA (1 or 2): with MT = 2 a weight tile fetched from L2 feeds two 128-pixel tiles (two TMEM
// accumulators), which cuts the operand traffic of the L2-bound 64 / 128-channel layers by a quarter.
template <int BN, int MT>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ConvTcArgs a) {
  constexpr int kTmemCols = MT * BN < 32 ? 32 : MT * BN;
  constexpr uint32_t kABytes = BM * BK * 2, kWBytes = BN * BK * 2, kStageBytes = MT * kABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * (BM * MT), nc0 = blockIdx.y * BN, z = blockIdx.z;
  const int pa = z >> 1, pb = z & 1;
  const int HW = a.H * a.W;
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
  tc::pdl_wait();      // everything above overlaps the previous kernel's tail (no-op without the PDL launch attribute)

  if (warp == 0) {
    if (tc::elect_one()) {
      int tn0[MT], ty0[MT];     // image and first row of each pixel tile (hoisted: the producer loop must stay short)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int mm = m0 + mt * BM;
        tn0[mt] = mm / HW;
        ty0[mt] = (mm - tn0[mt] * HW) / a.W;
      }
      int tap = 0, cb = 0;
      for (int kb = 0; kb < nkb; ++kb, ++cb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        if (!tc::mbar_wait(&empty_bar[s], ph ^ 1u, 1)) break;
        if (cb == cblocks) { cb = 0; ++tap; }
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
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tc::tma_load_5d(sa + mt * kABytes, &map_a, &full_bar[s], px * a.in_pitch + cb * BK, dx, py, ty0[mt] + dy, tn0[mt]);
          tc::tma_load_2d(sa + MT * kABytes, &map_w, &full_bar[s], tap * a.Cin + cb * BK, nc0);
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
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)      // a tile past the last pixel is out of bounds in n: zero-filled, never stored
          tc::tma_load_4d(sa + mt * kABytes, &map_a, &full_bar[s], cb * BK, dx, ty0[mt] + dy, tn0[mt]);
        tc::tma_load_2d(sa + MT * kABytes, &map_w, &full_bar[s], tap * a.Cin + cb * BK, z * a.Cout + nc0);
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
        const uint64_t dw = tc::make_desc_sw128(a_addr + MT * kABytes);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint64_t da = tc::make_desc_sw128(a_addr + mt * kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16(tmem_base + (uint32_t)(mt * BN), da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        }
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    tc::mbar_wait(&tmem_full_bar, 0, 3);
    tc::fence_after_sync();
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
    const int p = m0 + mt * BM + q * 32 + lane;
    const bool valid = p < a.total_pix;
    size_t op = (size_t)p;
    if (a.up == 2) {
      const int n = p / HW, rem = p - n * HW, y = rem / a.W, x = rem - y * a.W;
      op = ((size_t)n * (2 * a.H) + (size_t)(2 * y + pa)) * (size_t)(2 * a.W) + (size_t)(2 * x + pb);
    }
    bf16* dst = a.out + op * (size_t)a.out_pitch + nc0;
    float* dst32 = a.out32 ? a.out32 + op * (size_t)a.out_pitch + nc0 : nullptr;
    const float* prow = a.post ? a.post + (size_t)(valid ? p / HW : 0) * a.post_stride + nc0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + c0), v);
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
        if (dst32) {
          float4* d4 = reinterpret_cast<float4*>(dst32 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
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
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------------
// conv_halo_kernel: 3x3 convolution with Cin = 64 for LARGE images (64 x 64 and up), where conv_tc_kernel is bound by
// L2 -> shared-memory traffic (every tap reloads the 16 KB pixel tile and every CTA reloads the weights: 216 KB per
// 128 pixels against 1152 tensor-pipe cycles).  Here a persistent CTA keeps all nine weight tiles resident and loads
// each pixel tile ONCE, with its halo, as one TMA box (64 ch, W + 2, rows, 1) at (0, -1, r - 1, n): the out-of-bounds
// columns / rows are zero-filled, so the box lands as rows of W + 2 "padded slots".  In that padded row-major order a
// tap (dy, dx) is a pure shift by dy (W + 2) + dx slots, i.e. the SAME shared-memory tile read through a UMMA descriptor
// whose start address is moved by whole 128-byte rows (the descriptor's base-offset field carries the swizzle phase
// of a start that is not 1024-byte aligned).  A work unit is 128 consecutive padded slots of one image (W = 64: 33
// units per image, 2 of 66 slots per row are border slots whose results are dropped): 42 KB of loads instead of 216.
//   warp 0: TMA producer (2-stage ring of halo tiles)   warp 1: MMA issuer (9 taps x 4 UMMA 128 x BN x 16 per unit,
//   two TMEM accumulators)                               warps 2-5: epilogue of unit j under the MMAs of unit j + 1
// MODE 0: bias / ReLU / time term, bf16 NHWC store.  MODE 1 / 2 (BN = 16, three real output channels): out_conv of the
// pixel path, eps store / posterior update (pix_out.cuh).
// ------------------------------------------------------------------------------------------------------------------
struct HaloArgs {
  int H, W, Wp;            // Wp = W + 2
  int nrows;               // rows of the halo box
  int units_per_img;       // ceil(H * Wp / 128)
  int total_units;         // B * units_per_img
  int a_bytes, a_stride;   // bytes of one halo tile; stage pitch (1024-aligned, plus one guard row in front)
  int out_pitch, relu, post_stride;
  const float* post;
  const float* bias;
  bf16* out;
  int stages;              // halo tiles in flight (<= kHaloStages)
  PixOutArgs fin;          // MODE 1 / 2
};

constexpr int kHaloStages = 3;       // halo tiles in flight (one 42 KB TMA box has ~2 us of latency from a cold L2)
constexpr int kHaloTmemStages = 4;

template <int BN, int MODE, int NCH>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const HaloArgs a) {
  constexpr uint32_t kWTap = BN * BK * 2, kWBytes = 9 * kWTap, kWRegion = (kWBytes + 1023) & ~1023u;
  // Consecutive tcgen05.mma into ONE accumulator are issued ~100 cycles apart whatever N is (measured: 36 dependent
  // MMAs per unit took 3.2 k cycles at N = 16 and 3.7 k at N = 64), so the nine taps are dealt round-robin to NCH
  // independent accumulators that the epilogue adds up.
  constexpr int kAcc1 = BN < 32 ? 32 : BN;     // TMEM columns of one accumulator
  constexpr int kAcc = NCH * kAcc1;            // TMEM columns of one unit (stage)
  constexpr int kTS = kHaloTmemStages;         // accumulator stages: the epilogue of unit j runs under the MMAs of j+1 .. j+3
  constexpr int kCols = kTS * kAcc <= 64 ? 64 : (kTS * kAcc <= 128 ? 128 : (kTS * kAcc <= 256 ? 256 : 512));
  static_assert(kTS * kAcc <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + kWRegion + 1024;        // one guard KB: tap (-1, -1) of slot 0 reads 128 bytes in front of the tile
  __shared__ __align__(8) uint64_t full_bar[kHaloStages], empty_bar[kHaloStages], tfull_bar[kTS], tempty_bar[kTS], w_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[BN], post_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_a);
    tc::prefetch_tmap(&map_w);
    for (int s = 0; s < kHaloStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kTS; ++s) {
      tc::mbar_init(&tfull_bar[s], 1);
      tc::mbar_init(&tempty_bar[s], 4);
    }
    tc::mbar_init(&w_bar, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(&w_bar, kWBytes);      // weights do not depend on the previous kernel: fetched before the PDL wait
    for (int tap = 0; tap < 9; ++tap) tc::tma_load_2d(w_s + tap * kWTap, &map_w, &w_bar, tap * BK, 0);
  }
  if (warp == 1) tc::tmem_alloc<kCols>(&tmem_slot);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      bias_s[i] = a.bias ? a.bias[i] : 0.f;
      post_s[i] = (a.post && a.post_stride == 0) ? a.post[i] : 0.f;      // one time term for the whole batch (the sampler)
    }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int rowslots = a.Wp;
  tc::pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      for (int u = blockIdx.x; u < a.total_units; u += gridDim.x, ++it) {
        const int s = it % a.stages;
        if (!tc::mbar_wait(&empty_bar[s], (uint32_t)((it / a.stages) & 1) ^ 1u, 11)) break;
        const int n = u / a.units_per_img, s0 = (u - n * a.units_per_img) * BM, r = s0 / rowslots;
        tc::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)a.a_bytes);
        tc::tma_load_4d(a_s + (size_t)s * a.a_stride, &map_a, &full_bar[s], 0, -1, r - 1, n);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
      bool ok = tc::mbar_wait(&w_bar, 0, 12);
      const uint32_t w_addr = tc::smem_u32(w_s);
      int it = 0;
      for (int u = blockIdx.x; u < a.total_units && ok; u += gridDim.x, ++it) {
        const int s = it % a.stages, ts = it % kTS;
        const uint32_t ph = (uint32_t)((it / a.stages) & 1), tph = (uint32_t)((it / kTS) & 1);
        const int s0 = (u % a.units_per_img) * BM;
        const int first = s0 % rowslots + rowslots;      // tile slot of output slot 0 (the tile starts one row above)
        ok = tc::mbar_wait(&tempty_bar[ts], tph ^ 1u, 13) && tc::mbar_wait(&full_bar[s], ph, 14);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(a_s + (size_t)s * a.a_stride);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ts * kAcc);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          // row-shifted view of the SAME tile: the start address moves by whole 128-byte rows.  No descriptor base offset:
          // on B200 the 128-byte swizzle is a function of absolute shared-memory address bits (with the base-offset field
          // set to (address >> 7) & 7 the results are wrong; measured, see profiles/README.md)
          const int shift = first + (tap / 3 - 1) * rowslots + (tap % 3 - 1);
          const uint64_t da = tc::make_desc_sw128(a_addr + (uint32_t)shift * 128u);
          const uint64_t dw = tc::make_desc_sw128(w_addr + tap * kWTap);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc::umma_bf16(d_tmem + (uint32_t)((tap % NCH) * kAcc1), da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc,
                          (uint32_t)((tap >= NCH) || k != 0));
        }
        tc::umma_commit(&empty_bar[s]);
        tc::umma_commit(&tfull_bar[ts]);
      }
    }
  } else {
    const int q = warp & 3;
    const int HW = a.H * a.W;
    int it = 0;
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x, ++it) {
      const int s = it % kTS;
      const uint32_t ph = (uint32_t)((it / kTS) & 1);
      const int n = u / a.units_per_img, slot = (u - n * a.units_per_img) * BM + q * 32 + lane;
      const int y = slot / rowslots, cx = slot - y * rowslots;
      const bool valid = y < a.H && cx >= 1 && cx <= a.W;
      const int rem = y * a.W + cx - 1;
      if (!tc::mbar_wait(&tfull_bar[s], ph, 15)) break;
      tc::fence_after_sync();
      const uint32_t t_addr = tmem_base + (uint32_t)(s * kAcc) + ((uint32_t)(q * 32) << 16);
      if (MODE == 0) {
        bf16* dst = a.out + ((size_t)n * HW + (valid ? rem : 0)) * (size_t)a.out_pitch;
        const float* prow = (a.post && a.post_stride) ? a.post + (size_t)n * a.post_stride : nullptr;   // per-sample time terms
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          tc::tmem_ld16(t_addr + (uint32_t)c0, v);
#pragma unroll
          for (int ch = 1; ch < NCH; ++ch) {
            float v2[16];
            tc::tmem_ld16(t_addr + (uint32_t)(ch * kAcc1 + c0), v2);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += v2[j];
          }
          if (valid) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[j] += bias_s[c0 + j];
              if (a.relu) v[j] = fmaxf(v[j], 0.f);
              v[j] += post_s[c0 + j];
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
      } else {
        float v[16];
        tc::tmem_ld16(t_addr, v);
#pragma unroll
        for (int ch = 1; ch < NCH; ++ch) {
          float v2[16];
          tc::tmem_ld16(t_addr + (uint32_t)(ch * kAcc1), v2);
          v[0] += v2[0]; v[1] += v2[1]; v[2] += v2[2];
        }
        if (valid) {
          float eps[3] = {v[0] + bias_s[0], v[1] + bias_s[1], v[2] + bias_s[2]};
          float z[3] = {0.f, 0.f, 0.f};
          if (MODE == 2) pix_noise(a.fin, n, rem, HW, z);
          pix_finish<MODE == 2 ? 1 : 0>(a.fin, eps, n, rem, HW, z);
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[s]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// convt_halo_kernel<BN, KB>: one output parity (pa, pb) of ConvTranspose2d(4, 2, 1) with Cin = 64 KB, Cout = BN, in the halo
// scheme: a persistent CTA keeps the parity's 4 taps x KB weight tiles resident (64 KB at BN = 64, KB = 2) and loads every
// 128-slot pixel run ONCE with its halo (one box per 64-channel block), the four taps (0 | -1 or +1 per axis) being row /
// column shifts of that tile.  Measured (256 images): 152 -> 108 us; one ring slot per 64-channel tile (five slots) and two
// alternating accumulators changed nothing - about 105 clocks per MMA remain, as in conv_halo_kernel.  conv_tc_kernel reloads the pixel tile for every tap and the weight tile for every pixel tile:
// 160 - 192 KB of operands per 128 x 64 outputs against 2.1 k tensor clocks, which bound up1 (128 -> 64 at 64 x 64) at a
// third of the tensor rate.  blockIdx.y = parity; units are strided over blockIdx.x.
// ------------------------------------------------------------------------------------------------------------------
struct ConvTHaloArgs {
  int H, W, Wp;            // input grid; Wp = W + 2
  int units_per_img, total_units;
  int a_bytes, a_stride;   // one 64-channel halo tile; tile pitch (1024-aligned plus one guard KB)
  int Cin, Cout, out_pitch, relu;
  const float* bias;
  bf16* out;               // (B, 2H, 2W, out_pitch)
  int stages;
};
constexpr int kCtStagesMax = 3;

template <int BN, int KB>
__global__ void __launch_bounds__(kThreads, 1)
convt_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ConvTHaloArgs a) {
  constexpr uint32_t kWTap = BN * BK * 2, kWBytes = 4 * KB * kWTap;
  constexpr int kTS = 4;                       // accumulator stages: the epilogue of unit j runs under the MMAs of j + 1 .. j + 3
  constexpr int kCols = kTS * BN <= 256 ? 256 : 512;
  static_assert(kTS * BN <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + kWBytes + 1024;        // one guard KB: tap (-1, -1) of slot 0 reads 128 bytes in front of a tile
  __shared__ __align__(8) uint64_t full_bar[kCtStagesMax], empty_bar[kCtStagesMax], tfull_bar[kTS], tempty_bar[kTS], w_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.y, pa = z >> 1, pb = z & 1;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_a);
    tc::prefetch_tmap(&map_w);
    for (int s = 0; s < kCtStagesMax; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kTS; ++s) {
      tc::mbar_init(&tfull_bar[s], 1);
      tc::mbar_init(&tempty_bar[s], 4);
    }
    tc::mbar_init(&w_bar, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(&w_bar, kWBytes);      // weights do not depend on the previous kernel: fetched before the PDL wait
    for (int t = 0; t < 4; ++t)
      for (int cb = 0; cb < KB; ++cb)
        tc::tma_load_2d(w_s + (t * KB + cb) * kWTap, &map_w, &w_bar, t * a.Cin + cb * BK, z * a.Cout);
  }
  if (warp == 1) tc::tmem_alloc<kCols>(&tmem_slot);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_s[i] = a.bias ? a.bias[i] : 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int rowslots = a.Wp;
  const uint32_t stage_pitch = (uint32_t)(KB * a.a_stride);
  tc::pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      for (int u = blockIdx.x; u < a.total_units; u += gridDim.x, ++it) {
        const int s = it % a.stages;
        if (!tc::mbar_wait(&empty_bar[s], (uint32_t)((it / a.stages) & 1) ^ 1u, 11)) break;
        const int n = u / a.units_per_img, s0 = (u - n * a.units_per_img) * BM, r = s0 / rowslots;
        tc::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(KB * a.a_bytes));
        for (int cb = 0; cb < KB; ++cb)
          tc::tma_load_4d(a_s + (size_t)s * stage_pitch + (size_t)cb * a.a_stride, &map_a, &full_bar[s], cb * BK, -1, r - 1, n);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
      bool ok = tc::mbar_wait(&w_bar, 0, 12);
      const uint32_t w_addr = tc::smem_u32(w_s);
      int it = 0;
      for (int u = blockIdx.x; u < a.total_units && ok; u += gridDim.x, ++it) {
        const int s = it % a.stages, ts = it % kTS;
        const uint32_t ph = (uint32_t)((it / a.stages) & 1), tph = (uint32_t)((it / kTS) & 1);
        const int s0 = (u % a.units_per_img) * BM;
        const int first = s0 % rowslots + rowslots;      // tile slot of output slot 0 (the tile starts one row above)
        ok = tc::mbar_wait(&tempty_bar[ts], tph ^ 1u, 13) && tc::mbar_wait(&full_bar[s], ph, 14);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(a_s + (size_t)s * stage_pitch);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ts * BN);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // tap t of parity (pa, pb): offset 0 or (-1 | +1) per axis - a row / column shift of the SAME tile (conv_halo_kernel);
          // tap-major, channel blocks inside: the summation order of conv_tc_kernel, so both kernels give the same bits
          const int dy = (t >> 1) == 0 ? 0 : (pa == 0 ? -1 : 1), dx = (t & 1) == 0 ? 0 : (pb == 0 ? -1 : 1);
          const int shift = first + dy * rowslots + dx;
#pragma unroll
          for (int cb = 0; cb < KB; ++cb) {
            const uint64_t da = tc::make_desc_sw128(a_addr + (uint32_t)(cb * a.a_stride) + (uint32_t)shift * 128u);
            const uint64_t dw = tc::make_desc_sw128(w_addr + (uint32_t)(t * KB + cb) * kWTap);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_bf16(d_tmem, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((t | cb | k) != 0));
          }
        }
        tc::umma_commit(&empty_bar[s]);
        tc::umma_commit(&tfull_bar[ts]);
      }
    }
  } else {
    const int q = warp & 3;
    const int H2 = 2 * a.H, W2 = 2 * a.W;
    int it = 0;
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x, ++it) {
      const int s = it % kTS;
      const uint32_t ph = (uint32_t)((it / kTS) & 1);
      const int n = u / a.units_per_img, slot = (u - n * a.units_per_img) * BM + q * 32 + lane;
      const int y = slot / rowslots, cx = slot - y * rowslots;
      const bool valid = y < a.H && cx >= 1 && cx <= a.W;
      if (!tc::mbar_wait(&tfull_bar[s], ph, 15)) break;
      tc::fence_after_sync();
      const uint32_t t_addr = tmem_base + (uint32_t)(s * BN) + ((uint32_t)(q * 32) << 16);
      const size_t op = valid ? ((size_t)n * H2 + (size_t)(2 * y + pa)) * (size_t)W2 + (size_t)(2 * (cx - 1) + pb) : 0;
      bf16* dst = a.out + op * (size_t)a.out_pitch;
      // all column blocks of the accumulator are requested back to back and awaited once; the accumulator stage is handed back
      // to the MMA issuer before the arithmetic and the stores (the epilogue of a unit, one warp per 32 rows, is as long as the
      // unit's MMAs: a serial load - wait - finish chain per 16 columns left the tensor core waiting for free stages)
      uint32_t raw[BN / 16][16];
#pragma unroll
      for (int cb = 0; cb < BN / 16; ++cb) tc::tmem_ld16_nowait(t_addr + (uint32_t)(cb * 16), raw[cb]);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[s]);
      if (valid) {
#pragma unroll
        for (int cb = 0; cb < BN / 16; ++cb) {
          float v[16];
          uint32_t pk[8];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[cb * 16 + 4 * q4]);
            v[4 * q4] = __uint_as_float(raw[cb][4 * q4]) + b4.x;
            v[4 * q4 + 1] = __uint_as_float(raw[cb][4 * q4 + 1]) + b4.y;
            v[4 * q4 + 2] = __uint_as_float(raw[cb][4 * q4 + 2]) + b4.z;
            v[4 * q4 + 3] = __uint_as_float(raw[cb][4 * q4 + 3]) + b4.w;
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          uint4* d4 = reinterpret_cast<uint4*>(dst + cb * 16);
          d4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          d4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// conv_halo_stream_kernel<BN>: the halo scheme for Cin = 64 * KB > 64, where the nine weight tiles of every 64-channel
// block no longer fit next to the pixel tiles.  Pixel tiles keep the halo layout (ONE box per (unit, channel block)
// instead of nine), the weight tiles (BN x 64, one per (channel block, tap)) stream through their own TMA ring in
// consumption order.  Operand bytes per 128 x 128 output tile of a 128 -> 128 convolution at 32 x 32: 61 KB of pixels +
// 288 KB of weights instead of 288 + 288 (conv_tc_kernel is at the L2 bandwidth cap on those layers).
//   warp 0: TMA producer (pixel tile of the NEXT channel block one step ahead of the weight tiles)
//   warp 1: MMA issuer: per unit KB x 9 taps x 4 UMMA 128 x BN x 16 into one of four TMEM accumulators
//   warps 2-5: epilogue (bias, ReLU, time term, bf16 NHWC store) under the MMAs of the following units
// ------------------------------------------------------------------------------------------------------------------
struct HaloStreamArgs {
  int H, W, Wp, nrows, units_per_img, total_units;
  int a_bytes, a_stride;     // one pixel tile; stage pitch (1024-aligned + guard)
  int KB;                    // 64-channel blocks of the input
  int NA, NW;                // ring depths: pixel tiles, weight tiles
  int Cin;
  int out_pitch, relu, post_stride;
  const float* post;
  const float* bias;
  bf16* out;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_stream_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const HaloStreamArgs a) {
  constexpr uint32_t kWTile = BN * BK * 2;
  constexpr int kTS = 512 / BN < 4 ? 512 / BN : 4;      // TMEM accumulator stages
  constexpr int kMaxRing = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                                   // NW weight tiles
  uint8_t* a_s = smem + (size_t)a.NW * kWTile + 1024;    // NA pixel tiles, one guard KB in front
  __shared__ __align__(8) uint64_t afull[kMaxRing], aempty[kMaxRing], wfull[kMaxRing], wempty[kMaxRing], tfull[kTS], tempty[kTS];
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[BN], post_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NA = a.NA, NW = a.NW, KB = a.KB;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_a);
    tc::prefetch_tmap(&map_w);
    for (int s = 0; s < kMaxRing; ++s) {
      tc::mbar_init(&afull[s], 1); tc::mbar_init(&aempty[s], 1);
      tc::mbar_init(&wfull[s], 1); tc::mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < kTS; ++s) { tc::mbar_init(&tfull[s], 1); tc::mbar_init(&tempty[s], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<kTS * BN>(&tmem_slot);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      bias_s[i] = a.bias ? a.bias[i] : 0.f;
      post_s[i] = (a.post && a.post_stride == 0) ? a.post[i] : 0.f;
    }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const int rowslots = a.Wp;
  tc::pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      int ia = 0, iw = 0;        // running indices of the pixel / weight tiles issued
      auto load_a = [&](int u, int kb) -> bool {
        const int s = ia % NA;
        if (!tc::mbar_wait(&aempty[s], (uint32_t)((ia / NA) & 1) ^ 1u, 31)) return false;
        const int n = u / a.units_per_img, s0 = (u - n * a.units_per_img) * BM, r = s0 / rowslots;
        tc::mbar_arrive_expect_tx(&afull[s], (uint32_t)a.a_bytes);
        tc::tma_load_4d(a_s + (size_t)s * a.a_stride, &map_a, &afull[s], kb * BK, -1, r - 1, n);
        ++ia;
        return true;
      };
      bool ok = blockIdx.x < a.total_units ? load_a(blockIdx.x, 0) : true;
      for (int u = blockIdx.x; u < a.total_units && ok; u += gridDim.x) {
        for (int kb = 0; kb < KB && ok; ++kb) {
          // the pixel tile the MMA warp will need AFTER the nine weight tiles below
          if (kb + 1 < KB) ok = load_a(u, kb + 1);
          else if (u + (int)gridDim.x < a.total_units) ok = load_a(u + gridDim.x, 0);
          for (int tap = 0; tap < 9 && ok; ++tap) {
            const int s = iw % NW;
            ok = tc::mbar_wait(&wempty[s], (uint32_t)((iw / NW) & 1) ^ 1u, 32);
            tc::mbar_arrive_expect_tx(&wfull[s], kWTile);
            tc::tma_load_2d(w_s + (size_t)s * kWTile, &map_w, &wfull[s], tap * a.Cin + kb * BK, 0);
            ++iw;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
      bool ok = true;
      int ia = 0, iw = 0, it = 0;
      for (int u = blockIdx.x; u < a.total_units && ok; u += gridDim.x, ++it) {
        const int ts = it % kTS;
        const int s0 = (u % a.units_per_img) * BM;
        const int first = s0 % rowslots + rowslots;
        ok = tc::mbar_wait(&tempty[ts], (uint32_t)((it / kTS) & 1) ^ 1u, 33);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ts * BN);
        for (int kb = 0; kb < KB && ok; ++kb, ++ia) {
          const int sa = ia % NA;
          ok = tc::mbar_wait(&afull[sa], (uint32_t)((ia / NA) & 1), 34);
          const uint32_t a_addr = tc::smem_u32(a_s + (size_t)sa * a.a_stride);
#pragma unroll 1
          for (int tap = 0; tap < 9 && ok; ++tap, ++iw) {
            const int sw = iw % NW;
            ok = tc::mbar_wait(&wfull[sw], (uint32_t)((iw / NW) & 1), 35);
            tc::fence_after_sync();
            const int shift = first + (tap / 3 - 1) * rowslots + (tap % 3 - 1);
            const uint64_t da = tc::make_desc_sw128(a_addr + (uint32_t)shift * 128u);
            const uint64_t dw = tc::make_desc_sw128(tc::smem_u32(w_s + (size_t)sw * kWTile));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_bf16(d_tmem, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | tap | k) != 0));
            tc::umma_commit(&wempty[sw]);
          }
          tc::umma_commit(&aempty[sa]);
        }
        tc::umma_commit(&tfull[ts]);
      }
    }
  } else {
    const int q = warp & 3;
    const int HW = a.H * a.W;
    int it = 0;
    for (int u = blockIdx.x; u < a.total_units; u += gridDim.x, ++it) {
      const int ts = it % kTS;
      const int n = u / a.units_per_img, slot = (u - n * a.units_per_img) * BM + q * 32 + lane;
      const int y = slot / rowslots, cx = slot - y * rowslots;
      const bool valid = y < a.H && cx >= 1 && cx <= a.W;
      const int rem = y * a.W + cx - 1;
      if (!tc::mbar_wait(&tfull[ts], (uint32_t)((it / kTS) & 1), 36)) break;
      tc::fence_after_sync();
      const uint32_t t_addr = tmem_base + (uint32_t)(ts * BN) + ((uint32_t)(q * 32) << 16);
      bf16* dst = a.out + ((size_t)n * HW + (valid ? rem : 0)) * (size_t)a.out_pitch;
      const float* prow = (a.post && a.post_stride) ? a.post + (size_t)n * a.post_stride : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tc::tmem_ld16(t_addr + (uint32_t)c0, v);
        if (valid) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] += bias_s[c0 + j];
            if (a.relu) v[j] = fmaxf(v[j], 0.f);
            v[j] += post_s[c0 + j];
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
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[ts]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTS * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// conv_tc_pair_kernel: the 256-channel output tile as a 2-CTA MMA (tcgen05 cta_group::2, thread-block cluster of two).
// One tcgen05.mma of M = 128 reads its operands from shared memory at 64 B/clk (measured, tools/ubench.cu): a
// 128 x 256 x 16 tile step needs 4 KB of pixels + 8 KB of weights = 192 clocks against 128 clocks of tensor time.  A CTA
// pair computes 256 pixels x 256 channels per instruction: each CTA holds ITS 128 pixels (A) and HALF of the weight tile
// (B rows [128 r, 128 r + 128)), i.e. 4 KB + 4 KB per step = the tensor rate, and each weight byte is fetched from L2 once
// per pair.  Both CTAs run the TMA producer (cp.async.bulk.tensor ... cta_group::2, completing on the LEADER's barrier);
// the leader's elected thread issues the MMAs; tcgen05.commit multicasts the slot release and the accumulator-ready signal
// to both CTAs; each CTA's epilogue warps read their own TMEM (their 128 pixels x 256 channels).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pair_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pair_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address: the even (leader) CTA of the pair
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 2)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ConvTcArgs a) {
  constexpr int BN = 256, HALF = 128;
  constexpr uint32_t kABytes = BM * BK * 2, kWBytes = HALF * BK * 2, kStageBytes = kABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];      // used in the leader only: both CTAs' loads complete on it
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float bias_s[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = pair_rank();
  const int m0 = blockIdx.x * BM, nc0 = blockIdx.y * BN, z = blockIdx.z;
  const int pa = z >> 1, pb = z & 1;
  const int HW = a.H * a.W;
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
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_s[i] = a.bias ? a.bias[nc0 + i] : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  pair_sync();              // the peer's barriers exist before any remote completion / multicast commit reaches them
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  tc::pdl_wait();           // everything above overlaps the previous kernel's tail (no-op without the PDL launch attribute)

  if (warp == 0) {
    if (tc::elect_one()) {
      const int tn0 = m0 / HW, ty0 = (m0 - tn0 * HW) / a.W;
      const int wrow = z * a.Cout + nc0 + (int)rank * HALF;      // this CTA's half of the weight tile
      int tap = 0, cb = 0;
      for (int kb = 0; kb < nkb; ++kb, ++cb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        if (!tc::mbar_wait(&empty_bar[s], ph ^ 1u, 1)) break;
        if (cb == cblocks) { cb = 0; ++tap; }
        uint8_t* sa = smem + (size_t)s * kStageBytes;
        const uint32_t fb = tc::smem_u32(&full_bar[s]) & kPeerMask;      // the leader's barrier
        if (rank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * kStageBytes);
        if (a.up == 0) {
          const int ky = tap >> 2, kx = tap & 3;
          const int dy = ky == 0 ? -1 : (ky == 3 ? 1 : 0), dx = kx == 0 ? -1 : (kx == 3 ? 1 : 0);
          const int py = (ky == 0 || ky == 2) ? 1 : 0, px = (kx == 0 || kx == 2) ? 1 : 0;
          tma2_load_5d(sa, &map_a, fb, px * a.in_pitch + cb * BK, dx, py, ty0 + dy, tn0);
        } else {
          int dy, dx;
          if (a.up == 1) { dy = tap / 3 - 1; dx = tap - (tap / 3) * 3 - 1; }
          else { dy = (tap >> 1) == 0 ? 0 : (pa == 0 ? -1 : 1); dx = (tap & 1) == 0 ? 0 : (pb == 0 ? -1 : 1); }
          tma2_load_4d(sa, &map_a, fb, cb * BK, dx, ty0 + dy, tn0);
        }
        tma2_load_2d(sa + kABytes, &map_w, fb, tap * a.Cin + cb * BK, wrow);
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(2 * BM, BN);
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
          umma2_bf16(tmem_base, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        umma2_commit_both(&empty_bar[s]);
      }
      umma2_commit_both(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    tc::mbar_wait(&tmem_full_bar, 0, 3);
    tc::fence_after_sync();
    const int p = m0 + q * 32 + lane;
    const bool valid = p < a.total_pix;
    size_t op = (size_t)p;
    if (a.up == 2) {
      const int n = p / HW, rem = p - n * HW, y = rem / a.W, x = rem - y * a.W;
      op = ((size_t)n * (2 * a.H) + (size_t)(2 * y + pa)) * (size_t)(2 * a.W) + (size_t)(2 * x + pb);
    }
    bf16* dst = a.out + op * (size_t)a.out_pitch + nc0;
    float* dst32 = a.out32 ? a.out32 + op * (size_t)a.out_pitch + nc0 : nullptr;
    const float* prow = a.post ? a.post + (size_t)(valid ? p / HW : 0) * a.post_stride + nc0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
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
        if (dst32) {
          float4* d4 = reinterpret_cast<float4*>(dst32 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint32_t pk[8];
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
  }
  tc::fence_before_sync();
  __syncthreads();
  pair_sync();              // neither CTA frees TMEM / leaves while its peer's MMAs or commits may still address it
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
}

// final_conv[3] + Sigmoid (v2:277-278): Conv2d(32, 3, 3, padding 1) over NHWC bf16 -> NCHW fp32.  N = 3 output
// channels is no tensor-core shape: one thread per pixel on the CUDA cores.  A CTA owns an 8 x 32 pixel tile: its
// 10 x 34 halo is loaded ONCE with coalesced 16-byte reads into shared memory, channel-octet major ([4][340] uint4), so
// that the 32 lanes of a warp (32 neighbouring pixels) read consecutive 16-byte words for every tap (the direct
// per-thread global reads of the first version walked 64-byte strides through L1: 576 wavefronts per warp instead of 144).
constexpr int kO3TW = 32, kO3TH = 16, kO3HW = kO3TW + 2, kO3HH = kO3TH + 2, kO3HP = kO3HW * kO3HH;
__global__ void __launch_bounds__(256)
conv_out3_kernel(const bf16* __restrict__ in, const float* __restrict__ w /* (3, 9*32) */, const float* __restrict__ bias,
                 float* __restrict__ out, int H, int W) {
  __shared__ float4 ws[3 * 72];          // 16-byte broadcast reads of the weights
  __shared__ uint4 tile[4][kO3HP];       // 16 x 32 pixels + halo, channel-octet major
  for (int i = threadIdx.x; i < 3 * 288; i += 256) reinterpret_cast<float*>(ws)[i] = w[i];
  const int n = blockIdx.z, y0 = blockIdx.y * kO3TH, x0 = blockIdx.x * kO3TW, HW = H * W;
  const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)n * HW * 32);
  for (int i = threadIdx.x; i < 4 * kO3HP; i += 256) {
    const int hp = i >> 2, j = i & 3, hy = hp / kO3HW, hx = hp - hy * kO3HW;
    const int yy = y0 + hy - 1, xx = x0 + hx - 1;
    uint4 v = make_uint4(0, 0, 0, 0);     // zero padding
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(src + ((size_t)yy * W + xx) * 4 + j);
    tile[j][hp] = v;
  }
  __syncthreads();
  // every thread finishes TWO pixels (rows ty and ty + 8 of the tile): each weight word read from shared memory feeds both
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  float acc[2][3];
#pragma unroll
  for (int r = 0; r < 2; ++r) { acc[r][0] = bias[0]; acc[r][1] = bias[1]; acc[r][2] = bias[2]; }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int hp = (ty + tap / 3) * kO3HW + tx + tap % 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k4 = tap * 8 + j * 2;      // float4 index of channel j * 8 of this tap
      const float4 a0 = ws[k4], a1 = ws[k4 + 1], b0 = ws[72 + k4], b1 = ws[72 + k4 + 1], c0 = ws[144 + k4], c1 = ws[144 + k4 + 1];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint4 u = tile[j][hp + r * 8 * kO3HW];
        const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&wd[e]);
          f[2 * e] = __low2float(h2); f[2 * e + 1] = __high2float(h2);
        }
        acc[r][0] += f[0] * a0.x + f[1] * a0.y + f[2] * a0.z + f[3] * a0.w + f[4] * a1.x + f[5] * a1.y + f[6] * a1.z + f[7] * a1.w;
        acc[r][1] += f[0] * b0.x + f[1] * b0.y + f[2] * b0.z + f[3] * b0.w + f[4] * b1.x + f[5] * b1.y + f[6] * b1.z + f[7] * b1.w;
        acc[r][2] += f[0] * c0.x + f[1] * c0.y + f[2] * c0.z + f[3] * c0.w + f[4] * c1.x + f[5] * c1.y + f[6] * c1.z + f[7] * c1.w;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = y0 + ty + 8 * r, x = x0 + tx;
    if (y < H && x < W) {
      float* o = out + (size_t)n * 3 * HW + (size_t)y * W + x;
      o[0] = sigmoidf_(acc[r][0]);
      o[HW] = sigmoidf_(acc[r][1]);
      o[2 * HW] = sigmoidf_(acc[r][2]);
    }
  }
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
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
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

template <int BN, int MT>
int launch_bn_mt(ldm_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mw, const ConvTcArgs& a0, int nz, cudaStream_t st) {
  ConvTcArgs a = a0;
  const int nkb = a.taps * (a.Cin / BK);
  const size_t stage_bytes = (size_t)MT * BM * BK * 2 + (size_t)BN * BK * 2;
  static int limit_kb = -1;
  if (limit_kb < 0) {
    const char* e = getenv("LDM_CONV_SMEM_KB");
    limit_kb = e ? atoi(e) : 100;
  }
  int stages = nkb < kMaxStages ? nkb : kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes > (size_t)limit_kb * 1024) --stages;   // two CTAs per SM: one finishes while the other loads
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  dim3 grid(ceil_div(a.total_pix, BM * MT), a.Cout / BN, nz);
  LDM_CUDA(launch_maybe_pdl(conv_tc_kernel<BN, MT>, grid, kThreads, smem, st, ctx->use_pdl, ma, mw, a));
  ctx->launches++;
  ldm_kmark(ctx, "conv_tc");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

int conv_mt2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_CONV_MT2");
    v = e ? atoi(e) : 1;
  }
  return v;
}

int conv_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LDM_CONV_PAIR");
    v = e ? atoi(e) : 1;
  }
  return v;
}

// 256-channel tiles as CTA pairs (conv_tc_pair_kernel): `mw128` is the weight map with a 128-row box
int launch_pair(ldm_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mw128, const ConvTcArgs& a0, int nz, cudaStream_t st) {
  ConvTcArgs a = a0;
  const int nkb = a.taps * (a.Cin / BK);
  const size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)128 * BK * 2;
  static int want = -1;
  if (want < 0) {
    const char* e = getenv("LDM_CONV_PAIR_STAGES");
    want = e ? atoi(e) : 3;            // 3 x 32 KB: two CTAs per SM (one pair's epilogue under the other pair's main loop)
    if (want < 2) want = 2;
    if (want > 6) want = 6;
  }
  int stages = nkb < want ? nkb : want;
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool attr = false;
  if (!attr) {
    LDM_CUDA(cudaFuncSetAttribute(conv_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid(ceil_div(a.total_pix, BM), a.Cout / 256, nz);      // grid.x even (checked by the caller): __cluster_dims__(2, 1, 1)
  LDM_CUDA(launch_maybe_pdl(conv_tc_pair_kernel, grid, kThreads, smem, st, ctx->use_pdl, ma, mw128, a));
  ctx->launches++;
  ldm_kmark(ctx, "conv_tc_pair");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
int launch_bn(ldm_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mw, const ConvTcArgs& a, int nz, cudaStream_t st) {
  // two pixel tiles per CTA where that still leaves >= 1.5 CTAs per SM (64 / 128-channel tiles only: TMEM and
  // shared memory of two co-resident CTAs)
  if constexpr (BN == 64 || BN == 128) {
    if (conv_mt2() && (long long)ceil_div(a.total_pix, 2 * BM) * (a.Cout / BN) * nz * 2 >= 3ll * ctx->sm_count)
      return launch_bn_mt<BN, 2>(ctx, ma, mw, a, nz, st);
  }
  return launch_bn_mt<BN, 1>(ctx, ma, mw, a, nz, st);
}

}  // namespace

int conv_tc_pick_bn(int Cout) { return Cout >= 256 ? 256 : Cout; }

// in: (B, H, W, Cin of pitch in_pitch) bf16; out: (B, up*H, up*W, Cout of pitch out_pitch) bf16 (mode 0: (B, H/2, W/2, .)).
// L.w16 / L.map_w: (nz * Cout, taps * Cin), box (64, bn).  mode = ConvTcArgs::up.
static int conv_tc_launch(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, float* out32,
                          int out_pitch, int B, int H, int W, int mode, int relu, const float* post, int post_stride, cudaStream_t st) {
  LDM_TRY(conv_init(ctx));
  LDM_CHECK(L.w16 != nullptr, "conv_tc: layer not packed for the tensor-core path");
  LDM_CHECK(L.Cin % BK == 0, "conv_tc: Cin %% 64 == 0 required (Cin=%d)", L.Cin);
  LDM_CHECK((mode == 1 && L.taps == 9) || (mode == 2 && L.taps == 4) || (mode == 0 && L.taps == 16), "conv_tc: unsupported taps/mode combination");
  LDM_CHECK(out_pitch % 8 == 0 && (((uintptr_t)out | (uintptr_t)out32) & 15) == 0, "conv_tc: output must be 16-byte aligned (pitch %d)", out_pitch);
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
  a.bias = bias; a.out = out; a.out32 = out32;
  const int nz = mode == 2 ? 4 : 1;
  int bn = L.bn ? L.bn : conv_tc_pick_bn(L.Cout);
  const CUtensorMap* mw = &L.map_w;
  if (L.bn_alt && L.bn_alt < bn && ceil_div(a.total_pix, BM) * (L.Cout / bn) * nz < ctx->sm_count) {
    bn = L.bn_alt;      // fewer tiles than SMs: halve the tile (measured: 128 tiles of 256 channels 33 us -> 256 tiles of 128 channels 24 us)
    mw = &L.map_w_alt;
  }
  if (bn == 256 && conv_pair() && L.bn_alt == 128 && ceil_div(a.total_pix, BM) % 2 == 0)
    return launch_pair(ctx, ma, L.map_w_alt, a, nz, st);      // 2-CTA MMA: 256 pixels x 256 channels per CTA pair
  switch (bn) {
    case 32: return launch_bn<32>(ctx, ma, *mw, a, nz, st);
    case 64: return launch_bn<64>(ctx, ma, *mw, a, nz, st);
    case 128: return launch_bn<128>(ctx, ma, *mw, a, nz, st);
    case 256: return launch_bn<256>(ctx, ma, *mw, a, nz, st);
  }
  ldm_set_error("conv_tc: unsupported Cout %d", L.Cout);
  return -1;
}

int launch_conv_tc_ex(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int mode, int relu, const float* post, int post_stride, cudaStream_t st) {
  return conv_tc_launch(ctx, in, in_pitch, L, bias, out, nullptr, out_pitch, B, H, W, mode, relu, post, post_stride, st);
}

// Strict mode: `in` holds the (hi, lo, hi) bf16 thirds of an fp32 activation (3 C channels per pixel), L the matching
// (hi, hi, lo) split of the weights (L.Cin = 3 C): the three-term product on the bf16 tensor cores, fp32 NHWC output.
int launch_conv_tc_f32out(ldm_ctx* ctx, const bf16* in, const ConvLayer& L, const float* bias, float* out32, int B, int H, int W,
                          int up, cudaStream_t st) {
  return conv_tc_launch(ctx, in, L.Cin, L, bias, nullptr, out32, L.Cout, B, H, W, up, 0, nullptr, 0, st);
}

int launch_conv_tc(ldm_ctx* ctx, const bf16* in, const ConvLayer& L, const float* bias, bf16* out, int B, int H, int W,
                   int up, cudaStream_t st) {
  return launch_conv_tc_ex(ctx, in, L.Cin, L, bias, out, L.Cout, B, H, W, up, 0, nullptr, 0, st);
}


// 3x3 convolution, Cin = 64, through conv_halo_kernel.  L.map_w: (Cout rows, 9 * 64), box (64, Cout); Cout = 64 (bf16
// NHWC output) or 16 (out_conv of the pixel path: fin != nullptr, rows 3..15 of the packed weight are zero).
int conv_halo_supported(int H, int W, int Cin, int Cout) {
  if (Cin != 64 || (Cout != 64 && Cout != 32 && Cout != 16) || W < 30 || W + 2 > 256) return 0;
  const int Wp = W + 2, nrows = (Wp - 1 + 128 + Wp - 1) / Wp + 2;
  const size_t a_stride = (((size_t)nrows * Wp * 128 + 1023) & ~(size_t)1023) + 1024;
  const size_t smem = (((size_t)9 * Cout * 128 + 1023) & ~(size_t)1023) + 1024 + kHaloStages * a_stride + 1024;
  return nrows <= 256 && smem <= 220 * 1024;
}

int launch_conv_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                     int B, int H, int W, int relu, const float* post, int post_stride, const PixOutArgs* fin, int ddpm,
                     cudaStream_t st) {
  LDM_TRY(conv_init(ctx));
  LDM_CHECK(conv_halo_supported(H, W, L.Cin, L.Cout), "conv_halo: unsupported shape (H=%d W=%d Cin=%d Cout=%d)", H, W, L.Cin, L.Cout);
  LDM_CHECK(((uintptr_t)in & 15) == 0 && in_pitch % 8 == 0, "conv_halo: input must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<32, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<16, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<16, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    attr_set = true;
  }
  HaloArgs a = {};
  a.H = H; a.W = W; a.Wp = W + 2;
  a.nrows = (a.Wp - 1 + 128 + a.Wp - 1) / a.Wp + 2;      // rows a 128-slot run can touch, plus one above and one below
  a.units_per_img = ceil_div(H * a.Wp, BM);
  a.total_units = B * a.units_per_img;
  a.a_bytes = a.nrows * a.Wp * 128;
  a.a_stride = ((a.a_bytes + 1023) & ~1023) + 1024;
  a.out_pitch = out_pitch; a.relu = relu; a.post = post; a.post_stride = post_stride; a.bias = bias; a.out = out;
  if (fin) a.fin = *fin;
  CUtensorMap ma;
  {
    cuuint64_t dims[4] = {(cuuint64_t)64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)a.Wp, (cuuint32_t)a.nrows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode4(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ldm_set_error("cuTensorMapEncodeTiled (halo box %d x %d) failed: CUresult %d", a.Wp, a.nrows, (int)r);
      return (int)r;
    }
  }
  // out_conv (16 weight rows): two stages leave room for TWO CTAs per SM, which hides the exposed tile latency of the
  // single persistent CTA (26 -> measured below); the 64-channel variant keeps its 72 KB of weights and three stages
  static int two = -1;
  if (two < 0) {
    const char* e = getenv("LDM_HALO_TWO_CTAS");
    two = e ? atoi(e) : 1;
  }
  const bool pair = fin && two;
  a.stages = pair ? 2 : kHaloStages;
  const size_t smem = (((size_t)9 * L.Cout * 128 + 1023) & ~(size_t)1023) + 1024 + a.stages * (size_t)a.a_stride + 1024;
  const int slots = pair ? 2 * ctx->sm_count : ctx->sm_count;
  const int grid = a.total_units < slots ? a.total_units : slots;
  LDM_CHECK(fin ? L.Cout == 16 : (L.Cout == 64 || L.Cout == 32), "conv_halo: %d output channels in this mode", L.Cout);
  if (!fin && L.Cout == 32) LDM_CUDA(launch_maybe_pdl(conv_halo_kernel<32, 0, 1>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, L.map_w, a));
  else if (!fin) LDM_CUDA(launch_maybe_pdl(conv_halo_kernel<64, 0, 1>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, L.map_w, a));
  else if (ddpm) LDM_CUDA(launch_maybe_pdl(conv_halo_kernel<16, 2, 2>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, L.map_w, a));
  else LDM_CUDA(launch_maybe_pdl(conv_halo_kernel<16, 1, 2>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, L.map_w, a));
  ctx->launches++;
  ldm_kmark(ctx, "conv_halo");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

// ConvTranspose2d(4, 2, 1) with Cin = 128, Cout = 64 through convt_halo_kernel (the four parities as blockIdx.y).  L: the stacked
// sub-pixel weights of decoder.cu / pixel.cu ((4 Cout) rows, 4 taps x Cin columns), L.map_w boxed (64, 64).
static int convt_halo_plan(int H, int W, int Cin, int Cout, ConvTHaloArgs* a, size_t* smem) {
  if (Cin != 128 || Cout != 64 || W < 30 || W + 2 > 256) return 0;
  const int Wp = W + 2, nrows = (Wp - 1 + 128 + Wp - 1) / Wp + 2;
  if (nrows > 256) return 0;
  const int a_bytes = nrows * Wp * 128, a_stride = ((a_bytes + 1023) & ~1023) + 1024, KB = Cin / 64;
  const size_t fixed = (size_t)4 * KB * Cout * 128 + 1024 + 1024;
  int stages = (int)((220 * 1024 - fixed) / ((size_t)KB * a_stride));
  if (stages > kCtStagesMax) stages = kCtStagesMax;
  if (stages < 2) return 0;
  if (a) {
    a->H = H; a->W = W; a->Wp = Wp;
    a->units_per_img = ceil_div(H * Wp, BM);
    a->a_bytes = a_bytes; a->a_stride = a_stride; a->Cin = Cin; a->Cout = Cout; a->stages = stages;
  }
  if (smem) *smem = fixed + (size_t)stages * KB * a_stride;
  return 1;
}
int convt_halo_supported(int H, int W, int Cin, int Cout) {
  static const bool on = !(getenv("LDM_CONVT_HALO") && atoi(getenv("LDM_CONVT_HALO")) == 0);
  return on && convt_halo_plan(H, W, Cin, Cout, nullptr, nullptr);
}
int launch_convt_halo(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const float* bias, bf16* out, int out_pitch,
                      int B, int H, int W, int relu, cudaStream_t st) {
  LDM_TRY(conv_init(ctx));
  ConvTHaloArgs a = {};
  size_t smem = 0;
  LDM_CHECK(convt_halo_plan(H, W, L.Cin, L.Cout, &a, &smem), "convt_halo: unsupported shape (H=%d W=%d Cin=%d Cout=%d)", H, W, L.Cin, L.Cout);
  LDM_CHECK(L.taps == 4 && L.w16 != nullptr && ((uintptr_t)in & 15) == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0,
            "convt_halo: packed sub-pixel weights and 16-byte aligned buffers required");
  static bool attr_set = false;
  if (!attr_set) {
    LDM_CUDA(cudaFuncSetAttribute(convt_halo_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024));
    attr_set = true;
  }
  a.total_units = B * a.units_per_img;
  a.out_pitch = out_pitch; a.relu = relu; a.bias = bias; a.out = out;
  CUtensorMap ma;
  {
    const int nrows = a.a_bytes / (a.Wp * 128);
    cuuint64_t dims[4] = {(cuuint64_t)L.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)a.Wp, (cuuint32_t)nrows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode4(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ldm_set_error("cuTensorMapEncodeTiled (convt halo box %d x %d) failed: CUresult %d", a.Wp, nrows, (int)r);
      return (int)r;
    }
  }
  const int per = ctx->sm_count / 4 < 1 ? 1 : ctx->sm_count / 4;      // one CTA per SM, the parities side by side
  const int gx = a.total_units < per ? a.total_units : per;
  LDM_CUDA(launch_maybe_pdl(convt_halo_kernel<64, 2>, dim3(gx, 4), kThreads, smem, st, ctx->use_pdl, ma, L.map_w, a));
  ctx->launches++;
  ldm_kmark(ctx, "conv_tc");      // keeps the per-kernel tables of bench.py: the same layer on another kernel
  LDM_CUDA(cudaGetLastError());
  return 0;
}

// 3x3 convolution with Cin = 64 * KB (KB >= 1) and Cout = 64 or 128 through conv_halo_stream_kernel.  L.map_w must be
// boxed (64, Cout).  Returns 0 from *_supported when the shape does not fit (the caller falls back to conv_tc_kernel).
static void halo_stream_plan(int H, int W, int Cin, int Cout, int* nrows, int* a_stride, int* NA, int* NW, size_t* smem) {
  const int Wp = W + 2, KB = Cin / 64;
  *nrows = (Wp - 1 + 128 + Wp - 1) / Wp + 2;
  *a_stride = (((*nrows) * Wp * 128 + 1023) & ~1023) + 1024;
  const size_t budget = 218 * 1024, wtile = (size_t)Cout * 128;
  int nw = 4, na = (int)((budget - 2048 - nw * wtile) / (size_t)(*a_stride));
  if (na > 2 * KB) na = 2 * KB;
  if (na > 8) na = 8;
  while (nw < 8 && 2048 + (nw + 1) * wtile + (size_t)na * (*a_stride) <= budget) ++nw;
  *NA = na; *NW = nw;
  *smem = 2048 + nw * wtile + (size_t)na * (*a_stride) + 1024;
}

int conv_halo_stream_supported(int H, int W, int Cin, int Cout) {
  if (Cin % 64 != 0 || Cin < 64 || (Cout != 64 && Cout != 128) || W < 30 || W + 2 > 256) return 0;
  int nrows, a_stride, NA, NW;
  size_t smem;
  halo_stream_plan(H, W, Cin, Cout, &nrows, &a_stride, &NA, &NW, &smem);
  return nrows <= 256 && NA >= 2 && smem <= 222 * 1024;
}

int launch_conv_halo_stream(ldm_ctx* ctx, const bf16* in, int in_pitch, const ConvLayer& L, const CUtensorMap& map_w, const float* bias,
                            bf16* out, int out_pitch, int B, int H, int W, int relu, const float* post, int post_stride, cudaStream_t st) {
  LDM_TRY(conv_init(ctx));
  LDM_CHECK(conv_halo_stream_supported(H, W, L.Cin, L.Cout), "conv_halo_stream: unsupported shape (H=%d W=%d Cin=%d Cout=%d)", H, W, L.Cin, L.Cout);
  LDM_CHECK(((uintptr_t)in & 15) == 0 && in_pitch % 8 == 0, "conv_halo_stream: input must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_stream_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024));
    LDM_CUDA(cudaFuncSetAttribute(conv_halo_stream_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024));
    attr_set = true;
  }
  HaloStreamArgs a = {};
  size_t smem;
  a.H = H; a.W = W; a.Wp = W + 2; a.KB = L.Cin / 64; a.Cin = L.Cin;
  halo_stream_plan(H, W, L.Cin, L.Cout, &a.nrows, &a.a_stride, &a.NA, &a.NW, &smem);
  a.units_per_img = ceil_div(H * a.Wp, BM);
  a.total_units = B * a.units_per_img;
  a.a_bytes = a.nrows * a.Wp * 128;
  a.out_pitch = out_pitch; a.relu = relu; a.post = post; a.post_stride = post_stride; a.bias = bias; a.out = out;
  CUtensorMap ma;
  {
    cuuint64_t dims[4] = {(cuuint64_t)L.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)a.Wp, (cuuint32_t)a.nrows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode4(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ldm_set_error("cuTensorMapEncodeTiled (halo box %d x %d) failed: CUresult %d", a.Wp, a.nrows, (int)r);
      return (int)r;
    }
  }
  const int grid = a.total_units < ctx->sm_count ? a.total_units : ctx->sm_count;
  if (L.Cout == 64) LDM_CUDA(launch_maybe_pdl(conv_halo_stream_kernel<64>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, map_w, a));
  else LDM_CUDA(launch_maybe_pdl(conv_halo_stream_kernel<128>, dim3(grid), kThreads, smem, st, ctx->use_pdl, ma, map_w, a));
  ctx->launches++;
  ldm_kmark(ctx, "conv_halo_stream");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

int launch_conv_out3(ldm_ctx* ctx, const bf16* in, const float* w, const float* bias, float* out, int B, int H, int W,
                     cudaStream_t st) {
  conv_out3_kernel<<<dim3(ceil_div(W, kO3TW), ceil_div(H, kO3TH), B), 256, 0, st>>>(in, w, bias, out, H, W);
  ctx->launches++;
  ldm_kmark(ctx, "conv_out3");
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
