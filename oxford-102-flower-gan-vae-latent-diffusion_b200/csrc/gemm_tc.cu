// bf16 tensor-core GEMM for sm_100a:  C(M,N) = epi(A(M,K) . W(N,K)^T), fp32 accumulation in TMEM.
//   warp 0 : TMA producer  - cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64), 128-byte swizzle,
//                            multi-stage mbarrier ring
//   warp 1 : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16), commits free the ring
//   warps 2-5 : epilogue   - tcgen05.ld accumulators (lane = row), fused bias / table / residual /
//                            DDPM-update epilogue (epilogue.cuh)
// Used for every dense contraction of the denoiser step (v2:539-561) and Decoder.fc (v2:246-250).
#include "common.cuh"
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct TcArgs {
  int M, N, K;
  int stages;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TcArgs args,
               const Epilogue epi) {
  constexpr int kTmemCols = BN < 32 ? 32 : BN;           // power of two >= 32
  constexpr uint32_t kABytes = BM * BK * 2, kWBytes = BN * BK * 2, kStageBytes = kABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // Split-K over a thread-block cluster along z: CTA kr of KS contracts k-blocks [kr * nkb, (kr + 1) * nkb); the partial
  // accumulators of ranks 1.. are parked in their own shared memory and folded in by rank 0 through DSMEM.  A small-M GEMM
  // (M = 128 rows: N / BN CTAs, each streaming the whole A panel at ~44 B/clk/SM) is bound by that stream, not by MMAs.
  const int KS = gridDim.z, kr = blockIdx.z;
  const int nkb = args.K / BK / KS, kb0 = kr * nkb, S = args.stages;

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
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  // PDL (no-ops without the launch attribute): the NEXT kernel of the stream may start its prologue now, on the SMs this
  // small grid leaves idle; this kernel fetches its first weight tiles (they do not depend on the previous kernel) and
  // only then waits for the previous kernel's results.
  tc::pdl_launch_dependents();

  if (warp == 0) {
    if (tc::elect_one()) {
      const int pre = nkb < S ? nkb : S;
      for (int kb = 0; kb < pre; ++kb) {
        uint8_t* sa = smem + (size_t)kb * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[kb], kStageBytes);
        tc::tma_load_2d(sa + kABytes, &map_w, &full_bar[kb], (kb0 + kb) * BK, n0);
      }
      tc::pdl_wait();
      for (int kb = 0; kb < pre; ++kb) tc::tma_load_2d(smem + (size_t)kb * kStageBytes, &map_a, &full_bar[kb], (kb0 + kb) * BK, m0);
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        if (!tc::mbar_wait(&empty_bar[s], ph ^ 1u, 1)) break;
        uint8_t* sa = smem + (size_t)s * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        tc::tma_load_2d(sa, &map_a, &full_bar[s], (kb0 + kb) * BK, m0);
        tc::tma_load_2d(sa + kABytes, &map_w, &full_bar[s], (kb0 + kb) * BK, n0);
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
        for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K step = +2 in the 16-byte address field
          tc::umma_bf16(tmem_base, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(&tmem_full_bar);
    }
  } else {
    // epilogue warps 2..5 own TMEM lanes [32*(warp%4), +32)
    const int q = warp & 3;
    tc::pdl_wait();                       // the epilogue reads residuals / state written by earlier kernels
    tc::mbar_wait(&tmem_full_bar, 0, 3);  // every MMA of this CTA has completed: the stage buffers are free
    tc::fence_after_sync();
    if (kr != 0) {                        // park the partial accumulator, column-major (lanes = consecutive rows)
      float* red = reinterpret_cast<float*>(smem);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) red[(c0 + j) * BM + q * 32 + lane] = v[j];
      }
    }
  }
  if (KS > 1) cooperative_groups::this_cluster().sync();      // partials visible cluster-wide
  if (warp >= 2 && kr == 0 && epi.stage_f32 && KS == 1) {
    // bias + fp32 store through shared memory (the ring is free: every MMA has completed).  Row pitch BN * 4 + 16 bytes:
    // the 16-byte accesses of a quarter-warp fall into distinct banks both when a lane writes its own row and when a warp
    // reads one row.
    const int q = warp & 3;
    constexpr int kPitch = BN * 4 + 16;
    uint8_t* stg = smem;
    uint8_t* mine = stg + (size_t)(q * 32 + lane) * kPitch;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 b = epi.bias ? *reinterpret_cast<const float4*>(epi.bias + n0 + c0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(mine + (size_t)(c0 + j) * 4) = make_float4(v[j] + b.x, v[j + 1] + b.y, v[j + 2] + b.z, v[j + 3] + b.w);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int r = q; r < BM; r += 4) {
      const int row = m0 + r;
      if (row >= args.M) break;
      const uint8_t* src = stg + (size_t)r * kPitch;
      float* dst = epi.out_f32 + (size_t)row * epi.ld_of + n0;
      for (int c = lane; c < BN / 4; c += 32) reinterpret_cast<float4*>(dst)[c] = *reinterpret_cast<const float4*>(src + (size_t)c * 16);
    }
  } else if (warp >= 2 && kr == 0) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      for (int r = 1; r < KS; ++r) {
        const float* rem = cooperative_groups::this_cluster().map_shared_rank(reinterpret_cast<float*>(smem), r);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += rem[(c0 + j) * BM + q * 32 + lane];
      }
      if (row < args.M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int col = n0 + c0 + j;
          if (col < args.N) epi_finish4(epi, row, col, args.N, &v[j]);
        }
      }
    }
  }
  if (KS > 1) cooperative_groups::this_cluster().sync();      // ranks 1.. keep their shared memory alive until rank 0 has read it
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------------
// attn_tc_kernel<HD>: softmax(Q K^T / sqrt(HD)) V on the tensor cores, one CTA per (128 queries, head, batch element).
// v3's nn.MultiheadAttention over the rows of a call (v3:832-838: L = batch, 8 heads) and the spatial self-attention of
// UNetAttentionBlock (v2:434-459: L = H W tokens, 4 heads) are both this.
//   S = Q K^T : UMMA 128 x 128 x 16, HD / 16 steps, Q and K tiles (tokens x 64-column atoms, K-major) by TMA -> TMEM cols 0..127
//   softmax   : warps 2-5, thread = query row: two passes over the S row in TMEM (max, then exp / sum), P written as bf16
//               into shared memory in the 128-byte-swizzled K-major operand layout, online rescale across key tiles
//   O_j = P V : UMMA 128 x HD x 16, 8 steps; B operand = V^T tile (HD rows x keys), from the transposed copy the prep
//               kernel writes -> TMEM cols 128..128+HD; the row threads fold O_j into their fp32 accumulator registers
// Keys beyond L are masked by index; queries beyond L are not stored.
// ------------------------------------------------------------------------------------------------------------------
struct AttnArgs {
  int L, heads;
  int q_col0, k_col0;        // first column of Q / K of head 0 in the QK matrix (head h adds h * HD)
  float scale;
  void* out;                 // (batches * L, out_pitch): column of (head h, dim e) = h * out_sh + e * out_se
  int out_bf16, out_pitch, out_sh, out_se;
};

template <int HD>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_vt, const AttnArgs a) {
  constexpr int NA = (HD + 63) / 64;                 // 64-column atoms of a Q / K tile
  constexpr uint32_t kAtom = 128 * 128;              // 128 rows x 128 bytes
  constexpr uint32_t kQ = NA * kAtom, kVAtom = HD * 128, kV = 2 * kVAtom, kP = 2 * kAtom;
  constexpr uint32_t kVRegion = (kV + 1023) & ~1023u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + kQ;
  uint8_t* v_s = k_s + kQ;
  uint8_t* p_s = v_s + kVRegion;
  __shared__ __align__(8) uint64_t bar_q, bar_kv, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, z = blockIdx.z;
  const int L = a.L, ntiles = (L + 127) / 128;
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_qk);
    tc::prefetch_tmap(&map_vt);
    tc::mbar_init(&bar_q, 1);
    tc::mbar_init(&bar_kv, 1);
    tc::mbar_init(&bar_s, 1);
    tc::mbar_init(&bar_p, 128);
    tc::mbar_init(&bar_o, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(&tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  tc::pdl_launch_dependents();
  tc::pdl_wait();

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(&bar_q, kQ);
      for (int at = 0; at < NA; ++at) tc::tma_load_2d(q_s + at * kAtom, &map_qk, &bar_q, a.q_col0 + h * HD + at * 64, z * L + q0);
      for (int j = 0; j < ntiles; ++j) {
        if (j > 0 && !tc::mbar_wait(&bar_o, (uint32_t)((j - 1) & 1), 21)) break;     // K / V^T / P of the previous tile consumed
        tc::mbar_arrive_expect_tx(&bar_kv, kQ + kV);
        for (int at = 0; at < NA; ++at) tc::tma_load_2d(k_s + at * kAtom, &map_qk, &bar_kv, a.k_col0 + h * HD + at * 64, z * L + j * 128);
        for (int at = 0; at < 2; ++at) tc::tma_load_2d(v_s + at * kVAtom, &map_vt, &bar_kv, j * 128 + at * 64, (z * a.heads + h) * HD);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, 128), idesc_o = tc::make_idesc_bf16(128, HD);
      bool ok = tc::mbar_wait(&bar_q, 0, 22);
      const uint32_t qa = tc::smem_u32(q_s), ka = tc::smem_u32(k_s), va = tc::smem_u32(v_s), pa = tc::smem_u32(p_s);
      for (int j = 0; j < ntiles && ok; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        ok = tc::mbar_wait(&bar_kv, ph, 23);
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (uint32_t)(ks / 4) * kAtom + (uint32_t)(ks % 4) * 32u;
          tc::umma_bf16(tmem_base, tc::make_desc_sw128(qa + off), tc::make_desc_sw128(ka + off), idesc_s, (uint32_t)(ks != 0));
        }
        tc::umma_commit(&bar_s);
        ok = ok && tc::mbar_wait(&bar_p, ph, 24);
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t kk = (uint32_t)(ks % 4) * 32u;
          tc::umma_bf16(tmem_base + 128u, tc::make_desc_sw128(pa + (uint32_t)(ks / 4) * kAtom + kk),
                        tc::make_desc_sw128(va + (uint32_t)(ks / 4) * kVAtom + kk), idesc_o, (uint32_t)(ks != 0));
        }
        tc::umma_commit(&bar_o);
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY, l = 0.f;
    float o[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) o[e] = 0.f;
    uint8_t* p_row = p_s + r * 128;
    bool ok = true;
    for (int j = 0; j < ntiles && ok; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      const int nvalid = L - j * 128;            // keys of this tile that exist
      ok = tc::mbar_wait(&bar_s, ph, 25);
      tc::fence_after_sync();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tc::tmem_ld16(t_row + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < nvalid) mx = fmaxf(mx, v[i] * a.scale);
      }
      const float m_new = fmaxf(m, mx);
      const float corr = __expf(m - m_new);
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tc::tmem_ld16(t_row + (uint32_t)c0, v);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float p0 = (c0 + 2 * i < nvalid) ? __expf(v[2 * i] * a.scale - m_new) : 0.f;
          const float p1 = (c0 + 2 * i + 1 < nvalid) ? __expf(v[2 * i + 1] * a.scale - m_new) : 0.f;
          sum += p0 + p1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          pk[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        // keys c0 .. c0 + 15 = 16-byte chunks (c0 % 64) / 8 and the next one of atom c0 / 64, XOR-swizzled with the row
        uint8_t* base = p_row + (c0 / 64) * kAtom;
        const int ch = (c0 % 64) / 8;
        *reinterpret_cast<uint4*>(base + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(base + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      l = l * corr + sum;
      m = m_new;
      tc::fence_proxy_async();       // the generic-proxy writes of P must be visible to the tensor core's async proxy
      tc::fence_before_sync();
      tc::mbar_arrive(&bar_p);
      ok = ok && tc::mbar_wait(&bar_o, ph, 26);
      tc::fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        float v[16];
        tc::tmem_ld16(t_row + 128u + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c0 + i] = o[c0 + i] * corr + v[i];
      }
      tc::fence_before_sync();
    }
    if (ok && q0 + r < L) {
      const float inv = 1.0f / l;
      const size_t row = (size_t)z * L + q0 + r;
      if (a.out_bf16) {
        bf16* dst = reinterpret_cast<bf16*>(a.out) + row * a.out_pitch + (size_t)h * a.out_sh;
        if (a.out_se == 1) {
#pragma unroll
          for (int e = 0; e < HD; e += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(o[e + 2 * i] * inv, o[e + 2 * i + 1] * inv);
              pk[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(dst + e) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < HD; ++e) dst[(size_t)e * a.out_se] = __float2bfloat16_rn(o[e] * inv);
        }
      } else {
        float* dst = reinterpret_cast<float*>(a.out) + row * a.out_pitch + (size_t)h * a.out_sh;
#pragma unroll
        for (int e = 0; e < HD; ++e) dst[(size_t)e * a.out_se] = o[e] * inv;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem_base);
}

// qkv fp32 (rows, 3 d) = [Q | K | V] (the in_proj output, v3:832) -> QK bf16 (rows, 2 d) and V^T bf16 (d, ldv)
__global__ void __launch_bounds__(256)
attn_prep_kernel(const float* __restrict__ qkv, bf16* __restrict__ qk, bf16* __restrict__ vt, int rows, int d, int ldv) {
  ldm_pdl_launch_dependents();
  ldm_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 3 * d) return;
  const int row = i / (3 * d), col = i - row * 3 * d;
  const bf16 v = __float2bfloat16_rn(qkv[i]);
  if (col < 2 * d) qk[(size_t)row * 2 * d + col] = v;
  else vt[(size_t)(col - 2 * d) * ldv + row] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

// 2-D bf16 tensor (rows, cols) with row pitch ld (elements), box (box_rows, 64 cols), 128-byte swizzle
int make_map_2d(const void* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  LDM_CHECK(g_encode != nullptr, "tensor-core path not initialised (cuTensorMapEncodeTiled unavailable)");
  LDM_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "TMA needs 16-byte aligned base and pitch");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ldm_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%d cols=%d ld=%d box_rows=%d)", (int)r, rows, cols, ld, box_rows);
    return (int)r;
  }
  return 0;
}

template <int BN>
int launch_bn(ldm_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mw, int M, int N, int K, const Epilogue& epi,
              cudaStream_t st) {
  const int nkb_all = K / BK;
  const size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)BN * BK * 2;
  // split-K cluster (portable sizes): as many k-slices as keep >= 2 k-blocks per CTA and the grid within the SM count.
  // Off by default: parity-green, but measured neutral to slightly slower on the M = 128 GEMMs of the v3 step (526 -> 508
  // samples/s) - those kernels are bound by their fixed launch / prologue / first-tile latency (~8 us), not by the A stream.
  static int splitk = -1;
  if (splitk < 0) {
    const char* e = getenv("LDM_GEMM_SPLITK");
    splitk = e ? atoi(e) : 0;
  }
  int ks = 1;
  const int ctas = (N / BN) * ceil_div(M, BM);
  if (splitk)
    for (int cand = 8; cand >= 2; cand >>= 1)
      if (nkb_all % cand == 0 && nkb_all / cand >= 2 && ctas * cand <= ctx->sm_count) { ks = cand; break; }
  const int nkb = nkb_all / ks;
  int stages = nkb < kMaxStages ? nkb : kMaxStages;
  while ((size_t)stages * stage_bytes > 200 * 1024) --stages;
  if (epi.stage_f32)      // store-bound GEMM: two CTAs per SM, one writes its tile while the other runs its main loop
    while (stages > 2 && (size_t)stages * stage_bytes > 100 * 1024) --stages;
  size_t smem = (size_t)stages * stage_bytes + 1024;
  if (ks > 1 && smem < (size_t)BM * BN * 4 + 1024) smem = (size_t)BM * BN * 4 + 1024;       // room for the parked partial
  if (epi.stage_f32 && smem < (size_t)BM * (BN * 4 + 16) + 1024) smem = (size_t)BM * (BN * 4 + 16) + 1024;      // ... or for the staged output tile
  TcArgs a{M, N, K, stages};
  dim3 grid(N / BN, ceil_div(M, BM), ks);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ks > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = ks;
      ++na;
    }
    if (ctx->use_pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    LDM_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, ma, mw, a, epi));
  }
  ctx->launches++;
  ldm_kmark(ctx, "gemm_tc");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int tc_init(ldm_ctx* ctx) {
  (void)ctx;
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not found in the driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  // opt in to large dynamic shared memory once, outside any stream capture
  LDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  return 0;
}

int tc_pick_bn(int M, int N) {
  // keep >= ~32 CTAs in flight for the skinny (M = batch) denoiser GEMMs; wide tiles for big-N layers
  const int mt = ceil_div(M, BM);
  if (N % 128 == 0 && (N / 128) * mt >= 96) return 128;
  if (N % 64 == 0 && (N / 64) * mt >= 48) return 64;
  return 32;
}

int tc_make_weight_map(ldm_ctx* ctx, const bf16* w, int N, int K, int bn, CUtensorMap* out) {
  LDM_TRY(tc_init(ctx));
  LDM_CHECK(K % BK == 0 && N % bn == 0, "tensor-core GEMM needs K %% 64 == 0 and N %% %d == 0 (N=%d K=%d)", bn, N, K);
  return make_map_2d(w, N, K, K, bn, out);
}


// ------------------------------------------------------------------------------------------------------------------
// attention launchers
// ------------------------------------------------------------------------------------------------------------------
int attn_tc_supported(int hd) { return hd == 16 || hd == 32 || hd == 64 || hd == 128; }

int launch_attn_prep(ldm_ctx* ctx, const float* qkv, bf16* qk, bf16* vt, int rows, int d, int ldv, cudaStream_t st) {
  const int n = rows * 3 * d;
  LDM_CUDA(launch_maybe_pdl(attn_prep_kernel, dim3(ceil_div(n, 256)), 256, 0, st, ctx->use_pdl, qkv, qk, vt, rows, d, ldv));
  ctx->launches++;
  ldm_kmark(ctx, "attn_prep");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

template <int HD>
static int launch_attn_hd(ldm_ctx* ctx, const CUtensorMap& mqk, const CUtensorMap& mvt, const AttnArgs& a, int batches, cudaStream_t st) {
  constexpr int NA = (HD + 63) / 64;
  const size_t smem = (size_t)2 * NA * 16384 + (((size_t)2 * HD * 128 + 1023) & ~(size_t)1023) + 2 * 16384 + 1024;
  static bool attr = false;
  if (!attr) {
    LDM_CUDA(cudaFuncSetAttribute(attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  LDM_CUDA(launch_maybe_pdl(attn_tc_kernel<HD>, dim3(ceil_div(a.L, 128), a.heads, batches), kThreads, smem, st, ctx->use_pdl, mqk, mvt, a));
  ctx->launches++;
  ldm_kmark(ctx, "attn_tc");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

// qk: bf16 (batches * L, ld_qk) holding Q at columns q_col0 + h * hd and K at k_col0 + h * hd; vt: bf16
// (batches * heads * hd, ldv) = V transposed (row = (batch, head, dim), column = token).
int launch_attn_tc(ldm_ctx* ctx, const bf16* qk, int ld_qk, int qk_cols, const bf16* vt, int ldv, int L, int batches, int heads, int hd,
                   int q_col0, int k_col0, void* out, int out_bf16, int out_pitch, int out_sh, int out_se, cudaStream_t st) {
  LDM_TRY(tc_init(ctx));
  LDM_CHECK(attn_tc_supported(hd), "attn_tc: head_dim %d unsupported", hd);
  CUtensorMap mqk, mvt;
  LDM_TRY(make_map_2d(qk, batches * L, qk_cols, ld_qk, 128, &mqk));
  LDM_TRY(make_map_2d(vt, batches * heads * hd, L, ldv, hd, &mvt));
  AttnArgs a;
  a.L = L; a.heads = heads; a.q_col0 = q_col0; a.k_col0 = k_col0; a.scale = 1.0f / sqrtf((float)hd);
  a.out = out; a.out_bf16 = out_bf16; a.out_pitch = out_pitch; a.out_sh = out_sh; a.out_se = out_se;
  switch (hd) {
    case 16: return launch_attn_hd<16>(ctx, mqk, mvt, a, batches, st);
    case 32: return launch_attn_hd<32>(ctx, mqk, mvt, a, batches, st);
    case 64: return launch_attn_hd<64>(ctx, mqk, mvt, a, batches, st);
    default: return launch_attn_hd<128>(ctx, mqk, mvt, a, batches, st);
  }
}

int launch_gemm_tc(ldm_ctx* ctx, const bf16* A, int lda, int M, const DenseLayer& L, const Epilogue& epi,
                   cudaStream_t st) {
  LDM_CHECK(L.w16 != nullptr && L.bn > 0, "layer not packed for the tensor-core path");
  auto key = std::make_tuple((const void*)A, M, L.K, lda, 0);
  auto it = ctx->act_maps.find(key);
  if (it == ctx->act_maps.end()) {
    CUtensorMap m;
    LDM_TRY(make_map_2d(A, M, L.K, lda, BM, &m));
    it = ctx->act_maps.emplace(key, m).first;
  }
  switch (L.bn) {
    case 32: return launch_bn<32>(ctx, it->second, L.map_w, M, L.N, L.K, epi, st);
    case 64: return launch_bn<64>(ctx, it->second, L.map_w, M, L.N, L.K, epi, st);
    case 128: return launch_bn<128>(ctx, it->second, L.map_w, M, L.N, L.K, epi, st);
    case 256: return launch_bn<256>(ctx, it->second, L.map_w, M, L.N, L.K, epi, st);
  }
  ldm_set_error("unsupported BN %d", L.bn);
  return -1;
}

int tc_error_flag(int* out) {
  LDM_CUDA(cudaMemcpyFromSymbol(out, g_tc_error, sizeof(int)));
  return 0;
}
int tc_error_reset() {
  int z = 0;
  LDM_CUDA(cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int)));
  return 0;
}

// The same matrix as a 3-D tensor (64 columns, rows, cols / 64 k-blocks): one box = `kblocks` consecutive 64-column tiles of
// box_rows rows, which land as consecutive swizzled tiles in shared memory (one TMA instruction for several k-blocks)
int tc_make_kblock_map(const void* base, int rows, int cols, int ld, int box_rows, int kblocks, CUtensorMap* out) {
  LDM_CHECK(g_encode != nullptr, "tensor-core path not initialised (cuTensorMapEncodeTiled unavailable)");
  LDM_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0 && cols % BK == 0, "TMA needs 16-byte aligned base and pitch, whole k-blocks");
  cuuint64_t dims[3] = {(cuuint64_t)BK, (cuuint64_t)rows, (cuuint64_t)(cols / BK)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)BK * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ldm_set_error("cuTensorMapEncodeTiled (k-block map) failed: CUresult %d (rows=%d cols=%d ld=%d box_rows=%d x %d)", (int)r, rows, cols, ld, box_rows, kblocks);
    return (int)r;
  }
  return 0;
}

// 2-D bf16 activation (rows, cols) with row pitch ld, box (box_rows, 64 columns), 128-byte swizzle; rows read past
// `rows` are zero-filled
int tc_make_act_map(const void* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  return make_map_2d(base, rows, cols, ld, box_rows, out);
}
