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
  const int nkb = args.K / BK, S = args.stages;

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

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        if (!tc::mbar_wait(&empty_bar[s], ph ^ 1u, 1)) break;
        uint8_t* sa = smem + (size_t)s * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        tc::tma_load_2d(sa, &map_a, &full_bar[s], kb * BK, m0);
        tc::tma_load_2d(sa + kABytes, &map_w, &full_bar[s], kb * BK, n0);
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
    const int row = m0 + q * 32 + lane;
    tc::mbar_wait(&tmem_full_bar, 0, 3);
    tc::fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (row < args.M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int col = n0 + c0 + j;
          if (col < args.N) epi_finish4(epi, row, col, args.N, &v[j]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
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
  const int nkb = K / BK;
  const size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)BN * BK * 2;
  int stages = nkb < kMaxStages ? nkb : kMaxStages;
  while ((size_t)stages * stage_bytes > 200 * 1024) --stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  TcArgs a{M, N, K, stages};
  dim3 grid(N / BN, ceil_div(M, BM));
  gemm_tc_kernel<BN><<<grid, kThreads, smem, st>>>(ma, mw, a, epi);
  ctx->launches++;
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

// 2-D bf16 activation (rows, cols) with row pitch ld, box (box_rows, 64 columns), 128-byte swizzle; rows read past
// `rows` are zero-filled
int tc_make_act_map(const void* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  return make_map_2d(base, rows, cols, ld, box_rows, out);
}
