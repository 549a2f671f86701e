// Strict-precision path: fp32 CUDA-core GEMM  C = epi(A . W^T)  and the same tile engine used as an
// implicit-GEMM convolution over NHWC activations.  Register-prefetched shared-memory tiles, 4x4
// micro-tiles, 128-bit global loads.  This is the LDM_PRECISION_FP32 mode (eps within 1e-3 of the
// reference); the bf16 tensor-core kernels live in gemm_tc.cu.
#include "common.cuh"
#include "epilogue.cuh"

namespace {

constexpr int BK = 16;

template <int BM, int BN>
struct TileEngine {
  static constexpr int kThreads = (BM / 4) * (BN / 4);
  static constexpr int kALoads = BM * BK / 4 / kThreads;  // float4 loads per thread for the A tile
  static constexpr int kWLoads = BN * BK / 4 / kThreads;
  static_assert(kALoads >= 1 && kWLoads >= 1, "tile too small for the thread count");
  float (*As)[BM + 4];
  float (*Ws)[BN + 4];
  float acc[4][4];

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  __device__ __forceinline__ void store_a(int l, float4 v) {
    const int idx = threadIdx.x + l * kThreads, r = idx >> 2, kq = (idx & 3) * 4;
    As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
  }
  __device__ __forceinline__ void store_w(int l, float4 v) {
    const int idx = threadIdx.x + l * kThreads, r = idx >> 2, kq = (idx & 3) * 4;
    Ws[kq + 0][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
  }
  __device__ __forceinline__ void mma_tile() {
    const int ty = threadIdx.x / (BN / 4), tx = threadIdx.x % (BN / 4);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
};

// ------------------------------------------------------------------------------------------
// dense:  C(M,N) = epi(A(M,K; lda) . W(N,K)^T)
// ------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(TileEngine<BM, BN>::kThreads)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int M, int N, int K,
                const Epilogue epi) {
  using TE = TileEngine<BM, BN>;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  TE te;
  te.As = As;
  te.Ws = Ws;
  te.zero();
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float4 ra[TE::kALoads], rw[TE::kWLoads];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < TE::kALoads; ++l) {
      const int idx = threadIdx.x + l * TE::kThreads, r = m0 + (idx >> 2), kq = k0 + (idx & 3) * 4;
      ra[l] = r < M ? *reinterpret_cast<const float4*>(A + (size_t)r * lda + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int l = 0; l < TE::kWLoads; ++l) {
      const int idx = threadIdx.x + l * TE::kThreads, r = n0 + (idx >> 2), kq = k0 + (idx & 3) * 4;
      rw[l] = r < N ? *reinterpret_cast<const float4*>(W + (size_t)r * K + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int l = 0; l < TE::kALoads; ++l) te.store_a(l, ra[l]);
#pragma unroll
    for (int l = 0; l < TE::kWLoads; ++l) te.store_w(l, rw[l]);
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
    te.mma_tile();
    __syncthreads();
  }
  const int ty = threadIdx.x / (BN / 4), tx = threadIdx.x % (BN / 4);
  const int col = n0 + tx * 4;
  if (col >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row < M) epi_finish4(epi, row, col, N, te.acc[i]);
  }
}

// ------------------------------------------------------------------------------------------
// implicit-GEMM convolution over NHWC fp32: rows = input-grid pixels, K = taps * Cin
// ------------------------------------------------------------------------------------------
template <int BM, int BN>
__global__ void __launch_bounds__(TileEngine<BM, BN>::kThreads)
conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ out, const ConvGeom g) {
  using TE = TileEngine<BM, BN>;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  TE te;
  te.As = As;
  te.Ws = Ws;
  te.zero();
  const int M = g.B * g.H * g.W, N = g.Cout, K = g.taps * g.Cin;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int sd = g.stride > 1 ? g.stride : 1, IH = g.H * sd, IW = g.W * sd;
  const int ipitch = g.in_pitch ? g.in_pitch : g.Cin, opitch = g.out_pitch ? g.out_pitch : g.Cout;
  // pixel coordinates of the A rows this thread stages
  int pn[TE::kALoads], py[TE::kALoads], px[TE::kALoads];
#pragma unroll
  for (int l = 0; l < TE::kALoads; ++l) {
    const int idx = threadIdx.x + l * TE::kThreads, r = m0 + (idx >> 2);
    if (r < M) {
      pn[l] = r / (g.H * g.W);
      const int rem = r - pn[l] * g.H * g.W;
      py[l] = rem / g.W;
      px[l] = rem - py[l] * g.W;
    } else {
      pn[l] = -1; py[l] = 0; px[l] = 0;
    }
  }
  float4 ra[TE::kALoads], rw[TE::kWLoads];
  auto fetch = [&](int k0) {
    const int tap = k0 / g.Cin, c0 = k0 - tap * g.Cin;
    const int dy = g.dy[tap], dx = g.dx[tap];
#pragma unroll
    for (int l = 0; l < TE::kALoads; ++l) {
      const int idx = threadIdx.x + l * TE::kThreads;
      const int y = py[l] * sd + dy, x = px[l] * sd + dx;
      const bool ok = pn[l] >= 0 && y >= 0 && y < IH && x >= 0 && x < IW;
      ra[l] = ok ? *reinterpret_cast<const float4*>(in + (((size_t)pn[l] * IH + y) * IW + x) * ipitch + c0 + (idx & 3) * 4)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int l = 0; l < TE::kWLoads; ++l) {
      const int idx = threadIdx.x + l * TE::kThreads, r = n0 + (idx >> 2), kq = k0 + (idx & 3) * 4;
      rw[l] = r < N ? *reinterpret_cast<const float4*>(W + (size_t)r * K + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int l = 0; l < TE::kALoads; ++l) te.store_a(l, ra[l]);
#pragma unroll
    for (int l = 0; l < TE::kWLoads; ++l) te.store_w(l, rw[l]);
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
    te.mma_tile();
    __syncthreads();
  }
  const int ty = threadIdx.x / (BN / 4), tx = threadIdx.x % (BN / 4);
  const int OH = g.H * g.up, OW = g.W * g.up;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
    const int n = r / (g.H * g.W), rem = r - n * g.H * g.W, y = rem / g.W, x = rem - y * g.W;
    const int oy = y * g.up + (g.up == 2 ? g.pa : 0), ox = x * g.up + (g.up == 2 ? g.pb : 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= N) continue;
      float v = te.acc[i][j] + (bias ? bias[co] : 0.f);
      if (g.act == LDM_ACT_SWISH) v = swishf(v);
      else if (g.act == LDM_ACT_SIGMOID) v = sigmoidf_(v);
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.post) v += g.post[(size_t)n * g.post_stride + co];
      if (g.nchw_out) out[(((size_t)n * N + co) * OH + oy) * OW + ox] = v;
      else out[(((size_t)n * OH + oy) * OW + ox) * opitch + co] = v;
    }
  }
}

}  // namespace

int launch_gemm_f32(ldm_ctx* ctx, const float* A, int lda, const float* W, int M, int N, int K, const Epilogue& epi,
                    cudaStream_t st) {
  LDM_CHECK(K % BK == 0 && N % 4 == 0 && lda % 4 == 0, "gemm_f32: K %% 16, N %% 4, lda %% 4 required (M=%d N=%d K=%d)", M, N, K);
  if (M <= 1024) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 32));
    gemm_f32_kernel<32, 64><<<grid, TileEngine<32, 64>::kThreads, 0, st>>>(A, lda, W, M, N, K, epi);
  } else {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    gemm_f32_kernel<64, 64><<<grid, TileEngine<64, 64>::kThreads, 0, st>>>(A, lda, W, M, N, K, epi);
  }
  ctx->launches++;
  ldm_kmark(ctx, "gemm_f32");
  LDM_CUDA(cudaGetLastError());
  return 0;
}

int launch_conv_f32(ldm_ctx* ctx, const float* in, const float* w, const float* bias, float* out, const ConvGeom& g,
                    cudaStream_t st) {
  LDM_CHECK(g.Cin % BK == 0 && g.taps >= 1 && g.taps <= 16, "conv_f32: Cin %% 16 and taps in [1,16] required");
  const int M = g.B * g.H * g.W;
  if (g.Cout >= 64) {
    dim3 grid(ceil_div(g.Cout, 64), ceil_div(M, 64));
    conv_f32_kernel<64, 64><<<grid, TileEngine<64, 64>::kThreads, 0, st>>>(in, w, bias, out, g);
  } else {
    dim3 grid(ceil_div(g.Cout, 32), ceil_div(M, 64));
    conv_f32_kernel<64, 32><<<grid, TileEngine<64, 32>::kThreads, 0, st>>>(in, w, bias, out, g);
  }
  ctx->launches++;
  ldm_kmark(ctx, "conv_f32");
  LDM_CUDA(cudaGetLastError());
  return 0;
}
