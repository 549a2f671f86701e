// The v3 reverse-diffusion loop as ONE persistent sm_100a kernel (bf16 path).
//
//   v3 ConditionalDenoiseDiffusion.sample / p_sample (v3:876-893) calling ConditionalUNet.forward (v3:804-853)
//
// v3's nn.MultiheadAttention runs over the ROWS of a call (v3:832: L = batch, 8 heads), so the rows of a call are coupled
// and the per-cluster decomposition of chain.cu does not apply.  One call (B <= 128 rows = one UMMA M tile) is instead run by
// a grid of G <= 128 co-resident CTAs that walk the same phase list for all T steps and meet at a grid-wide barrier (one
// global counter, release / acquire) between phases; the per-layer path (27 launches per step, each ~8 us of launch,
// prologue and first-tile latency) becomes 4 phases per stage:
//
//   GEMM  [h_i | u_i] = in . [Wh ; Wb Wh]^T + tables     in = x (i = 0) or [h2_{i-1} | a_{i-1}]; Wh = [W_down | W_down W_o]
//                                                         folds out_proj, the residual and the down projection of the
//                                                         previous stage with this stage's block Linear (v3:818-825,838-841)
//   MID   h2 = swish(LN_a(u)) + h ; n = LN_b(h2)          one row per CTA, 128 threads (v3:826-830)
//   GEMM  [Q | K | V] = in_proj(n)                         epilogue writes [Q | K] row-major and V transposed (v3:832)
//   ATTN  a = softmax(Q K^T / sqrt(hd)) V                  one CTA per head: S and O in TMEM (v3:832-836)
//
// and GEMM h_S -> LN_f -> GEMM eps + posterior update for the tail (v3:844-853, 880-893).
//
// A GEMM phase is M = 128 rows x N outputs, cut into 32-column tiles and, for the long reductions, into k-parts whose fp32
// partials are added by the following row phase (fixed order: deterministic).  Work item j of a phase goes to CTA j mod G.
//   warp 0    : TMA producer.  The ring of 4 stages x two k-blocks (one 3-D box for the 128 x 128 operand tile, one for the
//               32 x 128 weight tile: a cp.async.bulk.tensor costs ~150 clocks of the TMA unit plus ~1.3 per 128-byte row) runs
//               across phases; it also stages the NEXT phase's descriptor in shared memory (the kernel parameters are reached
//               through generic loads) and runs the grid barrier
//   warp 1    : tcgen05.mma issuer (UMMA 128 x 32 x 16, fp32 accumulator in TMEM); during a barrier it requests the weight
//               tiles of the next phase's first item (weights do not depend on activations), off the barrier's path
//   warps 2-5 : epilogue (one TMEM load per tile, additive rows requested before the accumulator wait), row phases, softmax,
//               the step's Philox draws (CTAs without a head, last attention phase)
// Activations produced inside the kernel are read with ld.global.cg or by TMA (L2), never through L1.  The kernel is
// launched cooperatively (every CTA resident).  Every wait is bounded; a timeout raises an abort flag that drains the grid.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "epilogue.cuh"
#include "tc_ptx.cuh"

int tc_init(ldm_ctx* ctx);
int tc_make_act_map(const void* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
int attn_tc_supported(int hd);
int tc_make_kblock_map(const void* base, int rows, int cols, int ld, int box_rows, int kblocks, CUtensorMap* out);

namespace {

constexpr int BM = 128, BK = 64, BN = 32;
constexpr int kThreads = 192;
constexpr int kKB = 2;                    // k-blocks per ring stage: one TMA instruction each for the operand and the weight tiles
constexpr int kStages = 4;
constexpr uint32_t kATile = BM * BK * 2, kWTile = BN * BK * 2;
constexpr uint32_t kABytes = kKB * kATile, kWBytes = kKB * kWTile, kStageBytes = kABytes + kWBytes;
constexpr uint32_t kRingBytes = kStages * kStageBytes;                  // 160 KiB; the attention tiles reuse the same region
constexpr int kHeads = 8;
constexpr int kMaxPhases = 4 * LDM_MAX_STAGES + 3;
constexpr int kMaxMaps = 6 * LDM_MAX_STAGES + 6;
constexpr int kMaxParts = 8;
constexpr int kTraceSlots = 8;
constexpr uint32_t kColS = 128, kColO = 256;                            // TMEM columns: GEMM accumulator 0..31, S, O
enum { PH_GEMM = 0, PH_MID = 1, PH_ATTN = 2, PH_LNF = 3 };
enum { EP_HU = 0, EP_QKV = 1, EP_FIN = 2 };

struct alignas(16) LoopPhase {
  int kind;              // PH_*
  int ekind;             // GEMM: EP_*
  int amap, wmap;        // GEMM: operand / weight tensor map;  ATTN: [Q | K] / V^T tensor map
  int K, N, ks;          // GEMM: reduction length, outputs, k-parts
  int d, hd;             // MID / LNF: row width;  ATTN: model width and head width
  int parts;             // MID / LNF: fp32 partials to add
  int noise;             // ATTN: the CTAs without a head draw the step's Philox noise (consumed by EP_FIN)
  const float* tab_t;    // EP_HU: (n_t, N) per-timestep term, bias folded in
  const float* tab_c;    // EP_HU: (pairs, N) per-condition term
  const float* bias;     // EP_QKV / EP_FIN: [N]
  float* part;           // [parts][128][N] fp32: EP_HU output / input of MID and LNF
  int ld_part;
  const float *ga, *ba, *gb, *bb;
  bf16* o1; int ld1;     // MID: h2;  LNF: LN_f(h);  ATTN: a
  bf16* o2; int ld2;     // MID: LN_b(h2)
};

struct LoopParams {
  LoopPhase ph[kMaxPhases];
  int n_phases;
  int B, latent, n_t;
  int n_iter, t_start, sample;
  const int64_t* t_idx;        // forward(): device timesteps (t_len = 1 or B); sampling: null
  int t_len;
  const int32_t* cls;          // [B] condition pair of each row, or null
  const float4* coef;          // [n_steps] (c2, sqrt_alpha, sigma, 0)
  const float* noise;          // explicit draws: n_iter slabs of (B, latent), or null (Philox)
  size_t noise_slab;
  const unsigned long long* rng;   // {seed, sample_offset}
  float* x;                    // (B, latent) fp32 chain state (sampling)
  float* eps_out;              // (B, latent) fp32 (forward())
  bf16* x0;                    // (128, latent) bf16 operand copy of x
  bf16* qk;                    // (128, 2 d) [Q | K]
  bf16* vt;                    // (d, 128) V^T
  float* zbuf;                 // (128, latent) the step's draws
  const CUtensorMap* gmaps;    // tensor maps of all phases (global memory, 128-byte entries)
  unsigned int* sync;          // grid barrier counter (zero at launch)
  int* err;                    // [2]: abort code, block
  long long* trace;            // debug (LDM_V3LOOP_TRACE=1): [G][kMaxPhases][kTraceSlots] globaltimer stamps of step 1, null = off
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four epilogue warps

__device__ __forceinline__ void flag_abort(const LoopParams& P, int code) {
  if (atomicCAS(P.err, 0, code) == 0) P.err[1] = (int)blockIdx.x;
}
// bounded mbarrier wait that also raises the grid-wide abort flag: the probe is inline, the polling loop is one shared copy
// (instruction-cache footprint; a call costs register traffic, which only matters when the barrier is already complete)
__device__ __noinline__ bool wait_bar_slow(const LoopParams& P, uint64_t* bar, uint32_t parity, int code) {
  if (tc::mbar_wait(bar, parity, code)) return true;
  flag_abort(P, code);
  return false;
}
__device__ __forceinline__ bool wait_bar(const LoopParams& P, uint64_t* bar, uint32_t parity, int code) {
  if (tc::mbar_try_wait(bar, parity)) return true;
  return wait_bar_slow(P, bar, parity, code);
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void store_bf16x4(bf16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2(a, b), pack2(c, d));
}
// 32 lanes x 32 consecutive fp32 columns in one TMEM load
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------------------------- epilogues
// thread = row of the 128 x 32 accumulator tile.  The additive rows of a tile (tables, bias, chain state) are requested BEFORE
// the wait for the accumulator, so their L2 latency (the L1 does not survive the acquire of a grid barrier) hides under the MMAs.
__device__ __forceinline__ void epi_prefetch(const LoopParams& P, const LoopPhase& ph, int row, int n0, int kp, int t, float4 (&pre)[8]) {
  if (ph.ekind == EP_HU) {
    if (kp == 0) {   // per-timestep row (bias folded in) + per-condition row
      long long tr = t;
      if (P.t_idx) tr = P.t_idx[P.t_len == 1 ? 0 : row];
      tr = tr < 0 ? 0 : (tr >= P.n_t ? P.n_t - 1 : tr);          // validated upstream; the clamp keeps reads in bounds
      const float* tt = ph.tab_t + (size_t)tr * ph.N + n0;
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] = ldg4(tt + 4 * j);
      if (P.cls) {
        const float* tcp = ph.tab_c + (size_t)P.cls[row] * ph.N + n0;
        float4 c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = ldg4(tcp + 4 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) pre[j] = add4(pre[j], c[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = ldg4(ph.bias + n0 + 4 * j);
  }
}
// [h | u] partial of k-part kp
__device__ __forceinline__ void epi_hu(const LoopPhase& ph, const float (&v)[32], const float4 (&pre)[8], int row, int n0, int kp) {
  float4* out = reinterpret_cast<float4*>(ph.part + ((size_t)kp * BM + row) * ph.ld_part + n0);
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = make_float4(v[4 * j] + pre[j].x, v[4 * j + 1] + pre[j].y, v[4 * j + 2] + pre[j].z, v[4 * j + 3] + pre[j].w);
}
// in_proj: [Q | K] bf16 row-major, V transposed (consecutive lanes hold consecutive rows: every column is one contiguous run)
__device__ __forceinline__ void epi_qkv(const LoopParams& P, const LoopPhase& ph, float (&v)[32], const float4 (&pre)[8], int row, int n0) {
  const int two_d = 2 * ph.K;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[4 * j] += pre[j].x; v[4 * j + 1] += pre[j].y; v[4 * j + 2] += pre[j].z; v[4 * j + 3] += pre[j].w;
  }
  if (n0 < two_d) {
    uint4* dst = reinterpret_cast<uint4*>(P.qk + (size_t)row * two_d + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      dst[j] = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
  } else {
    bf16* dst = P.vt + (size_t)(n0 - two_d) * BM + row;
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[(size_t)i * BM] = __float2bfloat16_rn(v[i]);
  }
}
// eps = final(LN_f(h)): stored (forward()) or consumed by the posterior update (v3:880-887), which also refreshes the bf16 operand
__device__ __forceinline__ void epi_fin(const LoopParams& P, float (&v)[32], const float4 (&pre)[8], int row, int n0, int t, int step) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[4 * j] += pre[j].x; v[4 * j + 1] += pre[j].y; v[4 * j + 2] += pre[j].z; v[4 * j + 3] += pre[j].w;
  }
  const size_t off = (size_t)row * P.latent + n0;
  if (!P.sample) {
#pragma unroll
    for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(P.eps_out + off)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    return;
  }
  const float4 cf = P.coef[t];
  float4 z[8], xv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) xv[j] = reinterpret_cast<const float4*>(P.x + off)[j];     // written by this thread one step ago
  if (cf.z > 0.0f) {
    const float* zs = (P.noise ? P.noise + (size_t)step * P.noise_slab : P.zbuf) + off;
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = ldcg4(zs + 4 * j);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[4 * j] = ddpm_update_one(xv[j].x, v[4 * j], cf.x, cf.y, cf.z, z[j].x);
    v[4 * j + 1] = ddpm_update_one(xv[j].y, v[4 * j + 1], cf.x, cf.y, cf.z, z[j].y);
    v[4 * j + 2] = ddpm_update_one(xv[j].z, v[4 * j + 2], cf.x, cf.y, cf.z, z[j].z);
    v[4 * j + 3] = ddpm_update_one(xv[j].w, v[4 * j + 3], cf.x, cf.y, cf.z, z[j].w);
    reinterpret_cast<float4*>(P.x + off)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  uint4* dst = reinterpret_cast<uint4*>(P.x0 + off);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    dst[j] = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
}

// ---------------------------------------------------------------------------------------------------------------- row phases
// One row per CTA and pass: the four epilogue warps hold the row (<= 2 float4 per thread and array), statistics through
// shared memory.  All loads of a pass are independent (a warp-per-row version spent 11 us on 32 serialised L2 round trips).
__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
  v = warp_sum(v);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  epi_bar();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
__device__ __forceinline__ void block_stats(const float4 (&v)[2], const bool (&act)[2], int d, float* red, int tid, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (act[k]) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  mean = block_sum(s, red, tid) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (act[k]) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, e = v[k].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  rstd = 1.0f / sqrtf(block_sum(q, red + 4, tid) / (float)d + 1e-5f);
}
__device__ __forceinline__ float4 ln4(float4 v, float mean, float rstd, float4 g, float4 b) {
  return make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y, (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w);
}

// h2 = swish(LN_a(u)) + h ; n = LN_b(h2)    (v3:826-830); [h | u] = sum of the k-part partials
__device__ __noinline__ void mid_phase(const LoopPhase& ph, int B, int tid, float* red) {
  const int d = ph.d, nq = d >> 2, ld = ph.ld_part;
  const bool act[2] = {tid < nq, tid + 128 < nq};
  for (int r = (int)blockIdx.x; r < B; r += (int)gridDim.x) {
    const float* base = ph.part + (size_t)r * ld;
    float4 h[2], u[2], ga[2], ba[2], gb[2], bb[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      h[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      u[k] = h[k];
      if (act[k]) {      // with the row loads: one L2 round trip instead of three (no L1 hit after a grid barrier)
        const int q = tid + 128 * k;
        ga[k] = ldg4(ph.ga + 4 * q); ba[k] = ldg4(ph.ba + 4 * q); gb[k] = ldg4(ph.gb + 4 * q); bb[k] = ldg4(ph.bb + 4 * q);
      }
    }
#pragma unroll 1
    for (int pp = 0; pp < ph.parts; pp += 2) {      // two partials in flight
      const float* p = base + (size_t)pp * BM * ld;
      const bool two = pp + 1 < ph.parts;
      float4 a[2], b[2], a2[2], b2[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (act[k]) {
          const float* pk = p + (tid + 128 * k) * 4;
          a[k] = ldcg4(pk);
          b[k] = ldcg4(pk + d);
          a2[k] = two ? ldcg4(pk + (size_t)BM * ld) : make_float4(0.f, 0.f, 0.f, 0.f);
          b2[k] = two ? ldcg4(pk + (size_t)BM * ld + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (act[k]) {
          h[k] = add4(add4(h[k], a[k]), a2[k]);
          u[k] = add4(add4(u[k], b[k]), b2[k]);
        }
    }
    float mean, rstd;
    block_stats(u, act, d, red, tid, mean, rstd);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (act[k]) {
        const int q = tid + 128 * k;
        const float4 n = ln4(u[k], mean, rstd, ga[k], ba[k]);
        h[k] = make_float4(swishf(n.x) + h[k].x, swishf(n.y) + h[k].y, swishf(n.z) + h[k].z, swishf(n.w) + h[k].w);
        store_bf16x4(ph.o1 + (size_t)r * ph.ld1 + 4 * q, h[k].x, h[k].y, h[k].z, h[k].w);
      }
    block_stats(h, act, d, red + 8, tid, mean, rstd);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (act[k]) {
        const int q = tid + 128 * k;
        const float4 n = ln4(h[k], mean, rstd, gb[k], bb[k]);
        store_bf16x4(ph.o2 + (size_t)r * ph.ld2 + 4 * q, n.x, n.y, n.z, n.w);
      }
    epi_bar();     // `red` is reused by the next row
  }
}

// LN_f(h_S + final projections)    (v3:844-850)
__device__ __noinline__ void lnf_phase(const LoopPhase& ph, int B, int tid, float* red) {
  const int d = ph.d, nq = d >> 2, ld = ph.ld_part;
  const bool act[2] = {tid < nq, tid + 128 < nq};
  for (int r = (int)blockIdx.x; r < B; r += (int)gridDim.x) {
    const float* base = ph.part + (size_t)r * ld;
    float4 h[2], ga[2], ba[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      h[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act[k]) {
        ga[k] = ldg4(ph.ga + 4 * (tid + 128 * k));
        ba[k] = ldg4(ph.ba + 4 * (tid + 128 * k));
      }
    }
#pragma unroll 1
    for (int pp = 0; pp < ph.parts; pp += 2) {      // two partials in flight
      const float* p = base + (size_t)pp * BM * ld;
      const bool two = pp + 1 < ph.parts;
      float4 a[2], b[2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (act[k]) {
          a[k] = ldcg4(p + (tid + 128 * k) * 4);
          b[k] = two ? ldcg4(p + (size_t)BM * ld + (tid + 128 * k) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (act[k]) h[k] = add4(add4(h[k], a[k]), b[k]);
    }
    float mean, rstd;
    block_stats(h, act, d, red, tid, mean, rstd);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (act[k]) {
        const int q = tid + 128 * k;
        const float4 n = ln4(h[k], mean, rstd, ga[k], ba[k]);
        store_bf16x4(ph.o1 + (size_t)r * ph.ld1 + 4 * q, n.x, n.y, n.z, n.w);
      }
    epi_bar();
  }
}

// the step's standard-normal draws (v3:884), by the CTAs that have no head in the last attention phase: same counters as
// the per-layer path's fused update (philox.cuh), so both paths produce the same chain
__device__ __noinline__ void noise_phase(const LoopParams& P, int t, int tid) {
  const int qpr = P.latent >> 2, total = P.B * qpr;
  const unsigned long long seed = P.rng[0], off = P.rng[1];
  for (int i = ((int)blockIdx.x - kHeads) * 128 + tid; i < total; i += ((int)gridDim.x - kHeads) * 128) {
    const int row = i / qpr, quad = i - row * qpr;
    const float4 z = philox_normal4(seed, off + (unsigned long long)row, (uint32_t)t, (uint32_t)quad);
    reinterpret_cast<float4*>(P.zbuf + (size_t)row * P.latent)[quad] = z;
  }
}

// ---------------------------------------------------------------------------------------------------------------- attention
// One head per CTA, one key tile (L = B <= 128): S = Q K^T (UMMA 128 x 128 x 16) -> softmax (thread = query row, two passes
// over the S row in TMEM) -> P as bf16 in the swizzled operand layout -> O = P V (UMMA 128 x HD x 16) -> a.  The scheme of
// attn_tc_kernel (gemm_tc.cu) with the barriers of a persistent CTA; the head width is a run-time value (one copy of the code).
struct AttnBars {
  uint64_t *kv, *s, *p, *o;
};
__device__ __noinline__ void attn_phase(const LoopParams& P, const LoopPhase& ph, uint8_t* smem, const AttnBars& br, uint32_t tmem_base,
                                        uint32_t att_n, int warp, int lane, long long* tr) {
  constexpr uint32_t kAtom = 128 * 128;
  const int HD = ph.hd, NA = (HD + 63) / 64;
  const uint32_t kQ = (uint32_t)NA * kAtom, kVAtom = (uint32_t)HD * 128u, kV = 2 * kVAtom;
  const uint32_t kVRegion = (kV + 1023u) & ~1023u;
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + kQ;
  uint8_t* v_s = k_s + kQ;
  uint8_t* p_s = v_s + kVRegion;
  const int h = (int)blockIdx.x, d = ph.d, L = P.B;
  const uint32_t par = att_n & 1u;
  const CUtensorMap* mqk = &P.gmaps[ph.amap];
  const CUtensorMap* mvt = &P.gmaps[ph.wmap];
  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_arrive_expect_tx(br.kv, 2 * kQ + kV);
      for (int at = 0; at < NA; ++at) tc::tma_load_2d(q_s + at * kAtom, mqk, br.kv, h * HD + at * 64, 0);
      for (int at = 0; at < NA; ++at) tc::tma_load_2d(k_s + at * kAtom, mqk, br.kv, d + h * HD + at * 64, 0);
      for (int at = 0; at < 2; ++at) tc::tma_load_2d(v_s + at * kVAtom, mvt, br.kv, at * 64, h * HD);
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc_s = tc::make_idesc_bf16(128, 128), idesc_o = tc::make_idesc_bf16(128, HD);
    wait_bar(P, br.kv, par, 40);
    tc::fence_after_sync();
    if (tr && lane == 0) tr[2] = gtime();
    const uint32_t qa = tc::smem_u32(q_s), ka = tc::smem_u32(k_s), va = tc::smem_u32(v_s), pa = tc::smem_u32(p_s);
    if (tc::elect_one()) {
#pragma unroll 1
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t off = (uint32_t)(ks / 4) * kAtom + (uint32_t)(ks % 4) * 32u;
        tc::umma_bf16(tmem_base + kColS, tc::make_desc_sw128(qa + off), tc::make_desc_sw128(ka + off), idesc_s, (uint32_t)(ks != 0));
      }
      tc::umma_commit(br.s);
    }
    __syncwarp();
    wait_bar(P, br.p, par, 41);
    tc::fence_after_sync();
    if (tc::elect_one()) {
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t kk = (uint32_t)(ks % 4) * 32u;
        tc::umma_bf16(tmem_base + kColO, tc::make_desc_sw128(pa + (uint32_t)(ks / 4) * kAtom + kk),
                      tc::make_desc_sw128(va + (uint32_t)(ks / 4) * kVAtom + kk), idesc_o, (uint32_t)(ks != 0));
      }
      tc::umma_commit(br.o);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = rsqrtf((float)HD);
    uint8_t* p_row = p_s + r * 128;
    wait_bar(P, br.s, par, 42);
    tc::fence_after_sync();
    if (tr && threadIdx.x == 64) tr[3] = gtime();
    // exp((s - max) / sqrt(hd)) = 2^(s c - max c), c = log2(e) / sqrt(hd): one FFMA and one MUFU per element; the key mask only
    // exists for a short call (L < 128)
    const float c = scale * 1.4426950408889634f;
    const bool full = L >= 128;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(t_row + kColS + (uint32_t)c0, v);
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < L) mx = fmaxf(mx, v[i]);
      }
    }
    const float mc = mx * c;
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(t_row + kColS + (uint32_t)c0, v);
      uint32_t pk[16];
      if (full) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2(fmaf(v[2 * i], c, -mc)), p1 = ex2(fmaf(v[2 * i + 1], c, -mc));
          sum += p0 + p1;
          pk[i] = pack2(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = (c0 + 2 * i < L) ? ex2(fmaf(v[2 * i], c, -mc)) : 0.f;
          const float p1 = (c0 + 2 * i + 1 < L) ? ex2(fmaf(v[2 * i + 1], c, -mc)) : 0.f;
          sum += p0 + p1;
          pk[i] = pack2(p0, p1);
        }
      }
      // keys c0 .. c0 + 31 = four 16-byte chunks of atom c0 / 64, XOR-swizzled with the row
      uint8_t* base = p_row + (c0 / 64) * kAtom;
      const int ch = (c0 % 64) / 8;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(base + (((ch + j) ^ (r & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    tc::fence_proxy_async();       // the generic-proxy writes of P must be visible to the tensor core's async proxy
    tc::fence_before_sync();
    tc::mbar_arrive(br.p);
    if (tr && threadIdx.x == 64) tr[4] = gtime();
    wait_bar(P, br.o, par, 43);
    tc::fence_after_sync();
    if (tr && threadIdx.x == 64) tr[5] = gtime();
    const float inv = 1.0f / sum;
    bf16* dst = ph.o1 + (size_t)r * ph.ld1 + d + h * HD;   // [h2 | a]: a starts at column d
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += 16) {
      float v[16];
      tc::tmem_ld16(t_row + kColO + (uint32_t)c0, v);
      if (r < L) {
        uint4* o = reinterpret_cast<uint4*>(dst + c0);
        o[0] = make_uint4(pack2(v[0] * inv, v[1] * inv), pack2(v[2] * inv, v[3] * inv), pack2(v[4] * inv, v[5] * inv), pack2(v[6] * inv, v[7] * inv));
        o[1] = make_uint4(pack2(v[8] * inv, v[9] * inv), pack2(v[10] * inv, v[11] * inv), pack2(v[12] * inv, v[13] * inv), pack2(v[14] * inv, v[15] * inv));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 1) unet3_loop_kernel(const __grid_constant__ LoopParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t acc_full, acc_empty, bar_kv, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  __shared__ float s_red[16];
  // The phase descriptors live in the kernel parameters, which this code reaches through generic loads (~0.4 us each, and the
  // fields of a phase are read in dependent chains): warp 0 copies the NEXT phase into s_next while its own phase runs, and the
  // barrier moves it to s_ph.
  __shared__ LoopPhase s_ph, s_next;
  __shared__ int s_pref;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (int)gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_init(&acc_empty, 128);
    tc::mbar_init(&bar_kv, 1);
    tc::mbar_init(&bar_s, 1);
    tc::mbar_init(&bar_p, 128);
    tc::mbar_init(&bar_o, 1);
    tc::fence_barrier_init();
    s_abort = 0;
  }
  static_assert(sizeof(LoopPhase) % 16 == 0 && sizeof(LoopPhase) <= 32 * 16, "LoopPhase is copied by one warp, 16 bytes per lane");
  auto stage_phase = [&](LoopPhase* dst, int p) {      // one warp
    if (lane < (int)(sizeof(LoopPhase) / 16)) reinterpret_cast<uint4*>(dst)[lane] = reinterpret_cast<const uint4*>(&P.ph[p])[lane];
    __syncwarp();
  };
  if (warp == 1) tc::tmem_alloc<512>(&tmem_slot);
  if (warp == 0) {
    stage_phase(&s_ph, 0);
    stage_phase(&s_next, 0);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const AttnBars abars{&bar_kv, &bar_s, &bar_p, &bar_o};

  uint32_t it = 0;        // k-blocks through the ring so far (warps 0 and 1 walk the same sequence)
  uint32_t acc_n = 0;     // accumulators so far (warp 1 and the epilogue warps)
  uint32_t att_n = 0;     // attention tiles so far
  uint32_t sync_n = 0;    // grid barriers so far
  int pref = 0;           // warp 0: leading k-blocks of the coming GEMM phase whose weight tiles are already in flight

#define V3_STAMP(slot) do { if (P.trace && step == 1) P.trace[((size_t)blockIdx.x * kMaxPhases + p) * kTraceSlots + (slot)] = gtime(); } while (0)

  // weight tiles of the first item of GEMM phase `ph` (they do not depend on activations), in ring order.  Run by warp 1
  // while warp 0 sits in the grid barrier: both warps carry the same ring position `it`.
  auto prefetch_w = [&](const LoopPhase& ph) -> int {
    if (ph.kind != PH_GEMM) return 0;
    const int items = (ph.N / BN) * ph.ks, nks = ph.K / (BK * kKB) / ph.ks;
    if ((int)blockIdx.x >= items) return 0;
    const int tile = (int)blockIdx.x / ph.ks, kp = (int)blockIdx.x - tile * ph.ks;
    const int npre = nks < kStages ? nks : kStages;
    const int rot = (int)blockIdx.x % nks;
#pragma unroll 1
    for (int kb = 0; kb < npre; ++kb) {
      const uint32_t i2 = it + (uint32_t)kb;
      const int s = (int)(i2 % kStages);
      const uint32_t par = (i2 / kStages) & 1u;
      wait_bar(P, &empty_bar[s], par ^ 1u, 10);
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        tc::tma_load_3d(smem + (size_t)s * kStageBytes + kABytes, &P.gmaps[ph.wmap], &full_bar[s], 0, tile * BN, (kp * nks + (kb + rot) % nks) * kKB);
      }
      __syncwarp();
    }
    return npre;
  };

  if (warp == 1) s_pref = prefetch_w(s_next);
  __syncthreads();
  if (warp == 0) pref = s_pref;

  for (int step = 0; step < P.n_iter; ++step) {
    const int t = P.t_start - step;
    for (int p = 0; p < P.n_phases; ++p) {
      const LoopPhase& ph = s_ph;
      if (ph.kind == PH_GEMM) {
        const int items = (ph.N / BN) * ph.ks, nks = ph.K / (BK * kKB) / ph.ks;     // ring stages (kKB k-blocks) per item
        if (warp == 0) {
          bool first = true;
          for (int j = (int)blockIdx.x; j < items; j += G) {
            const int tile = j / ph.ks, kp = j - tile * ph.ks;
            // the CTAs of a phase all stream the SAME operand panel: each starts at its own k-block so that they do not hit
            // the same L2 lines at the same moment (the fp32 summation order is still fixed per tile)
            const int rot = j % nks;
#pragma unroll 1
            for (int kb = 0; kb < nks; ++kb, ++it) {
              const int s = (int)(it % kStages);
              const uint32_t par = (it / kStages) & 1u;
              uint8_t* sa = smem + (size_t)s * kStageBytes;
              const int kc = (kp * nks + (kb + rot) % nks) * kKB;
              const bool pre = first && kb < pref;
              if (!pre) wait_bar(P, &empty_bar[s], par ^ 1u, 11);
              if (tc::elect_one()) {
                if (!pre) {
                  tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                  tc::tma_load_3d(sa + kABytes, &P.gmaps[ph.wmap], &full_bar[s], 0, tile * BN, kc);
                }
                tc::tma_load_3d(sa, &P.gmaps[ph.amap], &full_bar[s], 0, 0, kc);
              }
              __syncwarp();
            }
            first = false;
          }
          pref = 0;
        } else if (warp == 1) {
          constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
          for (int j = (int)blockIdx.x; j < items; j += G) {
            wait_bar(P, &acc_empty, (acc_n & 1u) ^ 1u, 12);      // the epilogue has drained the previous accumulator
            tc::fence_after_sync();
#pragma unroll 1
            for (int kb = 0; kb < nks; ++kb, ++it) {
              const int s = (int)(it % kStages);
              const uint32_t par = (it / kStages) & 1u;
              wait_bar(P, &full_bar[s], par, 13);
              tc::fence_after_sync();
              if (lane == 0 && (kb == 0 || kb == nks - 1)) V3_STAMP(kb == 0 ? 2 : 3);
              if (tc::elect_one()) {
                const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * kStageBytes);
#pragma unroll
                for (int b = 0; b < kKB; ++b) {
                  const uint64_t da = tc::make_desc_sw128(a_addr + b * kATile), dw = tc::make_desc_sw128(a_addr + kABytes + b * kWTile);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    tc::umma_bf16(tmem_base, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (uint32_t)((kb | b | k) != 0));
                }
                tc::umma_commit(&empty_bar[s]);
                if (kb == nks - 1) tc::umma_commit(&acc_full);
              }
              __syncwarp();
            }
            ++acc_n;
          }
        } else {
          const int q = warp & 3, row = q * 32 + lane;
          for (int j = (int)blockIdx.x; j < items; j += G) {
            const int tile = j / ph.ks, kp = j - tile * ph.ks, n0 = tile * BN;
            float4 pre[8];
            if (row < P.B) epi_prefetch(P, ph, row, n0, kp, t, pre);
            wait_bar(P, &acc_full, acc_n & 1u, 14);
            tc::fence_after_sync();
            if (threadIdx.x == 64) V3_STAMP(4);
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
            tc::fence_before_sync();
            tc::mbar_arrive(&acc_empty);      // the accumulator is in registers: the next item's MMAs may start
            ++acc_n;
            if (row < P.B) {
              if (ph.ekind == EP_HU) epi_hu(ph, v, pre, row, n0, kp);
              else if (ph.ekind == EP_QKV) epi_qkv(P, ph, v, pre, row, n0);
              else epi_fin(P, v, pre, row, n0, t, step);
            }
            if (threadIdx.x == 64) V3_STAMP(5);
          }
        }
      } else if (ph.kind == PH_MID) {
        if (threadIdx.x == 64) V3_STAMP(4);
        if (warp >= 2) mid_phase(ph, P.B, (int)threadIdx.x - 64, s_red);
        if (threadIdx.x == 64) V3_STAMP(5);
      } else if (ph.kind == PH_LNF) {
        if (warp >= 2) lnf_phase(ph, P.B, (int)threadIdx.x - 64, s_red);
      } else {   // PH_ATTN
        if ((int)blockIdx.x < kHeads) {
          attn_phase(P, ph, smem, abars, tmem_base, att_n, warp, lane,
                     (P.trace && step == 1) ? P.trace + ((size_t)blockIdx.x * kMaxPhases + p) * kTraceSlots : nullptr);
          ++att_n;
        } else if (ph.noise && warp >= 2 && P.sample && !P.noise && t > 0) {
          noise_phase(P, t, (int)threadIdx.x - 64);
        }
      }

      if (step == P.n_iter - 1 && p == P.n_phases - 1) break;
      if (warp == 0) stage_phase(&s_next, p + 1 < P.n_phases ? p + 1 : 0);     // under the rest of this phase

      // ---- grid-wide barrier: every CTA has finished phase p (its global writes released) before any CTA starts the next
      tc::fence_before_sync();
      __syncthreads();
      if (warp == 0) {
        ++sync_n;
        if (lane == 0) {
          V3_STAMP(0);
          // release at gpu scope: the CTA's writes (ordered before this thread by the barrier above) become visible first
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.sync) : "memory");
          V3_STAMP(6);
          const unsigned int target = sync_n * (unsigned int)G;
          const long long t0 = clock64();
          int bad = 0;
          // one L2 round trip per poll: the abort flag and the clock are looked at every 32nd poll only (a second dependent
          // load per iteration doubled the poll period: ~0.35 us of detection delay per barrier, 19 barriers per step)
          unsigned int polls = 0;
          while (ld_acquire(P.sync) < target) {
            if ((++polls & 31u) == 0) {
              if (ld_relaxed(P.err) != 0) { bad = 1; break; }
              if (clock64() - t0 > 4000000000LL) { flag_abort(P, 20); bad = 1; break; }
            }
          }
          if (!bad && ld_relaxed(P.err) != 0) bad = 1;
          s_abort = bad;
          fence_proxy_async_all();   // peers' released generic-proxy writes -> this CTA's TMA (async-proxy) reads
          V3_STAMP(1);
        }
        __syncwarp();
      } else if (warp == 1) {       // next phase: its descriptor, and the weight tiles of its first item (off the barrier's path)
        if (lane < (int)(sizeof(LoopPhase) / 16)) reinterpret_cast<uint4*>(&s_ph)[lane] = reinterpret_cast<const uint4*>(&s_next)[lane];
        __syncwarp();
        const int npre = prefetch_w(s_next);
        if (lane == 0) s_pref = npre;
      }
      __syncthreads();
      tc::fence_after_sync();
      if (s_abort) goto done;
      if (warp == 0) pref = s_pref;
    }
  }
done:
#undef V3_STAMP
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------- pack
// C(M,N) = A(M,K) . op(B) in fp64 (run once per pack); transB = 0: B (K,N), transB = 1: B (N,K); all row-major with pitches
__global__ void __launch_bounds__(256) fold_mm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int transB,
                                                      float* __restrict__ C, int ldc, int M, int N, int K) {
  __shared__ float As[16][17], Bs[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  double s = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    As[ty][tx] = (m < M && k0 + tx < K) ? A[(size_t)m * lda + k0 + tx] : 0.f;
    if (transB) {
      const int bn = blockIdx.x * 16 + ty;
      Bs[tx][ty] = (bn < N && k0 + tx < K) ? Bm[(size_t)bn * ldb + k0 + tx] : 0.f;
    } else {
      Bs[ty][tx] = (k0 + ty < K && n < N) ? Bm[(size_t)(k0 + ty) * ldb + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) s += (double)As[ty][k] * (double)Bs[k][tx];
    __syncthreads();
  }
  if (m < M && n < N) C[(size_t)m * ldc + n] = (float)s;
}
__global__ void fold_mv_kernel(const float* __restrict__ A, int lda, const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ y,
                               int M, int K) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s = add ? (double)add[m] : 0.0;
  for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * lda + k] * (double)x[k];
  y[m] = (float)s;
}
__global__ void fold_copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
}

int mm(ldm_ctx* ctx, const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K, cudaStream_t st) {
  fold_mm_kernel<<<dim3(ceil_div(N, 16), ceil_div(M, 16)), 256, 0, st>>>(A, lda, B, ldb, transB, C, ldc, M, N, K);
  LDM_LAUNCHED_AS(ctx, "v3loop_pack");
  return 0;
}
int mv(ldm_ctx* ctx, const float* A, int lda, const float* x, const float* add, float* y, int M, int K, cudaStream_t st) {
  fold_mv_kernel<<<ceil_div(M, 128), 128, 0, st>>>(A, lda, x, add, y, M, K);
  LDM_LAUNCHED_AS(ctx, "v3loop_pack");
  return 0;
}
__global__ void fold_add_rows_kernel(float* __restrict__ M, const float* __restrict__ b, int rows, int cols) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)rows * cols) M[i] += b[i % cols];
}
int add_rows(ldm_ctx* ctx, float* M, const float* b, int rows, int cols, cudaStream_t st) {
  const size_t n = (size_t)rows * cols;
  fold_add_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(M, b, rows, cols);
  LDM_LAUNCHED_AS(ctx, "v3loop_pack");
  return 0;
}
int copy2d(ldm_ctx* ctx, const float* src, int lds, float* dst, int ldd, int rows, int cols, cudaStream_t st) {
  const size_t n = (size_t)rows * cols;
  fold_copy2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, rows, cols);
  LDM_LAUNCHED_AS(ctx, "v3loop_pack");
  return 0;
}

struct FoldedGemm {
  int N = 0, K = 0, ks = 1;
  bf16* w = nullptr;          // (N, K) bf16
  float* bias = nullptr;      // [N]
  float* tab_t = nullptr;     // (n_t, N)
  float* tab_c = nullptr;     // (ncls, N)
  int wmap = -1;
};

struct V3Loop {
  std::vector<void*> allocs;
  int nst = 0, grid = 0;
  size_t smem = 0;
  FoldedGemm ga[LDM_MAX_STAGES + 1];      // [h_i | u_i] for i < nst; h_S for i = nst
  bf16* x0 = nullptr;                     // (128, latent) bf16 copy of the chain state
  bf16* ha[LDM_MAX_STAGES] = {nullptr};   // (128, 2 d_i): [h2_i | a_i]
  bf16* nb = nullptr;                     // (128, dmax): LN_b(h2)
  bf16* qk = nullptr;                     // (128, 2 dmax): [Q | K]
  bf16* vt = nullptr;                     // (dmax, 128): V^T
  bf16* lnf = nullptr;                    // (128, d_S): LN_f(h_S)
  float* part = nullptr;                  // [kMaxParts][128][2 dmax] fp32
  float* zbuf = nullptr;                  // (128, latent): the step's standard-normal draws
  CUtensorMap* gmaps = nullptr;           // device copy of `maps`
  CUtensorMap maps[kMaxMaps];
  unsigned int* sync = nullptr;
  int* err = nullptr;
  LoopParams P;                           // maps and the launch-independent phase fields
};

// k-parts of a GEMM phase whose fp32 partials the next row phase adds: enough to keep every CTA's operand stream short
// without leaving the grid (items = tiles x parts <= G) and with at least two k-blocks per part
int pick_parts(int N, int K, int G) {
  const int tiles = N / BN;
  int ks = 1;
  const int nst = K / (BK * kKB);       // ring stages of the whole reduction
  while (ks * 2 <= kMaxParts && tiles * ks * 2 <= G && nst % (ks * 2) == 0 && nst / (ks * 2) >= 2) ks *= 2;
  return ks;
}

}  // namespace

void v3loop_free(ldm_ctx* ctx) {
  V3Loop* M = reinterpret_cast<V3Loop*>(ctx->v3loop);
  if (!M) return;
  for (void* p : M->allocs) cudaFree(p);
  delete M;
  ctx->v3loop = nullptr;
}

int v3loop_supported(ldm_ctx* ctx, int B) {
  return ctx->v3loop != nullptr && ctx->use_v3loop && ctx->precision == LDM_PRECISION_BF16 && ctx->unet.variant == 3 && B >= 1 && B <= BM;
}

int v3loop_error(ldm_ctx* ctx, int* out) {
  *out = 0;
  V3Loop* M = reinterpret_cast<V3Loop*>(ctx->v3loop);
  if (!M) return 0;
  int e[2] = {0, 0};
  LDM_CUDA(cudaMemcpy(e, M->err, sizeof(e), cudaMemcpyDeviceToHost));
  if (e[0]) {
    *out = (e[0] << 16) | (e[1] & 0xFFFF);
    LDM_CUDA(cudaMemset(M->err, 0, sizeof(e)));
  }
  return 0;
}

// Folds the packed v3 layers (UnetModel, fp32 copies) into the phase list.  Leaves ctx->v3loop null (the per-layer path runs)
// when the architecture is outside what the kernel covers.
int v3loop_pack(ldm_ctx* ctx, cudaStream_t st) {
  v3loop_free(ctx);
  UnetModel& U = ctx->unet;
  if (ctx->precision != LDM_PRECISION_BF16 || U.variant != 3) return 0;
  {
    const char* e = getenv("LDM_V3LOOP");
    ctx->use_v3loop = e ? atoi(e) : 1;
    if (!ctx->use_v3loop) return 0;
  }
  const int nst = U.nst;
  if (U.latent % (BK * kKB) != 0 || U.latent % BN != 0) return 0;
  for (int i = 0; i < nst; ++i)
    if (!attn_tc_supported(U.hid[i] / kHeads) || U.hid[i] % (kHeads * 16) != 0) return 0;
  if (4 * nst + 3 > kMaxPhases || 4 * nst + 6 + nst > kMaxMaps) return 0;
  LDM_TRY(tc_init(ctx));

  V3Loop* M = new V3Loop();
  ctx->v3loop = M;
  std::vector<void*> tmp;      // fp32 staging of the folds, released before returning (also on failure)
  auto done_tmp = [&]() { cudaStreamSynchronize(st); for (void* p : tmp) cudaFree(p); tmp.clear(); };
  auto fail = [&](int r) { done_tmp(); v3loop_free(ctx); return r; };
#define V3_TRY(x) do { int r_ = (x); if (r_ != 0) return fail(r_); } while (0)
  M->nst = nst;
  M->smem = kRingBytes + 1024;
  {
    static bool attr = false;
    if (!attr) {
      cudaError_t ce = cudaFuncSetAttribute(unet3_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M->smem);
      if (ce != cudaSuccess) { ldm_set_error("v3loop: %s", cudaGetErrorString(ce)); return fail(-1); }
      attr = true;
    }
    int per_sm = 0;
    cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, unet3_loop_kernel, kThreads, M->smem);
    if (ce != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); v3loop_free(ctx); return 0; }
    int G = ctx->sm_count < 128 ? ctx->sm_count : 128;
    const char* e = getenv("LDM_V3LOOP_GRID");
    if (e && atoi(e) >= 2 * kHeads && atoi(e) <= ctx->sm_count) G = atoi(e);
    if (G < 2 * kHeads) { v3loop_free(ctx); return 0; }   // the CTAs without a head draw the noise
    M->grid = G;
  }
  auto& A = M->allocs;
  const int dmax = U.dmax, L = U.latent;
  auto alloc_z = [&](void** out, size_t bytes) -> int {
    LDM_TRY(ldm_alloc(ctx, A, out, bytes));
    LDM_CUDA(cudaMemsetAsync(*out, 0, bytes, st));
    return 0;
  };
  // ---- workspace: 128 rows, zero once (rows >= B are read by the 128-row TMA boxes and must stay finite)
  V3_TRY(alloc_z((void**)&M->x0, (size_t)BM * L * sizeof(bf16)));
  for (int i = 0; i < nst; ++i) V3_TRY(alloc_z((void**)&M->ha[i], (size_t)BM * 2 * U.hid[i] * sizeof(bf16)));
  V3_TRY(alloc_z((void**)&M->nb, (size_t)BM * dmax * sizeof(bf16)));
  V3_TRY(alloc_z((void**)&M->qk, (size_t)BM * 2 * dmax * sizeof(bf16)));
  V3_TRY(alloc_z((void**)&M->vt, (size_t)dmax * BM * sizeof(bf16)));
  V3_TRY(alloc_z((void**)&M->lnf, (size_t)BM * U.hid[nst] * sizeof(bf16)));
  V3_TRY(alloc_z((void**)&M->part, (size_t)kMaxParts * BM * 2 * dmax * sizeof(float)));
  V3_TRY(alloc_z((void**)&M->zbuf, (size_t)BM * L * sizeof(float)));
  V3_TRY(alloc_z((void**)&M->sync, 64));
  V3_TRY(alloc_z((void**)&M->err, 64));

  LoopParams& P = M->P;
  int nmaps = 0, np = 0;
  auto add_map = [&](const void* base, int rows, int cols, int ld, int box_rows) -> int {
    if (nmaps >= kMaxMaps) { ldm_set_error("v3loop: tensor map table full"); return -1; }
    if (tc_make_act_map(base, rows, cols, ld, box_rows, &M->maps[nmaps]) != 0) return -1;
    return nmaps++;
  };
#define V3_MAP(dst, ...) do { (dst) = add_map(__VA_ARGS__); if ((dst) < 0) { done_tmp(); return fail(-1); } } while (0)
  // GEMM operands and weights: one box = kKB consecutive k-blocks
  auto add_kmap = [&](const void* base, int rows, int cols, int ld, int box_rows) -> int {
    if (nmaps >= kMaxMaps) { ldm_set_error("v3loop: tensor map table full"); return -1; }
    if (tc_make_kblock_map(base, rows, cols, ld, box_rows, kKB, &M->maps[nmaps]) != 0) return -1;
    return nmaps++;
  };
#define V3_KMAP(dst, ...) do { (dst) = add_kmap(__VA_ARGS__); if ((dst) < 0) { done_tmp(); return fail(-1); } } while (0)

  // ---- folded contractions
  for (int i = 0; i <= nst; ++i) {
    const int d = U.hid[i], K = i == 0 ? L : 2 * U.hid[i - 1], N = i < nst ? 2 * d : d;
    FoldedGemm& g = M->ga[i];
    g.N = N; g.K = K;
    g.ks = pick_parts(N, K, M->grid);
    float *wf, *bh;
    if (ldm_alloc_t(ctx, tmp, &wf, (size_t)N * K) != 0 || ldm_alloc_t(ctx, tmp, &bh, (size_t)d) != 0) { done_tmp(); return fail(-1); }
    V3_TRY(ldm_alloc_t(ctx, A, &g.bias, (size_t)N));
    if (i == 0) {   // h_0 = latent_proj(x)                                                   (v3:813)
      V3_TRY(copy2d(ctx, U.latent_proj.w32, L, wf, K, d, L, st));
      V3_TRY(copy2d(ctx, U.latent_proj.b, d, bh, d, 1, d, st));
    } else {        // h_i = down(h2 + out_proj(a) + b_o) = [W_d | W_d W_o] [h2 | a] + (b_d + W_d b_o)   (v3:838-841)
      const int dp = U.hid[i - 1];
      const DenseLayer &Wd = U.down[i - 1], &Wo = U.attn_o[i - 1];
      V3_TRY(copy2d(ctx, Wd.w32, dp, wf, K, d, dp, st));
      V3_TRY(mm(ctx, Wd.w32, dp, Wo.w32, dp, 0, wf + dp, K, d, dp, dp, st));
      V3_TRY(mv(ctx, Wd.w32, dp, Wo.b, Wd.b, bh, d, dp, st));
    }
    V3_TRY(copy2d(ctx, bh, d, g.bias, d, 1, d, st));
    if (i < nst) {  // u_i = block Linear of (h_i + T_i[t] + C_i[c]): every term goes through W_b      (v3:818-825)
      const DenseLayer& Wb = U.block[i];
      V3_TRY(mm(ctx, Wb.w32, d, wf, K, 0, wf + (size_t)d * K, K, d, K, d, st));
      V3_TRY(mv(ctx, Wb.w32, d, bh, Wb.b, g.bias + d, d, d, st));
      V3_TRY(ldm_alloc_t(ctx, A, &g.tab_t, (size_t)U.n_t * N));
      V3_TRY(ldm_alloc_t(ctx, A, &g.tab_c, (size_t)U.ncls * N));
      V3_TRY(copy2d(ctx, U.tab_t[i], d, g.tab_t, N, U.n_t, d, st));
      V3_TRY(mm(ctx, U.tab_t[i], d, Wb.w32, d, 1, g.tab_t + d, N, U.n_t, d, d, st));
      V3_TRY(copy2d(ctx, U.tab_c[i], d, g.tab_c, N, U.ncls, d, st));
      V3_TRY(mm(ctx, U.tab_c[i], d, Wb.w32, d, 1, g.tab_c + d, N, U.ncls, d, d, st));
    } else {        // h_S + final_time_proj + final_class_proj                                        (v3:844-846)
      V3_TRY(ldm_alloc_t(ctx, A, &g.tab_t, (size_t)U.n_t * N));
      V3_TRY(copy2d(ctx, U.tab_t[nst], d, g.tab_t, N, U.n_t, d, st));
      g.tab_c = U.tab_c[nst];
    }
    V3_TRY(add_rows(ctx, g.tab_t, g.bias, U.n_t, N, st));      // the bias rides in the per-timestep rows
    V3_TRY(ldm_alloc_t(ctx, A, &g.w, (size_t)N * K));
    V3_TRY(launch_to_bf16(ctx, wf, g.w, (size_t)N * K, st));
    V3_KMAP(g.wmap, g.w, N, K, K, BN);
  }
  {
    cudaError_t ce = cudaStreamSynchronize(st);
    done_tmp();
    if (ce != cudaSuccess) { ldm_set_error("v3loop pack: %s", cudaGetErrorString(ce)); return fail(-1); }
  }

  // ---- phase list
  int map_x0;
  V3_KMAP(map_x0, M->x0, BM, L, L, BM);
  int map_prev = map_x0;
  auto hu_phase = [&](const FoldedGemm& g, int amap) {
    LoopPhase& ph = P.ph[np++];
    ph.kind = PH_GEMM; ph.ekind = EP_HU; ph.amap = amap; ph.wmap = g.wmap; ph.K = g.K; ph.N = g.N; ph.ks = g.ks;
    ph.part = M->part; ph.ld_part = g.N;
    ph.tab_t = g.tab_t; ph.tab_c = g.tab_c;
  };
  for (int i = 0; i < nst; ++i) {
    const int d = U.hid[i], hd = d / kHeads;
    const FoldedGemm& g = M->ga[i];
    hu_phase(g, map_prev);                                        // [h | u]
    {   // h2, n
      LoopPhase& ph = P.ph[np++];
      ph.kind = PH_MID; ph.d = d; ph.parts = g.ks; ph.part = M->part; ph.ld_part = g.N;
      ph.ga = U.ln_a_w[i]; ph.ba = U.ln_a_b[i]; ph.gb = U.ln_b_w[i]; ph.bb = U.ln_b_b[i];
      ph.o1 = M->ha[i]; ph.ld1 = 2 * d; ph.o2 = M->nb; ph.ld2 = d;
    }
    {   // [Q | K | V]
      LoopPhase& ph = P.ph[np++];
      ph.kind = PH_GEMM; ph.ekind = EP_QKV; ph.K = d; ph.N = 3 * d; ph.ks = 1;
      V3_KMAP(ph.amap, M->nb, BM, d, d, BM);
      V3_KMAP(ph.wmap, U.qkv[i].w16, 3 * d, d, d, BN);
      ph.bias = U.qkv[i].b;
    }
    {   // a
      LoopPhase& ph = P.ph[np++];
      ph.kind = PH_ATTN; ph.d = d; ph.hd = hd; ph.noise = i == nst - 1;
      V3_MAP(ph.amap, M->qk, BM, 2 * d, 2 * d, BM);
      V3_MAP(ph.wmap, M->vt, d, BM, BM, hd);
      ph.o1 = M->ha[i]; ph.ld1 = 2 * d;
    }
    V3_KMAP(map_prev, M->ha[i], BM, 2 * d, 2 * d, BM);
  }
  {
    const int d = U.hid[nst];
    const FoldedGemm& g = M->ga[nst];
    hu_phase(g, map_prev);                                        // h_S
    {   // LN_f
      LoopPhase& ph = P.ph[np++];
      ph.kind = PH_LNF; ph.d = d; ph.parts = g.ks; ph.part = M->part; ph.ld_part = g.N;
      ph.ga = U.ln_f_w; ph.ba = U.ln_f_b;
      ph.o1 = M->lnf; ph.ld1 = d;
    }
    {   // eps = final(LN_f(h_S))  (`return out`, v3:853) [+ posterior update v3:880-893]
      LoopPhase& ph = P.ph[np++];
      ph.kind = PH_GEMM; ph.ekind = EP_FIN; ph.K = d; ph.N = L; ph.ks = 1;
      V3_KMAP(ph.amap, M->lnf, BM, d, d, BM);
      V3_KMAP(ph.wmap, U.fin.w16, L, d, U.fin.K, BN);      // the first half of [W_f | s W_f]
      ph.bias = U.fin.b;
    }
  }
  P.n_phases = np;
  P.latent = L; P.n_t = U.n_t;
  P.x0 = M->x0; P.qk = M->qk; P.vt = M->vt; P.zbuf = M->zbuf;
  V3_TRY(ldm_alloc(ctx, A, (void**)&M->gmaps, sizeof(M->maps)));
  if (cudaMemcpy(M->gmaps, M->maps, sizeof(M->maps), cudaMemcpyHostToDevice) != cudaSuccess) { ldm_set_error("v3loop: tensor map upload failed"); return fail(-1); }
  P.gmaps = M->gmaps;
  P.sync = M->sync;
  P.err = M->err;
#undef V3_TRY
#undef V3_MAP
#undef V3_KMAP
  return 0;
}

// One launch = n_iter steps (sample = 1: t = t_start .. t_start - n_iter + 1 with the posterior update on `x`, which also
// holds the result) or one forward (sample = 0: eps of (x, t_idx) to eps_out).  The bf16 copy of x must NOT be staged by the
// caller: it is made here.
__global__ void v3loop_stage_x_kernel(const float* __restrict__ x, bf16* __restrict__ dst, int n4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  reinterpret_cast<uint2*>(dst)[i] = pk;
}

int launch_v3loop(ldm_ctx* ctx, int B, int n_iter, int t_start, int sample, const int64_t* t_idx, int t_len, float* x, float* eps_out,
                  const float* noise, cudaStream_t st) {
  V3Loop* M = reinterpret_cast<V3Loop*>(ctx->v3loop);
  LDM_CHECK(M != nullptr && v3loop_supported(ctx, B), "v3loop: not available for this call (batch %d)", B);
  LDM_CHECK(n_iter >= 1 && x != nullptr && (sample || (eps_out != nullptr && t_idx != nullptr)), "v3loop: bad arguments");
  UnetModel& U = ctx->unet;
  LoopParams P = M->P;
  P.B = B; P.n_iter = n_iter; P.t_start = t_start; P.sample = sample;
  P.coef = ctx->coef_dev;
  P.noise = noise; P.noise_slab = (size_t)B * U.latent;
  P.cls = ctx->has_cls ? ctx->cls : nullptr;
  P.t_idx = sample ? nullptr : t_idx; P.t_len = t_len;
  P.rng = ctx->rng_dev;
  if (sample) {
    LDM_CHECK(ctx->coef_dev != nullptr, "v3loop: schedule not set");
    P.x = x;
  } else {
    P.eps_out = eps_out;
  }
  const int n4 = B * U.latent / 4;
  v3loop_stage_x_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(x, M->x0, n4);
  LDM_LAUNCHED_AS(ctx, "v3loop_stage_x");
  LDM_CUDA(cudaMemsetAsync(M->sync, 0, sizeof(unsigned int), st));
  static int want_trace = -1;
  if (want_trace < 0) { const char* e = getenv("LDM_V3LOOP_TRACE"); want_trace = e ? atoi(e) : 0; }
  long long* trace = nullptr;
  const size_t trace_n = (size_t)M->grid * kMaxPhases * 8;
  if (want_trace && !ctx->capturing && n_iter > 2) {
    LDM_CUDA(cudaMalloc(&trace, trace_n * sizeof(long long)));
    LDM_CUDA(cudaMemsetAsync(trace, 0, trace_n * sizeof(long long), st));
    P.trace = trace;
  }
  {
    // cooperative launch: the grid barrier needs every CTA resident at once, also when other work shares the device
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(M->grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = M->smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, unet3_loop_kernel, P);
    if (le == cudaErrorCooperativeLaunchTooLarge) {      // the device cannot hold the grid at once (shared / partitioned GPU):
      (void)cudaGetLastError();                          // this context runs v3 through the per-layer sequence from now on
      ctx->use_v3loop = 0;
      return LDM_V3LOOP_UNAVAILABLE;
    }
    LDM_CUDA(le);
  }
  LDM_LAUNCHED_AS(ctx, "unet3_loop");
  if (trace) {   // per phase of step 1: this CTA's work, then its wait at the barrier (ns), for a few CTAs
    std::vector<long long> h(trace_n);
    LDM_CUDA(cudaStreamSynchronize(st));
    LDM_CUDA(cudaMemcpy(h.data(), trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    const int show[4] = {0, 7, 8, M->grid - 1};
    for (int k = 0; k < 4; ++k) {
      const int c = show[k];
      fprintf(stderr, "v3loop trace CTA %3d:", c);
      for (int p = 0; p + 1 < P.n_phases; ++p) {   // relative to the exit of the previous barrier: first tile in, last tile in, accumulator ready, epilogue done | work + wait
        const long long* a = &h[((size_t)c * kMaxPhases + p) * 8];
        const long long prev = p == 0 ? a[0] : h[((size_t)c * kMaxPhases + p - 1) * 8 + 1];
        auto rel = [&](long long v) { return v ? v - prev : -1; };
        fprintf(stderr, " [%d k%d %lld %lld %lld %lld | %lld+%lld(%lld)]", p, P.ph[p].kind, rel(a[2]), rel(a[3]), rel(a[4]), rel(a[5]), a[0] - prev, a[1] - a[0], a[6] - a[0]);
      }
      fprintf(stderr, "\n");
    }
  }
  return 0;
}
