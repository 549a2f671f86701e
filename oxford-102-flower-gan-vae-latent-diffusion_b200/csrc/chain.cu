// The reverse-diffusion chain as ONE persistent sm_100a kernel (bf16 path).
//
//   ConditionalDenoiseDiffusion.sample / p_sample (v2:580-598) calling ConditionalUNet.forward (v2:535-561)
//
// Work decomposition.  Samples are independent, so a thread-block CLUSTER of 16 CTAs owns 32
// batch rows for the WHOLE chain: all T steps run inside one launch and nothing ever crosses a cluster, so there
// is no grid-wide synchronisation and no kernel boundary on the dependency path.  A step is a sequence of
// "phases", one folded dense contraction each (api.cu folds every Linear that sits between two LayerNorms):
//
//   phase 0          x            -> [h_0 | u_0]            u_j = Linear_b,j(h_j) comes out of the SAME GEMM as h_j
//   phase j = 1..S-1 [h2 | n]     -> [h_j | u_j]            h2 = swish(LN_a(u)) + h ; n = LN_b(h2)   (v2:546-553)
//   phase S          [h2 | n]     -> h_S -> LN_f            (v2:554-559)
//   phase S+1        [LN_f | x]   -> eps -> x_{t-1}         (v2:560-561, 584-592; Philox noise generated here)
//
// Inside a phase the OUTPUT FEATURES are split over the CTAs of the cluster in 128-row tiles (UMMA M = 128 rows
// of the weight, UMMA N = the 32 batch rows), accumulators in TMEM.  Weight tiles stream from L2 through a TMA
// ring that never waits for a phase boundary (weights do not depend on activations); the 32 x K bf16 operand is
// exchanged through global memory (it stays in L2), copied by the epilogue warps into the 128-byte-swizzled
// K-major layout the UMMA shared-memory descriptor expects.  LayerNorm needs whole-row statistics: every warp
// publishes per-row partial (mean, M2) pairs into the shared memory of its peers (DSMEM) and the partials are
// merged with Chan's formula.  The cluster-wide rendezvous is an mbarrier in every CTA that the 16 CTAs
// arrive on remotely (release/acquire at cluster scope): only the epilogue warps take part, so the TMA and
// MMA warps keep running ahead.
//
//   warp 0    : TMA producer of weight tiles (all steps, all phases of this CTA, 5-stage ring)
//   warp 1    : TMEM allocator + tcgen05.mma issuer
//   warps 2-5 : operand copy, epilogue (tables, LayerNorm, Swish, residual, DDPM update), cluster rendezvous
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int CS = LDM_CHAIN_CLUSTER;   // CTAs per cluster
constexpr int BNB = LDM_CHAIN_ROWS;     // batch rows per cluster
constexpr int BK = 64;
constexpr int kStages = 5;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t kWBytes = 128 * BK * 2;               // one weight k-block: 16 KiB
constexpr uint32_t kActKbBytes = BNB * BK * 2;           // one operand k-block: 4 KiB
constexpr int kMaxK = LDM_CHAIN_MAX_K;
constexpr uint32_t kActBytes = (uint32_t)BNB * kMaxK * 2;  // 128 KiB
constexpr int kSlots = 32;                               // partial-statistics slots per buffer
constexpr uint32_t kSlotBytes = 2u * kSlots * BNB * sizeof(float2);   // two buffers: 16 KiB
constexpr size_t kSmemBytes = 1024 + kActBytes + (size_t)kStages * kWBytes + kSlotBytes;

struct ChainPhase {
  int type;            // LDM_PH_*
  int K;               // reduction length, multiple of 64
  int tiles;           // 128-row weight tiles
  int first;           // cluster rank of the CTA that owns tile 0 (tile i -> rank first + i)
  int d;               // LayerNorm width (stage: d_j; final-LN / eps: latent)
  int rows;            // tiles * 128: leading dimension of the tables
  const float* bias;   // [rows]       tile order
  const float* tab_t;  // [n_t, rows]  tile order (null: none)
  const float* tab_c;  // [ncls, rows] tile order (null: none)
  const float *ga, *ba, *gb, *bb;   // LayerNorm affine parameters, natural feature order
  const bf16* in;      // operand of this phase (B, ld_in); null -> af[step parity] (+ in_off)
  int ld_in, in_off;
  bf16* out;           // operand this phase produces (stage phases)
  int ld_out;
};

struct ChainParams {
  CUtensorMap wmap[LDM_CHAIN_MAX_PHASES];
  ChainPhase ph[LDM_CHAIN_MAX_PHASES];
  int n_phases;
  int B;
  int n_iter;                 // reverse steps (sample) or 1 (forward)
  int t_start;                // sample: step `it` runs timestep t_start - it
  int sample;                 // 1: fused posterior update on x; 0: write eps_out
  int latent;
  int n_t;
  const int64_t* t_idx;       // forward: device timesteps, t_len = 1 or B
  int t_len;
  const int32_t* cls;         // validated class of each row, or null (c = None)
  float* x;                   // (B, latent) fp32 chain state
  float* eps_out;             // (B, latent)
  const float* noise;         // explicit draws (n_iter, B, latent) or null -> Philox
  const unsigned long long* rng;   // {seed, sample_offset}
  const float4* coef;         // [n_steps]: (c2, sqrt_alpha, sigma, 0)
  bf16* af[2];                // (B, ld_af): [LN_f(h) | x] operand of the last phase, double-buffered over steps
  int ld_af;
  int* err;                   // [2]: first failure code, detail
  long long* trace;           // profiling aid: [CS][64] clock64 stamps of cluster 0 in step trace_step (null: off)
  int trace_step;
};

// ------------------------------------------------------------------------------------------- cluster PTX
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void remote_st_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void remote_st_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Every wait in this kernel is bounded and abortable: the first timeout raises the abort flag of all CTAs of the
// cluster, after which every wait returns at once and the kernel drains to its end (no hung GPU).  The epilogue
// warps share named barriers, so they never leave individually: they finish the step on whatever data they have
// and take ONE uniform decision per step (end of the step loop body).
struct Waiter {
  volatile int* abort_flag;   // this CTA's flag (shared memory)
  int* err;
  __device__ __forceinline__ bool aborted() const { return *abort_flag != 0; }
  __device__ __noinline__ void fail(int code) const {
    if (atomicCAS(err, 0, code) == 0) err[1] = (int)(blockIdx.x + blockIdx.y * gridDim.x);
    const uint32_t a = tc::smem_u32(const_cast<int*>(abort_flag));
    for (uint32_t r = 0; r < (uint32_t)CS; ++r) remote_st_u32(mapa_u32(a, r), 1u);
  }
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code) const {
    if (aborted()) return false;
    for (uint32_t it = 0; it < (1u << 19); ++it) {
      if (tc::mbar_try_wait(bar, parity)) return true;
      if (it > 32) {
        if (aborted()) return false;
        __nanosleep(it > 2048 ? 128 : 20);
      }
    }
    fail(code);
    return false;
  }
  __device__ __forceinline__ bool wait_cluster(uint64_t* bar, uint32_t parity, int code) const {
    if (aborted()) return false;
    for (uint32_t it = 0; it < (1u << 19); ++it) {
      if (try_wait_cluster(bar, parity)) return true;
      if (it > 32) {
        if (aborted()) return false;
        __nanosleep(it > 2048 ? 128 : 20);
      }
    }
    fail(code);
    return false;
  }
};

// 32 lanes x 32 consecutive fp32 accumulator columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Transposed warp reduction: every lane holds 32 values (one per batch row j); on return lane j holds the sum over
// the 32 lanes of value j.  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int n = 16; n >= 1; n >>= 1) {
    const bool upper = (lane & n) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? v[i] : v[i + n];
      const float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
    }
  }
  return v[0];
}

// (sum, sum of squares) over this warp's 32 features for each of the 32 rows -> lane j: (mean, M2) of row j
__device__ __forceinline__ float2 warp_row_stats(const float (&v)[32], int lane) {
  float a[32], b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = v[j]; b[j] = v[j] * v[j]; }
  const float s1 = warp_transpose_sum(a, lane);
  const float s2 = warp_transpose_sum(b, lane);
  const float m = s1 * (1.0f / 32.0f);
  return make_float2(m, fmaxf(s2 - s1 * m, 0.0f));
}

// publish this warp's partial statistics (lane j = row j) into slot `slot` of buffer `buf` of ranks [first, first+n)
__device__ __forceinline__ void publish_stats(float2* slots, int buf, int slot, int lane, float2 st, int first, int n) {
  const uint32_t local = tc::smem_u32(slots + ((size_t)buf * kSlots + slot) * BNB + lane);
  for (int r = 0; r < n; ++r) remote_st_f2(mapa_u32(local, (uint32_t)(first + r)), st.x, st.y);
}

// merge `n` partials of 32 features each (Chan et al.): lane j -> (mean, rstd) of row j over d = 32 n features
__device__ __forceinline__ float2 combine_stats(const float2* slots, int buf, int n, int lane) {
  const float2* p = slots + (size_t)buf * kSlots * BNB + lane;
  float msum = 0.f, m2 = 0.f;
  for (int k = 0; k < n; ++k) msum += p[(size_t)k * BNB].x;
  const float mean = msum / (float)n;
  for (int k = 0; k < n; ++k) {
    const float2 s = p[(size_t)k * BNB];
    const float dm = s.x - mean;
    m2 += s.y + 32.0f * dm * dm;
  }
  const float var = m2 / (32.0f * (float)n);
  return make_float2(mean, 1.0f / sqrtf(var + 1e-5f));
}

// global (B, ld) bf16 rows [row0, row0+32) x K  ->  shared K-major operand, k-blocks of 64 elements, 128-byte swizzle
__device__ __forceinline__ void load_operand(uint8_t* act_s, const bf16* __restrict__ in, int ld, int K, int row0, int B,
                                             int et) {
  const int cpr = K >> 3;                 // 16-byte chunks per row
  const int total = BNB * cpr;
  for (int base = 0; base < total; base += kEpiThreads * 8) {
    uint4 buf[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * kEpiThreads + et;
      buf[u] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < total) {
        const int r = idx / cpr, kc = idx - r * cpr;
        if (row0 + r < B) buf[u] = __ldcg(reinterpret_cast<const uint4*>(in + (size_t)(row0 + r) * ld + (size_t)kc * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * kEpiThreads + et;
      if (idx < total) {
        const int r = idx / cpr, kc = idx - r * cpr;
        const uint32_t off = (uint32_t)(kc >> 3) * kActKbBytes + (uint32_t)r * 128u + (uint32_t)(((kc & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(act_s + off) = buf[u];
      }
    }
  }
}

__device__ __forceinline__ int clamp_t(long long t, int n_t) { return (int)(t < 0 ? 0 : (t >= n_t ? n_t - 1 : t)); }

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_s = smem;
  uint8_t* ring = smem + kActBytes;
  float2* slots = reinterpret_cast<float2*>(ring + (size_t)kStages * kWBytes);
  float* ysm = reinterpret_cast<float*>(act_s);     // [32 rows][64 features]: aliases the operand once the MMAs are done
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t act_full_bar, tmem_full_bar, xbar;
  __shared__ uint32_t tmem_slot;
  __shared__ int abort_flag, abort_decision;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x;                 // gridDim.x == CS: rank in cluster
  const int row0 = blockIdx.y * BNB;           // first batch row of this cluster
  const Waiter W{&abort_flag, P.err};

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&act_full_bar, kEpiThreads);
    tc::mbar_init(&tmem_full_bar, 1);
    tc::mbar_init(&xbar, CS);
    abort_flag = 0;
    tc::fence_barrier_init();
    for (int p = 0; p < P.n_phases; ++p) tc::prefetch_tmap(&P.wmap[p]);
  }
  if (warp == 1) tc::tmem_alloc<32>(&tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  cluster_sync_all();   // every CTA's barriers exist before any remote arrive / store

  if (warp == 0) {
    // ------------------------------------------------------------------ weight-tile producer
    if (lane == 0) {
      uint32_t n = 0;
      bool ok = true;
      for (int it = 0; it < P.n_iter && ok; ++it) {
        for (int p = 0; p < P.n_phases && ok; ++p) {
          const int tile = rank - P.ph[p].first;
          if (tile < 0 || tile >= P.ph[p].tiles) continue;
          const int nkb = P.ph[p].K / BK;
          for (int kb = 0; kb < nkb; ++kb, ++n) {
            const uint32_t s = n % kStages, par = (n / kStages) & 1u;
            if (!W.wait(&empty_bar[s], par ^ 1u, 1)) { ok = false; break; }
            tc::mbar_arrive_expect_tx(&full_bar[s], kWBytes);
            tc::tma_load_2d(ring + (size_t)s * kWBytes, &P.wmap[p], &full_bar[s], kb * BK, tile * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, BNB);
      uint32_t n = 0, an = 0;
      bool ok = true;
      for (int it = 0; it < P.n_iter && ok; ++it) {
        for (int p = 0; p < P.n_phases && ok; ++p) {
          const int tile = rank - P.ph[p].first;
          if (tile < 0 || tile >= P.ph[p].tiles) continue;
          if (!W.wait(&act_full_bar, an & 1u, 2)) { ok = false; break; }
          ++an;
          tc::fence_after_sync();
          const int nkb = P.ph[p].K / BK;
          for (int kb = 0; kb < nkb; ++kb, ++n) {
            const uint32_t s = n % kStages, par = (n / kStages) & 1u;
            if (!W.wait(&full_bar[s], par, 3)) { ok = false; break; }
            tc::fence_after_sync();
            const uint64_t dw = tc::make_desc_sw128(tc::smem_u32(ring + (size_t)s * kWBytes));
            const uint64_t dx = tc::make_desc_sw128(tc::smem_u32(act_s + (size_t)kb * kActKbBytes));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_bf16(tmem_base, dw + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            tc::umma_commit(&empty_bar[s]);
          }
          if (ok) tc::umma_commit(&tmem_full_bar);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ operand copy + epilogue + rendezvous
    const int et = threadIdx.x - 64;          // 0..127
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int lrow = q * 32 + lane;           // row of the 128-row weight tile this thread finishes
    const uint32_t xbar_local = tc::smem_u32(&xbar);
    uint32_t xpar = 0, tpar = 0;
    int tr_n = 0;
    bool tr_on = false;
    auto stamp = [&]() {
      if (tr_on && tr_n < 64) P.trace[rank * 64 + tr_n++] = clock64();
    };

    auto rendezvous = [&](int code) {
      epi_bar_sync();                         // this CTA's epilogue threads are done writing
      if (et < CS) remote_arrive(mapa_u32(xbar_local, (uint32_t)et));
      W.wait_cluster(&xbar, xpar, code);
      xpar ^= 1u;
      stamp();
    };

    for (int it = 0; it < P.n_iter; ++it) {
      const int par = it & 1;
      tr_on = P.trace != nullptr && blockIdx.y == 0 && it == P.trace_step && et == 0;
      stamp();
      const int t_uni = P.sample ? P.t_start - it : (P.t_len == 1 ? clamp_t(P.t_idx[0], P.n_t) : -1);
      for (int p = 0; p < P.n_phases; ++p) {
        const ChainPhase& ph = P.ph[p];
        const int tile = rank - ph.first;
        const bool active = tile >= 0 && tile < ph.tiles;
        float v[32];
        if (active) {
          const bf16* in = ph.in ? ph.in : P.af[par] + ph.in_off;
          load_operand(act_s, in, ph.ld_in, ph.K, row0, P.B, et);
          tc::fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)
          tc::mbar_arrive(&act_full_bar);
          stamp();
          // additive terms of this thread's weight row while the MMAs run
          const int grow = tile * 128 + lrow;
          float add[32];
          {
            const float b = ph.bias ? ph.bias[grow] : 0.f;
            const float tt = (ph.tab_t && t_uni >= 0) ? ph.tab_t[(size_t)t_uni * ph.rows + grow] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float a = b + tt;
              const int r = row0 + j;
              if (r < P.B) {
                if (ph.tab_t && t_uni < 0) a += ph.tab_t[(size_t)clamp_t(P.t_idx[r], P.n_t) * ph.rows + grow];
                if (ph.tab_c && P.cls) a += ph.tab_c[(size_t)P.cls[r] * ph.rows + grow];
              }
              add[j] = a;
            }
          }
          stamp();
          W.wait(&tmem_full_bar, tpar, 4);
          stamp();
          tpar ^= 1u;
          tc::fence_after_sync();
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
          tc::fence_before_sync();            // ordered before the next phase's MMAs through act_full_bar
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += add[j];
        }

        if (ph.type == LDM_PH_STAGE) {
          // lanes 0..63 of the tile: h features; lanes 64..127: u = Linear_b(h) of the SAME features
          const bool is_u = q >= 2;
          const int fl = (q & 1) * 32 + lane;       // feature inside the 64-feature slice
          const int f = tile * 64 + fl;
          const int nparts = ph.tiles * 2;
          if (active && is_u) publish_stats(slots, 0, tile * 2 + (q & 1), lane, warp_row_stats(v, lane), ph.first, ph.tiles);
          rendezvous(5);
          if (active) {
            if (is_u) {   // y = swish(LN_a(u))                                            (v2:520-522)
              const float2 st = combine_stats(slots, 0, nparts, lane);
              const float g = ph.ga[f], b = ph.ba[f];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float m = __shfl_sync(0xffffffffu, st.x, j), r = __shfl_sync(0xffffffffu, st.y, j);
                ysm[j * 64 + fl] = swishf((v[j] - m) * r * g + b);
              }
            }
            epi_bar_sync();
            if (!is_u) {  // h2 = y + h                                                    (v2:547)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += ysm[j * 64 + fl];
              publish_stats(slots, 1, tile * 2 + (q & 1), lane, warp_row_stats(v, lane), ph.first, ph.tiles);
            }
          }
          rendezvous(6);
          if (active && !is_u) {   // n = LN_b(h2); operand of the next phase is [h2 | n]    (v2:548-553)
            const float2 st = combine_stats(slots, 1, nparts, lane);
            const float g = ph.gb[f], b = ph.bb[f];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float m = __shfl_sync(0xffffffffu, st.x, j), r = __shfl_sync(0xffffffffu, st.y, j);
              if (row0 + j < P.B) {
                bf16* o = ph.out + (size_t)(row0 + j) * ph.ld_out;
                o[f] = __float2bfloat16_rn(v[j]);
                o[ph.d + f] = __float2bfloat16_rn((v[j] - m) * r * g + b);
              }
            }
          }
          rendezvous(7);
        } else if (ph.type == LDM_PH_FINAL_LN) {
          const int f = tile * 128 + lrow;
          if (active) publish_stats(slots, 0, tile * 4 + q, lane, warp_row_stats(v, lane), ph.first, ph.tiles);
          rendezvous(8);
          if (active) {   // LN_f(h + T_f[t] + C_f[c])                                     (v2:554-559)
            const float2 st = combine_stats(slots, 0, ph.tiles * 4, lane);
            const float g = ph.ga[f], b = ph.ba[f];
            bf16* o = P.af[par] + f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float m = __shfl_sync(0xffffffffu, st.x, j), r = __shfl_sync(0xffffffffu, st.y, j);
              if (row0 + j < P.B) o[(size_t)(row0 + j) * P.ld_af] = __float2bfloat16_rn((v[j] - m) * r * g + b);
            }
          }
          rendezvous(9);
        } else {   // LDM_PH_EPS: v is eps_theta                                           (v2:560-561)
          if (active) {
            const int f = tile * 128 + lrow;
            if (!P.sample) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (row0 + j < P.B) P.eps_out[(size_t)(row0 + j) * P.latent + f] = v[j];
            } else {
              const int t = P.t_start - it;
              const float4 cf = P.coef[t];
              float z[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) z[j] = 0.f;
              if (cf.z > 0.0f) {
                if (P.noise) {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (row0 + j < P.B) z[j] = P.noise[((size_t)it * P.B + row0 + j) * P.latent + f];
                } else {
                  // the 4 lanes that share a Philox quad split the rows between them, then trade components
                  const unsigned long long seed = P.rng[0], off = P.rng[1] + (unsigned long long)row0;
                  const int sub = lane & 3, base = lane & ~3;
#pragma unroll
                  for (int g = 0; g < 8; ++g) {
                    const float4 z4 = philox_normal4(seed, off + (unsigned long long)(4 * g + sub), (uint32_t)t, (uint32_t)(f >> 2));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const float gx = __shfl_sync(0xffffffffu, z4.x, base + i), gy = __shfl_sync(0xffffffffu, z4.y, base + i);
                      const float gz = __shfl_sync(0xffffffffu, z4.z, base + i), gw = __shfl_sync(0xffffffffu, z4.w, base + i);
                      z[4 * g + i] = sub == 0 ? gx : (sub == 1 ? gy : (sub == 2 ? gz : gw));
                    }
                  }
                }
              }
              bf16* o = P.af[par ^ 1] + P.latent + f;   // x operand of the NEXT step
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (row0 + j < P.B) {
                  float* xp = P.x + (size_t)(row0 + j) * P.latent + f;
                  const float xn = ddpm_update_one(*xp, v[j], cf.x, cf.y, cf.z, z[j]);
                  *xp = xn;
                  o[(size_t)(row0 + j) * P.ld_af] = __float2bfloat16_rn(xn);
                }
              }
            }
          }
          rendezvous(10);
        }
      }
      // uniform abort decision for the 128 epilogue threads
      epi_bar_sync();
      if (et == 0) abort_decision = abort_flag;
      epi_bar_sync();
      if (abort_decision) break;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while a peer may still address its shared memory
  if (warp == 1) tc::tmem_dealloc<32>(tmem_base);
}

// ------------------------------------------------------------------------------------------- pack kernels
// C(M,N) = A(M,K) . op(B) [+ bias per row of C's column?]: fp64 accumulation, run once at pack time.
//   transB = 0: B is (K,N) row-major with pitch ldb;  transB = 1: B is (N,K) row-major with pitch ldb
__global__ void pack_mm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int transB,
                               float* __restrict__ C, int ldc, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  double s = 0.0;
  if (transB) for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * lda + k] * (double)Bm[(size_t)n * ldb + k];
  else for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * lda + k] * (double)Bm[(size_t)k * ldb + n];
  C[(size_t)m * ldc + n] = (float)s;
}
// y(M) = A(M,K) x + add
__global__ void pack_mv_kernel(const float* __restrict__ A, int lda, const float* __restrict__ x, const float* __restrict__ add,
                               float* __restrict__ y, int M, int K) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s = add ? (double)add[m] : 0.0;
  for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * lda + k] * (double)x[k];
  y[m] = (float)s;
}
__global__ void pack_copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
}
// natural row order [h_0..h_{d-1} | u_0..u_{d-1}] -> tile order (tile s: h[64s..64s+63] | u[64s..64s+63]); half = 0: identity
__device__ __forceinline__ int tile_src_row(int rt, int d, int stage) {
  if (!stage) return rt;
  const int s = rt >> 7, l = rt & 127;
  return l < 64 ? s * 64 + l : d + s * 64 + (l - 64);
}
__global__ void pack_rows_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int K, int d, int stage) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * K) return;
  const int rt = (int)(i / K), k = (int)(i % K);
  dst[i] = __float2bfloat16_rn(src[(size_t)tile_src_row(rt, d, stage) * K + k]);
}
// tables: dst[t][rt] = src[t][src_row(rt)]
__global__ void pack_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int rows, int d, int stage) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * rows) return;
  const int t = (int)(i / rows), rt = (int)(i % rows);
  dst[i] = src[(size_t)t * rows + tile_src_row(rt, d, stage)];
}

#define LDM_LAUNCHED(ctx)             \
  do {                                \
    (ctx)->launches++;                \
    LDM_CUDA(cudaGetLastError());     \
  } while (0)

int mm(ldm_ctx* ctx, const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K,
       cudaStream_t st) {
  pack_mm_kernel<<<dim3(ceil_div(N, 128), M), 128, 0, st>>>(A, lda, B, ldb, transB, C, ldc, M, N, K);
  LDM_LAUNCHED(ctx);
  return 0;
}
int mv(ldm_ctx* ctx, const float* A, int lda, const float* x, const float* add, float* y, int M, int K, cudaStream_t st) {
  pack_mv_kernel<<<ceil_div(M, 128), 128, 0, st>>>(A, lda, x, add, y, M, K);
  LDM_LAUNCHED(ctx);
  return 0;
}
int copy2d(ldm_ctx* ctx, const float* src, int lds, float* dst, int ldd, int rows, int cols, cudaStream_t st) {
  const size_t n = (size_t)rows * cols;
  pack_copy2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, rows, cols);
  LDM_LAUNCHED(ctx);
  return 0;
}

bool g_chain_attr_set = false;

}  // namespace

// Can a 16-CTA cluster of this kernel be scheduled on this device?  (called once per context)
int chain_init(ldm_ctx* ctx) {
  LDM_TRY(tc_init(ctx));
  if (!g_chain_attr_set) {
    LDM_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    LDM_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    g_chain_attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  LDM_CUDA(cudaOccupancyMaxActiveClusters(&n, chain_kernel, &cfg));
  LDM_CHECK(n >= 1, "the device cannot co-schedule a cluster of %d CTAs with %zu bytes of shared memory each", CS, kSmemBytes);
  ctx->chain_max_clusters = n;
  return 0;
}

// Fold the packed fp32 layers of ctx->unet (api.cu) into the per-phase weight tiles and tables.
int chain_pack(ldm_ctx* ctx, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  ChainModel& C = ctx->chain;
  for (void* p : C.allocs) cudaFree(p);
  C = ChainModel();
  const int nst = U.nst, L = U.latent;
  LDM_CHECK(nst + 2 <= LDM_CHAIN_MAX_PHASES, "chain: too many stages");
  LDM_CHECK(L % 128 == 0 && 2 * L <= kMaxK && L / 128 * 4 <= kSlots, "chain: latent_dim %d unsupported", L);
  for (int i = 0; i < nst; ++i)
    LDM_CHECK(U.hid[i] % 64 == 0 && U.hid[i] / 64 <= CS && 2 * U.hid[i] <= kMaxK, "chain: hidden dim %d unsupported", U.hid[i]);
  auto& PA = C.allocs;
  std::vector<void*> tmp;
  auto free_tmp = [&]() { for (void* p : tmp) cudaFree(p); tmp.clear(); };
  int rc = 0;
  auto body = [&]() -> int {
    C.n_phases = nst + 2;
    // ---- natural-order folded matrices, biases and tables, then the tile-order / bf16 copies
    for (int j = 0; j <= nst + 1; ++j) {
      ChainPhaseHost& H = C.ph[j];
      int rows, K, d, stage;
      float *Gn = nullptr, *bn = nullptr, *Tn = nullptr, *Cn = nullptr;   // natural order (temporaries)
      if (j == 0) {
        // [h_0 | u_0] = [W_lp ; W_b0 W_lp] x + [b_lp ; W_b0 b_lp + b_b0] + [T_0 ; W_b0 T_0][t] + [C_0 ; W_b0 C_0][c]
        d = U.hid[0]; rows = 2 * d; K = L; stage = 1;
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Gn, (size_t)rows * K));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &bn, (size_t)rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Tn, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Cn, (size_t)U.ncls * rows));
        LDM_TRY(copy2d(ctx, U.latent_proj.w32, K, Gn, K, d, K, st));
        LDM_TRY(mm(ctx, U.block[0].w32, d, U.latent_proj.w32, K, 0, Gn + (size_t)d * K, K, d, K, d, st));
        LDM_TRY(copy2d(ctx, U.latent_proj.b, 1, bn, 1, d, 1, st));
        LDM_TRY(mv(ctx, U.block[0].w32, d, U.latent_proj.b, U.block[0].b, bn + d, d, d, st));
        LDM_TRY(copy2d(ctx, U.tab_t[0], d, Tn, rows, U.n_t, d, st));
        LDM_TRY(mm(ctx, U.tab_t[0], d, U.block[0].w32, d, 1, Tn + d, rows, U.n_t, d, d, st));
        LDM_TRY(copy2d(ctx, U.tab_c[0], d, Cn, rows, U.ncls, d, st));
        LDM_TRY(mm(ctx, U.tab_c[0], d, U.block[0].w32, d, 1, Cn + d, rows, U.ncls, d, d, st));
        H.type = LDM_PH_STAGE;
      } else if (j <= nst) {
        // D = [W_d | W_d A] (dn x 2dp), db = W_d a + b_d with A, a the folded L = 1 attention of stage j-1
        const int i = j - 1, dp = U.hid[i], dn = U.hid[i + 1];
        const bool last = j == nst;
        d = dn; K = 2 * dp; rows = last ? dn : 2 * dn; stage = last ? 0 : 1;
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Gn, (size_t)rows * K));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &bn, (size_t)rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Tn, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Cn, (size_t)U.ncls * rows));
        LDM_TRY(copy2d(ctx, U.down[i].w32, dp, Gn, K, dn, dp, st));
        LDM_TRY(mm(ctx, U.down[i].w32, dp, U.ov[i].w32, dp, 0, Gn + dp, K, dn, dp, dp, st));
        LDM_TRY(mv(ctx, U.down[i].w32, dp, U.ov[i].b, U.down[i].b, bn, dn, dp, st));
        LDM_TRY(copy2d(ctx, U.tab_t[j], dn, Tn, rows, U.n_t, dn, st));
        LDM_TRY(copy2d(ctx, U.tab_c[j], dn, Cn, rows, U.ncls, dn, st));
        if (!last) {   // u_j rows: W_b,j applied to everything above
          const float* Wb = U.block[j].w32;
          LDM_TRY(mm(ctx, Wb, dn, Gn, K, 0, Gn + (size_t)dn * K, K, dn, K, dn, st));
          LDM_TRY(mv(ctx, Wb, dn, bn, U.block[j].b, bn + dn, dn, dn, st));
          LDM_TRY(mm(ctx, U.tab_t[j], dn, Wb, dn, 1, Tn + dn, rows, U.n_t, dn, dn, st));
          LDM_TRY(mm(ctx, U.tab_c[j], dn, Wb, dn, 1, Cn + dn, rows, U.ncls, dn, dn, st));
        }
        H.type = last ? LDM_PH_FINAL_LN : LDM_PH_STAGE;
      } else {
        // eps = [W_f | s W_f] [LN_f(h) ; x] + (1 + s) b_f : already folded by api.cu (U.fin)
        d = L; rows = U.fin.N; K = U.fin.K; stage = 0;
        Gn = U.fin.w32; bn = U.fin.b;
        H.type = LDM_PH_EPS;
      }
      LDM_CHECK(rows % 128 == 0 && K % BK == 0 && K <= kMaxK && rows / 128 <= CS, "chain: phase %d shape (%d x %d) unsupported", j, rows, K);
      H.K = K; H.rows = rows; H.tiles = rows / 128; H.d = d;
      LDM_TRY(ldm_alloc_t(ctx, PA, &H.w, (size_t)rows * K));
      LDM_TRY(ldm_alloc_t(ctx, PA, &H.bias, (size_t)rows));
      {
        const size_t n = (size_t)rows * K;
        pack_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Gn, H.w, rows, K, d, stage);
        LDM_LAUNCHED(ctx);
        pack_cols_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(bn, H.bias, 1, rows, d, stage);
        LDM_LAUNCHED(ctx);
      }
      if (Tn) {
        LDM_TRY(ldm_alloc_t(ctx, PA, &H.tab_t, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, PA, &H.tab_c, (size_t)U.ncls * rows));
        size_t n = (size_t)U.n_t * rows;
        pack_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Tn, H.tab_t, U.n_t, rows, d, stage);
        LDM_LAUNCHED(ctx);
        n = (size_t)U.ncls * rows;
        pack_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Cn, H.tab_c, U.ncls, rows, d, stage);
        LDM_LAUNCHED(ctx);
      }
      LDM_TRY(tc_make_weight_map(ctx, H.w, rows, K, 128, &H.map));   // box = 128 rows x 64 k
      LDM_CUDA(cudaStreamSynchronize(st));
      free_tmp();
    }
    return 0;
  };
  rc = body();
  if (rc != 0) { cudaStreamSynchronize(st); free_tmp(); return rc; }

  // ---- assign tiles to cluster ranks: heaviest phase first, each into the window with the smallest resulting peak
  {
    double load[CS] = {0};
    int order[LDM_CHAIN_MAX_PHASES];
    for (int j = 0; j < C.n_phases; ++j) order[j] = j;
    auto bytes = [&](int j) { return (double)C.ph[j].K * 256.0 + (double)C.ph[j].K * BNB * 2.0; };   // weight tile + operand
    for (int a = 0; a < C.n_phases; ++a)
      for (int b = a + 1; b < C.n_phases; ++b)
        if (bytes(order[b]) * C.ph[order[b]].tiles > bytes(order[a]) * C.ph[order[a]].tiles) { int t = order[a]; order[a] = order[b]; order[b] = t; }
    for (int a = 0; a < C.n_phases; ++a) {
      ChainPhaseHost& H = C.ph[order[a]];
      int best = 0;
      double best_peak = 1e300, best_sum = 1e300;
      for (int f = 0; f + H.tiles <= CS; ++f) {
        double peak = 0, sum = 0;
        for (int r = f; r < f + H.tiles; ++r) { peak = load[r] > peak ? load[r] : peak; sum += load[r]; }
        if (peak < best_peak - 1e-9 || (peak < best_peak + 1e-9 && sum < best_sum)) { best_peak = peak; best_sum = sum; best = f; }
      }
      H.first = best;
      for (int r = best; r < best + H.tiles; ++r) load[r] += bytes(order[a]);
    }
    C.peak_bytes_per_step = 0;
    for (int r = 0; r < CS; ++r) C.peak_bytes_per_step = load[r] > C.peak_bytes_per_step ? load[r] : C.peak_bytes_per_step;
  }
  C.ready = true;
  return 0;
}

// Run `n_iter` reverse steps (sample = 1) or one forward evaluation (sample = 0) for `B` rows.
// The bf16 operand copy of x must already sit in ctx->af_op[0] (columns [latent, 2 latent)).
int launch_chain(ldm_ctx* ctx, int B, int n_iter, int t_start, int sample, const int64_t* t_idx, int t_len, float* x,
                 float* eps_out, const float* noise, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  ChainModel& C = ctx->chain;
  LDM_CHECK(C.ready, "chain: weights not packed");
  LDM_CHECK(!sample || ctx->coef_dev != nullptr, "chain: schedule not set");
  ChainParams P;
  memset(&P, 0, sizeof(P));
  const int nst = U.nst, L = U.latent;
  for (int j = 0; j < C.n_phases; ++j) {
    const ChainPhaseHost& H = C.ph[j];
    ChainPhase& D = P.ph[j];
    P.wmap[j] = H.map;
    D.type = H.type; D.K = H.K; D.tiles = H.tiles; D.first = H.first; D.d = H.d; D.rows = H.rows;
    D.bias = H.bias; D.tab_t = H.tab_t; D.tab_c = ctx->has_cls ? H.tab_c : nullptr;
    if (j < nst) { D.ga = U.ln_a_w[j]; D.ba = U.ln_a_b[j]; D.gb = U.ln_b_w[j]; D.bb = U.ln_b_b[j]; }
    else if (j == nst) { D.ga = U.ln_f_w; D.ba = U.ln_f_b; }
    if (j == 0) { D.in = nullptr; D.in_off = L; D.ld_in = 2 * L; }
    else if (j <= nst) { D.in = ctx->opbuf[j - 1]; D.ld_in = 2 * U.hid[j - 1]; D.in_off = 0; }
    else { D.in = nullptr; D.in_off = 0; D.ld_in = 2 * L; }
    if (j < nst) { D.out = ctx->opbuf[j]; D.ld_out = 2 * U.hid[j]; }
  }
  P.n_phases = C.n_phases;
  P.B = B; P.n_iter = n_iter; P.t_start = t_start; P.sample = sample; P.latent = L; P.n_t = U.n_t;
  P.t_idx = t_idx; P.t_len = t_len;
  P.cls = ctx->has_cls ? ctx->cls : nullptr;
  P.x = x; P.eps_out = eps_out; P.noise = noise; P.rng = ctx->rng_dev;
  P.coef = ctx->coef_dev;
  P.af[0] = (bf16*)ctx->af_op[0]; P.af[1] = (bf16*)ctx->af_op[1]; P.ld_af = 2 * L;
  P.err = ctx->chain_err;
  P.trace = ctx->chain_trace;
  P.trace_step = ctx->chain_trace_step;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, ceil_div(B, BNB));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LDM_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel, P));
  ctx->launches++;
  return 0;
}
