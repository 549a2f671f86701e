// The reverse-diffusion chain as ONE persistent sm_100a kernel (bf16 path).
//
//   ConditionalDenoiseDiffusion.sample / p_sample (v2:580-598) calling ConditionalUNet.forward (v2:535-561)
//
// Work decomposition.  Samples are independent, so a thread-block CLUSTER of 16 CTAs owns NB = 16 NW
// batch rows (32 / 48 / 64) for the WHOLE chain: all T steps run inside one launch and nothing ever crosses a cluster,
// so there is no grid-wide synchronisation and no kernel boundary on the dependency path.  A step is FIVE "phases", one
// folded dense contraction each (chain_pack below folds every Linear that sits between two non-linearities):
//
//   phase 0 (MERGED)  [x~ | -c_b LN_f(h) | -c_b x]  -> [h_0 | u_0] of the NEXT forward, and eps of the previous one
//                                                       (eps only feeds the fp32 posterior update, off the critical path)
//   phase j = 1..3    h2_{j-1} (raw)                 -> [h_j | u_j]       two accumulators W1 h2, W2 h2: LayerNorm_b (v2:549) is
//                                                       applied AFTER the contraction, out = acc1 + r (acc2 - mu q)
//   phase 4 (FINAL_LN) h2_3 (raw)                    -> h_S -> -c_b LN_f(h_S + T_f[t] + C_f[c])                 (v2:554-559)
//
//   u_j = Linear_b,j(h_j) comes out of the SAME contraction as h_j; h2 = swish(LN_a(u)) + h (v2:546-548) is the epilogue.
//
// Inside a phase the OUTPUT FEATURES are split over the CTAs in 128-row weight tiles (UMMA M = 128 weight rows, N = NB
// batch rows, fp32 accumulators in TMEM); phases with few tiles and a long reduction give each accumulator / K half to its
// own CTA ("units", ks = 2) and the helper ships its accumulator to the tile owner with st.async.
//
// Data movement.  Weights (11 MB bf16, L2 resident) stream through a ring of 32 KiB slots (two tiles = 8 MMAs per slot) that
// runs ahead of the phases.  The NB x K bf16 operand of a phase is written to global memory by the producing CTAs, handed
// over with one release / acquire round on a cluster mbarrier (16 remote arrives per CTA), and TMA-loaded into an operand
// ring of 8 k-block slots: every k-block of the unit is requested the moment the hand-over is through, four k-blocks per TMA
// instruction (a cp.async.bulk.tensor costs the TMA unit ~150 clocks plus ~1.3 per 128-byte row; the first request is armed
// before the hand-over wait).  LayerNorm needs
// whole-row statistics: per-row (mean, M2) partials travel as 8-byte st.async into the peers' shared memory (completion on
// a transaction mbarrier) and are merged with Chan's formula; the statistics of h2 travel in the BACKGROUND to the next
// phase's tile owners.  Measured limits that shape all this (tools/ubench.cu): TMA inbound 51.7 B/clk per
// SM whatever the grid, DSMEM bulk copies 15.5 B/clk per SM (so operands cannot be broadcast through DSMEM), one
// tcgen05.mma of M = 128, N <= 128 every 74 clocks whether A comes from shared memory or from TMEM.
//
// The additive terms of a tile (bias + time-table row + per-sample terms) do not wait for the contraction: the epilogue warps
// write them into accumulator 0 of the COMING phase while the hand-over is in flight and that phase's first MMA accumulates
// (LDM_CHAIN_PREINIT; a monotonic shared-memory counter orders the writes before the MMA thread).
//
//   warp 0    : TMA producer of weight tiles (elected thread; per-launch unit plans, no per-slot address arithmetic)
//   warp 1    : operand producer: waits for the phase hand-over, requests the unit's k-blocks
//   warp 2    : TMEM allocator + tcgen05.mma issuer
//   warp 3    : set-up (phase table and unit plans into shared memory)
//   warps 4.. : 4 NW epilogue warps: tables, LayerNorm, Swish, residual, statistics exchange, operand stores, hand-over,
//               posterior update (eps owners)
// Every wait is a hardware-suspended mbarrier try_wait, bounded by a clock and abortable by any CTA of the cluster.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int CS = LDM_CHAIN_CLUSTER;   // CTAs per cluster
constexpr int BK = 64;
constexpr int kCtlThreads = 128;        // warp 0: weight TMA, warp 1: operand TMA, warp 2: TMEM + MMA, warp 3: set-up
constexpr uint32_t kWTile = 128 * BK * 2;                // one weight k-block of a tile: 16 KiB
constexpr uint32_t kWSlotBytes = 2 * kWTile;             // a weight-ring slot carries the two weight tiles of one chunk (= 8 MMAs)
constexpr int kMaxWSlots = 6;
constexpr int kXGroup = 4;                               // operand k-blocks per grouped TMA box (kXSlots % kXGroup == 0)
constexpr int kXSlots = 8;                               // operand ring: k-blocks of NB rows, loaded as soon as the hand-over is through
constexpr int kSlots = CS;                               // partial-statistics slots per buffer: one per tile of a phase
constexpr int kChains = 2;                               // TMEM column blocks of NB fp32 accumulators: acc1, acc2 (dual phases)
constexpr int kTmemCols = 512;
constexpr int kMaxXMaps = LDM_MAX_STAGES + 2;

// per-variant geometry: NW epilogue warps per TMEM lane quadrant, 16 batch rows each
// LDM_CHAIN_PREINIT (build switch, A/B): the additive terms of a tile (bias + time-table row + per-sample terms) are written INTO
// the accumulator by the epilogue warps before the phase's MMAs start (first MMA with accumulate = 1), so that their TMEM read and
// the additions leave the dependency path between the last MMA and the statistics exchange.  NW >= 3 only (terms in TMEM).
// Measured (same box, us per step): B = 256: 36.34 -> 35.86, B = 300: 36.17 -> 35.14.  LDM_CHAIN_PREINIT=0 at build time restores
// the addition in the epilogue.
#ifndef LDM_CHAIN_PREINIT
#define LDM_CHAIN_PREINIT 1
#endif

template <int NW> struct Geo {
  static constexpr int NB = 16 * NW;                                   // batch rows per cluster
  static constexpr int kEpiThreads = 128 * NW;
  static constexpr int kThreads = kCtlThreads + kEpiThreads;
  static constexpr uint32_t kXBytes = NB * BK * 2;                     // one operand k-block (multiple of 1024)
  static constexpr uint32_t kXRingBytes = kXSlots * kXBytes;
  static constexpr uint32_t kSlotBytes = 2u * kSlots * NB * sizeof(float2);
  static constexpr uint32_t kGstatBytes = 2u * NW * 4u * 16u * sizeof(float2);   // per-warp partials of a row group, double-buffered
  static constexpr uint32_t kRowStatBytes = 4u * NW * 16u * sizeof(float2);      // (mean, rstd) of 16 rows per epilogue warp
  static constexpr uint32_t kPbufBytes = 128u * NB * sizeof(float);               // split-K partner's partial accumulator
  static constexpr uint32_t kAuxBytes = kSlotBytes + kGstatBytes + kRowStatBytes + kPbufBytes;
  static constexpr int kWSlotsFit = (int)((226u * 1024u - kAuxBytes - kXRingBytes) / kWSlotBytes);
  static constexpr int kWSlots = kWSlotsFit < kMaxWSlots ? kWSlotsFit : kMaxWSlots;
  static constexpr size_t kSmemBytes = 1024 + (size_t)kWSlots * kWSlotBytes + kXRingBytes + kAuxBytes;
  static_assert(kWSlots >= 2, "weight ring too small");
  // Per-sample additive terms (class-table rows; per-row timesteps of forward()): with 168 registers per thread (NW = 2)
  // they are re-fetched from the L2-resident tables in every phase's preamble, which saves their TMEM read (the
  // accumulator loads of an epilogue run at 64 B/clk per SM); at NW >= 3 (128 registers or fewer) holding 16 more values
  // across the waits spills, and they stay parked in TMEM columns for the whole launch (measured: 38.5 vs 37.0 us / step)
  static constexpr bool kCaddInTmem = NW >= 3;
  static constexpr bool kPreinit = LDM_CHAIN_PREINIT != 0;      // every variant alike: results do not depend on the rows per cluster
};

struct ChainPhase {
  int type;            // LDM_PH_*
  int K;               // reduction length, multiple of 64
  int tiles;           // 128-row weight tiles
  int first;           // cluster rank of the CTA that owns unit 0; unit = tile * ks + k-part -> rank first + unit
  int ks;              // 1, or 2 units per tile (the odd unit sends its accumulator to the even one): a K split (plain
                       // phases) or one unit per accumulator (dual phases)
  int dual;            // 1: operand = raw h2 of the previous stage, LayerNorm applied AFTER the contraction: two weight
                       //    blocks [W1 | W2] (columns [0,K) and [K,2K)) and two accumulators, out = acc1 + r (acc2 - mu q)
  const float* q;      // dual: q = W2 . 1 (tile order)
  int prev_tiles;      // dual: stage tiles of the phase that produced the operand (partials of its row statistics)
  int d;               // LayerNorm width (stage: d_j; final-LN / eps: latent)
  int rows;            // tiles * 128: leading dimension of the tables
  int xmap;            // index of the operand's tensor map (+ 1 on odd steps when xmap_alt)
  int xmap_alt;        // 1: the operand alternates between xmap and xmap + 1 with the step parity ([LN_f(h) | x])
  int xcol;            // first column of the operand inside that buffer
  int cadd_col;        // >= 0: the phase has a per-sample additive term (class table row; per-row timestep of forward()), parked at
                       // this TMEM column when Geo::kCaddInTmem; -1: none
  int nst_tiles;       // MERGED: tiles [0, nst_tiles) are stage tiles, the rest finish eps
  int eps_kb0;         // MERGED: first k-block the eps tiles read (they skip the x~ block)
  const float* g0b;    // MERGED: G_0 . b_fin (tile order)
  const float* bias;   // [rows]       tile order
  const float* tab_t;  // [n_t, rows]  tile order (null: none)
  const float* tab_c;  // [ncls, rows] tile order (null: none)
  const float *ga, *ba, *gb, *bb;   // LayerNorm affine parameters, natural feature order
  bf16* out;           // operand this phase produces (stage phases)
  int ld_out;          // its pitch; with ChainParams::split the row is [hi | lo | hi], each `wout` wide
  int wout;
};

struct ChainParams {
  CUtensorMap wmap[LDM_CHAIN_MAX_PHASES];
  CUtensorMap xmaps[kMaxXMaps];   // [0], [1]: af[0], af[1]; [2 + j]: opbuf[j]
  CUtensorMap xmaps4[kMaxXMaps];  // the same operands as (64, rows, k-block) tensors: one box = kXGroup consecutive k-blocks
  int xgroup;                     // 1: operand k-blocks are requested kXGroup at a time where the ring position allows
  ChainPhase ph[LDM_CHAIN_MAX_PHASES];
  int n_phases;
  int B;                      // rows of the whole batch (leading dimension of `noise` slabs)
  int row_begin, row_end;     // rows this launch works on
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
  const float4* coef_or_one;  // sample: coef; forward(): one entry (1, 1, 0, 0)
  bf16* af[2];                // (B, ld_af): [x~ | -c_b LN_f(h) | -c_b x] operand of the merged phase, double-buffered over steps
  int ld_af;
  int split;                  // strict mode: every operand is stored as bf16 (hi, lo, hi) thirds against weights (hi, hi, lo): the
                              // three-term split hi.hi + hi.lo + lo.hi keeps ~16 bits of the fp32 product on the bf16 tensor cores
  int writer_fence;           // 1: every writer thread issues fence.proxy.async before the hand-over; 0: the operand producer does
  int z_col;                  // TMEM column where the eps owners park the step's noise
  int x_col;                  // TMEM column of the fp32 chain state of the eps owners
  int* err;                   // [2]: first failure code, detail
  long long* trace;           // profiling aid (LDM_CHAIN_TRACE builds): [CS][LDM_CHAIN_TRACE_TRACKS][LDM_CHAIN_TRACE_LEN] tagged clock64 stamps of
                              // cluster 0 in step trace_step (null: off); tracks: 0 h-warp thread, 1 u-warp thread, 2 weight producer, 3 operand producer, 4 MMA issuer
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
__device__ __forceinline__ void remote_st_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// 8-byte store into a peer's shared memory that signals the peer's mbarrier (complete_tx of 8 bytes) when it lands
__device__ __forceinline__ void st_async_f2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void st_async_f4(uint32_t cluster_addr, float a, float b, float c, float d, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(cluster_bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of spinning through issue slots that the working warps of the same scheduler need
constexpr uint32_t kWaitHintNs = 1u << 15;
__device__ __forceinline__ bool try_wait_cta(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(kWaitHintNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_cluster(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(kWaitHintNs)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns, thread i <-> lane (base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
      "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
      "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Every wait in this kernel is bounded and abortable: the first timeout raises the abort flag of all CTAs of the
// cluster, after which every wait returns at once and the kernel drains to its end (no hung GPU).
struct Waiter {
  volatile int* abort_flag;   // this CTA's flag (shared memory)
  int* err;
  __device__ __forceinline__ bool aborted() const { return *abort_flag != 0; }
  __device__ __noinline__ void fail(int code) const {
    if (atomicCAS(err, 0, code) == 0) err[1] = (int)(blockIdx.x + blockIdx.y * gridDim.x);
    const uint32_t a = tc::smem_u32(const_cast<int*>(abort_flag));
    for (uint32_t r = 0; r < (uint32_t)CS; ++r) remote_st_u32(mapa_u32(a, r), 1u);
  }
  // slow path: out of line, so that the fast path is one try_wait and one branch
  __device__ __noinline__ bool wait_slow(uint32_t bar_addr, uint32_t parity, int code, int cluster) const {
    const long long t0 = clock64();
    for (;;) {
      if (cluster ? try_wait_cluster(bar_addr, parity) : try_wait_cta(bar_addr, parity)) return true;
      if (aborted()) return false;
      if (clock64() - t0 > (4ll << 30)) break;   // ~2 s: a barrier of this kernel completes within microseconds
    }
    fail(code);
    return false;
  }
  __device__ __forceinline__ bool wait(uint32_t bar_addr, uint32_t parity, int code) const {
    if (try_wait_cta(bar_addr, parity)) return true;
    return wait_slow(bar_addr, parity, code, 0);
  }
  __device__ __forceinline__ bool wait(uint64_t* bar, uint32_t parity, int code) const { return wait(tc::smem_u32(bar), parity, code); }
  __device__ __forceinline__ bool wait_cluster(uint32_t bar_addr, uint32_t parity, int code) const {
    if (try_wait_cluster(bar_addr, parity)) return true;
    return wait_slow(bar_addr, parity, code, 1);
  }
  __device__ __forceinline__ bool wait_cluster(uint64_t* bar, uint32_t parity, int code) const { return wait_cluster(tc::smem_u32(bar), parity, code); }
};

// Tagged clock stamps of one thread (LDM_CHAIN_TRACE builds): value = clock64 | phase << 52 | tag << 56
struct Tracer {
  long long* buf;
  int n;
  bool on;
  __device__ __forceinline__ void stamp(int tag, int p) {
#ifdef LDM_CHAIN_TRACE
    if (on && n < LDM_CHAIN_TRACE_LEN) buf[n++] = (clock64() & 0xFFFFFFFFFFFFFll) | ((long long)p << 52) | ((long long)tag << 56);
#endif
  }
};

// Transposed warp reduction of 16 values per lane: on return EVERY lane holds the 32-lane total of element (lane >> 1).
// 16 shuffles instead of 16 x 5.
__device__ __forceinline__ float tsum16(float (&a)[16], int lane) {
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? a[i] : a[i + 8], keep = up ? a[i + 8] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? a[i] : a[i + 2], keep = up ? a[i + 2] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = (lane & 2) != 0;
    const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// Transposed reduction of 8 values per lane inside each 16-lane half of the warp: on return every lane holds the
// 16-lane total of element ((lane >> 1) & 7) of its half.  8 shuffles.
__device__ __forceinline__ float tsum8_half(float (&a)[8], int lane) {
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? a[i] : a[i + 2], keep = up ? a[i + 2] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = (lane & 2) != 0;
    const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}
// (mean, M2) over the 16 features of this half-warp, of the half's row ((lane >> 1) & 7)
__device__ __forceinline__ float2 half_row_stats8(const float (&v)[8], int lane) {
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = v[j]; b[j] = v[j] * v[j]; }
  const float s1 = tsum8_half(a, lane);
  const float s2 = tsum8_half(b, lane);
  const float m = s1 * (1.0f / 16.0f);
  return make_float2(m, fmaxf(s2 - s1 * m, 0.0f));
}

// (mean, M2) over this warp's 32 features of batch row (lane >> 1)
__device__ __forceinline__ float2 warp_row_stats16(const float (&v)[16], int lane) {
  float a[16], b[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { a[j] = v[j]; b[j] = v[j] * v[j]; }
  const float s1 = tsum16(a, lane);
  const float s2 = tsum16(b, lane);
  const float m = s1 * (1.0f / 32.0f);
  return make_float2(m, fmaxf(s2 - s1 * m, 0.0f));
}

// v sigmoid(v) = h + h tanh(h), h = v / 2: one MUFU (tanh.approx, relative error ~2^-11, below the bf16 rounding of the result)
__device__ __forceinline__ float swish_fast(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float swish_precise(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// bf16 (hi, lo) parts of an fp32 value: v = hi + lo up to 2^-17 relative
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// operand store: one bf16, or the (hi, lo, hi) thirds `w` columns apart
__device__ __forceinline__ void store_operand(bf16* o, float v, int split, int w) {
  if (!split) { *o = __float2bfloat16_rn(v); return; }
  bf16 hi, lo;
  split_bf16(v, hi, lo);
  o[0] = hi; o[w] = lo; o[2 * w] = hi;
}
__device__ __forceinline__ int clamp_t(long long t, int n_t) { return (int)(t < 0 ? 0 : (t >= n_t ? n_t - 1 : t)); }

// work unit of cluster rank `rank` in a phase: tile, unit-in-tile and its share of the reduction; false: no unit.
// `tail`: the extra merged phase after the last step, in which only the eps tiles work.
//   mode 0 (plain phase)          : k-blocks [kb0, kb0 + nkb) of [W | X] -> accumulator 0
//   mode 1 (dual phase, one unit) : every k-block against W1 (-> accumulator 0) and W2 (-> accumulator 1)
//   mode 2 (dual phase, two units): unit kp multiplies the whole operand with W_{kp+1} into its own accumulator 0
struct Unit { int tile, kp, kb0, nkb, mode; bool is_eps; };
__device__ __forceinline__ bool unit_of(const ChainPhase& ph, int rank, bool tail, Unit& u) {
  const int un = rank - ph.first;
  if (un < 0 || un >= ph.tiles * ph.ks) return false;
  u.tile = un >> (ph.ks - 1);          // ks is 1 or 2
  u.kp = un & (ph.ks - 1);
  u.is_eps = false;
  const int nkb = ph.K / BK;
  if (ph.dual) {
    u.mode = ph.ks == 1 ? 1 : 2;
    u.kb0 = 0;
    u.nkb = nkb;
    return true;
  }
  u.mode = 0;
  int lo = 0;
  if (ph.type == LDM_PH_MERGED) {
    u.is_eps = u.tile >= ph.nst_tiles;
    if (u.is_eps) lo = ph.eps_kb0;
    else if (tail) return false;
  }
  u.nkb = (nkb - lo) >> (ph.ks - 1);
  u.kb0 = lo + u.kp * u.nkb;
  return true;
}

// What the three control threads need to know about this CTA's unit of a phase, worked out ONCE per launch (the threads
// run a long dependent instruction stream per ring slot; everything that can be hoisted out of it is).
// A chunk is the work behind one weight-ring slot: two 128 x 64 weight tiles and 8 MMAs.
//   mode 1      : chunk c = k-block c against W1 and W2 (one operand slot)
//   modes 0, 2  : chunk c = k-blocks 2c, 2c + 1 of one weight matrix (two operand slots; the last chunk may hold one)
struct __align__(16) UnitPlan {
  int valid;        // this CTA has a unit in the phase
  int valid_tail;   // ... and in the tail pass (eps tiles only)
  int tile_row;     // first weight row of the tile
  int nkb;          // operand k-blocks the unit reads
  int mode;
  int wcol0;        // column of the first weight tile (accumulator 0)
  int wcol1;        // mode 1: column of W2's first tile
  int xcol0;        // operand column of k-block 0
  int pre;          // the unit owns its tile and is no eps tile: its accumulator 0 may arrive pre-initialised (Geo::kPreinit)
  int pad_[3];
};

// k-blocks the operand producer requests with one box at ring slot xs when `rem` k-blocks of the unit remain (the MMA issuer
// applies the same rule).  Measured at B = 256: groups of 4 take 5.4 % off the step, the whole ring (8) in one box only 3 % - the
// first MMA of a phase waits for the first box to land.
__device__ __forceinline__ int x_group(int enabled, uint32_t xs, int rem) {
  if (!enabled) return 1;
  if ((xs % kXGroup) == 0 && rem >= kXGroup) return kXGroup;
  return 1;
}

// control-thread PTX on raw shared-memory addresses (no generic -> shared conversions inside the loops)
__device__ __forceinline__ void expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma2d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int NW>
__global__ void __launch_bounds__(Geo<NW>::kThreads, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  using G = Geo<NW>;
  constexpr int NB = G::NB;
  constexpr int S = G::kWSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wring = smem;                                                                                  // [S][2] weight tiles
  uint8_t* xring = smem + (size_t)S * kWSlotBytes;                                                        // [kXSlots] operand k-blocks
  float2* slots = reinterpret_cast<float2*>(xring + G::kXRingBytes);                                      // [2][kSlots][NB]
  float2* gstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(slots) + G::kSlotBytes);          // [2][NW][4 warps][16 rows]
  float2* rowstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(gstat) + G::kGstatBytes);       // [4 NW warps][16 rows]
  float4* pbuf = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(rowstat) + G::kRowStatBytes);      // [NW][4 chunks][128 rows] x 4 columns
  __shared__ __align__(8) uint64_t wfull_bar[kMaxWSlots];
  __shared__ __align__(8) uint64_t wempty_bar[kMaxWSlots];
  __shared__ __align__(8) uint64_t xfull_bar[kXSlots];
  __shared__ __align__(8) uint64_t xempty_bar[kXSlots];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t obar[2];   // phase hand-over: 16 remote arrives (one per CTA) + the local operand producer
  __shared__ __align__(8) uint64_t sbar[2];   // LayerNorm statistics: transaction barrier fed by the peers' st.async
  __shared__ __align__(8) uint64_t pbar;      // split-K: the partner's partial accumulator has landed in pbuf
  __shared__ unsigned int icnt;               // kPreinit: epilogue warps that are through with a coming phase's accumulator (one count per warp and
                                              // phase, monotonic: a parity barrier would alias when the MMA thread skips phases without a unit)
  __shared__ uint32_t tmem_slot;
  __shared__ int abort_flag;
  __shared__ ChainPhase sphase[LDM_CHAIN_MAX_PHASES];   // shared-memory copy: indexed constant-bank reads are slow
  __shared__ UnitPlan splan[LDM_CHAIN_MAX_PHASES];
  __shared__ int scls[64], strow[64];                   // class / clamped per-row timestep of the cluster's rows (-1: no term)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x;                 // gridDim.x == CS: rank in cluster
  const int row0 = P.row_begin + blockIdx.y * NB;   // first batch row of this cluster
  const Waiter W{&abort_flag, P.err};
  const int NP = P.n_phases;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&wfull_bar[s], 1);         // the weight producer's arrive.expect_tx
      tc::mbar_init(&wempty_bar[s], 1);        // tcgen05.commit
    }
    for (int s = 0; s < kXSlots; ++s) {
      tc::mbar_init(&xfull_bar[s], 1);
      tc::mbar_init(&xempty_bar[s], 1);
    }
    tc::mbar_init(&tmem_full_bar, 1);
    tc::mbar_init(&obar[0], CS + 1);
    tc::mbar_init(&obar[1], CS + 1);
    tc::mbar_init(&sbar[0], 1);
    tc::mbar_init(&sbar[1], 1);
    tc::mbar_init(&pbar, 1);
    icnt = 0;
    abort_flag = 0;
    tc::fence_barrier_init();
    for (int p = 0; p < NP; ++p) tc::prefetch_tmap(&P.wmap[p]);
    for (int p = 0; p < kMaxXMaps; ++p) tc::prefetch_tmap(&P.xmaps[p]);
    for (int p = 0; p < kMaxXMaps; ++p) tc::prefetch_tmap(&P.xmaps4[p]);
  }
  if (warp == 2) tc::tmem_alloc<kTmemCols>(&tmem_slot);
  if (warp == 3) {
    const int nw = (int)(sizeof(ChainPhase) / 4) * P.n_phases;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.ph[0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sphase[0]);
    for (int i = lane; i < nw; i += 32) dst[i] = src[i];
    for (int j = lane; j < NB; j += 32) {
      const int r = row0 + j;
      const bool ok = r < P.row_end;
      scls[j] = ok && P.cls ? P.cls[r] : -1;
      strow[j] = ok && !P.sample && P.t_len != 1 ? clamp_t(P.t_idx[r], P.n_t) : -1;
    }
    __syncwarp();
    if (lane < NP) {
      const ChainPhase& ph = sphase[lane];
      UnitPlan pl;
      Unit un, ut;
      pl.valid = unit_of(ph, rank, false, un) ? 1 : 0;
      pl.valid_tail = unit_of(ph, rank, true, ut) ? 1 : 0;
      pl.tile_row = pl.valid ? un.tile * 128 : 0;
      pl.nkb = pl.valid ? un.nkb : 0;
      pl.mode = pl.valid ? un.mode : 0;
      pl.wcol0 = pl.valid ? (un.mode == 0 ? un.kb0 * BK : (un.mode == 2 ? un.kp * ph.K : 0)) : 0;
      pl.wcol1 = ph.K;
      pl.xcol0 = pl.valid && un.mode == 0 ? ph.xcol + un.kb0 * BK : 0;
      pl.pre = pl.valid && un.kp == 0 && !un.is_eps ? 1 : 0;
      pl.pad_[0] = pl.pad_[1] = pl.pad_[2] = 0;
      splan[lane] = pl;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  cluster_sync_all();   // every CTA's barriers exist before any remote arrive / store
  if (warp == 0) {
    // ------------------------------------------------------------------ weight-tile producer (runs ahead of the phases)
    if (tc::elect_one()) {
      const uint32_t ring_a = tc::smem_u32(wring), full_a = tc::smem_u32(&wfull_bar[0]), empty_a = tc::smem_u32(&wempty_bar[0]);
      uint32_t ws = 0, wpar = 0;
      bool ok = true;
      Tracer TR{P.trace ? P.trace + ((size_t)rank * LDM_CHAIN_TRACE_TRACKS + 2) * LDM_CHAIN_TRACE_LEN : nullptr, 0, false};
      for (int it = 0; it <= P.n_iter && ok; ++it) {
        const bool tail = it == P.n_iter;
        TR.on = P.trace != nullptr && blockIdx.y == 0 && it == P.trace_step;
        for (int p = 0; p < (tail ? 1 : NP) && ok; ++p) {
          const UnitPlan pl = splan[p];
          if (!(tail ? pl.valid_tail : pl.valid)) continue;
          const CUtensorMap* wm = &P.wmap[p];
          const int nchunks = pl.mode == 1 ? pl.nkb : (pl.nkb + 1) >> 1;
          const int step = pl.mode == 1 ? BK : 2 * BK;
          int col_a = pl.wcol0, col_b = pl.mode == 1 ? pl.wcol1 : pl.wcol0 + BK;
          for (int c = 0; c < nchunks; ++c, col_a += step, col_b += step) {
            const bool two = pl.mode == 1 || 2 * c + 1 < pl.nkb;
            if (!W.wait(empty_a + 8u * ws, wpar ^ 1u, 1)) { ok = false; break; }
            const uint32_t fb = full_a + 8u * ws, dst = ring_a + ws * kWSlotBytes;
            expect_tx_a(fb, two ? kWSlotBytes : kWTile);
            tma2d_a(dst, wm, fb, col_a, pl.tile_row);
            if (two) tma2d_a(dst + kWTile, wm, fb, col_b, pl.tile_row);
            if (c == 0) TR.stamp(20, p);
            if (c == nchunks - 1) TR.stamp(21, p);
            if (++ws == (uint32_t)S) { ws = 0; wpar ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ operand producer: waits for the phase hand-over,
    //                                                                    then requests EVERY k-block of the unit at once
    if (tc::elect_one()) {
      const uint32_t ring_a = tc::smem_u32(xring), full_a = tc::smem_u32(&xfull_bar[0]), empty_a = tc::smem_u32(&xempty_bar[0]);
      uint32_t xs = 0, xpar = 0, gp = 0;
      bool ok = true;
      Tracer TR{P.trace ? P.trace + ((size_t)rank * LDM_CHAIN_TRACE_TRACKS + 3) * LDM_CHAIN_TRACE_LEN : nullptr, 0, false};
      tc::mbar_arrive(&obar[0]);   // this thread's share of hand-overs 0 and 1
      tc::mbar_arrive(&obar[1]);
      for (int it = 0; it <= P.n_iter && ok; ++it) {
        const bool tail = it == P.n_iter;
        TR.on = P.trace != nullptr && blockIdx.y == 0 && it == P.trace_step;
        for (int p = 0; p < (tail ? 1 : NP) && ok; ++p, ++gp) {
          const UnitPlan pl = splan[p];
          const CUtensorMap* xm = &P.xmaps[sphase[p].xmap + (sphase[p].xmap_alt ? (it & 1) : 0)];
          // the first request of the phase is armed BEFORE the hand-over is awaited (slot free, bytes expected): after it only
          // the proxy fence and the TMA instruction stand between the hand-over and the first operand bytes
          const bool valid = tail ? pl.valid_tail : pl.valid;
          const int grp0 = valid ? x_group(P.xgroup, xs, pl.nkb) : 0;
          if (grp0 > 0) {
            if (!W.wait(empty_a + 8u * (xs + grp0 - 1), xpar ^ 1u, 3)) { ok = false; break; }
            expect_tx_a(full_a + 8u * xs, (uint32_t)grp0 * G::kXBytes);
          }
          if (gp > 0) {
            const uint32_t f = gp - 1;   // hand-over that publishes this phase's operand
            if (!W.wait_cluster(&obar[f & 1u], (f >> 1) & 1u, 2)) { ok = false; break; }
            TR.stamp(30, p);
            tc::mbar_arrive(&obar[f & 1u]);   // share of hand-over f + 2 (same barrier, next phase)
            if (!P.writer_fence) fence_proxy_async_global();   // peers' generic-proxy global writes (released above) -> async-proxy reads below
          }
          if (!(tail ? pl.valid_tail : pl.valid)) continue;
          const CUtensorMap* xm4 = &P.xmaps4[sphase[p].xmap + (sphase[p].xmap_alt ? (it & 1) : 0)];
          int xcol = pl.xcol0;
          for (int kb = 0; kb < pl.nkb;) {
            const uint32_t fb = full_a + 8u * xs;
            const int grp = x_group(P.xgroup, xs, pl.nkb - kb);
            const bool armed = kb == 0;      // the unit's first request: see above
            if (grp > 1) {
              // `grp` k-blocks as ONE box: the TMA unit spends ~150 clocks per instruction plus ~1.3 per 128-byte row, so a
              // 6 KB box per k-block held the first MMAs of a phase back.  The slots of a ring round are freed in order: the
              // last slot of the group is the one to wait for.  The group completes on its first slot's barrier; the other
              // slots' barriers are not used in this round (the MMA issuer keeps one parity bit per barrier).
              if (!armed) {
                if (!W.wait(empty_a + 8u * (xs + grp - 1), xpar ^ 1u, 3)) { ok = false; break; }
                expect_tx_a(fb, (uint32_t)grp * G::kXBytes);
              }
              tma3d_a(ring_a + xs * G::kXBytes, xm4, fb, 0, row0, xcol / BK);
              if (kb == 0) TR.stamp(32, p);
              kb += grp; xcol += grp * BK; xs += grp;
              if (xs == (uint32_t)kXSlots) { xs = 0; xpar ^= 1u; }
              continue;
            }
            if (!armed) {
              if (!W.wait(empty_a + 8u * xs, xpar ^ 1u, 3)) { ok = false; break; }
              expect_tx_a(fb, G::kXBytes);
            }
            tma2d_a(ring_a + xs * G::kXBytes, xm, fb, xcol, row0);
            if (kb == 0) TR.stamp(32, p);
            ++kb; xcol += BK;
            if (++xs == (uint32_t)kXSlots) { xs = 0; xpar ^= 1u; }
          }
          TR.stamp(33, p);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, NB);
      const uint32_t wfull_a = tc::smem_u32(&wfull_bar[0]), wempty_a = tc::smem_u32(&wempty_bar[0]);
      const uint32_t xfull_a = tc::smem_u32(&xfull_bar[0]), xempty_a = tc::smem_u32(&xempty_bar[0]);
      const uint32_t tfull_a = tc::smem_u32(&tmem_full_bar);
      const uint64_t wdesc0 = tc::make_desc_sw128(tc::smem_u32(wring)), xdesc0 = tc::make_desc_sw128(tc::smem_u32(xring));
      constexpr uint64_t kWSlotD = kWSlotBytes >> 4, kWTileD = kWTile >> 4, kXD = G::kXBytes >> 4;   // descriptor start-address units (16 B)
      const uint32_t acc0 = tmem_base, acc1 = tmem_base + (uint32_t)NB;
      uint32_t ws = 0, wpar = 0, xs = 0, xpar = 0;
      uint32_t xfpar = 0;       // bit s: parity of the next phase of operand barrier s (a barrier is only used by the first slot of a group)
      bool ok = true;
      Tracer TR{P.trace ? P.trace + ((size_t)rank * LDM_CHAIN_TRACE_TRACKS + 4) * LDM_CHAIN_TRACE_LEN : nullptr, 0, false};
      uint32_t gph = 0;         // phases so far (all CTAs count alike): icnt grows by 4 NW per phase
      for (int it = 0; it <= P.n_iter && ok; ++it) {
        const bool tail = it == P.n_iter;
        TR.on = P.trace != nullptr && blockIdx.y == 0 && it == P.trace_step;
        for (int p = 0; p < (tail ? 1 : NP) && ok; ++p, ++gph) {
          const UnitPlan pl = splan[p];
          if (!(tail ? pl.valid_tail : pl.valid)) continue;
          // accumulator 0 already holds the tile's additive terms: written by the 4 NW epilogue warps after the previous phase
          const uint32_t pre = G::kPreinit && gph > 0 && !tail && pl.pre ? 1u : 0u;
          if (pre) {
            const uint32_t want = gph * (uint32_t)(4 * NW), ia = tc::smem_u32(&icnt);
            uint32_t seen;
            const long long t0 = clock64();
            for (;;) {
              asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(ia) : "memory");
              if (seen >= want) break;
              if (W.aborted()) { ok = false; break; }
              if (clock64() - t0 > (4ll << 30)) { W.fail(12); ok = false; break; }
            }
            if (!ok) break;
          }
          if (pl.mode == 1) {
            // dual: both weight tiles of a chunk against the same operand k-block; consecutive MMAs never depend on each other
            int grp = 0;      // k-blocks of the current operand group that are still to come (their barrier is the group's first)
            for (int c = 0; c < pl.nkb; ++c) {
              if (grp == 0) {
                if (!W.wait(xfull_a + 8u * xs, (xfpar >> xs) & 1u, 4)) { ok = false; break; }
                xfpar ^= 1u << xs;
                grp = x_group(P.xgroup, xs, pl.nkb - c);
              }
              --grp;
              if (!W.wait(wfull_a + 8u * ws, wpar, 4)) { ok = false; break; }
              if (c == 0) TR.stamp(40, p);
              if (c == pl.nkb / 2) TR.stamp(42, p);
              tc::fence_after_sync();
              const uint64_t dwa = wdesc0 + (uint64_t)ws * kWSlotD, dwb = dwa + kWTileD, dx = xdesc0 + (uint64_t)xs * kXD;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint32_t acc = (uint32_t)(c != 0 || k != 0);
                tc::umma_bf16(acc0, dwa + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, acc | pre);
                tc::umma_bf16(acc1, dwb + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, acc);
              }
              commit_a(wempty_a + 8u * ws);
              commit_a(xempty_a + 8u * xs);
              if (++ws == (uint32_t)S) { ws = 0; wpar ^= 1u; }
              if (++xs == (uint32_t)kXSlots) { xs = 0; xpar ^= 1u; }
            }
          } else {
            // one accumulator: a chunk is two consecutive k-blocks (the last chunk of an odd count holds one)
            const int nchunks = (pl.nkb + 1) >> 1;
            int grp = 0;
            for (int c = 0; c < nchunks; ++c) {
              const bool two = 2 * c + 1 < pl.nkb;
              const uint32_t xs2 = xs + 1 == (uint32_t)kXSlots ? 0u : xs + 1, xpar2 = xs + 1 == (uint32_t)kXSlots ? xpar ^ 1u : xpar;
              if (grp == 0) {
                if (!W.wait(xfull_a + 8u * xs, (xfpar >> xs) & 1u, 4)) { ok = false; break; }
                xfpar ^= 1u << xs;
                grp = x_group(P.xgroup, xs, pl.nkb - 2 * c);
              }
              --grp;
              if (two) {
                if (grp == 0) {
                  if (!W.wait(xfull_a + 8u * xs2, (xfpar >> xs2) & 1u, 4)) { ok = false; break; }
                  xfpar ^= 1u << xs2;
                  grp = x_group(P.xgroup, xs2, pl.nkb - (2 * c + 1));
                }
                --grp;
              }
              if (!W.wait(wfull_a + 8u * ws, wpar, 4)) { ok = false; break; }
              if (c == 0) TR.stamp(40, p);
              if (c == nchunks / 2) TR.stamp(42, p);
              tc::fence_after_sync();
              const uint64_t dwa = wdesc0 + (uint64_t)ws * kWSlotD, dxa = xdesc0 + (uint64_t)xs * kXD, dxb = xdesc0 + (uint64_t)xs2 * kXD;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                tc::umma_bf16(acc0, dwa + (uint64_t)(2 * k), dxa + (uint64_t)(2 * k), idesc, (uint32_t)(c != 0 || k != 0) | pre);
              commit_a(xempty_a + 8u * xs);
              if (two) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  tc::umma_bf16(acc0, dwa + kWTileD + (uint64_t)(2 * k), dxb + (uint64_t)(2 * k), idesc, 1u);
                commit_a(xempty_a + 8u * xs2);
              }
              commit_a(wempty_a + 8u * ws);
              if (++ws == (uint32_t)S) { ws = 0; wpar ^= 1u; }
              xs = two ? xs2 : xs;
              xpar = two ? xpar2 : xpar;
              if (++xs == (uint32_t)kXSlots) { xs = 0; xpar ^= 1u; }
            }
          }
          if (ok) commit_a(tfull_a);
          TR.stamp(41, p);
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue + statistics exchange + hand-over
    const int et = threadIdx.x - kCtlThreads;   // 0 .. 128 NW - 1
    const int q = warp & 3;                     // TMEM lane quadrant of this warp
    const int g = (warp - 4) >> 2;              // batch-row group: rows [16 g, 16 g + 16) of the cluster
    const int lrow = q * 32 + lane;             // row of the 128-row weight tile this thread finishes
    const int s0 = 16 * g;
    const int srow = s0 + (lane >> 1);          // batch row (in cluster) this lane owns in the statistics exchange
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sbar_local[2] = {tc::smem_u32(&sbar[0]), tc::smem_u32(&sbar[1])};
    const uint32_t obar_local[2] = {tc::smem_u32(&obar[0]), tc::smem_u32(&obar[1])};
    uint32_t tpar = 0;            // parity of tmem_full_bar
    uint32_t ppar = 0;            // parity of pbar
    uint32_t sidx = 0;            // statistics exchanges so far (same count in every CTA)
    uint32_t sph[2] = {0, 0};     // completed phases of sbar[b] in THIS CTA
    uint32_t oidx = 0;            // hand-overs so far
    Tracer TR{P.trace ? P.trace + ((size_t)rank * LDM_CHAIN_TRACE_TRACKS + (et == 0 ? 0 : 1)) * LDM_CHAIN_TRACE_LEN : nullptr, 0, false};
    auto stamp = [&](int tag, int p) { TR.stamp(tag, p); };
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(G::kEpiThreads) : "memory"); };
    // Every epilogue warp needs the same barrier: ONE warp polls the mbarrier, the others block in a named barrier (bar.sync
    // costs no issue slots, a polling warp does; measured: try_wait returns within ~45 clocks whatever the suspend hint, and
    // twelve polling warps took 12 % of the SM's issue slots away from the MMA / TMA threads and from each other)
    auto epi_wait = [&](uint64_t* bar, uint32_t parity, int code) {
      if (warp == 4) W.wait(bar, parity, code);
      asm volatile("bar.sync 8, %0;" ::"n"(G::kEpiThreads) : "memory");
    };
    auto group_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); };   // the 4 warps of row group g

    // Statistics exchange, sender side.  Lane l holds (mean, M2) of group row (l >> 1) over `cnt` features of this
    // warp.  The 4 warps of the row group merge their partials (Chan et al.) in shared memory, then every thread
    // sends the merged (mean, M2) of one row over 4 cnt features to 1/8 of the CTAs that own a tile of the phase:
    // an 8-byte st.async that also signals the receiver's transaction barrier.
    const int tg = q * 32 + lane;                 // thread index inside the row group
    const int xrow = tg >> 3, xds = tg & 7;       // row / destination slice this thread publishes
    auto publish = [&](uint32_t buf, int tile, float2 st, float cnt, int first, int ntiles, int ks) {
      float2* gs = gstat + ((size_t)buf * NW + g) * 64;
      if ((lane & 1) == 0) gs[q * 16 + (lane >> 1)] = st;
      stamp(13, 0);
      group_bar();
      stamp(14, 0);
      const float2 p0 = gs[xrow], p1 = gs[16 + xrow], p2 = gs[32 + xrow], p3 = gs[48 + xrow];
      const float mean = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
      const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
      const float m2 = ((p0.y + p1.y) + (p2.y + p3.y)) + cnt * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
      const uint32_t slot = tc::smem_u32(slots + ((size_t)buf * kSlots + tile) * NB + s0 + xrow);
      for (int k = xds; k < ntiles; k += 8)
        st_async_f2(mapa_u32(slot, (uint32_t)(first + k * ks)), mean, m2, mapa_u32(sbar_local[buf], (uint32_t)(first + k * ks)));
    };
    // wait until the partials of all `nparts` tiles for all NB rows have landed in buffer `buf`
    auto exchange_wait = [&](uint32_t buf, int nparts, int code) {
      if (et == 0) tc::mbar_arrive_expect_tx(&sbar[buf], (uint32_t)nparts * NB * (uint32_t)sizeof(float2));
      epi_wait(&sbar[buf], sph[buf] & 1u, code);
      sph[buf]++;
    };
    // receiver side: merge the `nparts` per-tile partials (`cnt` features each) into (mean, rstd) of the warp's 16
    // rows, left in shared memory (rs[row], broadcast reads).  Lane l merges half of the partials of row (l >> 1).
    float2* rs = rowstat + (size_t)(warp - 4) * 16;
    auto combine = [&](uint32_t buf, int nparts, float cnt) {
      const float2* p = slots + ((size_t)buf * kSlots + (lane & 1)) * NB + srow;
      const int half = (nparts + 1 - (lane & 1)) >> 1;   // partials lane & 1, lane & 1 + 2, ...
      const float inv_n = 1.0f / (float)nparts;
      float2 s[kSlots / 2];
#pragma unroll
      for (int k = 0; k < kSlots / 2; ++k) s[k] = k < half ? p[(size_t)(2 * k) * NB] : make_float2(0.f, 0.f);
      float msum = 0.f;
#pragma unroll
      for (int k = 0; k < kSlots / 2; ++k) msum += s[k].x;
      msum += __shfl_xor_sync(0xffffffffu, msum, 1);
      const float mean = msum * inv_n;
      float m2 = 0.f;
#pragma unroll
      for (int k = 0; k < kSlots / 2; ++k) {
        const float dm = s[k].x - mean;
        m2 += k < half ? s[k].y + cnt * dm * dm : 0.f;
      }
      m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
      const float var = m2 * inv_n / cnt;
      __syncwarp();
      if ((lane & 1) == 0) rs[lane >> 1] = make_float2(mean, rsqrtf(var + 1e-5f));
      __syncwarp();
    };

    const bool per_row_t = !P.sample && P.t_len != 1;
    if constexpr (G::kCaddInTmem) {
      // ---- per-sample additive terms live in TMEM for the whole chain (Geo::kCaddInTmem)
      for (int p = 0; p < NP; ++p) {
        const ChainPhase& ph = sphase[p];
        Unit un;
        if (!unit_of(ph, rank, false, un) || un.kp != 0 || un.is_eps || ph.cadd_col < 0) continue;
        const int grow = un.tile * 128 + lrow;
        float c[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ci = scls[s0 + j], ti = strow[s0 + j];
          float a = 0.f;
          if (ph.tab_c && ci >= 0) a += __ldg(ph.tab_c + (size_t)ci * ph.rows + grow);
          if (ph.tab_t && per_row_t && ti >= 0) a += __ldg(ph.tab_t + (size_t)ti * ph.rows + grow);
          c[j] = a;
        }
        tmem_st16(lane_taddr + (uint32_t)(ph.cadd_col + s0), c);
      }
      tc::fence_before_sync();
    }

    // ---- the chain state of the rows / features this thread finishes in the eps tiles stays in registers
    // (kept in spare TMEM columns between the eps tiles' visits, so that it costs no registers in the other phases)
    const ChainPhase& ph0 = sphase[0];
    Unit eun;
    const bool eps_owner = unit_of(ph0, rank, true, eun);
    const int ef = eps_owner ? (eun.tile - ph0.nst_tiles) * 128 + lrow : 0;   // latent feature of this thread
    const float bfin = eps_owner ? __ldg(ph0.bias + eun.tile * 128 + lrow) : 0.f;   // (1 + s) b_f (v2:560-561)
    if (eps_owner) {
      float x0[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = row0 + s0 + j;
        x0[j] = r < P.row_end ? P.x[(size_t)r * P.latent + ef] : 0.f;
      }
      tmem_st16(lane_taddr + (uint32_t)(P.x_col + s0), x0);
    }

    for (int it = 0; it <= P.n_iter; ++it) {
      const bool tail = it == P.n_iter;
      const int par = it & 1;
      TR.on = P.trace != nullptr && blockIdx.y == 0 && it == P.trace_step && (et == 0 || et == 64);
      stamp(1, 0);
      // timestep of this iteration's forward (t_uni) and posterior coefficients of this and the previous iteration
      const int t_uni = P.sample ? P.t_start - it : (P.t_len == 1 ? clamp_t(P.t_idx[0], P.n_t) : -1);
      // unconditional loads (clamped index), patched AFTER the first wait of the iteration: no stall on their L2 latency here
      const int tp = P.sample ? P.t_start - it + (it > 0 ? 1 : 0) : 0, tc_ = P.sample ? P.t_start - it + (tail ? 1 : 0) : 0;
      float4 cf_prev = P.coef_or_one[P.sample ? tp : 0], cf_cur = P.coef_or_one[P.sample ? tc_ : 0];   // (c2, sqrt_alpha, sigma)
      // (the coefficients are only CONSUMED after the first accumulator wait of the iteration: their L2 latency is hidden)
      for (int p = 0; p < (tail ? 1 : NP); ++p) {
        const ChainPhase& ph = sphase[p];
        Unit un;
        const bool has_unit = unit_of(ph, rank, tail, un);
        const bool active = has_unit && un.kp == 0;       // owns the tile: finishes it
        const bool partner = has_unit && un.kp != 0;      // split-K helper: ships its partial accumulator to the owner
        const bool eps_tile = active && un.is_eps;
        const int tile = has_unit ? un.tile : 0;
        const int grow = tile * 128 + lrow;
        if (partner) {
          epi_wait(&tmem_full_bar, tpar, 5);
          tpar ^= 1u;
          tc::fence_after_sync();
          float pv[16];
          tc::tmem_ld16(lane_taddr + (uint32_t)s0, pv);
          tc::fence_before_sync();
          const uint32_t owner = (uint32_t)(ph.first + tile * ph.ks);
          const uint32_t dst = mapa_u32(tc::smem_u32(pbuf + (size_t)(g * 4) * 128 + lrow), owner);
          const uint32_t dbar = mapa_u32(tc::smem_u32(&pbar), owner);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_async_f4(dst + (uint32_t)(c * 128 * sizeof(float4)), pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3], dbar);
        }
        float v[16];
        float h2k[8];                 // h2 of this lane's feature x 8 rows: its statistics are published AFTER the hand-over
        bool pub_b = false;
        const uint32_t sidx0 = sidx;  // exchange index before this phase: sidx0 - 1 is the previous phase's background exchange
        if (active) {
          // everything that does not depend on the accumulator is fetched before the waits
          float t_b = 0.f, t_t = 0.f, t_g = 0.f, t_q = 0.f;   // loaded now, summed after the wait (no dependent use before it)
          // kPreinit: bias, time row and per-sample terms are already IN the accumulator (written after the previous phase)
          const bool pre = G::kPreinit && !eps_tile && !(it == 0 && p == 0);
          if (!eps_tile) {
            if (ph.bias && !pre) t_b = __ldg(ph.bias + grow);
            if (ph.tab_t && t_uni >= 0 && !pre) t_t = __ldg(ph.tab_t + (size_t)t_uni * ph.rows + grow);
            if (ph.type == LDM_PH_MERGED) t_g = __ldg(ph.g0b + grow);
            if (ph.dual) t_q = __ldg(ph.q + grow);
          }
          const bool has_c = ph.cadd_col >= 0 && !eps_tile && !pre;
          float cadd[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) cadd[j] = 0.f;
          if (has_c && !G::kCaddInTmem) {
            // 16 independent L2 loads per table, issued back to back (indices from shared memory; a missing term reads
            // row 0 and is multiplied by zero): their latency hides under the accumulator wait
            if (ph.tab_c) {
              const float* tc_ = ph.tab_c + grow;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int ci = scls[s0 + j];
                cadd[j] = 0.f + (ci >= 0 ? 1.f : 0.f) * __ldg(tc_ + (size_t)(ci >= 0 ? ci : 0) * ph.rows);
              }
            }
            if (ph.tab_t && per_row_t) {
              const float* tt_ = ph.tab_t + grow;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int ti = strow[s0 + j];
                cadd[j] += (ti >= 0 ? 1.f : 0.f) * __ldg(tt_ + (size_t)(ti >= 0 ? ti : 0) * ph.rows);
              }
            }
          }
          if (ph.dual) {
            // (mu, r) of the operand rows: the previous phase's background statistics exchange lands while the tensor core
            // is still busy with this phase; merged here, used after the accumulator wait
            const uint32_t bb = (sidx0 - 1u) & 1u;
            exchange_wait(bb, ph.prev_tiles, 11);
            combine(bb, ph.prev_tiles, 64.0f);
          }
          stamp(2, p);
          epi_wait(&tmem_full_bar, tpar, 5);
          tpar ^= 1u;
          stamp(3, p);
          tc::fence_after_sync();
          const float cb_prev = it > 0 ? cf_prev.x / cf_prev.y : 0.f;   // c2 / sqrt(alpha) of the previous step (forward(): 1)
          const float t_e = cb_prev * t_g;                              // the eps bias of the previous step, seen through G_0
          float a2[16];
          tc::tmem_ld16(lane_taddr + (uint32_t)s0, v);
          if (pre) {
            // the accumulator started from b + (T + C): the same association in every mode (see below), only the eps term is left
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] - t_e;
          } else if (has_c) {
            // one association for every mode (uniform t: t_t = T[t], c = C[c_r]; per-row t: t_t = 0, c = C[c_r] + T[t_r]), so
            // that a row's result does not depend on how its timestep was passed: acc + (b + (T + C)) - e
            if constexpr (G::kCaddInTmem) tc::tmem_ld16(lane_taddr + (uint32_t)(ph.cadd_col + s0), cadd);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] + (t_b + (t_t + cadd[j]))) - t_e;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (v[j] + (t_b + t_t)) - t_e;
          }
          if (ph.dual && ph.ks == 1) tc::tmem_ld16(lane_taddr + (uint32_t)(NB + s0), a2);
          tc::fence_before_sync();            // ordered before the next phase's MMAs through the hand-over
          if (ph.ks > 1) {   // the partner unit's accumulator: acc2 of a dual phase (or the other K half of a plain one)
            if (et == 0) tc::mbar_arrive_expect_tx(&pbar, G::kPbufBytes);
            epi_wait(&pbar, ppar, 10);
            ppar ^= 1u;
            const float4* pp = pbuf + (size_t)(g * 4) * 128 + lrow;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 a = pp[c * 128];
              if (ph.dual) { a2[4 * c] = a.x; a2[4 * c + 1] = a.y; a2[4 * c + 2] = a.z; a2[4 * c + 3] = a.w; }
              else { v[4 * c] += a.x; v[4 * c + 1] += a.y; v[4 * c + 2] += a.z; v[4 * c + 3] += a.w; }
            }
          }
          stamp(4, p);
          if (ph.dual) {
            // LayerNorm of the operand rows applied after the contraction: W2 . LN_b(h2) = r (W2' h2 - mu q) + const
            // (gamma, beta folded into W2' / the bias at pack time); (mu, r) of the 16 rows were merged before the wait
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 mr = rs[j];
              v[j] += mr.y * (a2[j] - mr.x * t_q);
            }
            __syncwarp();
          }
          stamp(5, p);
        }

        if (ph.type == LDM_PH_STAGE || ph.type == LDM_PH_MERGED) {
          // Tile rows: quadrant q holds 16 h features in lanes 0..15 and u = Linear_b(h) of the SAME 16 features in lanes
          // 16..31.  The two halves trade 8 rows, after which every lane owns feature f for 8 rows, h AND u: the
          // LayerNorm / Swish / residual chain of v2:546-548 runs in registers on all 32 lanes.
          const int f = tile * 64 + q * 16 + (lane & 15);
          const int hr = (lane >> 4) * 8;             // first of this lane's 8 rows inside the row group
          const int ntile = ph.type == LDM_PH_MERGED ? ph.nst_tiles : ph.tiles;
          const uint32_t b0 = sidx & 1u;
          sidx += 2;                                  // [A: statistics of u, waited here] [B: statistics of h2, background]
          if (active && !eps_tile) {
            const float ga = __ldg(ph.ga + f), ba = __ldg(ph.ba + f);
            float u[8];
            {
              const bool lo = lane < 16;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float recv = __shfl_xor_sync(0xffffffffu, lo ? v[8 + i] : v[i], 16);
                h2k[i] = lo ? v[i] : recv;
                u[i] = lo ? recv : v[8 + i];
              }
            }
            stamp(12, p);
            const float2 ust = half_row_stats8(u, lane);
            publish(b0, tile, ust, 16.0f, ph.first, ntile, ph.ks);
            stamp(6, p);
            exchange_wait(b0, ntile, 6);
            stamp(7, p);
            combine(b0, ntile, 64.0f);
            stamp(15, p);
            bf16* o = ph.out + (size_t)(row0 + s0 + hr) * ph.ld_out + f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // h2 = swish(LN_a(u)) + h: the next phase's operand (its LayerNorm is applied there)
              const float2 mr = rs[hr + i];
              const float a = mr.y * ga;                      // LayerNorm_a as one FMA: u a + (beta - mu a)
              const float pre = fmaf(u[i], a, fmaf(-mr.x, a, ba));
              h2k[i] += P.split ? swish_precise(pre) : swish_fast(pre);
              store_operand(o + (size_t)i * ph.ld_out, h2k[i], P.split, ph.wout);   // rows beyond the batch land in the buffers' padding
            }
            pub_b = true;
          }
        } else {   // LDM_PH_FINAL_LN
          const int f = grow;
          const int nparts = ph.tiles;
          const uint32_t b0 = sidx & 1u;
          sidx += 1;
          if (active) {   // -c_b LN_f(h + T_f[t] + C_f[c]): block 1 of the next merged operand     (v2:554-559)
            const float ga = __ldg(ph.ga + f), be = __ldg(ph.ba + f);
            publish(b0, tile, warp_row_stats16(v, lane), 32.0f, ph.first, ph.tiles, ph.ks);
            exchange_wait(b0, nparts, 8);
            stamp(11, p);
            combine(b0, nparts, 128.0f);
            const float cb_cur = cf_cur.x / cf_cur.y;
            bf16* o = P.af[par ^ 1] + (size_t)(row0 + s0) * P.ld_af + P.latent + f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 mr = rs[j];
              store_operand(o + (size_t)j * P.ld_af, -cb_cur * ((v[j] - mr.x) * mr.y * ga + be), P.split, 3 * P.latent);
            }
          }
        }

        // ---- hand-over: this CTA's share of the phase is in global memory; tell every CTA of the cluster, then wait
        //      for all of them (the operand producer waits on the same barrier and starts the next phase's loads)
        stamp(8, p);
        if (P.writer_fence) fence_proxy_async_all();
        epi_bar();
        const uint32_t ob = oidx & 1u, opar = (oidx >> 1) & 1u;
        if (et < CS) remote_arrive(mapa_u32(obar_local[ob], (uint32_t)et));
        stamp(9, p);

        if constexpr (G::kPreinit) {
          // ---- the coming phase's accumulator 0 starts from its additive terms b + (T[t] + C[c_r] (+ T[t_r])): read from the
          //      TMEM-resident per-sample terms and written while the hand-over is in flight, ~2 k clocks before that phase's first
          //      MMA (which waits for icnt and accumulates).  Every epilogue warp arrives once per phase, unit or not.
          const bool last_p = p + 1 == (tail ? 1 : NP);
          const int itn = last_p ? it + 1 : it, pn = last_p ? 0 : p + 1;
          if (itn <= P.n_iter) {
            const ChainPhase& nx = sphase[pn];
            Unit nu;
            if (itn != P.n_iter && unit_of(nx, rank, false, nu) && nu.kp == 0 && !nu.is_eps) {
              const int grow_n = nu.tile * 128 + lrow;
              const int t_n = P.sample ? P.t_start - itn : (P.t_len == 1 ? clamp_t(P.t_idx[0], P.n_t) : -1);
              const float xb = nx.bias ? __ldg(nx.bias + grow_n) : 0.f;
              const float xt = (nx.tab_t && t_n >= 0) ? __ldg(nx.tab_t + (size_t)t_n * nx.rows + grow_n) : 0.f;
              float xi[16];
              if (nx.cadd_col >= 0) {
                if constexpr (G::kCaddInTmem) {
                  tc::tmem_ld16(lane_taddr + (uint32_t)(nx.cadd_col + s0), xi);
                } else {      // the same sums as the TMEM-resident terms: (0 + C[c_r]) + T[t_r]
#pragma unroll
                  for (int j = 0; j < 16; ++j) xi[j] = 0.f;
                  if (nx.tab_c) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const int ci = scls[s0 + j];
                      xi[j] = 0.f + (ci >= 0 ? 1.f : 0.f) * __ldg(nx.tab_c + grow_n + (size_t)(ci >= 0 ? ci : 0) * nx.rows);
                    }
                  }
                  if (nx.tab_t && per_row_t) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const int ti = strow[s0 + j];
                      xi[j] += (ti >= 0 ? 1.f : 0.f) * __ldg(nx.tab_t + grow_n + (size_t)(ti >= 0 ? ti : 0) * nx.rows);
                    }
                  }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) xi[j] = xb + (xt + xi[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) xi[j] = xb + xt;
              }
              tmem_st16(lane_taddr + (uint32_t)s0, xi);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(tc::smem_u32(&icnt)) : "memory");
          }
        }

        if (eps_tile) {
          // ---- posterior update (v2:584-592), OFF the critical path: nothing in the cluster needs x_{t-1} itself before
          //      the merged phase of the next step; its pieces reach global memory long before that hand-over.
          //      v = W_f' . (-c_b [LN_f(h) | x]) = -c_b (eps - b_fin) of the previous iteration's forward.
          const float cb_prev = it > 0 ? cf_prev.x / cf_prev.y : 0.f;
          const float cb_cur = cf_cur.x / cf_cur.y;
          float xr[16];
          tc::tmem_ld16(lane_taddr + (uint32_t)(P.x_col + s0), xr);
          if (it > 0) {
            if (!P.sample) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int row = row0 + s0 + j;
                if (row < P.row_end) P.eps_out[(size_t)row * P.latent + ef] = bfin - v[j];
              }
            } else {
              float zz[16];
              tc::tmem_ld16(lane_taddr + (uint32_t)(P.z_col + s0), zz);
              const float inv_cb = 1.0f / cb_prev;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                xr[j] = ddpm_update_one(xr[j], bfin - v[j] * inv_cb, cf_prev.x, cf_prev.y, cf_prev.z, zz[j]);
              if (tail) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int row = row0 + s0 + j;
                  if (row < P.row_end) P.x[(size_t)row * P.latent + ef] = xr[j];
                }
              } else {
                tmem_st16(lane_taddr + (uint32_t)(P.x_col + s0), xr);
              }
            }
          }
          if (!tail) {
            // pieces of the next merged operand: x~ = x / sqrt(alpha_t) + sigma_t z_t (block 0) and -c_b x (block 2)
            float z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0.f;
            if (P.sample && cf_cur.z > 0.0f) {
              const int t = P.t_start - it;
              if (P.noise) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int r = row0 + s0 + j;
                  if (r < P.row_end) z[j] = P.noise[((size_t)it * P.B + r) * P.latent + ef];
                }
              } else {
                // the 4 lanes that share a Philox quad split the rows between them, then trade components
                const unsigned long long seed = P.rng[0], off = P.rng[1] + (unsigned long long)(row0 + s0);
                const int sub = lane & 3, base = lane & ~3;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  const float4 z4 = philox_normal4(seed, off + (unsigned long long)(4 * gq + sub), (uint32_t)t, (uint32_t)(ef >> 2));
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float gx = __shfl_sync(0xffffffffu, z4.x, base + i), gy = __shfl_sync(0xffffffffu, z4.y, base + i);
                    const float gz = __shfl_sync(0xffffffffu, z4.z, base + i), gw = __shfl_sync(0xffffffffu, z4.w, base + i);
                    z[4 * gq + i] = sub == 0 ? gx : (sub == 1 ? gy : (sub == 2 ? gz : gw));
                  }
                }
              }
            }
            if (P.sample) tmem_st16(lane_taddr + (uint32_t)(P.z_col + s0), z);   // kept for the update one iteration later
            bf16* o = P.af[par ^ 1] + (size_t)(row0 + s0) * P.ld_af + ef;
            const float inv_sa = 1.0f / cf_cur.y;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              store_operand(o + (size_t)j * P.ld_af, xr[j] * inv_sa + cf_cur.z * z[j], P.split, 3 * P.latent);
              store_operand(o + (size_t)j * P.ld_af + 2 * P.latent, -cb_cur * xr[j], P.split, 3 * P.latent);
            }
          }
        }

        epi_wait(&obar[ob], opar, 9);
        oidx++;
        stamp(10, p);
        if (pub_b) {
          // ---- background exchange, issued once the hand-over is through (the operand producer's proxy fence must not
          //      queue behind these stores): (mean, M2) of h2 over this warp's 16 features -> the tile owners of the NEXT
          //      phase, who apply LayerNorm_b (v2:549) after their contraction
          const ChainPhase& nx = sphase[p + 1];
          publish((sidx0 + 1u) & 1u, tile, half_row_stats8(h2k, lane), 16.0f, nx.first, nx.tiles, nx.ks);
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while a peer may still address its shared memory
  if (warp == 2) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------- pack kernels
// C(M,N) = A(M,K) . op(B): fp64 accumulation, run once per pack.  16 x 16 output tile per block, 16-deep k tiles staged in
// shared memory (the first version read every operand element from global memory once per output element: 8.4 ms per pack).
//   transB = 0: B is (K,N) row-major with pitch ldb;  transB = 1: B is (N,K) row-major with pitch ldb
__global__ void __launch_bounds__(256) pack_mm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int transB,
                                                      float* __restrict__ C, int ldc, int M, int N, int K) {
  __shared__ float As[16][17], Bs[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  double s = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, am = blockIdx.y * 16 + ty;
    As[ty][tx] = (am < M && ka < K) ? A[(size_t)am * lda + ka] : 0.f;
    if (transB) {       // Bs[k][n] = B[n][k]: thread (ty, tx) loads B[n0 + ty][k0 + tx]
      const int bn = blockIdx.x * 16 + ty, kb = k0 + tx;
      Bs[tx][ty] = (bn < N && kb < K) ? Bm[(size_t)bn * ldb + kb] : 0.f;
    } else {
      const int kb = k0 + ty, bn = blockIdx.x * 16 + tx;
      Bs[ty][tx] = (kb < K && bn < N) ? Bm[(size_t)kb * ldb + bn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) s += (double)As[ty][k] * (double)Bs[k][tx];
    __syncthreads();
  }
  if (m < M && n < N) C[(size_t)m * ldc + n] = (float)s;
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
// natural row order [h_0..h_{d-1} | u_0..u_{d-1}] -> tile order (tile s, quadrant q: h[64s+16q ..+15] | u[64s+16q ..+15]); stage = 0: identity
__device__ __forceinline__ int tile_src_row(int rt, int d, int stage, int nsr) {
  if (!stage || rt >= nsr) return rt;   // rows [nsr, rows): the eps rows of the merged phase, natural order
  const int s = rt >> 7, l = rt & 127, q = l >> 5, w = l & 31;   // quadrant q: 16 h rows, then the 16 u rows of the same features
  return w < 16 ? s * 64 + q * 16 + w : d + s * 64 + q * 16 + (w - 16);
}
__global__ void pack_rows_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int K, int d, int stage, int nsr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * K) return;
  const int rt = (int)(i / K), k = (int)(i % K);
  dst[i] = __float2bfloat16_rn(src[(size_t)tile_src_row(rt, d, stage, nsr) * K + k]);
}
// strict mode: each accumulator block of Kp columns becomes [hi | hi | lo] (3 Kp columns), against operands stored as [hi | lo | hi]
__global__ void pack_rows_split_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int K, int Kp, int d, int stage, int nsr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * K) return;
  const int rt = (int)(i / K), k = (int)(i % K), a = k / Kp, kk = k - a * Kp;
  const float v = src[(size_t)tile_src_row(rt, d, stage, nsr) * K + k];
  const bf16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  bf16* o = dst + (size_t)rt * 3 * K + (size_t)a * 3 * Kp + kk;
  o[0] = hi; o[Kp] = hi; o[2 * Kp] = lo;
}
// first merged operand of a launch: [x | 0 | 0] (bf16), or its (hi, lo, hi) thirds in strict mode
__global__ void chain_stage_x_kernel(const float* __restrict__ x, bf16* __restrict__ dst, int B, int L, int split) {
  const int W3 = 3 * L, ld = split ? 3 * W3 : W3;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * W3) return;
  const int r = (int)(i / W3), c = (int)(i % W3);
  const float v = c < L ? x[(size_t)r * L + c] : 0.f;
  store_operand(dst + (size_t)r * ld + c, v, split, W3);
}
// tables: dst[t][rt] = src[t][src_row(rt)]
__global__ void pack_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int rows, int d, int stage, int nsr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * rows) return;
  const int t = (int)(i / rows), rt = (int)(i % rows);
  dst[i] = src[(size_t)t * rows + tile_src_row(rt, d, stage, nsr)];
}

// M[r][c] *= g[c] on a (rows x cols) block with pitch ld
__global__ void pack_scale_cols_kernel(float* __restrict__ M, int ld, const float* __restrict__ g, int rows, int cols) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  M[(size_t)r * ld + c] *= g[c];
}
__global__ void pack_fill_kernel(float* __restrict__ p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}


int mm(ldm_ctx* ctx, const float* A, int lda, const float* B, int ldb, int transB, float* C, int ldc, int M, int N, int K,
       cudaStream_t st) {
  pack_mm_kernel<<<dim3(ceil_div(N, 16), ceil_div(M, 16)), 256, 0, st>>>(A, lda, B, ldb, transB, C, ldc, M, N, K);
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

template <int NW>
int chain_occupancy(int* out) {
  using G = Geo<NW>;
  LDM_CUDA(cudaFuncSetAttribute(chain_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmemBytes));
  LDM_CUDA(cudaFuncSetAttribute(chain_kernel<NW>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, 1);
  cfg.blockDim = dim3(G::kThreads);
  cfg.dynamicSmemBytes = G::kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, chain_kernel<NW>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *out = n;
  return 0;
}

int g_chain_clusters[5] = {0, 0, 0, 0, 0};   // co-resident clusters per variant NW (index)

template <int NW>
int launch_variant(const ChainParams& P, int nclusters, cudaStream_t st) {
  using G = Geo<NW>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, nclusters);
  cfg.blockDim = dim3(G::kThreads);
  cfg.dynamicSmemBytes = G::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LDM_CUDA(cudaLaunchKernelEx(&cfg, chain_kernel<NW>, P));
  return 0;
}

}  // namespace

int tc_make_act_map(const void* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out);
int tc_make_kblock_map(const void* base, int rows, int cols, int ld, int box_rows, int kblocks, CUtensorMap* out);

// Can 16-CTA clusters of this kernel be scheduled on this device?  (called once per context)
int chain_init(ldm_ctx* ctx) {
  LDM_TRY(tc_init(ctx));
  if (!g_chain_attr_set) {
    LDM_TRY(chain_occupancy<2>(&g_chain_clusters[2]));
    LDM_TRY(chain_occupancy<3>(&g_chain_clusters[3]));
    LDM_TRY(chain_occupancy<4>(&g_chain_clusters[4]));
    g_chain_attr_set = true;
  }
  LDM_CHECK(g_chain_clusters[2] >= 1 || g_chain_clusters[3] >= 1,
            "the device cannot co-schedule a cluster of %d CTAs with %zu bytes of shared memory each", CS, Geo<3>::kSmemBytes);
  ctx->chain_max_clusters = g_chain_clusters[3] > 0 ? g_chain_clusters[3] : g_chain_clusters[2];
  return 0;
}

// Fold the packed fp32 layers of ctx->unet (api.cu) into the per-phase weight tiles and tables.
int chain_pack(ldm_ctx* ctx, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  ChainModel& C = ctx->chain;
  for (void* p : C.allocs) cudaFree(p);
  C = ChainModel();
  const int nst = U.nst, L = U.latent;
  const int split = ctx->precision == LDM_PRECISION_FP32 ? 1 : 0;   // strict mode: three-term bf16 split (see ChainParams::split)
  LDM_CHECK(nst + 1 <= LDM_CHAIN_MAX_PHASES && nst <= 5, "chain: more stages (%d) than the persistent kernel is validated for: the per-layer path takes over", nst);
  LDM_CHECK(L % 128 == 0 && L / 128 <= kSlots, "chain: latent_dim %d unsupported", L);
  for (int i = 0; i < nst; ++i)
    LDM_CHECK(U.hid[i] % 64 == 0 && U.hid[i] / 64 <= CS && U.hid[i] / 64 <= kSlots, "chain: hidden dim %d unsupported", U.hid[i]);
  auto& PA = C.allocs;
  std::vector<void*> tmp;
  auto free_tmp = [&]() { for (void* p : tmp) cudaFree(p); tmp.clear(); };
  int rc = 0;
  auto body = [&]() -> int {
    C.n_phases = nst + 1;
    // ---- natural-order folded matrices, biases and tables, then the tile-order / bf16 copies
    for (int j = 0; j <= nst; ++j) {
      ChainPhaseHost& H = C.ph[j];
      int rows, K, d, stage, nsr;
      float *Gn = nullptr, *bn = nullptr, *Tn = nullptr, *Cn = nullptr, *g0n = nullptr, *qn = nullptr;   // natural order (temporaries)
      if (j == 0) {
        // MERGED phase.  With x_{t-1} = x~ - c_b eps, x~ = x_t / sqrt(alpha_t) + sigma_t z_t, c_b = c2 / sqrt(alpha_t)
        // and eps = W_f' [LN_f(h) ; x_t] + b_fin (v2:560-561, 584-592), the first stage of the NEXT forward,
        //   [h_0 | u_0] = G_0 x_{t-1} + ...,   G_0 = [W_lp ; W_b0 W_lp]                         (v2:539, 546)
        // is linear in the operand [x~ | -c_b LN_f(h) | -c_b x_t]:
        //   stage rows : [G_0 | G_0 W_f']        (+ tables, - c_b G_0 b_fin added by the epilogue)
        //   eps rows   : [ 0  |   W_f'  ]        -> -c_b (eps - b_fin): finishes the PREVIOUS forward, off the critical path
        d = U.hid[0]; nsr = 2 * d; rows = nsr + L; K = 3 * L; stage = 1;
        LDM_CHECK(U.fin.N == L && U.fin.K == 2 * L, "chain: final layer shape");
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Gn, (size_t)rows * K));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &bn, (size_t)rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Tn, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Cn, (size_t)U.ncls * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &g0n, (size_t)rows));
        LDM_CUDA(cudaMemsetAsync(Gn, 0, sizeof(float) * (size_t)rows * K, st));
        LDM_CUDA(cudaMemsetAsync(Tn, 0, sizeof(float) * (size_t)U.n_t * rows, st));
        LDM_CUDA(cudaMemsetAsync(Cn, 0, sizeof(float) * (size_t)U.ncls * rows, st));
        LDM_CUDA(cudaMemsetAsync(g0n, 0, sizeof(float) * (size_t)rows, st));
        LDM_TRY(copy2d(ctx, U.latent_proj.w32, L, Gn, K, d, L, st));                                           // W_lp
        LDM_TRY(mm(ctx, U.block[0].w32, d, U.latent_proj.w32, L, 0, Gn + (size_t)d * K, K, d, L, d, st));     // W_b0 W_lp
        LDM_TRY(mm(ctx, Gn, K, U.fin.w32, 2 * L, 0, Gn + L, K, nsr, 2 * L, L, st));                           // G_0 W_f'
        LDM_TRY(copy2d(ctx, U.fin.w32, 2 * L, Gn + (size_t)nsr * K + L, K, L, 2 * L, st));                    // eps rows
        LDM_TRY(copy2d(ctx, U.latent_proj.b, 1, bn, 1, d, 1, st));
        LDM_TRY(mv(ctx, U.block[0].w32, d, U.latent_proj.b, U.block[0].b, bn + d, d, d, st));
        LDM_TRY(copy2d(ctx, U.fin.b, 1, bn + nsr, 1, L, 1, st));                                               // b_fin on the eps rows
        LDM_TRY(mv(ctx, Gn, K, U.fin.b, nullptr, g0n, nsr, L, st));                                            // G_0 b_fin
        LDM_TRY(copy2d(ctx, U.tab_t[0], d, Tn, rows, U.n_t, d, st));
        LDM_TRY(mm(ctx, U.tab_t[0], d, U.block[0].w32, d, 1, Tn + d, rows, U.n_t, d, d, st));
        LDM_TRY(copy2d(ctx, U.tab_c[0], d, Cn, rows, U.ncls, d, st));
        LDM_TRY(mm(ctx, U.tab_c[0], d, U.block[0].w32, d, 1, Cn + d, rows, U.ncls, d, d, st));
        H.type = LDM_PH_MERGED;
        H.nst_tiles = nsr / 128;
        H.eps_kb0 = split ? 0 : L / BK;   // strict mode: the x~ block recurs in every third of the operand (its eps weights are zero)
      } else {
        // Operand: raw h2 of stage i = j-1 (dp wide).  With n = LN_b(h2) = r Gamma (h2 - mu) + beta (v2:549) and the folded
        // L = 1 attention A n + a (v2:550-552), h_{j} = D (h2 + A n + a) + b_d (v2:553) becomes
        //   h_j = W1 h2 + r (W2 h2 - mu q) + const,  W1 = D, W2 = D A Gamma, q = W2 1, const = D (A beta + a) + b_d
        // so LayerNorm_b is applied AFTER the contraction from the per-row (mu, r) that the previous phase exchanges in the
        // background.  Weights: [W1 | W2] (rows x 2dp), two accumulators.
        const int i = j - 1, dp = U.hid[i], dn = U.hid[i + 1];
        const bool last = j == nst;
        d = dn; K = 2 * dp; rows = last ? dn : 2 * dn; stage = last ? 0 : 1; nsr = rows;
        float *wv = nullptr, *ones = nullptr;
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Gn, (size_t)rows * K));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &bn, (size_t)rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Tn, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &Cn, (size_t)U.ncls * rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &qn, (size_t)rows));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &wv, (size_t)dp));
        LDM_TRY(ldm_alloc_t(ctx, tmp, &ones, (size_t)dp));
        LDM_TRY(copy2d(ctx, U.down[i].w32, dp, Gn, K, dn, dp, st));                                   // W1 = D
        LDM_TRY(mm(ctx, U.down[i].w32, dp, U.ov[i].w32, dp, 0, Gn + dp, K, dn, dp, dp, st));          // D A
        {
          const size_t n = (size_t)dn * dp;
          pack_scale_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Gn + dp, K, U.ln_b_w[i], dn, dp);   // W2 = D A Gamma
          LDM_LAUNCHED(ctx);
          pack_fill_kernel<<<ceil_div(dp, 256), 256, 0, st>>>(ones, 1.0f, dp);
          LDM_LAUNCHED(ctx);
        }
        LDM_TRY(mv(ctx, U.ov[i].w32, dp, U.ln_b_b[i], U.ov[i].b, wv, dp, dp, st));                    // A beta + a
        LDM_TRY(mv(ctx, U.down[i].w32, dp, wv, U.down[i].b, bn, dn, dp, st));                         // D (A beta + a) + b_d
        LDM_TRY(copy2d(ctx, U.tab_t[j], dn, Tn, rows, U.n_t, dn, st));
        LDM_TRY(copy2d(ctx, U.tab_c[j], dn, Cn, rows, U.ncls, dn, st));
        if (!last) {   // u_j rows: W_b,j applied to everything above
          const float* Wb = U.block[j].w32;
          LDM_TRY(mm(ctx, Wb, dn, Gn, K, 0, Gn + (size_t)dn * K, K, dn, K, dn, st));
          LDM_TRY(mv(ctx, Wb, dn, bn, U.block[j].b, bn + dn, dn, dn, st));
          LDM_TRY(mm(ctx, U.tab_t[j], dn, Wb, dn, 1, Tn + dn, rows, U.n_t, dn, dn, st));
          LDM_TRY(mm(ctx, U.tab_c[j], dn, Wb, dn, 1, Cn + dn, rows, U.ncls, dn, dn, st));
        }
        LDM_TRY(mv(ctx, Gn + dp, K, ones, nullptr, qn, rows, dp, st));                                 // q = W2 1 (all rows)
        H.dual = 1;
        H.type = last ? LDM_PH_FINAL_LN : LDM_PH_STAGE;
      }
      LDM_CHECK(rows % 128 == 0 && K % BK == 0 && rows / 128 <= CS, "chain: phase %d shape (%d x %d) unsupported", j, rows, K);
      const int Kp = H.dual ? K / 2 : K, mult = split ? 3 : 1;      // per-accumulator reduction length, natural and as the kernel sees it
      H.K = Kp * mult; H.rows = rows; H.tiles = rows / 128; H.d = d;
      // two units per tile when few tiles carry a long reduction: one per accumulator (dual) / one per K half (plain)
      H.ks = (K >= 1024 && H.tiles * 2 <= CS && (K / BK) % 2 == 0 && !getenv("LDM_CHAIN_NO_SPLITK")) ? 2 : 1;
      LDM_TRY(ldm_alloc_t(ctx, PA, &H.w, (size_t)rows * K * mult));
      LDM_TRY(ldm_alloc_t(ctx, PA, &H.bias, (size_t)rows));
      {
        const size_t n = (size_t)rows * K;
        if (split) pack_rows_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Gn, H.w, rows, K, Kp, d, stage, nsr);
        else pack_rows_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Gn, H.w, rows, K, d, stage, nsr);
        LDM_LAUNCHED(ctx);
        pack_cols_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(bn, H.bias, 1, rows, d, stage, nsr);
        LDM_LAUNCHED(ctx);
        if (g0n) {
          LDM_TRY(ldm_alloc_t(ctx, PA, &H.g0b, (size_t)rows));
          pack_cols_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(g0n, H.g0b, 1, rows, d, stage, nsr);
          LDM_LAUNCHED(ctx);
        }
        if (qn) {
          LDM_TRY(ldm_alloc_t(ctx, PA, &H.q, (size_t)rows));
          pack_cols_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(qn, H.q, 1, rows, d, stage, nsr);
        }
        LDM_LAUNCHED(ctx);
      }
      if (Tn) {
        LDM_TRY(ldm_alloc_t(ctx, PA, &H.tab_t, (size_t)U.n_t * rows));
        LDM_TRY(ldm_alloc_t(ctx, PA, &H.tab_c, (size_t)U.ncls * rows));
        size_t n = (size_t)U.n_t * rows;
        pack_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Tn, H.tab_t, U.n_t, rows, d, stage, nsr);
        LDM_LAUNCHED(ctx);
        n = (size_t)U.ncls * rows;
        pack_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Cn, H.tab_c, U.ncls, rows, d, stage, nsr);
        LDM_LAUNCHED(ctx);
      }
      LDM_TRY(tc_make_weight_map(ctx, H.w, rows, K * mult, 128, &H.map));   // box = 128 rows x 64 k
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
    auto bytes = [&](int j) {   // weight tile(s) + operand of one unit
      const ChainPhaseHost& H = C.ph[j];
      if (H.dual) return H.ks == 1 ? (double)H.K * (512.0 + 96.0) : (double)H.K * (256.0 + 96.0);
      return ((double)H.K * 256.0 + (double)H.K * 96.0) / H.ks;
    };
    auto units = [&](int j) { return C.ph[j].tiles * C.ph[j].ks; };
    for (int a = 0; a < C.n_phases; ++a)
      for (int b = a + 1; b < C.n_phases; ++b)
        if (bytes(order[b]) * units(order[b]) > bytes(order[a]) * units(order[a])) { int t = order[a]; order[a] = order[b]; order[b] = t; }
    for (int a = 0; a < C.n_phases; ++a) {
      ChainPhaseHost& H = C.ph[order[a]];
      int best = 0;
      double best_peak = 1e300, best_sum = 1e300;
      const int nu = H.tiles * H.ks;
      const bool avoid0 = false;
      const int p0a = 0, p0b = 0;
      for (int pass = 0; pass < 2 && best_peak > 1e299; ++pass)
      for (int f = 0; f + nu <= CS; ++f) {
        if (pass == 0 && avoid0 && f < p0b && f + nu > p0a) continue;
        double peak = 0, sum = 0;
        for (int r = f; r < f + nu; ++r) { peak = load[r] > peak ? load[r] : peak; sum += load[r]; }
        if (peak < best_peak - 1e-9 || (peak < best_peak + 1e-9 && sum < best_sum)) { best_peak = peak; best_sum = sum; best = f; }
      }
      H.first = best;
      for (int r = best; r < best + nu; ++r) load[r] += bytes(order[a]);
    }
    C.peak_bytes_per_step = 0;
    for (int r = 0; r < CS; ++r) C.peak_bytes_per_step = load[r] > C.peak_bytes_per_step ? load[r] : C.peak_bytes_per_step;
  }
  if (getenv("LDM_CHAIN_DEBUG"))
    for (int j = 0; j < C.n_phases; ++j)
      fprintf(stderr, "chain phase %d: type %d rows %d K %d tiles %d ks %d first %d\n", j, C.ph[j].type, C.ph[j].rows, C.ph[j].K,
              C.ph[j].tiles, C.ph[j].ks, C.ph[j].first);
  C.ready = true;
  return 0;
}

// Batch rows per cluster for a batch of B rows: the variant with the fewest waves of co-resident clusters, then the
// fewest rows per cluster (shortest epilogue).
static int chain_pick_nw(int B, int n_phases) {
  int best = 0, best_waves = 1 << 30;
  if (const char* f = getenv("LDM_CHAIN_NW")) { const int nw = atoi(f); if (nw >= 2 && nw <= 4 && g_chain_clusters[nw] >= 1) return nw; }
  for (int nw = 2; nw <= 4; ++nw) {
    if ((kChains + 2 + (nw >= 3 ? n_phases : 0)) * 16 * nw > kTmemCols) continue;   // accumulators + parked noise + state (+ per-sample terms) must fit the TMEM columns
    if (g_chain_clusters[nw] < 1) continue;
    const int waves = ceil_div(ceil_div(B, 16 * nw), g_chain_clusters[nw]);
    if (waves < best_waves) { best_waves = waves; best = nw; }
  }
  return best;
}

// Run `n_iter` reverse steps (sample = 1) or one forward evaluation (sample = 0) for `B` rows.
// `x`: (B, latent) fp32 input; in sample mode it is also the state that receives x_{t_end - 1}.
int launch_chain(ldm_ctx* ctx, int B, int n_iter, int t_start, int sample, const int64_t* t_idx, int t_len, float* x,
                 float* eps_out, const float* noise, cudaStream_t st) {
  UnetModel& U = ctx->unet;
  ChainModel& C = ctx->chain;
  LDM_CHECK(C.ready, "chain: weights not packed");
  LDM_CHECK(!sample || ctx->coef_dev != nullptr, "chain: schedule not set");
  const int nw = chain_pick_nw(B, C.n_phases);
  LDM_CHECK(nw >= 2, "chain: no cluster variant fits this device");
  const int NB = 16 * nw;
  ChainParams P;
  memset(&P, 0, sizeof(P));
  const int nst = U.nst, L = U.latent;
  LDM_CHECK(nst + 2 <= kMaxXMaps, "chain: too many stages");
  // stage the first merged operand: [x | 0 | 0] (the stage tiles see G_0 x; the eps tiles have nothing to finish yet)
  const int split = ctx->precision == LDM_PRECISION_FP32 ? 1 : 0, mult = split ? 3 : 1;
  {
    const size_t n = (size_t)B * 3 * L;
    chain_stage_x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, ctx->caf[0], B, L, split);
    LDM_LAUNCHED_AS(ctx, "chain_stage_x");
  }
  // operand descriptors: rows beyond the batch are zero-filled by the TMA unit
  LDM_TRY(tc_make_act_map(ctx->caf[0], B, 3 * L * mult, 3 * L * mult, NB, &P.xmaps[0]));
  LDM_TRY(tc_make_act_map(ctx->caf[1], B, 3 * L * mult, 3 * L * mult, NB, &P.xmaps[1]));
  for (int j = 0; j < nst; ++j) LDM_TRY(tc_make_act_map(ctx->opbuf[j], B, U.hid[j] * mult, U.hid[j] * mult, NB, &P.xmaps[2 + j]));
  for (int j = nst; j < kMaxXMaps - 2; ++j) P.xmaps[2 + j] = P.xmaps[0];
  LDM_TRY(tc_make_kblock_map(ctx->caf[0], B, 3 * L * mult, 3 * L * mult, NB, kXGroup, &P.xmaps4[0]));
  LDM_TRY(tc_make_kblock_map(ctx->caf[1], B, 3 * L * mult, 3 * L * mult, NB, kXGroup, &P.xmaps4[1]));
  for (int j = 0; j < nst; ++j) LDM_TRY(tc_make_kblock_map(ctx->opbuf[j], B, U.hid[j] * mult, U.hid[j] * mult, NB, kXGroup, &P.xmaps4[2 + j]));
  for (int j = nst; j < kMaxXMaps - 2; ++j) P.xmaps4[2 + j] = P.xmaps4[0];
  { const char* xg = getenv("LDM_CHAIN_XGROUP"); P.xgroup = xg ? atoi(xg) : 1; }
  int cadd = (kChains + 2) * NB;      // TMEM: accumulators, parked noise, chain state, then the per-sample terms (NW >= 3)
  for (int j = 0; j < C.n_phases; ++j) {
    const ChainPhaseHost& H = C.ph[j];
    ChainPhase& D = P.ph[j];
    P.wmap[j] = H.map;
    D.type = H.type; D.K = H.K; D.tiles = H.tiles; D.first = H.first; D.ks = H.ks; D.d = H.d; D.rows = H.rows;
    D.nst_tiles = H.nst_tiles; D.eps_kb0 = H.eps_kb0; D.g0b = H.g0b;
    D.dual = H.dual; D.q = H.q;
    D.prev_tiles = j == 0 ? 0 : (j == 1 ? C.ph[0].nst_tiles : C.ph[j - 1].tiles);
    D.bias = H.bias; D.tab_t = H.tab_t; D.tab_c = ctx->has_cls ? H.tab_c : nullptr;
    D.cadd_col = -1;
    if (H.tab_t) { D.cadd_col = nw >= 3 ? cadd : 0; cadd += nw >= 3 ? NB : 0; }
    if (j < nst) { D.ga = U.ln_a_w[j]; D.ba = U.ln_a_b[j]; D.gb = U.ln_b_w[j]; D.bb = U.ln_b_b[j]; }
    else { D.ga = U.ln_f_w; D.ba = U.ln_f_b; }
    if (j == 0) { D.xmap = 0; D.xmap_alt = 1; D.xcol = 0; }
    else { D.xmap = 2 + (j - 1); D.xmap_alt = 0; D.xcol = 0; }
    if (j < nst) { D.out = ctx->opbuf[j]; D.ld_out = U.hid[j] * mult; D.wout = U.hid[j]; }
  }
  for (int j = C.n_phases; j < LDM_CHAIN_MAX_PHASES; ++j) P.wmap[j] = C.ph[0].map;
  P.z_col = kChains * NB;
  P.x_col = kChains * NB + NB;
  { const char* wf = getenv("LDM_CHAIN_WRITER_FENCE"); P.writer_fence = wf ? atoi(wf) : 0; }
  LDM_CHECK(cadd <= kTmemCols, "chain: TMEM columns exhausted (%d phases x %d rows)", C.n_phases, NB);
  P.n_phases = C.n_phases;
  P.B = B; P.row_begin = 0; P.row_end = B;
  P.n_iter = n_iter; P.t_start = t_start; P.sample = sample; P.latent = L; P.n_t = U.n_t;
  P.t_idx = t_idx; P.t_len = t_len;
  P.cls = ctx->has_cls ? ctx->cls : nullptr;
  P.x = x; P.eps_out = eps_out; P.noise = noise; P.rng = ctx->rng_dev;
  P.coef = ctx->coef_dev;
  P.coef_or_one = sample ? ctx->coef_dev : ctx->coef_one;
  P.af[0] = ctx->caf[0]; P.af[1] = ctx->caf[1]; P.ld_af = 3 * L * mult; P.split = split;
  P.err = ctx->chain_err;
  P.trace = ctx->chain_trace;
  P.trace_step = ctx->chain_trace_step;
  const int nclusters = ceil_div(B, NB);
  int rc = -1;
  if (nw == 2) rc = launch_variant<2>(P, nclusters, st);
  else if (nw == 3) rc = launch_variant<3>(P, nclusters, st);
  else rc = launch_variant<4>(P, nclusters, st);
  LDM_TRY(rc);
  ctx->launches++;
  ldm_kmark(ctx, "chain_kernel");
  return 0;
}
