// Micro-benchmarks behind the chain-kernel design decisions.  Stand-alone: no library code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench.bin tools/ubench.cu -lcuda
//   tools/ubench.bin            (on a B200)
// 1. TMA streaming rate L2 -> shared memory per SM (16 KiB 128B-swizzled boxes through a ring) for several grid sizes
// 2. DSMEM all-to-all with cp.async.bulk shared::cta -> shared::cluster (6 KiB per peer, 16-CTA cluster)
// 3. DSMEM st.async (8 B) ping-pong latency and remote mbarrier arrive ping-pong latency
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1, 0x989680;\n\t"
      "@!P bra W_%=;\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%0], %1, 0x989680;\n\t"
      "@!P bra W_%=;\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// ------------------------------------------------------------------------------------------- 1. TMA streaming
constexpr int kBox = 16384;
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int rows, int kblocks, int n_loads, int stages,
                                                       long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[16], empty[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles = rows / 128;
  long long t0 = 0;
  if (threadIdx.x == 0) {   // producer
    t0 = clock64();
    for (int n = 0; n < n_loads; ++n) {
      const int s = n % stages, par = (n / stages) & 1;
      if (n >= stages) mbar_wait(&empty[s], par ^ 1);
      const int lin = (blockIdx.x * 7 + n) % (tiles * kblocks);
      mbar_expect_tx(&full[s], kBox);
      tma_load_2d(ring + (size_t)s * kBox, &map, &full[s], (lin % kblocks) * 64, (lin / kblocks) * 128);
    }
  } else if (threadIdx.x == 32) {   // consumer: frees the slot as soon as it is full
    for (int n = 0; n < n_loads; ++n) {
      const int s = n % stages, par = (n / stages) & 1;
      mbar_wait(&full[s], par);
      mbar_arrive(&empty[s]);
    }
    out[blockIdx.x] = clock64();
  }
  if (threadIdx.x == 0) out[gridDim.x + blockIdx.x] = t0;
}

// ------------------------------------------------------------------------------------------- 2. DSMEM bulk all-to-all
constexpr int kCS = 16;
__global__ void __launch_bounds__(128, 1) bulk_a2a_kernel(int chunk, int reps, int ndst, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* src = base;                       // chunk bytes
  uint8_t* dst = base + 8192;                // kCS x chunk
  __shared__ __align__(8) uint64_t bar;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < chunk; i += blockDim.x) src[i] = (uint8_t)(i + rank);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (uint32_t)(ndst * chunk));
      for (int d = 0; d < ndst; ++d) {
        const uint32_t peer = (rank + d) % kCS;
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa(smem_u32(dst + (size_t)rank * chunk), peer)), "r"(smem_u32(src)), "r"(chunk), "r"(mapa(smem_u32(&bar), peer)) : "memory");
      }
      mbar_wait_cluster(&bar, r & 1);
    }
    __syncthreads();
  }
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; }
}

// ------------------------------------------------------------------------------------------- 3. ping-pong latencies
// mode 0: st.async 8 bytes + complete_tx; mode 1: remote mbarrier arrive (release.cluster) / wait (acquire.cluster)
__global__ void __launch_bounds__(32, 1) pingpong_kernel(int mode, int reps, int peer_rank, long long* out) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) float2 slot;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  cluster_sync();
  if (threadIdx.x == 0 && (rank == 0 || rank == (uint32_t)peer_rank)) {
    const uint32_t other = rank == 0 ? (uint32_t)peer_rank : 0u;
    const uint32_t rbar = mapa(smem_u32(&bar), other), rslot = mapa(smem_u32(&slot), other);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (rank == 0) {
        if (mode == 0) {
          asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(rslot), "f"(1.0f), "f"(2.0f), "r"(rbar) : "memory");
          mbar_expect_tx(&bar, 8);
          mbar_wait(&bar, r & 1);
        } else {
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
          mbar_wait_cluster(&bar, r & 1);
        }
      } else {
        if (mode == 0) {
          mbar_expect_tx(&bar, 8);
          mbar_wait(&bar, r & 1);
          asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(rslot), "f"(1.0f), "f"(2.0f), "r"(rbar) : "memory");
        } else {
          mbar_wait_cluster(&bar, r & 1);
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        }
      }
    }
    if (rank == 0) out[0] = clock64() - t0;
  }
  cluster_sync();
}


// ------------------------------------------------------------------------------------------- 4. tcgen05.mma operand paths
// mode 0: A and B from shared memory (SS); mode 1: A from tensor memory (TS), B from shared memory; mode 2: tcgen05.cp only
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// element (r, k) of a K-major 128B-swizzled [rows x 64] bf16 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15));
}
__global__ void __launch_bounds__(128, 1) mma_kernel(int mode, int N, int reps, long long* out, float* dout) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;             // A tile: 128 x 64 bf16
  uint8_t* sb = base + 16384;     // B tile: N x 64 bf16 (N <= 256)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sa + sw128_off(r, k)) = __float2bfloat16((float)(((r * 3 + k * 5) % 7) - 3));
  }
  for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sb + sw128_off(r, k)) = __float2bfloat16((float)(((r * 2 + k) % 5) - 2));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const uint32_t id = idesc_bf16(128, N);
  const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
  const uint32_t acc_ss = tm, acc_ts = tm + 256, wtm = tm + 480;   // 32 columns of A in TMEM
  uint32_t par = 0;
  if (threadIdx.x == 0) {
    // ---- correctness: D1 = SS over 4 k-steps; A -> TMEM by tcgen05.cp; D2 = TS over 4 k-steps
    for (int k = 0; k < 4; ++k) mma_ss(acc_ss, da + 2 * k, db + 2 * k, id, k != 0);
    for (int k = 0; k < 4; ++k) cp_128x256b(wtm + 8 * k, da + 2 * k);
    for (int k = 0; k < 4; ++k) mma_ts(acc_ts, wtm + 8 * k, db + 2 * k, id, k != 0);
    commit(&bar);
    mbar_wait(&bar, par); par ^= 1;
    // ---- timing
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0) { for (int k = 0; k < 4; ++k) mma_ss(acc_ss, da + 2 * k, db + 2 * k, id, 1); }
      else if (mode == 1) { for (int k = 0; k < 4; ++k) mma_ts(acc_ts, wtm + 8 * k, db + 2 * k, id, 1); }
      else if (mode == 2) { for (int k = 0; k < 4; ++k) cp_128x256b(wtm + 8 * k, da + 2 * k); }
      else { for (int k = 0; k < 4; ++k) cp_128x256b(wtm + 8 * k, da + 2 * k); for (int k = 0; k < 4; ++k) mma_ts(acc_ts, wtm + 8 * k, db + 2 * k, id, 1); }
    }
    const long long t1 = clock64();
    commit(&bar);
    mbar_wait(&bar, par); par ^= 1;
    out[0] = clock64() - t0;
    out[1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0 && dout) {   // dump both results of the correctness pass (first N columns): thread = lane (row)
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 8) {
      uint32_t v[8], w[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(lane_addr + c));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(lane_addr + 256 + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) {
        dout[(size_t)threadIdx.x * 256 + c + i] = __uint_as_float(v[i]);
        dout[(size_t)(128 + threadIdx.x) * 256 + c + i] = __uint_as_float(w[i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


// ------------------------------------------------------------------------------------------- 5. tcgen05.mma issue patterns
// SS MMAs of shape M x N x 16 rotating over `nacc` independent accumulators (column blocks of TMEM) and, with `nab` > 1,
// over distinct A k-slices: is the 74-clock floor of a small-N MMA a per-accumulator dependency or operand delivery?
__global__ void __launch_bounds__(128, 1) mma_rot_kernel(int M, int N, int nacc, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;
  uint8_t* sb = base + 16384;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const uint32_t id = idesc_bf16(M, N);
  const uint64_t da = desc_sw128(smem_u32(sa)), db = desc_sw128(smem_u32(sb));
  const uint32_t stride = (uint32_t)(512 / nacc);
  if (threadIdx.x == 0) {
    for (int j = 0; j < nacc; ++j) mma_ss(tm + j * stride, da, db, id, 0);
    commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    int j = 0;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k) {
        mma_ss(tm + j * stride, da + 2 * k, db + 2 * k, id, 1);
        j = j + 1 == nacc ? 0 : j + 1;
      }
    const long long t1 = clock64();
    commit(&bar);
    mbar_wait(&bar, 1);
    out[0] = clock64() - t0;
    out[1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  long long* out;
  CK(cudaMalloc(&out, sizeof(long long) * 1024));
  std::vector<long long> h(1024);

  // ---- 1. streaming: an 11 MiB bf16 matrix (5632 rows x 1024), L2 resident
  const int rows = 5632, K = 1024;
  __nv_bfloat16* w;
  CK(cudaMalloc(&w, (size_t)rows * K * 2));
  CK(cudaMemset(w, 0, (size_t)rows * K * 2));
  CUtensorMap map;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int n_loads = 4000;
  for (int stages : {2, 4, 8, 12}) {
    for (int grid : {1, 16, 64, 96, 112, 148}) {
      for (int rep = 0; rep < 2; ++rep) {
        stream_kernel<<<grid, 64, stages * kBox + 1024>>>(map, rows, K / 64, n_loads, stages, out);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
      double worst = 0, mean = 0;
      for (int b = 0; b < grid; ++b) { double c = (double)(h[b] - h[grid + b]); mean += c / grid; worst = c > worst ? c : worst; }
      printf("stream stages %2d grid %3d: %.1f B/clk/SM mean, %.1f worst-CTA; chip %.0f B/clk; %.0f cyc per 16 KiB box\n", stages, grid,
             (double)n_loads * kBox / mean, (double)n_loads * kBox / worst, (double)n_loads * kBox * grid / mean, mean / n_loads);
    }
  }

  // ---- 2. DSMEM bulk all-to-all in one 16-CTA cluster
  {
    CK(cudaFuncSetAttribute(bulk_a2a_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(bulk_a2a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    for (int nclusters : {1, 6}) {
      for (int chunk : {2048, 4096, 6144}) {
        for (int ndst : {4, 8, 16}) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(kCS * nclusters); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 8192 + kCS * 6144 + 1024;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          const int reps = 200;
          for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelEx(&cfg, bulk_a2a_kernel, chunk, reps, ndst, out)); CK(cudaDeviceSynchronize()); }
          CK(cudaMemcpy(h.data(), out, sizeof(long long) * kCS * nclusters, cudaMemcpyDeviceToHost));
          double worst = 0;
          for (int b = 0; b < kCS * nclusters; ++b) worst = (double)h[b] > worst ? (double)h[b] : worst;
          printf("bulk a2a clusters %d chunk %d B x %2d peers: %.0f cyc per round, %.1f B/clk/SM outbound\n", nclusters, chunk, ndst, worst / reps,
                 (double)chunk * ndst * reps / worst);
        }
      }
    }
  }
  // ---- 3. ping-pong latencies
  {
    CK(cudaFuncSetAttribute(pingpong_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int mode = 0; mode < 2; ++mode)
      for (int peer : {1, 8, 15}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCS); cfg.blockDim = dim3(32);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const int reps = 1000;
        CK(cudaLaunchKernelEx(&cfg, pingpong_kernel, mode, reps, peer, out));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, sizeof(long long), cudaMemcpyDeviceToHost));
        printf("pingpong %s rank 0 <-> %2d: %.0f cyc round trip (%.0f one way)\n", mode == 0 ? "st.async+complete_tx" : "remote mbarrier arrive", peer,
               (double)h[0] / reps, (double)h[0] / reps / 2);
      }
  }
  // ---- 4. tcgen05.mma: operands from shared memory vs A from tensor memory
  {
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    float* dout;
    CK(cudaMalloc(&dout, sizeof(float) * 256 * 256));
    std::vector<float> hd(256 * 256);
    for (int N : {48, 32, 64, 128, 256}) {
      CK(cudaMemset(dout, 0, sizeof(float) * 256 * 256));
      mma_kernel<<<1, 128, 16384 + 32768 + 1024>>>(0, N, 64, out, dout);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hd.data(), dout, sizeof(float) * 256 * 256, cudaMemcpyDeviceToHost));
      int bad_ss = 0, bad_ts = 0;
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < N; ++j) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += (float)(((r * 3 + k * 5) % 7) - 3) * (float)(((j * 2 + k) % 5) - 2);
          if (hd[(size_t)r * 256 + j] != ref) bad_ss++;
          if (hd[(size_t)(128 + r) * 256 + j] != ref) bad_ts++;
        }
      printf("mma N=%3d correctness: SS mismatches %d, cp + TS mismatches %d (of %d)\n", N, bad_ss, bad_ts, 128 * N);
      for (int mode = 0; mode < 4; ++mode) {
        const int reps = 64;
        mma_kernel<<<1, 128, 16384 + 32768 + 1024>>>(mode, N, reps, out, nullptr);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, sizeof(long long) * 2, cudaMemcpyDeviceToHost));
        const char* nm[4] = {"SS mma", "TS mma (A in TMEM)", "tcgen05.cp 128x256b", "cp + TS mma"};
        printf("mma N=%3d %-22s: %.1f cyc per K=16 step (issue alone %.1f)\n", N, nm[mode], (double)h[0] / (reps * 4), (double)h[1] / (reps * 4));
      }
    }
  }
  // ---- 5. issue patterns: rotating accumulators, M = 64
  {
    CK(cudaFuncSetAttribute(mma_rot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int M : {128, 64})
      for (int N : {48, 64, 128, 256})
        for (int nacc : {1, 2, 4}) {
          if (N * nacc > 512) continue;
          if (M == 64 && (N % 8)) continue;
          mma_rot_kernel<<<1, 128, 16384 + 32768 + 1024>>>(M, N, nacc, 64, out);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h.data(), out, sizeof(long long) * 2, cudaMemcpyDeviceToHost));
          printf("mma-rot M=%3d N=%3d accumulators %d: %.1f cyc per K=16 step (issue alone %.1f)\n", M, N, nacc, (double)h[0] / 256, (double)h[1] / 256);
        }
  }
  return 0;
}
