// Finishing of GEMM accumulators, shared by the CUDA-core and tcgen05 kernels (see Epilogue in common.cuh).
#pragma once
#include "common.cuh"
#include "philox.cuh"

__device__ __forceinline__ int epi_trow(const Epilogue& e, int row) {
  long long t = e.t_idx ? e.t_idx[e.t_len == 1 ? 0 : row] : (long long)e.t_const;
  t = t < 0 ? 0 : (t >= e.n_t ? e.n_t - 1 : t);   // range is validated upstream; clamp keeps reads in bounds
  return (int)t;
}

// Finish 4 consecutive columns [col, col+4) of one row (col % 4 == 0, all in range).
__device__ __forceinline__ void epi_finish4(const Epilogue& e, int row, int col, int N, float v[4]) {
  if (e.bias) {
    float4 b = *reinterpret_cast<const float4*>(e.bias + col);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (e.tab_t) {
    float4 b = *reinterpret_cast<const float4*>(e.tab_t + (size_t)epi_trow(e, row) * e.ld_t + col);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (e.tab_c && e.cls) {
    float4 b = *reinterpret_cast<const float4*>(e.tab_c + (size_t)e.cls[row] * e.ld_c + col);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (e.resid) {
    float4 b = *reinterpret_cast<const float4*>(e.resid + (size_t)row * e.ld_r + col);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (e.act == LDM_ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = swishf(v[i]);
  } else if (e.act == LDM_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sigmoidf_(v[i]);
  }
  if (e.ddpm) {
    float4 xv = *reinterpret_cast<const float4*>(e.x + (size_t)row * N + col);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.sigma > 0.0f) {
      if (e.noise) z = *reinterpret_cast<const float4*>(e.noise + (size_t)row * N + col);
      else z = philox_normal4(e.rng[0], e.rng[1] + (unsigned long long)row, (uint32_t)e.step, (uint32_t)(col >> 2));
    }
    v[0] = ddpm_update_one(xv.x, v[0], e.c2, e.sqrt_alpha, e.sigma, z.x);
    v[1] = ddpm_update_one(xv.y, v[1], e.c2, e.sqrt_alpha, e.sigma, z.y);
    v[2] = ddpm_update_one(xv.z, v[2], e.c2, e.sqrt_alpha, e.sigma, z.z);
    v[3] = ddpm_update_one(xv.w, v[3], e.c2, e.sqrt_alpha, e.sigma, z.w);
    *reinterpret_cast<float4*>(e.x + (size_t)row * N + col) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + (size_t)row * e.ld_of + col) = make_float4(v[0], v[1], v[2], v[3]);
  if (e.vt && col >= e.vt_col0) {      // consecutive lanes hold consecutive rows: each column is one contiguous run
#pragma unroll
    for (int i = 0; i < 4; ++i) e.vt[(size_t)(col - e.vt_col0 + i) * e.ld_vt + row] = __float2bfloat16_rn(v[i]);
    return;
  }
  if (e.out_bf16) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[2], v[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(e.out_bf16 + (size_t)row * e.ld_ob + col) = pk;
  }
}
