// out_conv epilogue of the v4 / v5 pixel path, shared by the CUDA-core kernels (pixel.cu) and the tensor-core halo
// kernel (conv_tc.cu): what happens to the three eps values of one pixel.
#pragma once
#include "common.cuh"
#include "philox.cuh"

struct PixOutArgs {
  const bf16* in;           // (B, H, W, C)
  const float* w;           // (3, 9 C), k = tap * C + ci
  const float* bias;        // (3)
  const float* res_ratio;   // device scalar or null
  const float* x_in;        // (B, 3, H, W): input of the forward (residual term of v5); the state when DDPM
  float* out;               // DDPM == 0: eps;  DDPM == 1: the state (== x_in)
  const float* noise;       // DDPM: explicit (B, 3, H, W) draws or null
  const unsigned long long* rng;   // DDPM: {seed, sample_offset}
  float c2, sqrt_alpha, sigma;
  int step, H, W, C, total_pix;
};

// What happens to the three eps values of pixel `rem` of sample n (shared by both out_conv kernels).  Philox: element
// e = c * HW + rem of the flattened (3, H, W) sample, counter quad e / 4, component e % 4 (oracle/philox.py).
template <int DDPM>
__device__ __forceinline__ void pix_finish(const PixOutArgs& a, float (&eps)[3], int n, int rem, int HW, const float (&z)[3]) {
  const size_t i0 = (size_t)n * 3 * HW + rem;
  if (a.res_ratio) {
    const float rr = __ldg(a.res_ratio);
#pragma unroll
    for (int c = 0; c < 3; ++c) eps[c] = __fadd_rn(eps[c], __fmul_rn(rr, a.x_in[i0 + (size_t)c * HW]));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (DDPM == 0) a.out[i0 + (size_t)c * HW] = eps[c];
    else a.out[i0 + (size_t)c * HW] = ddpm_update_one(a.x_in[i0 + (size_t)c * HW], eps[c], a.c2, a.sqrt_alpha, a.sigma, z[c]);
  }
}


// the three noise draws of pixel `rem` of sample n at this step (explicit tensor or the Philox stream), 0 when sigma == 0
__device__ __forceinline__ void pix_noise(const PixOutArgs& a, int n, int rem, int HW, float (&z)[3]) {
  z[0] = z[1] = z[2] = 0.f;
  if (!(a.sigma > 0.f)) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (a.noise) z[c] = a.noise[(size_t)n * 3 * HW + (size_t)c * HW + rem];
    else {
      const uint32_t e = (uint32_t)(c * HW + rem);
      const float4 z4 = philox_normal4(a.rng[0], a.rng[1] + (unsigned long long)n, (uint32_t)a.step, e >> 2);
      const int j = e & 3;
      z[c] = j == 0 ? z4.x : (j == 1 ? z4.y : (j == 2 ? z4.z : z4.w));
    }
  }
}
