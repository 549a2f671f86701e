// Philox4x32-10 + Box-Muller: device side of the noise stream specified in oracle/philox.py
// (tests/test_philox_spec.py pins the spec with known answers; tests/test_gpu_parity.py compares the in-kernel draws of
// ldm_randn / the chain with it).
//   key = (seed lo, seed hi); counter = (quad, sample lo, step, sample hi)
//   r0..r3 -> u(r) = ((r >> 8) + 0.5) * 2^-24 ; (z0, z1) = sqrt(-2 ln u(r0)) (cos, sin)(2 pi u(r1)), same for (r2, r3)
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float philox_u01(uint32_t r) {
  // (k + 0.5) 2^-24 needs 25 significant bits once k = r >> 8 reaches 2^23: the add then rounds to even (one fp32 rounding,
  // at most 2^-25 away from the float64 spec of oracle/philox.py; u may reach 1.0f, for which the radius is 0)
  return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-08f;
}

// four standard normals for elements [4*quad, 4*quad+4) of global sample `sample` at `step`
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long sample,
                                                 uint32_t step, uint32_t quad) {
  uint4 r = philox4x32_10(make_uint4(quad, (uint32_t)sample, step, (uint32_t)(sample >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float4 z;
  float rad0 = sqrtf(-2.0f * logf(philox_u01(r.x)));
  float rad1 = sqrtf(-2.0f * logf(philox_u01(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * philox_u01(r.y), &s0, &c0);
  sincospif(2.0f * philox_u01(r.w), &s1, &c1);
  z.x = rad0 * c0;
  z.y = rad0 * s0;
  z.z = rad1 * c1;
  z.w = rad1 * s1;
  return z;
}

// The posterior update of the reference, operation for operation (v2:586-592): separate fp32
// roundings (no FMA contraction) so that, given the same eps and noise, the result is bit-equal to
// torch's elementwise kernels:  mean = (x - c2 * eps) / sqrt(alpha_t);  x' = mean + sqrt(beta_t) * z
__device__ __forceinline__ float ddpm_update_one(float x, float eps, float c2, float sqrt_alpha, float sigma,
                                                 float z) {
  float mean = __fdiv_rn(__fsub_rn(x, __fmul_rn(c2, eps)), sqrt_alpha);
  return sigma > 0.0f ? __fadd_rn(mean, __fmul_rn(sigma, z)) : mean;
}
