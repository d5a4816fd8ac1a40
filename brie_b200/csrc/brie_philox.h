// Counter-based MC noise shared by the CUDA kernels and the host-side dump.
//
// The reference draws noise from TensorFlow's stateful global RNG
// (brie/models/model_TFProb.py:159 `self.Z.sample(MC_size)`, init at :18-31),
// which is unseeded and not reproducible.  This spec replaces it with a
// stateless generator keyed by (seed, phase, model, step, cell, event, sample)
// so the noise is independent of tiling, event sharding and launch order.
//
//   key     = (seed lo32, seed hi32)
//   counter = (event [global index], cell, step, stream)
//   stream  = phase << 28 | model << 16 | block          (block = sample / 4)
//   x0..x3  = Philox4x32-10(counter, key)               (Salmon et al., SC'11)
//   u(x)    = ((x >> 9) + 0.5) * 2^-23                   in (0,1)
//   normals = BoxMuller(u(x0), u(x1)) ++ BoxMuller(u(x2), u(x3))
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BRIE_HD __host__ __device__ __forceinline__
#else
#define BRIE_HD static inline
#endif

#define BRIE_PHASE_TRAIN 0u
#define BRIE_PHASE_EVAL 1u
#define BRIE_PHASE_INIT 2u

#define BRIE_INIT_Z_LOC 0u
#define BRIE_INIT_Z_STD_LOG 1u
#define BRIE_INIT_WC 2u
#define BRIE_INIT_WG 3u
#define BRIE_INIT_INTERCEPT 4u

BRIE_HD uint32_t brie_stream_word(uint32_t phase, uint32_t model, uint32_t block) {
  return (phase << 28) | (model << 16) | block;
}

BRIE_HD void brie_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#if defined(__CUDA_ARCH__)
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

BRIE_HD float brie_u01(uint32_t x) {
#if defined(__CUDA_ARCH__)
  // same value without an int->float conversion: [1,2) mantissa trick, exact subtraction
  return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994039535522f;  // 1 - 2^-24
#else
  return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f;  // 2^-23
#endif
}

// One Box-Muller pair.  On the device the fast SFU paths are used (lg2/sin/cos
// approx); on the host libm.  They agree to ~3e-6 absolute in the normals;
// tests that need the device's exact noise read it back with
// brie_philox_normals_device().
BRIE_HD void brie_box_muller(uint32_t a, uint32_t b, float* n0, float* n1) {
  const float u1 = brie_u01(a), u2 = brie_u01(b);
#if defined(__CUDA_ARCH__)
  float lg, r, sn, cs;
  const float t = 6.283185307179586f * u2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));  // -2 ln u1
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(t));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(t));
#else
  const float r = sqrtf(-2.0f * logf(u1));
  const float t = 6.283185307179586f * u2;
  const float sn = sinf(t), cs = cosf(t);
#endif
  *n0 = r * cs;
  *n1 = r * sn;
}

// Four normals for one (event, cell, step, stream) counter.
BRIE_HD void brie_normals4(uint32_t event, uint32_t cell, uint32_t step, uint32_t stream,
                           uint64_t seed, float out[4]) {
  uint32_t x[4];
  brie_philox4x32_10(event, cell, step, stream, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  brie_box_muller(x[0], x[1], &out[0], &out[1]);
  brie_box_muller(x[2], x[3], &out[2], &out[3]);
}
