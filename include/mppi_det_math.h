/*
 * mppi_det_math.h -- canonical, bit-reproducible fp32 sine/cosine.
 *
 * Why this exists: the MPPI rollout (reference: nav2_sortham_controller/src/optimizer.cpp:313-343)
 * feeds cos/sin of the running yaw into the x/y cumulative sums, and the trajectory points are then
 * turned into costmap cell indices.  Cell indices must agree BIT-EXACTLY between the CUDA path and
 * the CPU oracle, so both sides must produce the same float for cos/sin.  libm's sinf/cosf (glibc) and
 * CUDA's sinf/cosf are each accurate to ~1 ulp but are not the same function.  The reference itself
 * uses xsimd's polynomial kernels under -ffast-math (CMakeLists.txt:46-68), i.e. yet another ~1 ulp
 * implementation, so any faithful ~1 ulp sincos is an equally valid restatement.
 *
 * This header therefore defines ONE sincos built only from IEEE-754 correctly rounded operations
 * (add, mul, fma, round-to-nearest-even) so that gcc on x86 and nvcc on sm_100a give identical bits:
 *   - 3-term Cody-Waite reduction by pi/2 with explicit fmaf (|x| <= 1e5; k < 2^16 so k*c1 is exact
 *     inside the fma), a 2-term double-precision reduction beyond that;
 *   - Cephes single-precision minimax polynomials on [-pi/4, pi/4], Horner form with explicit fmaf.
 * Absolute error vs. the true sin/cos is < 1.2e-7 (checked against libm in tests/test_det_math.py).
 *
 * Both the product kernels (mpcholonavigation_b200/csrc) and the CPU oracle (oracle/) include this
 * file; it is the single shared definition of the canonical arithmetic, nothing else is shared.
 */
#ifndef MPPI_DET_MATH_H_
#define MPPI_DET_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPPI_HD __host__ __device__ __forceinline__
#else
#define MPPI_HD static inline
#endif

/* Non-contractable fp32 primitives: on the device the _rn intrinsics are never fused by ptxas; on the
 * host the oracle is built with -ffp-contract=off so plain operators are single IEEE operations. */
#if defined(__CUDA_ARCH__)
#define MPPI_FADD(a, b) __fadd_rn((a), (b))
#define MPPI_FSUB(a, b) __fsub_rn((a), (b))
#define MPPI_FMUL(a, b) __fmul_rn((a), (b))
#define MPPI_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#define MPPI_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define MPPI_DMUL(a, b) __dmul_rn((a), (b))
#else
#define MPPI_FADD(a, b) ((float)(a) + (float)(b))
#define MPPI_FSUB(a, b) ((float)(a) - (float)(b))
#define MPPI_FMUL(a, b) ((float)(a) * (float)(b))
#define MPPI_FFMA(a, b, c) fmaf((a), (b), (c))
#define MPPI_DFMA(a, b, c) fma((a), (b), (c))
#define MPPI_DMUL(a, b) ((double)(a) * (double)(b))
#endif

/* pi/2 = C1 + C2 + C3 (each the float nearest to the running remainder) */
#define MPPI_PIO2_C1 1.57079637e+00f      /* 0x3fc90fdb */
#define MPPI_PIO2_C2 (-4.37113883e-08f)   /* 0xb33bbd2e */
#define MPPI_PIO2_C3 (-1.71512451e-15f)   /* 0xa6f72ced */
#define MPPI_TWO_OVER_PI_F 6.36619747e-01f /* 0x3f22f983 */
#define MPPI_PIO2_D1 1.5707963267948966    /* 0x1.921fb54442d18p+0 */
#define MPPI_PIO2_D2 6.123233995736766e-17 /* 0x1.1a62633145c07p-54 */
#define MPPI_TWO_OVER_PI_D 0.6366197723675814

/* v with its sign bit flipped where bit 31 of m is set (v -> -v exactly, zeros and NaNs included) */
MPPI_HD float mppi_det_xor_sign(float v, uint32_t m)
{
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__float_as_uint(v) ^ (m & 0x80000000u));
#else
  union {float f; uint32_t u;} w;
  w.f = v;
  w.u ^= m & 0x80000000u;
  return w.f;
#endif
}

/* second half of the sincos: Cephes sinf/cosf kernels on the reduced argument |r| <= pi/4, then the quadrant */
MPPI_HD void mppi_det_sincosf_reduced(float r, int32_t q, float * s_out, float * c_out)
{
  const float z = MPPI_FMUL(r, r);
  float ps = MPPI_FFMA(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = MPPI_FFMA(ps, z, -1.6666654611e-1f);
  const float sr = MPPI_FFMA(MPPI_FMUL(ps, z), r, r);
  float pc = MPPI_FFMA(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = MPPI_FFMA(pc, z, 4.166664568298827e-2f);
  const float cr = MPPI_FFMA(MPPI_FMUL(pc, z), z, MPPI_FFMA(-0.5f, z, 1.0f));
  /* quadrant fix-up, branch-free: q&3 = 0: (sr, cr)  1: (cr, -sr)  2: (-sr, -cr)  3: (-cr, sr).
   * The swap is a select on bit 0; the signs are bit 1 of q (sine) and of q + 1 (cosine) moved onto the sign bit. */
  {
    const uint32_t uq = (uint32_t)q;
    const float s0 = (uq & 1u) ? cr : sr;
    const float c0 = (uq & 1u) ? sr : cr;
    *s_out = mppi_det_xor_sign(s0, uq << 30);
    *c_out = mppi_det_xor_sign(c0, (uq << 30) + 0x40000000u);
  }
}

/* first half for |x| <= 1e5 (the caller guarantees the range): 3-term Cody-Waite reduction by pi/2 */
MPPI_HD void mppi_det_reduce_small(float x, float * r_out, int32_t * q_out)
{
#if defined(__CUDA_ARCH__)
  /* rintf + float->int without conversion instructions: adding 1.5 * 2^23 rounds (to nearest even, like rintf) the
   * product to an integer that sits in the low mantissa bits; |x * 2/pi| < 2^16 here.  Same kf, same q. */
  const float tm = __fadd_rn(MPPI_FMUL(x, MPPI_TWO_OVER_PI_F), 12582912.0f);
  const float kf = __fsub_rn(tm, 12582912.0f);
  *q_out = __float_as_int(tm) - 0x4B400000;
#else
  const float kf = rintf(MPPI_FMUL(x, MPPI_TWO_OVER_PI_F));
  *q_out = (int32_t)kf;
#endif
  float r = MPPI_FFMA(-kf, MPPI_PIO2_C1, x);
  r = MPPI_FFMA(-kf, MPPI_PIO2_C2, r);
  *r_out = MPPI_FFMA(-kf, MPPI_PIO2_C3, r);
}

/* sin and cos of x for callers that have already checked |x| <= 1e5: the same bits as mppi_det_sincosf */
MPPI_HD void mppi_det_sincosf_small(float x, float * s_out, float * c_out)
{
  float r;
  int32_t q;
  mppi_det_reduce_small(x, &r, &q);
  mppi_det_sincosf_reduced(r, q, s_out, c_out);
}

/* sin and cos of x, both at once.  Deterministic across host and device. */
MPPI_HD void mppi_det_sincosf(float x, float * s_out, float * c_out)
{
  float r;
  int32_t q;
  if (fabsf(x) <= 1.0e5f) {
    mppi_det_reduce_small(x, &r, &q);
  } else if (fabsf(x) <= 2.0e9f) {
    const double xd = (double)x;
    const double kd = rint(MPPI_DMUL(xd, MPPI_TWO_OVER_PI_D));
    double rd = MPPI_DFMA(-kd, MPPI_PIO2_D1, xd);
    rd = MPPI_DFMA(-kd, MPPI_PIO2_D2, rd);
    r = (float)rd;
    q = (int32_t)((int64_t)kd & 3);
  } else {
    /* out of the supported range (never reached by a yaw angle); defined, not accurate */
    r = 0.0f;
    q = 0;
  }
  mppi_det_sincosf_reduced(r, q, s_out, c_out);
}

#endif  /* MPPI_DET_MATH_H_ */
