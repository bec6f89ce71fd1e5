// lattice.cuh -- map-lattice distances and the gaussian neighbourhood weight, op by op as the
// reference evaluates them (som_rout.c:434-468, 541-542).  Shared by K3 (training) and K4 (qerror2).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// Maps up to 1024 x 1024: dx^2 (a multiple of 1/4 below 2^20) and 0.75*dy^2 are exact in FP32,
// so is their sum, and (float)sqrt((double)r) == the correctly rounded sqrtf(r) (a double
// carries more than 2*24+2 bits).  The result equals the reference expression bit for bit.
__device__ __forceinline__ float hexa_dist_small(int bx, int by, int tx, int ty) {
  float dx = (float)(bx - tx);
  if (((by - ty) & 1) != 0) dx += ((by & 1) == 0) ? -0.5f : 0.5f;
  const float dy = (float)(by - ty);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(__fmul_rn(0.75f, dy), dy)));
}
__device__ __forceinline__ float rect_dist_small(int bx, int by, int tx, int ty) {
  const float dx = (float)(bx - tx), dy = (float)(by - ty);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// som_rout.c:434-455 -- the mixed float/double expression of the reference, op by op
__device__ __forceinline__ float hexa_dist_dev(int bx, int by, int tx, int ty) {
  float dx = (float)(bx - tx);
  if (((by - ty) % 2) != 0) {
    if ((by % 2) == 0) dx = (float)__dadd_rn((double)dx, -0.5);
    else dx = (float)__dadd_rn((double)dx, 0.5);
  }
  float r = __fmul_rn(dx, dx);
  float dy = (float)(by - ty);
  r = (float)__dadd_rn((double)r, __dmul_rn(__dmul_rn(0.75, (double)dy), (double)dy));
  return (float)__dsqrt_rn((double)r);
}
// som_rout.c:457-468
__device__ __forceinline__ float rect_dist_dev(int bx, int by, int tx, int ty) {
  float dx = (float)(bx - tx), dy = (float)(by - ty);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  return (float)__dsqrt_rn((double)r);
}
// som_rout.c:541-542 : alpha * (float)exp((double)(-dd*dd / (2.0*radius*radius)))
__device__ __forceinline__ float gauss_alpha_dev(float alpha, float dd, float radius) {
  float num = __fmul_rn(-dd, dd);
  double den = __dmul_rn(__dmul_rn(2.0, (double)radius), (double)radius);
  float w = (float)exp(__ddiv_rn((double)num, den));
  return __fmul_rn(alpha, w);
}

// The same value, cheaper.  The reference's factor is RN_float(exp(x)), x = num / den in double.  A double
// result with relative error eps rounds to the same float as the true exp(x) unless it lies within eps of a
// float rounding tie, i.e. unless the 29 mantissa bits that the conversion drops are within eps * 2^52 of
// 0x10000000.  So: x' = num * (1 / den) (inv_den is computed once per step, off the critical path), an
// 11-term Taylor polynomial after the usual k * ln2 reduction (|r| <= 0.347: truncation 6e-15, argument
// error <= 104 * 2^-52 = 2.3e-14, evaluation ~1e-15: eps < 2^-44, i.e. < 256 units of the dropped bits), and
// only results further than 4096 units from the tie are accepted; the rest -- and the float-denormal range
// -104 <= x <= -87, whose ties sit elsewhere -- take the exact IEEE division and the full double exp.  Below
// -104 exp(x) < 2^-150 rounds to +0.  About 16 FP64 instructions per unit instead of about 60, with results
// that are bit-identical to gauss_alpha_dev (tests/test_train_gpu.py compares the trained maps bit for bit).
__device__ __forceinline__ float gauss_alpha_fast(float alpha, float dd, float radius, double inv_den) {
  const float num = __fmul_rn(-dd, dd);
  const double x = __dmul_rn((double)num, inv_den);
  float w = 0.0f;
  bool ok = x < -104.0;
  if (!ok && x > -87.0) {
    const double kd = rint(__dmul_rn(x, 1.4426950408889634074));
    double r = __fma_rn(-kd, 6.93147180369123816490e-01, x);          // k * ln2_hi is exact (ln2_hi: 32 bits)
    r = __fma_rn(-kd, 1.90821492927058770002e-10, r);
    double q = 2.50521083854417187751e-08;                            // 1/11!
    q = __fma_rn(q, r, 2.75573192239858906526e-07);                   // 1/10!
    q = __fma_rn(q, r, 2.75573192239858906526e-06);                   // 1/9!
    q = __fma_rn(q, r, 2.48015873015873015873e-05);                   // 1/8!
    q = __fma_rn(q, r, 1.98412698412698412698e-04);                   // 1/7!
    q = __fma_rn(q, r, 1.38888888888888888889e-03);                   // 1/6!
    q = __fma_rn(q, r, 8.33333333333333333333e-03);                   // 1/5!
    q = __fma_rn(q, r, 4.16666666666666666667e-02);                   // 1/4!
    q = __fma_rn(q, r, 1.66666666666666666667e-01);                   // 1/3!
    q = __fma_rn(q, r, 0.5);
    q = __fma_rn(q, r, 1.0);
    q = __fma_rn(q, r, 1.0);                                          // exp(r), in [0.70, 1.42]
    const long long bits = __double_as_longlong(q) + ((long long)(int)kd << 52);   // * 2^k, k >= -126
    const unsigned lo = (unsigned)(bits & 0x1FFFFFFFll);
    const unsigned dist = lo > 0x10000000u ? lo - 0x10000000u : 0x10000000u - lo;
    if (dist > 4096u) {
      w = (float)__longlong_as_double(bits);
      ok = true;
    }
  }
  if (!ok) {
    const double den = __dmul_rn(__dmul_rn(2.0, (double)radius), (double)radius);
    w = (float)exp(__ddiv_rn((double)num, den));
  }
  return __fmul_rn(alpha, w);
}
