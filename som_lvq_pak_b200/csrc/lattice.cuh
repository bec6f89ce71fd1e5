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
