// k3_train.cu -- K3: the online training loop as ONE persistent kernel (sm_100a).
//
// Replaces the sequential loops of som_training (reference som_rout.c:556-671, with
// bubble_adapt 472-506 / gaussian_adapt 511-549) and lvq1/olvq1/lvq2/lvq3_training
// (lvq_rout.c:498-916).  Step t+1 reads the codebook written by step t, so the steps stay
// sequential; inside a step the units are spread over the CTAs of a co-resident grid:
//   * every CTA owns a fixed slice of units, kept in shared memory for the whole run
//     (component-major so that both the search and the update are conflict-free);
//   * search: each thread accumulates the reference's FP32 sum for its unit(s), the CTA
//     reduces to its best (or two best) packed keys;
//   * exchange: one relaxed 64-bit store per CTA into a tagged slot, every CTA polls all
//     slots of the step (the only grid-wide synchronisation, no atomics, no reset);
//   * update: elementwise c += a*(x-c) with separate sub, mul, add (lvq_pak.c:339-351) on
//     the owner CTA(s).
// Sample order, learning rates and radii arrive as per-step arrays computed on the host
// with the reference's own formulas (bmu_som_schedule / bmu_lvq_schedule).
#include "common.cuh"
#include "k3_train.h"

namespace bmu {

#define K3_NOKEY 0xFFFFFFFFFFFFFF00ull

__device__ __forceinline__ u64 make_key(float d, int idx, bool maxidx) {
  unsigned f = (unsigned)(maxidx ? (0xFFFFFF - idx) : idx) & 0xFFFFFFu;
  return ((u64)__float_as_uint(d) << 32) | ((u64)f << 8);
}
__device__ __forceinline__ float key_diff(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_idx(u64 k, bool maxidx) {
  int f = (int)((k >> 8) & 0xFFFFFFu);
  return maxidx ? (0xFFFFFF - f) : f;
}
// merge two sorted pairs (a1<=a2), (b1<=b2) into the two smallest
__device__ __forceinline__ void merge2(u64 &a1, u64 &a2, u64 b1, u64 b2) {
  u64 lo = a1 < b1 ? a1 : b1;
  u64 hi = a1 < b1 ? b1 : a1;
  u64 m = a2 < b2 ? a2 : b2;
  a1 = lo;
  a2 = hi < m ? hi : m;
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
// lvq_pak.c:339-351 on one component
__device__ __forceinline__ float adapt1(float c, float x, float a) {
  return __fadd_rn(c, __fmul_rn(a, __fsub_rn(x, c)));
}

__device__ __forceinline__ void stage_x(float *dst, const float *src, int D, int tid) {
  // cp.async (LDGSTS): 16-byte pieces when the row is 16-byte aligned, else 4-byte
  if ((D & 3) == 0) {
    for (int i = tid * 4; i < D; i += K3_THREADS * 4)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i)), "l"(src + i));
  } else {
    for (int i = tid; i < D; i += K3_THREADS)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + i)), "l"(src + i));
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <bool HAS_MASK, bool TOP2>
__global__ void __launch_bounds__(K3_THREADS, 1) k3_kernel(const K3Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = p.D, U = p.U, Us = p.Us;
  const int Dpad = (D + 3) & ~3;
  float *xs = reinterpret_cast<float *>(smem_raw);                // [2][Dpad]
  u64 *wred = reinterpret_cast<u64 *>(xs + 2 * Dpad);            // [2][16] per-warp keys
  u64 *gw = wred + 32;                                            // [2] global winners
  float *ua_s = reinterpret_cast<float *>(gw + 2);                // [U] OLVQ1 rates
  float *sl = p.slice_in_smem ? ua_s + ((U + 3) & ~3)
                              : p.gslice + (long)blockIdx.x * D * Us;   // [D][Us]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const long u0 = (long)blockIdx.x * U;
  const int ucount = (int)max(0L, min((long)U, p.M - u0));
  const int mode = p.mode;
  const bool is_som = mode <= K3_SOM_GAUSSIAN;

  // ---- load the slice, transposed to component-major
  for (long t = tid; t < (long)ucount * D; t += K3_THREADS) {
    int u = (int)(t / D), i = (int)(t % D);
    sl[(long)i * Us + u] = p.codes[(u0 + u) * D + i];
  }
  if (mode == K3_OLVQ1)
    for (int u = tid; u < ucount; u += K3_THREADS) ua_s[u] = p.unit_alpha[u0 + u];

  int cur = 0;
  unsigned bstep = 0;                       // number of grid exchanges done so far
  long s_cur = p.nsteps > 0 ? p.sample[0] : 0;
  long s_nxt = p.nsteps > 1 ? p.sample[1] : 0;
  if (p.nsteps > 0) stage_x(xs, p.data + s_cur * D, D, tid);

  for (long t = 0; t < p.nsteps; t++) {
    const float talp = p.talp ? p.talp[t] : 0.0f;
    const float trad = p.trad ? p.trad[t] : 0.0f;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                        // xs[cur] landed; previous update finished
    const float *x = xs + cur * Dpad;
    if (t + 1 < p.nsteps) stage_x(xs + (cur ^ 1) * Dpad, p.data + s_nxt * D, D, tid);
    const long s_this = s_cur;
    s_cur = s_nxt;
    s_nxt = (t + 2 < p.nsteps) ? p.sample[t + 2] : 0;
    cur ^= 1;

    if (HAS_MASK && p.valid[s_this] == 0) continue;    // som_rout.c:635-640: empty sample

    int bx = 0, by = 0;
    bool have_fixed = false;
    if (is_som && p.fixed_xy) {                          // som_rout.c:628-632
      short fx = p.fixed_xy[2 * s_this];
      if (fx >= 0) { have_fixed = true; bx = fx; by = p.fixed_xy[2 * s_this + 1]; }
    }

    u64 g1 = K3_NOKEY, g2 = K3_NOKEY;
    if (!have_fixed) {
      // ---- search: the reference's sum, component order, one rounding per operation
      u64 k1 = K3_NOKEY, k2 = K3_NOKEY;
      for (int u = tid; u < ucount; u += K3_THREADS) {
        const float *col = sl + u;
        float acc = 0.0f;
        if (HAS_MASK) {
          for (int i = 0; i < D; i++) {
            float xi = x[i];
            if (xi != xi) continue;
            acc = sq_acc(acc, col[(long)i * Us], xi);
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < D; i++) acc = sq_acc(acc, col[(long)i * Us], x[i]);
        }
        // k == 1: only d < FLT_MAX can win (lvq_pak.c:57,79); k == 2: d <= FLT_MAX is inserted
        const bool cand = TOP2 ? (acc <= FLT_MAX) : (acc < FLT_MAX);
        if (cand) {
          u64 key = make_key(acc, (int)(u0 + u), TOP2);
          if (key < k1) { k2 = k1; k1 = key; }
          else if (TOP2 && key < k2) k2 = key;
        }
      }
      // warp reduce, then across warps
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        u64 o1 = __shfl_xor_sync(0xffffffffu, k1, off);
        if (TOP2) {
          u64 o2 = __shfl_xor_sync(0xffffffffu, k2, off);
          merge2(k1, k2, o1, o2);
        } else {
          k1 = o1 < k1 ? o1 : k1;
        }
      }
      if (lane == 0) { wred[warp] = k1; if (TOP2) wred[16 + warp] = k2; }
      __syncthreads();
      if (warp == 0) {
        u64 b1 = lane < K3_THREADS / 32 ? wred[lane] : K3_NOKEY;
        u64 b2 = (TOP2 && lane < K3_THREADS / 32) ? wred[16 + lane] : K3_NOKEY;
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
          u64 o1 = __shfl_xor_sync(0xffffffffu, b1, off);
          if (TOP2) {
            u64 o2 = __shfl_xor_sync(0xffffffffu, b2, off);
            merge2(b1, b2, o1, o2);
          } else {
            b1 = o1 < b1 ? o1 : b1;
          }
        }
        if (G > 1) {
          // ---- grid exchange: tagged slot per CTA, double buffered by exchange parity
          const u64 tag = (u64)((bstep + 1) & 0xFFu);
          u64 *slot = p.slots + ((size_t)(bstep & 1) * G) * 2;
          if (lane == 0) {
            st_relaxed_u64(slot + 2 * blockIdx.x, b1 | tag);
            if (TOP2) st_relaxed_u64(slot + 2 * blockIdx.x + 1, b2 | tag);
          }
          u64 m1 = K3_NOKEY, m2 = K3_NOKEY;
          for (int c = lane; c < G; c += 32) {
            u64 v1, v2 = K3_NOKEY;
            do { v1 = ld_relaxed_u64(slot + 2 * c); } while ((v1 & 0xFFu) != tag);
            v1 &= ~0xFFull;
            if (TOP2) {
              do { v2 = ld_relaxed_u64(slot + 2 * c + 1); } while ((v2 & 0xFFu) != tag);
              v2 &= ~0xFFull;
              merge2(m1, m2, v1, v2);
            } else {
              m1 = v1 < m1 ? v1 : m1;
            }
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            u64 o1 = __shfl_xor_sync(0xffffffffu, m1, off);
            if (TOP2) {
              u64 o2 = __shfl_xor_sync(0xffffffffu, m2, off);
              merge2(m1, m2, o1, o2);
            } else {
              m1 = o1 < m1 ? o1 : m1;
            }
          }
          b1 = m1; b2 = m2;
        }
        if (lane == 0) { gw[0] = b1; gw[1] = b2; }
      }
      bstep++;
      __syncthreads();
      g1 = gw[0]; g2 = gw[1];
    }

    // ---- update
    if (is_som) {
      if (!have_fixed) {
        // no winner (every distance NaN/Inf): the reference would adapt around index -1,
        // i.e. garbage in, garbage out; we skip the step instead.
        if (g1 == K3_NOKEY) continue;
        int w = key_idx(g1, false);
        bx = w % p.xdim; by = w / p.xdim;                  // som_rout.c:641-642
      }
      for (int u = tid; u < ucount; u += K3_THREADS) {
        const int gidx = (int)(u0 + u);
        const int tx = gidx % p.xdim, ty = gidx / p.xdim;  // som_rout.c:493-494
        const float dd = p.topol == 4 ? rect_dist_dev(bx, by, tx, ty) : hexa_dist_dev(bx, by, tx, ty);
        float a;
        if (mode == K3_SOM_GAUSSIAN) a = gauss_alpha_dev(talp, dd, trad);
        else { if (!(dd <= trad)) continue; a = talp; }    // som_rout.c:496
        float *col = sl + u;
        if (HAS_MASK) {
          for (int i = 0; i < D; i++) {
            float xi = x[i];
            if (xi != xi) continue;
            col[(long)i * Us] = adapt1(col[(long)i * Us], xi, a);
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < D; i++) col[(long)i * Us] = adapt1(col[(long)i * Us], x[i], a);
        }
      }
    } else {
      if (g1 == K3_NOKEY) continue;
      const int dl = p.data_label[s_this];
      int w1 = key_idx(g1, TOP2), w2 = -1;
      float a1 = 0.0f, a2 = 0.0f;
      bool do1 = false, do2 = false;
      if (mode == K3_LVQ1) {                               // lvq_rout.c:552-555
        a1 = (p.code_label[w1] == dl) ? talp : -talp;
        do1 = true;
      } else if (mode == K3_OLVQ1) {                       // lvq_rout.c:657-673
        const bool own = w1 >= u0 && w1 < u0 + ucount;
        if (own) {
          const float ta = ua_s[w1 - u0];
          const bool correct = p.code_label[w1] == dl;
          a1 = correct ? ta : -ta;
          do1 = true;
          __syncthreads();                                 // everyone has read ua_s
          if (tid == 0) {
            float nt;
            if (correct) nt = __fdiv_rn(ta, __fadd_rn(1.0f, ta));
            else {
              nt = __fdiv_rn(ta, __fsub_rn(1.0f, ta));
              if (nt > p.alpha_cap) nt = p.alpha_cap;
            }
            ua_s[w1 - u0] = nt;
          }
        }
      } else {                                             // lvq_rout.c:765-781, 870-896
        if (g2 == K3_NOKEY) continue;
        w2 = key_idx(g2, true);
        const int l1 = p.code_label[w1], l2 = p.code_label[w2];
        if (l1 != l2) {
          if ((l1 == dl || l2 == dl) && __fdiv_rn(key_diff(g1), key_diff(g2)) > p.win_thr) {
            a1 = (l2 == dl) ? -talp : talp;                // the correct one moves towards x
            a2 = (l2 == dl) ? talp : -talp;
            do1 = do2 = true;
          }
        } else if (mode == K3_LVQ3 && l1 == dl) {
          a1 = a2 = __fmul_rn(talp, p.epsilon);
          do1 = do2 = true;
        }
      }
      // when both units move, the reference adapts the "best" (correct) one first; the two
      // updates touch different rows, so the order is immaterial
      if (do1 && w1 >= u0 && w1 < u0 + ucount) {
        float *col = sl + (w1 - u0);
        for (int i = tid; i < D; i += K3_THREADS) {
          float xi = x[i];
          if (HAS_MASK && xi != xi) continue;
          col[(long)i * Us] = adapt1(col[(long)i * Us], xi, a1);
        }
      }
      if (do2 && w2 >= u0 && w2 < u0 + ucount) {
        float *col = sl + (w2 - u0);
        for (int i = tid; i < D; i += K3_THREADS) {
          float xi = x[i];
          if (HAS_MASK && xi != xi) continue;
          col[(long)i * Us] = adapt1(col[(long)i * Us], xi, a2);
        }
      }
    }
  }

  // ---- write the slice back
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (long t = tid; t < (long)ucount * D; t += K3_THREADS) {
    int u = (int)(t / D), i = (int)(t % D);
    p.codes[(u0 + u) * D + i] = sl[(long)i * Us + u];
  }
  if (mode == K3_OLVQ1)
    for (int u = tid; u < ucount; u += K3_THREADS) p.unit_alpha[u0 + u] = ua_s[u];
}

// ---------------------------------------------------------------- mask encoding
__global__ void encode_mask_kernel(float *__restrict__ data, const unsigned char *__restrict__ mask,
                                   unsigned char *__restrict__ valid, long N, int D) {
  const int lane = threadIdx.x & 31;
  const long w0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nw = ((long)gridDim.x * blockDim.x) >> 5;
  for (long n = w0; n < N; n += nw) {
    int unmasked = 0;
    for (int i = lane; i < D; i += 32) {
      if (mask[n * D + i]) data[n * D + i] = __uint_as_float(K3_MASK_SENTINEL);
      else unmasked++;
    }
    unmasked = __reduce_add_sync(0xffffffffu, unmasked);
    if (lane == 0) valid[n] = unmasked > 0;
  }
}

cudaError_t k3_encode_mask(float *d_data, const unsigned char *d_mask, unsigned char *d_valid,
                           long N, int D, cudaStream_t st) {
  encode_mask_kernel<<<148 * 4, 256, 0, st>>>(d_data, d_mask, d_valid, N, D);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- planning + launch
static size_t k3_fixed_smem(int D, int U) {
  int Dpad = (D + 3) & ~3;
  return (size_t)2 * Dpad * 4 + 34 * 8 + (size_t)((U + 3) & ~3) * 4;
}

K3Plan k3_plan(long M, int D, int num_sms, size_t smem_optin) {
  K3Plan best{};
  double best_cost = 1e300;
  for (int G = 1; G <= num_sms; G = (G < num_sms && G * 2 > num_sms) ? num_sms : G * 2) {
    int U = (int)((M + G - 1) / G);
    int Us = (U & 1) ? U : U + 1;
    size_t fixed = k3_fixed_smem(D, U);
    size_t slice = (size_t)D * Us * 4;
    bool fits = fixed + slice <= smem_optin;
    // rough cycles per step: rounds of units per thread x (search + update) + exchange
    double rounds = (double)((U + K3_THREADS - 1) / K3_THREADS);
    double cost = rounds * D * 24.0 * (fits ? 1.0 : 6.0) + (G > 1 ? 1500.0 + 4.0 * G : 0.0);
    if (cost < best_cost) {
      best_cost = cost;
      best.grid = G; best.U = U; best.Us = Us; best.slice_in_smem = fits ? 1 : 0;
      best.smem_bytes = fixed + (fits ? slice : 0);
      best.gslice_floats = fits ? 0 : (size_t)G * D * Us;
    }
    if (G == num_sms) break;
  }
  return best;
}

cudaError_t k3_launch(const K3Params &p, const K3Plan &plan, bool has_mask, cudaStream_t st) {
  const bool top2 = p.mode == K3_LVQ2 || p.mode == K3_LVQ3;
  void *fn;
  if (has_mask) fn = top2 ? (void *)k3_kernel<true, true> : (void *)k3_kernel<true, false>;
  else fn = top2 ? (void *)k3_kernel<false, true> : (void *)k3_kernel<false, false>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)plan.smem_bytes);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(p.slots, 0, sizeof(u64) * 4 * plan.grid, st);
  if (e != cudaSuccess) return e;
  K3Params pp = p;
  void *args[] = {&pp};
  // cooperative launch only for its guarantee that all CTAs are co-resident (the slot
  // exchange spins on other CTAs)
  return cudaLaunchCooperativeKernel(fn, dim3(plan.grid), dim3(K3_THREADS), args, plan.smem_bytes, st);
}

}  // namespace bmu
