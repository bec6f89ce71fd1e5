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
#include "lattice.cuh"
#include "k3_train.h"

namespace bmu {

#define K3_NOKEY 0xFFFFFFFFFFFFFF00ull
#define K3_SLOT_STRIDE 16          // u64 per CTA slot: one 128-byte line each, spread over L2 slices
#define K3_POLL_WARPS 5           // generic kernel: the same split of the slots over five of its eight warps
#define K3F_POLL_WARPS 5          // fused kernel: 5 x 32 lanes >= K3_MAX_GRID slots, one per lane
#define K3F_SLOT_STRIDE 8          // fused kernel: two slots per line (measured 2-3 % of the C5 step over 16, 4 and 1)

__device__ __forceinline__ u64 make_key(float d, int idx, bool maxidx) {
  unsigned f = (unsigned)(maxidx ? (0xFFFFFF - idx) : idx) & 0xFFFFFFu;
  return ((u64)__float_as_uint(d) << 32) | ((u64)f << 8);
}
__device__ __forceinline__ float key_diff(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_idx(u64 k, bool maxidx) {
  int f = (int)((k >> 8) & 0xFFFFFFu);
  return maxidx ? (0xFFFFFF - f) : f;
}
// warp-wide minimum of a 64-bit key with two REDUX.MIN.U32 (hardware warp reduction) instead
// of five rounds of two dependent shuffles
__device__ __forceinline__ u64 warp_min_u64(u64 k) {
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return ((u64)mh << 32) | ml;
}

// merge two sorted pairs (a1<=a2), (b1<=b2) into the two smallest
__device__ __forceinline__ void merge2(u64 &a1, u64 &a2, u64 b1, u64 b2) {
  u64 lo = a1 < b1 ? a1 : b1;
  u64 hi = a1 < b1 ? b1 : a1;
  u64 m = a2 < b2 ? a2 : b2;
  a1 = lo;
  a2 = hi < m ? hi : m;
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

// Exact packed helpers for K3.  The codebook can decay towards zero during training, so the
// .ftz trick of K1 is not admissible here: sub and mul are packed (FADD2, FMUL2), the
// accumulating add stays scalar (ptxas does not fuse FMUL2 with a scalar FADD; verified in SASS).
__device__ __forceinline__ void sq_acc2(float &acc_a, float &acc_b, u64 c2, u64 x2) {
  const u64 d = sub2(c2, x2);
  float lo, hi;
  unpack2(mul2(d, d), lo, hi);
  acc_a = __fadd_rn(acc_a, lo);
  acc_b = __fadd_rn(acc_b, hi);
}
// c + a*(x-c) for a pair of units with per-unit rates
__device__ __forceinline__ u64 adapt2(u64 c2, u64 x2, u64 a2) {
  float cl, ch, pl, ph;
  unpack2(c2, cl, ch);
  unpack2(mul2(a2, sub2(x2, c2)), pl, ph);
  return pack2(__fadd_rn(cl, pl), __fadd_rn(ch, ph));
}

template <bool HAS_MASK, bool TOP2, bool SLICE_SMEM>
__global__ void __launch_bounds__(K3_THREADS, 1) k3_kernel(const K3Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = p.D, U = p.U, Us = p.Us;                        // Us is even: unit pairs are 8-byte aligned
  const int Dpad = (D + 3) & ~3;
  float *xs = reinterpret_cast<float *>(smem_raw);                // [2][Dpad]
  u64 *wred = reinterpret_cast<u64 *>(xs + 2 * Dpad);            // [2][16] per-warp keys
  u64 *gw = wred + 32;                                            // [K3_POLL_WARPS][2] partial winners of the exchange
  float *ua_s = reinterpret_cast<float *>(gw + 2 * K3_POLL_WARPS + 2);   // [U] OLVQ1 rates
  float *sl;                                                      // [D][Us] component-major slice
  if (SLICE_SMEM) sl = ua_s + ((U + 3) & ~3);
  else sl = p.gslice + (size_t)blockIdx.x * D * Us;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int u0 = blockIdx.x * U;
  const int ucount = max(0, min(U, (int)p.M - u0));
  const int npair = (ucount + 1) >> 1;
  const int mode = p.mode;
  const bool is_som = mode <= K3_SOM_GAUSSIAN;
  const bool small_map = p.xdim <= 1024 && p.ydim <= 1024 && p.xdim > 0;

  // ---- load the slice, transposed to component-major (pad column zeroed)
  for (int t = tid; t < D * Us; t += K3_THREADS) sl[t] = 0.0f;
  __syncthreads();
  for (int t = tid; t < ucount * D; t += K3_THREADS) {
    int u = t / D, i = t - u * D;
    sl[i * Us + u] = p.codes[(size_t)(u0 + u) * D + i];
  }
  if (mode == K3_OLVQ1)
    for (int u = tid; u < ucount; u += K3_THREADS) ua_s[u] = p.unit_alpha[u0 + u];

  // lattice coordinates of this thread's first unit pair never change (som_rout.c:493-494)
  int tx0[2] = {0, 0}, ty0[2] = {0, 0};
  if (is_som && p.xdim > 0) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int gidx = u0 + 2 * tid + h;
      tx0[h] = gidx % p.xdim; ty0[h] = gidx / p.xdim;
    }
  }

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
      // ---- search: the reference's sum, component order, one rounding per operation;
      //      one thread = two adjacent units (packed sub/mul, scalar accumulate)
      u64 k1 = K3_NOKEY, k2 = K3_NOKEY;
      for (int up = tid; up < npair; up += K3_THREADS) {
        const float *col = sl + 2 * up;
        float acc_a = 0.0f, acc_b = 0.0f;
        if (HAS_MASK) {
          for (int i = 0; i < D; i++) {
            const float xi = x[i];
            if (xi != xi) continue;
            sq_acc2(acc_a, acc_b, *reinterpret_cast<const u64 *>(col + i * Us), pack2(xi, xi));
          }
        } else {
          int i = 0;
          for (; i + 8 <= D; i += 8) {
            // all loads of the batch first, so that their latency overlaps the dependent adds
            u64 c[8];
#pragma unroll
            for (int q = 0; q < 8; q++) c[q] = *reinterpret_cast<const u64 *>(col + (i + q) * Us);
            const float4 xa = *reinterpret_cast<const float4 *>(x + i);
            const float4 xb = *reinterpret_cast<const float4 *>(x + i + 4);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int q = 0; q < 8; q++) sq_acc2(acc_a, acc_b, c[q], pack2(xv[q], xv[q]));
          }
          for (; i < D; i++)
            sq_acc2(acc_a, acc_b, *reinterpret_cast<const u64 *>(col + i * Us), pack2(x[i], x[i]));
        }
        // k == 1: only d < FLT_MAX can win (lvq_pak.c:57,79); k == 2: d <= FLT_MAX is inserted
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float acc = h ? acc_b : acc_a;
          const int u = 2 * up + h;
          const bool cand = u < ucount && (TOP2 ? (acc <= FLT_MAX) : (acc < FLT_MAX));
          if (cand) {
            u64 key = make_key(acc, u0 + u, TOP2);
            if (key < k1) { k2 = k1; k1 = key; }
            else if (TOP2 && key < k2) k2 = key;
          }
        }
      }
      // warp reduce, then across warps
      if (TOP2) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          u64 o1 = __shfl_xor_sync(0xffffffffu, k1, off);
          u64 o2 = __shfl_xor_sync(0xffffffffu, k2, off);
          merge2(k1, k2, o1, o2);
        }
      } else {
        k1 = warp_min_u64(k1);
      }
      if (lane == 0) { wred[warp] = k1; if (TOP2) wred[16 + warp] = k2; }
      __syncthreads();
      if (warp < K3_POLL_WARPS) {
        // CTA minimum (warp 0), then the grid exchange: tagged slot per CTA, double buffered by exchange parity.
        // Split polling as in the fused kernel: warp w watches the slots of CTAs 32 w .. 32 w + 31, one strong
        // load (two for the two keys of LVQ2/3) per lane and wave; the partial results meet in gw[] behind the barrier.
        u64 b1 = K3_NOKEY, b2 = K3_NOKEY;
        if (warp == 0) {
          b1 = lane < K3_THREADS / 32 ? wred[lane] : K3_NOKEY;
          b2 = (TOP2 && lane < K3_THREADS / 32) ? wred[16 + lane] : K3_NOKEY;
          if (TOP2) {
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {
              u64 o1 = __shfl_xor_sync(0xffffffffu, b1, off);
              u64 o2 = __shfl_xor_sync(0xffffffffu, b2, off);
              merge2(b1, b2, o1, o2);
            }
          } else {
            b1 = warp_min_u64(b1);
          }
        }
        if (G > 1) {
          const u64 tag = (u64)((bstep + 1) & 0xFFu);
          u64 *slot = p.slots + ((size_t)(bstep & 1) * G) * K3_SLOT_STRIDE;
          if (tid == 0) {
            st_relaxed_u64(slot + K3_SLOT_STRIDE * blockIdx.x, b1 | tag);
            if (TOP2) st_relaxed_u64(slot + K3_SLOT_STRIDE * blockIdx.x + 1, b2 | tag);
          }
          if (p.poll_delay_ns > 0) __nanosleep((unsigned)p.poll_delay_ns);          // see the fused kernel
          else if (p.poll_delay_ns < 0) { const long long w0 = clock64(); while (clock64() - w0 < -p.poll_delay_ns) { } }
          const int c = warp * 32 + lane;
          bool pend = c < G;
          u64 v1 = K3_NOKEY, v2 = K3_NOKEY;
          while (__any_sync(0xffffffffu, pend)) {
            if (pend) {
              v1 = ld_relaxed_u64(slot + K3_SLOT_STRIDE * c);
              if (TOP2) v2 = ld_relaxed_u64(slot + K3_SLOT_STRIDE * c + 1);
              if ((v1 & 0xFFu) == tag && (!TOP2 || (v2 & 0xFFu) == tag)) pend = false;
            }
          }
          b1 = c < G ? (v1 & ~0xFFull) : K3_NOKEY;
          b2 = (TOP2 && c < G) ? (v2 & ~0xFFull) : K3_NOKEY;
          if (TOP2) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              u64 o1 = __shfl_xor_sync(0xffffffffu, b1, off);
              u64 o2 = __shfl_xor_sync(0xffffffffu, b2, off);
              merge2(b1, b2, o1, o2);
            }
          } else {
            b1 = warp_min_u64(b1);
          }
        }
        if (lane == 0) { gw[2 * warp] = b1; gw[2 * warp + 1] = b2; }    // G == 1: warp 0's pair, the others K3_NOKEY
      }
      bstep++;
      __syncthreads();
      g1 = gw[0]; g2 = gw[1];
#pragma unroll
      for (int w = 1; w < K3_POLL_WARPS; w++) {
        if (TOP2) merge2(g1, g2, gw[2 * w], gw[2 * w + 1]);
        else { const u64 o = gw[2 * w]; g1 = o < g1 ? o : g1; }
      }
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
      for (int up = tid; up < npair; up += K3_THREADS) {
        float a[2];
        bool upd[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int gidx = u0 + 2 * up + h;
          int tx, ty;
          if (up == tid) { tx = tx0[h]; ty = ty0[h]; }
          else { tx = gidx % p.xdim; ty = gidx / p.xdim; }   // som_rout.c:493-494
          float dd;
          if (small_map) dd = p.topol == 4 ? rect_dist_small(bx, by, tx, ty) : hexa_dist_small(bx, by, tx, ty);
          else dd = p.topol == 4 ? rect_dist_dev(bx, by, tx, ty) : hexa_dist_dev(bx, by, tx, ty);
          upd[h] = 2 * up + h < ucount;
          if (mode == K3_SOM_GAUSSIAN) a[h] = gauss_alpha_dev(talp, dd, trad);
          else { a[h] = talp; upd[h] = upd[h] && (dd <= trad); }   // som_rout.c:496
        }
        if (!upd[0] && !upd[1]) continue;
        const u64 a2 = pack2(a[0], a[1]);
        float *col = sl + 2 * up;
        const bool both = upd[0] && upd[1];
        auto fix = [&](u64 oldv, u64 newv) {                 // leave the untouched unit bit-identical
          if (both) return newv;
          float ol, oh, nl, nh;
          unpack2(oldv, ol, oh);
          unpack2(newv, nl, nh);
          return pack2(upd[0] ? nl : ol, upd[1] ? nh : oh);
        };
        int i = 0;
        if (!HAS_MASK) {
          for (; i + 8 <= D; i += 8) {
            u64 c[8];
#pragma unroll
            for (int q = 0; q < 8; q++) c[q] = *reinterpret_cast<const u64 *>(col + (i + q) * Us);
            const float4 xa = *reinterpret_cast<const float4 *>(x + i);
            const float4 xb = *reinterpret_cast<const float4 *>(x + i + 4);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int q = 0; q < 8; q++)
              *reinterpret_cast<u64 *>(col + (i + q) * Us) = fix(c[q], adapt2(c[q], pack2(xv[q], xv[q]), a2));
          }
        }
        for (; i < D; i++) {
          const float xi = x[i];
          if (HAS_MASK && xi != xi) continue;
          u64 *cp = reinterpret_cast<u64 *>(col + i * Us);
          const u64 c2 = *cp;
          *cp = fix(c2, adapt2(c2, pack2(xi, xi), a2));
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
          col[i * Us] = adapt1(col[i * Us], xi, a1);
        }
      }
      if (do2 && w2 >= u0 && w2 < u0 + ucount) {
        float *col = sl + (w2 - u0);
        for (int i = tid; i < D; i += K3_THREADS) {
          float xi = x[i];
          if (HAS_MASK && xi != xi) continue;
          col[i * Us] = adapt1(col[i * Us], xi, a2);
        }
      }
    }
  }

  // ---- write the slice back
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int t = tid; t < ucount * D; t += K3_THREADS) {
    int u = t / D, i = t - u * D;
    p.codes[(size_t)(u0 + u) * D + i] = sl[i * Us + u];
  }
  if (mode == K3_OLVQ1)
    for (int u = tid; u < ucount; u += K3_THREADS) p.unit_alpha[u0 + u] = ua_s[u];
}

// ---------------------------------------------------------------- fused SOM kernel
// Specialisation for the large-map case (no masks, no fixed points, D >= 64, <= 512 units per
// CTA): the generic kernel above makes three passes over its shared-memory slice per step
// (search: read; update: read + write) and is bound by shared-memory bandwidth -- measured on
// C5 (443 units x 128 dims per CTA): search 3350 + update 6139 of 12200 cycles per step.  Here
//   * one thread owns one unit; its first K3F_DR components live in REGISTERS for the whole run,
//     the rest in shared memory as component pairs [(D-DR)/2][512] (conflict-free 8-byte lanes);
//   * the update of step t and the winner search of step t+1 are ONE pass: a component pair is
//     adapted towards x_t and, while still in registers, its squared difference to x_{t+1} is
//     accumulated (same operations and order as adapt_vector / find_winner_euc, lvq_pak.c:339-351,
//     63-73: the search of step t+1 sees exactly the codebook the update of step t left).
// The grid exchange is the generic kernel's (tagged slot per CTA, every CTA polls all slots).
constexpr int K3F_THREADS = 512;
constexpr int K3F_DR = 64;

// Four component pairs of one unit: optional update towards xt, then the squared differences to
// xn.  Written stage by stage over the four pairs so that the independent operations of different
// pairs are adjacent in the instruction stream (a warp issues in order; one pair at a time left
// the pass latency bound at ~80 cycles per pair).  Only the final accumulation is a chain.
__device__ __forceinline__ void k3f_quad(u64 (&c)[4], const float4 &ta, const float4 &tb, const float4 &na,
                                         const float4 &nb, bool upd, float a, float &acc) {
  if (upd) {                                            // c + a*(x - c), one rounding per operation
    const u64 a2 = pack2(a, a);
    const u64 xt[4] = {pack2(ta.x, ta.y), pack2(ta.z, ta.w), pack2(tb.x, tb.y), pack2(tb.z, tb.w)};
    u64 d[4];
#pragma unroll
    for (int k = 0; k < 4; k++) d[k] = sub2(xt[k], c[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) d[k] = mul2(a2, d[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float c0, c1, p0, p1;
      unpack2(c[k], c0, c1);
      unpack2(d[k], p0, p1);
      c[k] = pack2(__fadd_rn(c0, p0), __fadd_rn(c1, p1));
    }
  }
  const u64 xn[4] = {pack2(na.x, na.y), pack2(na.z, na.w), pack2(nb.x, nb.y), pack2(nb.z, nb.w)};
  u64 e[4];
#pragma unroll
  for (int k = 0; k < 4; k++) e[k] = sub2(c[k], xn[k]);  // code - sample (lvq_pak.c:70)
#pragma unroll
  for (int k = 0; k < 4; k++) e[k] = mul2(e[k], e[k]);
#pragma unroll
  for (int k = 0; k < 4; k++) {                          // component order
    float s0, s1;
    unpack2(e[k], s0, s1);
    acc = __fadd_rn(__fadd_rn(acc, s0), s1);
  }
}

__global__ void __launch_bounds__(K3F_THREADS, 1) k3_som_fused_kernel(const K3Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = p.D, U = p.U;
  const int Dp = (D + 3) & ~3;                                     // x rows padded with zeros
  const int nsp = (Dp - K3F_DR) / 2;                               // component pairs kept in shared memory
  float *xs = reinterpret_cast<float *>(smem_raw);                 // [3][Dp]
  u64 *wred = reinterpret_cast<u64 *>(xs + 3 * Dp);                // [16] per-warp keys
  u64 *gw = wred + 16;                                             // [K3F_POLL_WARPS] partial minima of the exchange (+pad)
  u64 *sl2 = gw + 8;                                               // [nsp][512] component pairs

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int u0 = blockIdx.x * U;
  const int ucount = max(0, min(U, (int)p.M - u0));
  const bool active = tid < ucount;
  const bool warp_live = (tid & ~31) < ucount;                    // this warp owns at least one unit
  const int gidx = u0 + tid;
  const bool gaussian = p.mode == K3_SOM_GAUSSIAN;
  const bool small_map = p.xdim <= 1024 && p.ydim <= 1024 && p.xdim > 0;
  const int tx = p.xdim > 0 ? gidx % p.xdim : 0, ty = p.xdim > 0 ? gidx / p.xdim : 0;   // som_rout.c:493-494

  // ---- load this thread's unit (components beyond D are zero: they add +0 to every sum)
  u64 cr[K3F_DR / 2];
  {
    const float *src = p.codes + (size_t)gidx * D;
#pragma unroll
    for (int j = 0; j < K3F_DR / 2; j++)
      cr[j] = active ? pack2(src[2 * j], src[2 * j + 1]) : pack2(0.0f, 0.0f);
    for (int j = 0; j < nsp; j++) {
      const int i = K3F_DR + 2 * j;
      const float v0 = (active && i < D) ? src[i] : 0.0f, v1 = (active && i + 1 < D) ? src[i + 1] : 0.0f;
      sl2[(size_t)j * K3F_THREADS + tid] = pack2(v0, v1);
    }
  }
  for (int i = tid; i < 3 * Dp; i += K3F_THREADS) xs[i] = 0.0f;    // pads stay zero: cp.async writes only D floats
  __syncthreads();

  auto stage = [&](int buf, long srow) {
    const float *src = p.data + srow * D;
    float *dst = xs + buf * Dp;
    if ((D & 3) == 0) {
      for (int i = tid * 4; i < D; i += K3F_THREADS * 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i)), "l"(src + i));
    } else {
      for (int i = tid; i < D; i += K3F_THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + i)), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // one pass over the unit: optional update towards xt, then the squared distance to xn
  auto pass = [&](const float *xt, const float *xn, bool upd, float a) -> float {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < K3F_DR / 2; j += 4) {
      const float4 ta = *reinterpret_cast<const float4 *>(xt + 2 * j), tb = *reinterpret_cast<const float4 *>(xt + 2 * j + 4);
      const float4 na = *reinterpret_cast<const float4 *>(xn + 2 * j), nb = *reinterpret_cast<const float4 *>(xn + 2 * j + 4);
      u64 c[4] = {cr[j], cr[j + 1], cr[j + 2], cr[j + 3]};
      k3f_quad(c, ta, tb, na, nb, upd, a, acc);
      cr[j] = c[0]; cr[j + 1] = c[1]; cr[j + 2] = c[2]; cr[j + 3] = c[3];
    }
    for (int j = 0; j + 4 <= nsp; j += 4) {
      u64 *q = sl2 + (size_t)j * K3F_THREADS + tid;
      u64 c[4] = {q[0], q[K3F_THREADS], q[2 * K3F_THREADS], q[3 * K3F_THREADS]};
      const float *xtj = xt + K3F_DR + 2 * j, *xnj = xn + K3F_DR + 2 * j;
      const float4 ta = *reinterpret_cast<const float4 *>(xtj), tb = *reinterpret_cast<const float4 *>(xtj + 4);
      const float4 na = *reinterpret_cast<const float4 *>(xnj), nb = *reinterpret_cast<const float4 *>(xnj + 4);
      k3f_quad(c, ta, tb, na, nb, upd, a, acc);
      if (upd) { q[0] = c[0]; q[K3F_THREADS] = c[1]; q[2 * K3F_THREADS] = c[2]; q[3 * K3F_THREADS] = c[3]; }
    }
    if (nsp & 2) {                                   // Dp - DR is a multiple of 4, so nsp is even: one pair of pairs left
      const int j = nsp - 2;
      u64 *q = sl2 + (size_t)j * K3F_THREADS + tid;
      u64 c[4] = {q[0], q[K3F_THREADS], pack2(0.0f, 0.0f), pack2(0.0f, 0.0f)};
      const float *xtj = xt + K3F_DR + 2 * j, *xnj = xn + K3F_DR + 2 * j;
      const float4 ta = *reinterpret_cast<const float4 *>(xtj), na = *reinterpret_cast<const float4 *>(xnj);
      const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      k3f_quad(c, ta, z, na, z, upd, a, acc);          // zero pairs add +0 to the sum
      if (upd) { q[0] = c[0]; q[K3F_THREADS] = c[1]; }
    }
    return acc;
  };

  unsigned bstep = 0;
  int b0 = 0;                                     // xs buffer of the current sample
  float acc = 0.0f;
  if (p.nsteps > 0) {
    stage(0, p.sample[0]);
    if (p.nsteps > 1) stage(1, p.sample[1]);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (warp_live) acc = pass(xs, xs, false, 0.0f);   // distances to the first sample, no update
  }
  // phase cycles of thread 0 ($BMU_K3_PROF): [0] CTA minimum, [1] grid exchange (own key stored -> global minimum
  // known), [2] barrier after the exchange (waiting for warp 0 + the staged sample), [3] lattice distance and
  // gaussian weight, [4] fused update + search pass
  long long pc[6] = {0, 0, 0, 0, 0, 0}, c0 = 0, c1;
  const bool prof = p.prof != nullptr && tid == 0;
#define K3_TICK(i) do { if (prof) { c1 = clock64(); pc[i] += c1 - c0; c0 = c1; } } while (0)
  if (prof) c0 = clock64();
  for (long t = 0; t < p.nsteps; t++) {
    const float talp = p.talp[t], trad = p.trad[t];
    // 1 / (2 r^2) of this step: needed only after the grid exchange, so the division runs while the CTA waits
    const double inv_den = gaussian ? __ddiv_rn(1.0, __dmul_rn(__dmul_rn(2.0, (double)trad), (double)trad)) : 0.0;
    const int b1 = b0 == 2 ? 0 : b0 + 1, b2 = b1 == 2 ? 0 : b1 + 1;
    // row of the sample two steps ahead: loaded HERE so that its L2 latency hides behind the exchange.  (Loaded
    // where it is used -- after the exchange, by warp 0, the warp every other one waits for -- it cost ~800 of
    // the ~9200 cycles of a step: $BMU_K3_PROF, profiles/r02_k3_phase_cycles.txt.)
    const long srow2 = (t + 2 < p.nsteps) ? (long)p.sample[t + 2] : 0;
    // ---- winner of step t: CTA minimum, then the grid exchange
    // The key carries the unit's lattice position instead of its index, (ty, tx) in index order (index = ty * xdim + tx,
    // tx < xdim), so the minimum still prefers the lower index (lvq_pak.c:79) and nobody divides after the exchange
    u64 k1 = (active && acc < FLT_MAX) ? make_key(acc, (ty << 12) | tx, false) : K3_NOKEY;   // lvq_pak.c:57,79
    k1 = warp_min_u64(k1);
    if (lane == 0) wred[warp] = k1;
    __syncthreads();
    if (warp < K3F_POLL_WARPS) {
      // Split polling (tools/ubench/grid_exchange.cu): warp w watches the slots of CTAs 32 w .. 32 w + 31, ONE strong
      // load per lane and wave, instead of one warp with five loads per lane; the five partial minima meet in shared
      // memory behind the barrier below.  Warp 0 publishes the CTA's key.
      u64 part = K3_NOKEY;
      if (warp == 0) {
        part = lane < K3F_THREADS / 32 ? wred[lane] : K3_NOKEY;
        part = warp_min_u64(part);
        K3_TICK(0);
      }
      if (G > 1) {
        const u64 tag = (u64)((bstep + 1) & 0xFFu);
        constexpr int SS = K3F_SLOT_STRIDE;
        u64 *slot = p.slots + ((size_t)(bstep & 1) * G) * K3_SLOT_STRIDE;
        if (tid == 0) st_relaxed_u64(slot + SS * blockIdx.x, part | tag);
        // All CTAs publish within ~50 cycles of each other and a store needs a few hundred cycles to reach L2:
        // polls issued at once arrive BEFORE the keys and cost a whole extra round trip (600-900 cycles), and
        // every poll wave in flight delays the stores it is waiting for, so the first poll waits.  A busy wait
        // in cycles (negative value) is steadier than __nanosleep, whose 100 ns are 260 cycles and whose 300 ns
        // are 1200; tuned with $BMU_K3_POLL_DELAY_NS, profiles/r02_k3_phase_cycles.txt, r02_k3_exchange_ubench.txt
        if (p.poll_delay_ns > 0) __nanosleep((unsigned)p.poll_delay_ns);
        else if (p.poll_delay_ns < 0) { const long long w0 = clock64(); while (clock64() - w0 < -p.poll_delay_ns) { } }
        const int c = warp * 32 + lane;
        bool pend = c < G;
        u64 v = K3_NOKEY;
        while (__any_sync(0xffffffffu, pend)) {
          if (prof) pc[5]++;
          if (pend) {
            v = ld_relaxed_u64(slot + SS * c);
            if ((v & 0xFFu) == tag) pend = false;
          }
        }
        part = warp_min_u64(c < G ? (v & ~0xFFull) : K3_NOKEY);
      }
      if (lane == 0) gw[warp] = part;                      // G == 1: warp 0's CTA minimum, the others K3_NOKEY
      K3_TICK(1);
    }
    bstep++;
    // sample t+1 (staged one step ago) must have landed before the fused pass reads it
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    K3_TICK(2);
    u64 g1 = gw[0];
#pragma unroll
    for (int w = 1; w < K3F_POLL_WARPS; w++) { const u64 o = gw[w]; g1 = o < g1 ? o : g1; }
    if (t + 2 < p.nsteps) stage(b2, srow2);                // buffer b2 was last read two passes ago
    // ---- update of step t fused with the search of step t+1
    bool upd = false;
    float a = talp;
    if (g1 != K3_NOKEY && warp_live) {                     // no winner (all distances NaN/Inf): step skipped
      const int bx = (int)(g1 >> 8) & 0xFFF, by = (int)(g1 >> 20) & 0xFFF;     // som_rout.c:641-642
      float dd;
      if (small_map) dd = p.topol == 4 ? rect_dist_small(bx, by, tx, ty) : hexa_dist_small(bx, by, tx, ty);
      else dd = p.topol == 4 ? rect_dist_dev(bx, by, tx, ty) : hexa_dist_dev(bx, by, tx, ty);
      if (gaussian) { a = gauss_alpha_fast(talp, dd, trad, inv_den); upd = active; }
      else upd = active && dd <= trad;                     // som_rout.c:496
    }
    const float *xt = xs + b0 * Dp;
    const float *xn = (t + 1 < p.nsteps) ? xs + b1 * Dp : xt;
    K3_TICK(3);
    if (warp_live) acc = pass(xt, xn, upd, a);             // warps without a unit (512 threads, 443 units at C5) sit out
    K3_TICK(4);
    b0 = b1;
  }
  if (prof)
    for (int i = 0; i < 6; i++) p.prof[(size_t)blockIdx.x * 8 + i] = pc[i];
#undef K3_TICK

  // ---- write the unit back
  __syncthreads();
  if (active) {
    float *dst = p.codes + (size_t)gidx * D;
#pragma unroll
    for (int j = 0; j < K3F_DR / 2; j++) {
      float c0, c1;
      unpack2(cr[j], c0, c1);
      dst[2 * j] = c0;
      dst[2 * j + 1] = c1;
    }
    for (int j = 0; j < nsp; j++) {
      float c0, c1;
      unpack2(sl2[(size_t)j * K3F_THREADS + tid], c0, c1);
      const int i = K3F_DR + 2 * j;
      if (i < D) dst[i] = c0;
      if (i + 1 < D) dst[i + 1] = c1;
    }
  }
}

static size_t k3f_smem_bytes(int D) {
  const int Dp = (D + 3) & ~3;
  return (size_t)3 * Dp * 4 + 24 * 8 + (size_t)((Dp - K3F_DR) / 2) * K3F_THREADS * 8;
}

bool k3_fused_eligible(const K3Params &p, const K3Plan &plan, bool has_mask, size_t smem_optin) {
  return p.mode <= K3_SOM_GAUSSIAN && !has_mask && p.fixed_xy == nullptr && p.D >= K3F_DR && plan.U <= K3F_THREADS &&
         plan.grid <= 32 * K3F_POLL_WARPS && p.talp && p.trad && p.xdim > 0 && p.xdim <= 4096 && p.ydim <= 4096 &&   // key: 12 bits each
         k3f_smem_bytes(p.D) <= smem_optin;
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
  return (size_t)2 * Dpad * 4 + (34 + 2 * K3_POLL_WARPS) * 8 + (size_t)((U + 3) & ~3) * 4;
}

K3Plan k3_plan(long M, int D, int num_sms, size_t smem_optin) {
  K3Plan best{};
  double best_cost = 1e300;
  // the grid exchange polls at most K3_MAX_GRID slots (NQ = 5 rounds of 32 lanes): a part with more SMs
  // than that simply leaves the extra ones idle (B200: 148)
  if (num_sms > K3_MAX_GRID) num_sms = K3_MAX_GRID;
  for (int G = 1; G <= num_sms; G = (G < num_sms && G * 2 > num_sms) ? num_sms : G * 2) {
    int U = (int)((M + G - 1) / G);
    int Us = (U + 1) & ~1;                 // even: unit pairs stay 8-byte aligned
    size_t fixed = k3_fixed_smem(D, U);
    size_t slice = (size_t)D * Us * 4;
    bool fits = fixed + slice <= smem_optin;
    // rough cycles per step: rounds of units per thread x (search + update) + exchange
    double rounds = (double)((U + 2 * K3_THREADS - 1) / (2 * K3_THREADS));
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
  if (plan.grid < 1 || plan.grid > K3_MAX_GRID) return cudaErrorInvalidConfiguration;   // winners of CTAs >= 160 would never be read
  const bool top2 = p.mode == K3_LVQ2 || p.mode == K3_LVQ3;
  void *fn;
  {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (k3_fused_eligible(p, plan, has_mask, (size_t)optin)) {
      const size_t smem = k3f_smem_bytes(p.D);
      cudaError_t e = cudaFuncSetAttribute((void *)k3_som_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) return e;
      e = cudaMemsetAsync(p.slots, 0, sizeof(u64) * 2 * K3_SLOT_STRIDE * plan.grid, st);
      if (e != cudaSuccess) return e;
      K3Params pp = p;
      if (pp.poll_delay_ns == K3_POLL_DELAY_AUTO) pp.poll_delay_ns = K3_POLL_DELAY_FUSED;
      void *args[] = {&pp};
      return cudaLaunchCooperativeKernel((void *)k3_som_fused_kernel, dim3(plan.grid), dim3(K3F_THREADS), args, smem, st);
    }
  }
  if (plan.slice_in_smem) {
    if (has_mask) fn = top2 ? (void *)k3_kernel<true, true, true> : (void *)k3_kernel<true, false, true>;
    else fn = top2 ? (void *)k3_kernel<false, true, true> : (void *)k3_kernel<false, false, true>;
  } else {
    if (has_mask) fn = top2 ? (void *)k3_kernel<true, true, false> : (void *)k3_kernel<true, false, false>;
    else fn = top2 ? (void *)k3_kernel<false, true, false> : (void *)k3_kernel<false, false, false>;
  }
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)plan.smem_bytes);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(p.slots, 0, sizeof(u64) * 2 * K3_SLOT_STRIDE * plan.grid, st);
  if (e != cudaSuccess) return e;
  K3Params pp = p;
  if (pp.poll_delay_ns == K3_POLL_DELAY_AUTO) pp.poll_delay_ns = plan.grid > 64 ? K3_POLL_DELAY_GENERIC : 0;
  void *args[] = {&pp};
  // cooperative launch only for its guarantee that all CTAs are co-resident (the slot
  // exchange spins on other CTAs)
  return cudaLaunchCooperativeKernel(fn, dim3(plan.grid), dim3(K3_THREADS), args, plan.smem_bytes, st);
}

}  // namespace bmu
