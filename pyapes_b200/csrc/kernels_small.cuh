// Small-grid CG: the WHOLE solve as one persistent cooperative kernel.
//
// Below ~10^6 cells an iteration of the fused kernels is launch- and latency-bound (two dependent
// launches with a grid-wide reduction each: ~30 us per iteration at 64^2 .. 1024^2, where the data
// would move in a few us).  Here every CTA stays resident, the vectors live in L2, the three
// dependencies of a CG iteration are grid barriers, and the Krylov scalars are recomputed
// redundantly -- and bit-identically -- by every CTA from the per-CTA partial sums, so there is no
// host round trip, no launch and no scalar broadcast inside the loop.
//
// Arithmetic is the generic kernels' (eval_equation, finalize_stage): same operation order, same
// roundings; only the reduction order differs (as between any two kernel variants).
// Preconditions (checked by the host): single GPU, static shell (every face Dirichlet, so no
// boundary work inside the loop), no nonlinear advection.
#pragma once
#include <cooperative_groups.h>

#include "kernels_generic.cuh"

namespace pa {

namespace cg = cooperative_groups;

constexpr int kSmallBlock = 256;

// Grid barrier on a monotonically increasing arrival counter (co-residency is guaranteed by the
// cooperative launch): one atomic and one spinning thread per CTA -- about a third of the latency of
// cooperative_groups' grid.sync() on this part.  `epoch` counts this CTA's barriers.
struct GridBarrier {
  unsigned int* counter;
  unsigned int epoch;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      __threadfence();
      atomicAdd(counter, 1u);
      while (*(volatile unsigned int*)counter < epoch) {
      }
      __threadfence();
    }
    __syncthreads();
  }
};

// deterministic all-CTA sum of NS per-CTA partials; every CTA gets the same bits
template <int NS>
__device__ __forceinline__ void coop_allsum(double (&v)[NS], double* partials, double* smem /* NS*32 */,
                                            GridBarrier& grid) {
  block_sum<NS>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) partials[s * kMaxPartials + blockIdx.x] = v[s];
  }
  grid.sync();
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    acc[s] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc[s] += __ldcg(&partials[s * kMaxPartials + b]);
  }
  block_sum<NS>(acc, smem);
  __shared__ double bc[NS];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) bc[s] = acc[s];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NS; ++s) v[s] = bc[s];
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kSmallBlock)
k_cg_persistent(GridDev g, EqDev<T> eq, T* __restrict__ xa, T* __restrict__ xb, T* __restrict__ r,
                T* __restrict__ d, SolverState* st, double* partA, double* partB) {
  GridBarrier grid{&st->ticket[7], 0u};  // ticket[7] is zero on entry (k_state_init)
  __shared__ SolverState ls;      // this CTA's copy of the solver state (identical in every CTA)
  __shared__ double red[2 * 32];
  if (threadIdx.x == 0) ls = *st;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  T* cur = xa;
  T* nxt = xb;
  while (!ls.done) {
    // d = r + beta d on the region                                       (linalg.py:141)
    {
      const T beta = (T)ls.scal[S_BETA];
      for (long long idx = first; idx < g.cells; idx += stride) {
        Cell c = decode(g, idx);
        if (in_region(g, c)) d[idx] = r[idx] + beta * d[idx];
      }
    }
    grid.sync();
    // alpha = rr / sum d*A(d)                                             (linalg.py:114-120)
    {
      double v[1] = {0.0};
      for (long long idx = first; idx < g.cells; idx += stride) {
        Cell c = decode(g, idx);
        if (in_region(g, c)) {
          T ad = eval_equation<T>(g, eq, c, [&](long long j) { return __ldcg(&d[j]); });
          T q = __ldcg(&d[idx]) * ad;
          v[0] += (double)q;
        }
      }
      coop_allsum<1>(v, partA, red, grid);
      if (threadIdx.x == 0) {
        ls.sum[R_A] = v[0];
        finalize_stage<T>(ST_CG_DAD, &ls);
      }
      __syncthreads();
    }
    // x_new = x + alpha d ; r -= alpha A(d) ; sums |r|^2 and |dx|^2      (linalg.py:122-137)
    {
      const T alpha = (T)ls.scal[S_ALPHA];
      double v[2] = {0.0, 0.0};
      for (long long idx = first; idx < g.cells; idx += stride) {
        Cell c = decode(g, idx);
        if (in_region(g, c)) {
          T ad = eval_equation<T>(g, eq, c, [&](long long j) { return __ldcg(&d[j]); });
          T xo = cur[idx];
          T xn = xo + alpha * __ldcg(&d[idx]);
          T rn = r[idx] - alpha * ad;
          r[idx] = rn;
          nxt[idx] = xn;
          T q = rn * rn;
          v[0] += (double)q;
          if (!on_shell(g, c)) {
            T df = xn - xo;
            T q2 = df * df;
            v[1] += (double)q2;
          }
        }
      }
      coop_allsum<2>(v, partB, red, grid);
      if (threadIdx.x == 0) {
        ls.sum[R_A] = v[0];
        ls.sum[R_B] = v[1];
        ls.sum[R_SHELL] = 0.0;  // static shell: the boundary cells never change
        finalize_stage<T>(ST_CG_FIN, &ls);
      }
      __syncthreads();
    }
    T* t = cur;
    cur = nxt;
    nxt = t;
  }
  grid.sync();  // nobody may still be spinning on the counter inside *st
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ls.ticket[7] = 0u;
    *st = ls;
  }
}

// co-resident CTAs of the persistent kernel on this device (0 if cooperative launch is unavailable)
template <typename T>
static int small_cg_max_blocks() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int dev = 0, coop = 0, per_sm = 0, sms = 0;
  cached = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent<T>, kSmallBlock, 0) != cudaSuccess)
    return 0;
  cached = per_sm * sms;
  return cached;
}

}  // namespace pa
