// Shared device-side definitions for the pyapes_b200 kernels (sm_100a).
//
// Arithmetic policy: the whole library is compiled with -fmad=false and every update is
// written in the reference's operation order, so that elementwise results are bit-identical
// to the reference's eager torch ops (one IEEE rounding per torch op).  See DESIGN.md §3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyapes_b200.h"

namespace pa {

constexpr int kMaxOps = PA_MAX_OPS;
constexpr int kBlock = 256;          // threads per CTA of the generic kernels
constexpr int kMaxPartials = 4096;   // upper bound on CTAs taking part in a reduction
constexpr int kNumSums = 4;          // reduction slots per launch
constexpr int kNumSMs = 148;         // B200
constexpr int kMaxDevices = 16;      // per-device caches (streams, scratch, function attributes)

// ordinal of the calling thread's current device, clamped into the per-device cache arrays
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) d = 0;
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

// ---- grid ------------------------------------------------------------------------------
struct GridDev {
  int n[3];
  int lo[3], hi[3];
  int gn0, goff0, olo0, ohi0;
  int act[3];  // 1 if the kernel axis carries a mesh axis
  long long cells;
};

inline GridDev make_grid(const pa_grid& g) {
  GridDev d;
  for (int a = 0; a < 3; ++a) {
    d.n[a] = g.n[a];
    d.lo[a] = g.lo[a];
    d.hi[a] = g.hi[a];
  }
  // mesh axis -> kernel axis: 3-D (0,1,2); 2-D (0,2) so that the march runs along the first mesh
  // axis and the second is the contiguous one; 1-D (2)
  d.act[0] = g.ndim >= 2;
  d.act[1] = g.ndim == 3;
  d.act[2] = 1;
  d.gn0 = g.gn0;
  d.goff0 = g.goff0;
  d.olo0 = g.olo0;
  d.ohi0 = g.ohi0;
  d.cells = (long long)g.n[0] * g.n[1] * g.n[2];
  return d;
}

// ---- equation --------------------------------------------------------------------------
template <typename T>
struct OpDev {
  int kind;
  int has_param;
  T sign;
  T param;
  T coef[3][3][3];
  const T* adv;
  T two_dx[3];
  T dx[3];
  int zero_am_lo[3];
  int zero_ap_hi[3];
  const T* param_field;
  int edge;
  T adv_const;
  const T* coef_tab[3];
  // fused CG kernels only: a second, purely diagonal operator c*phi (the implicit-Euler term
  // (1/dt)*phi that `ddt` appends) folded into this one:  A(phi) = (0 + star(phi)) + shift*phi
  int has_shift;
  T shift;
};

template <typename T>
struct EqDev {
  int nops;
  OpDev<T> op[kMaxOps];
};

template <typename T>
inline EqDev<T> make_eq(const pa_equation& e) {
  EqDev<T> d;
  d.nops = e.nops;
  for (int k = 0; k < e.nops && k < kMaxOps; ++k) {
    const pa_op& s = e.ops[k];
    OpDev<T>& o = d.op[k];
    o.kind = s.kind;
    o.has_param = s.has_param;
    o.sign = (T)s.sign;
    o.param = (T)s.param;
    for (int a = 0; a < 3; ++a) {
      for (int c = 0; c < 3; ++c)
        for (int q = 0; q < 3; ++q) o.coef[a][c][q] = (T)s.coef[a][c][q];
      o.two_dx[a] = (T)s.two_dx[a];
      o.dx[a] = (T)s.dx[a];
      o.zero_am_lo[a] = s.zero_am_lo[a];
      o.zero_ap_hi[a] = s.zero_ap_hi[a];
    }
    o.adv = (const T*)s.adv;
    o.param_field = (const T*)s.param_field;
    o.edge = s.edge;
    o.adv_const = (T)s.adv_const;
    for (int a = 0; a < 3; ++a) o.coef_tab[a] = (const T*)s.coef_tab[a];
    o.has_shift = 0;
    o.shift = (T)0;
  }
  return d;
}

// ---- faces -----------------------------------------------------------------------------
template <typename T>
struct FaceDev {
  int axis, side, kind;
  T value;
  const T* values;
};

// ---- index helpers ---------------------------------------------------------------------
struct Cell {
  int i[3];
  long long idx;
};

__device__ __forceinline__ Cell decode(const GridDev& g, long long idx) {
  Cell c;
  c.idx = idx;
  int n12 = g.n[1] * g.n[2];
  c.i[0] = (int)(idx / n12);
  int rem = (int)(idx - (long long)c.i[0] * n12);
  c.i[1] = rem / g.n[2];
  c.i[2] = rem - c.i[1] * g.n[2];
  return c;
}

__device__ __forceinline__ long long stride_of(const GridDev& g, int a) {
  return a == 2 ? 1LL : (a == 1 ? (long long)g.n[2] : (long long)g.n[1] * g.n[2]);
}

// coefficient class of index `i` along axis `a` (fdc.py:84-93 masks rolled one inward)
__device__ __forceinline__ int coef_class(const GridDev& g, int a, int i) {
  int gi = (a == 0) ? i + g.goff0 : i;
  int gn = (a == 0) ? g.gn0 : g.n[a];
  return gi == 1 ? 1 : (gi == gn - 2 ? 2 : 0);
}

__device__ __forceinline__ bool in_region(const GridDev& g, const Cell& c) {
  return c.i[0] >= g.lo[0] && c.i[0] < g.hi[0] && c.i[1] >= g.lo[1] && c.i[1] < g.hi[1] &&
         c.i[2] >= g.lo[2] && c.i[2] < g.hi[2];
}

// a cell on the outer shell of the GLOBAL box (any active axis index 0 or n-1)
__device__ __forceinline__ bool on_shell(const GridDev& g, const Cell& c) {
  bool s = false;
  if (g.act[0]) {
    int gi = c.i[0] + g.goff0;
    s |= (gi == 0) | (gi == g.gn0 - 1);
  }
  if (g.act[1]) s |= (c.i[1] == 0) | (c.i[1] == g.n[1] - 1);
  if (g.act[2]) s |= (c.i[2] == 0) | (c.i[2] == g.n[2] - 1);
  return s;
}

__device__ __forceinline__ bool owned(const GridDev& g, const Cell& c) {
  return c.i[0] >= g.olo0 && c.i[0] < g.ohi0;
}

// ---- the operator sum at one cell, given a functor that returns field values ------------
// `val(idx)` returns the field at linear index idx.  Neighbours wrap around like
// torch.roll (fdc.py:198); on non-periodic axes the wrapped values only reach cells outside
// the solver region.
template <typename T, typename F>
__device__ __forceinline__ T eval_equation(const GridDev& g, const EqDev<T>& eq, const Cell& c,
                                           F val) {
  T vc = val(c.idx);
  T vp[3], vm[3];
  long long ip[3], im[3];
  int cls[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (!g.act[a]) continue;
    long long st = stride_of(g, a);
    int i = c.i[a], n = g.n[a];
    ip[a] = (i + 1 == n) ? c.idx - (long long)(n - 1) * st : c.idx + st;
    im[a] = (i == 0) ? c.idx + (long long)(n - 1) * st : c.idx - st;
    vp[a] = val(ip[a]);
    vm[a] = val(im[a]);
    cls[a] = coef_class(g, a, i);
  }
  T res = (T)0;
  for (int k = 0; k < eq.nops; ++k) {
    const OpDev<T>& o = eq.op[k];
    T acc = (T)0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!g.act[a]) continue;
      T Ap, Ac, Am;
      if (o.kind == PA_OP_STAR && o.coef_tab[a] != nullptr) {  // per-index table (rz)
        const T* t = o.coef_tab[a] + 3 * c.i[a];
        Ap = t[0];
        Ac = t[1];
        Am = t[2];
      } else if (o.kind == PA_OP_STAR) {
        Ap = o.coef[a][cls[a]][0];
        Ac = o.coef[a][cls[a]][1];
        Am = o.coef[a][cls[a]][2];
      } else if (o.kind == PA_OP_DIV_CENTRAL_FIELD) {
        // Ap = 1*u[+1] / 2dx ; Ac = 0 ; Am = -1*u[-1] / 2dx  (fdc.py:736-738, 607-609)
        Ap = (cls[a] == 2 && o.zero_ap_hi[a]) ? (T)0 : o.adv[ip[a]] / o.two_dx[a];
        Ac = (T)0;
        Am = (cls[a] == 1 && o.zero_am_lo[a]) ? (T)0 : (-o.adv[im[a]]) / o.two_dx[a];
      } else if (o.kind == PA_OP_DIV_UPWIND_FIELD) {
        T u = o.adv[c.idx];  // fdc.py:765-770 (same u on every axis, no 1/dx)
        Ap = (T)2 * (u < (T)0 ? u : (T)0);
        Ac = (T)0;
        Am = (T)2 * (u > (T)0 ? u : (T)0);
      } else {
        T u = o.adv[c.idx];
        T up = u > (T)0 ? u : (T)0, um = u < (T)0 ? u : (T)0;
        Ap = um / o.dx[a];
        Ac = (up - um) / o.dx[a];
        Am = (-up) / o.dx[a];
      }
      // ((0 + 0*v[+2]) + Ap*v[+1]) + Ac*v) + Am*v[-1]) + 0*v[-2]   (fdc.py:188-198)
      T s = Ap * vp[a];
      s = s + Ac * vc;
      s = s + Am * vm[a];
      acc = acc + s;  // axes accumulate into zeros (fdc.py:103-108)
    }
    if (o.param_field != nullptr)
      acc = acc * o.param_field[c.idx];  // Tensor coefficient (fdm.py:130,169)
    else if (o.has_param)
      acc = acc * o.param;  // fdm.py:169
    acc = acc * o.sign;                    // ops.py:140-143
    res = res + acc;                       // ops.py:149
  }
  return res;
}

// ---- reductions ------------------------------------------------------------------------
// Device-resident solver state: every scalar of the Krylov recurrences lives here so that an
// iteration needs no host round trip (SURVEY §7 "hard parts").
struct SolverState {
  double sum[8];    // finalized reductions (meaning depends on the solver)
  double scal[8];   // alpha, beta, omega, rho, ... already rounded to the field dtype
  double tol;
  double tolerance;
  int itr;
  int max_it;
  int done;
  int status;
  int finished_flag;  // bicgstab's `finished`
  int swaps;          // BiCGSTAB: x updates performed (== ping-pong swaps); itr is bumped BEFORE the update there
  unsigned int ticket[8];
  unsigned long long epoch;  // sequence number of the next peer-memory all-reduce (p2p_allreduce)
  unsigned int halo_count;   // boundary CTAs of phase B that have written their planes (monotonic)
  unsigned int halo_target;  // value halo_count reaches when the current iteration's are all done
  unsigned int halo_cnt[2];  // peer-memory halo exchange: phase-B CTAs that have stored their share of the
                             // first / last owned plane into the neighbour's landing zone (reset by the last one)
};

// ---- all-reduce over NVLink peer memory -----------------------------------------------------
// One process per GPU; every rank owns a small mailbox in device memory that its peers map through
// CUDA IPC (api.cu pa_p2p_*).  A reduction is executed by ONE thread per rank -- the thread that has
// just finished the grid-wide reduction of a fused kernel -- so the sum over ranks and the scalar
// stage that consumes it happen inside the kernel that produced the local sum: no NCCL launch, no
// finalize launch.  Mailbox of the receiver: [slot 2][source rank 16][8 words]; words 0..3 carry the
// values, word 7 the epoch.  Epochs increase monotonically over the life of the process, a rank can
// be at most one reduction ahead of the slowest one, so two slots never collide.  Every rank adds the
// contributions in rank order: the result is bitwise identical everywhere.
struct P2PDev {
  unsigned long long* const* peers;  // device array [nranks]: mailbox of every rank (peers[me] is local)
  int me, nranks;
  int slot0, count;                  // which SolverState::sum entries are reduced
};

// ---- halo exchange over NVLink peer memory (CG on slabs) ------------------------------------------
// Every rank owns a landing zone next to its mailbox (same IPC allocation): [slot 2][side 2] planes, side 0 =
// the ghost plane below the owned range, side 1 the one above, plus one flag word per side.  Phase B of
// iteration k stores the new r of its first / last owned plane straight into the neighbour's landing zone
// (slot = iteration parity) while it writes its own copy, and the last CTA to finish a plane publishes the
// sequence number of the iteration in the neighbour's flag.  The producer lane of the neighbour's next phase A
// waits for that flag and TMA-loads the ghost plane from the landing zone.  No NCCL call, no second stream and
// no waiting kernel inside the loop: the only dependency is between kernels on DIFFERENT GPUs.
struct HaloDev {
  void* dst[2];                        // [0]: first owned plane -> lower neighbour; [1]: last -> upper (null: none);
                                       // addresses of landing slot 0, slot s lies s * slot_bytes further
  long long slot_bytes;
  unsigned long long* flag_dst[2];     // the neighbours' flag words
  const unsigned long long* flag_src;  // this rank's two flag words (volatile reads)
  int tiles;                           // CTAs that share one plane (tiles_y * tiles_z)
  int slot;                            // landing slot phase B writes / phase A reads
  int on;
};

// Watchdog of the flag wait: a peer that never arrives (crashed rank, a rank that left the solve on an
// error) must not leave this GPU spinning for ever.  After kP2PTimeoutNs the reduction gives up and
// returns false; the caller latches `done` with status PA_PEER_LOST and the host reports PA_ERR_NCCL.
constexpr unsigned long long kP2PTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool p2p_allreduce(const P2PDev& pp, unsigned long long epoch, double* vals) {
  const int slot = (int)(epoch & 1ull);
  const int off = (slot * 16 + pp.me) * 8;
  for (int p = 0; p < pp.nranks; ++p) {
    volatile unsigned long long* dst = pp.peers[p] + off;
    for (int k = 0; k < pp.count; ++k) dst[k] = (unsigned long long)__double_as_longlong(vals[k]);
  }
  __threadfence_system();
  for (int p = 0; p < pp.nranks; ++p) {
    volatile unsigned long long* dst = pp.peers[p] + off;
    dst[7] = epoch;
  }
  double tot[4] = {0.0, 0.0, 0.0, 0.0};
  unsigned long long t0 = 0ull;  // the clock is only read once a wait has lasted a while
  for (int q = 0; q < pp.nranks; ++q) {
    volatile unsigned long long* src = pp.peers[pp.me] + (slot * 16 + q) * 8;
    unsigned int spins = 0;
    while (src[7] != epoch) {
      if ((++spins & 0xfffu) == 0u) {
        const unsigned long long now = global_timer_ns();
        if (t0 == 0ull) t0 = now;
        else if (now - t0 > kP2PTimeoutNs) return false;
      }
    }
    __threadfence_system();
    for (int k = 0; k < pp.count; ++k) tot[k] += __longlong_as_double((long long)src[k]);
  }
  for (int k = 0; k < pp.count; ++k) vals[k] = tot[k];
  return true;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NS doubles; result valid in thread 0.  Fixed shuffle/smem order ->
// deterministic for a fixed launch geometry.
template <int NS>
__device__ __forceinline__ void block_sum(double (&v)[NS], double* smem /* NS*32 */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    double x = warp_sum(v[s]);
    if (lane == 0) smem[s * 32 + w] = x;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double x = lane < nw ? smem[s * 32 + lane] : 0.0;
      x = warp_sum(x);
      v[s] = x;
    }
  }
  __syncthreads();
}

// Grid-wide deterministic reduction: every CTA stores its partials, the last CTA to arrive
// (ticket) sums them in a fixed order and calls fin(sums) from thread 0.
template <int NS, typename Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NS], double* partials, int nblocks,
                                            int block_id, unsigned int* ticket, Fin fin) {
  __shared__ double red_smem[NS * 32];
  __shared__ bool is_last;
  block_sum<NS>(v, red_smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) partials[s * kMaxPartials + block_id] = v[s];
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == (unsigned int)nblocks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    acc[s] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
      acc[s] += __ldcg(&partials[s * kMaxPartials + b]);
  }
  block_sum<NS>(acc, red_smem);
  if (threadIdx.x == 0) {
    *ticket = 0u;
    fin(acc);
  }
}

template <typename T>
__device__ __forceinline__ T nan_to_num0(T x) {  // linalg.py:302-305
  return (isnan(x) || isinf(x)) ? (T)0 : x;
}

}  // namespace pa
