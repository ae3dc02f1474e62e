// Shared-memory-RESIDENT kernels for 2-D grids that fit the SMs' shared memory (1024^2 fp64 = 8 MB of the
// 33 MB that 148 SMs hold): explicit Euler stepping and the whole CG solve as ONE cooperative launch.
//
// Why: on such grids a time step / a CG phase of the streaming TMA kernels costs 8-18 us although its data
// would move in 2-3 us -- launch, pipeline fill and the row-by-row march of a 2-D "plane" dominate
// (BASELINE config 3 at 1024^2: 0.30 of the HBM-equivalent roofline, CG 0.27).  Here the field never leaves
// the SM between steps:
//   * CTA b (one per SM) owns R consecutive rows of the (n0, n2) grid and keeps them in shared memory for the
//     whole launch;
//   * the rows a neighbour needs -- the CTA's FIRST and LAST row -- are computed first in every step and stored
//     into a global exchange buffer as LL lines: every 8-byte word carries 4 bytes of data and a 4-byte sequence
//     number (the scheme of NCCL's LL protocol; aligned 8-byte accesses are single transactions).  The consumer --
//     the neighbour's thread that computes the same column of ITS boundary row in the next step -- loads the line
//     and retries until both sequence numbers are the ones it expects.  No flag word, no fence, no polling lane,
//     no barrier between CTAs: the dependency is per column, and the line is in flight while both CTAs compute
//     their interior rows.  (First version: red.release per warp / a publishing lane + bulk async copies into halo
//     rows -- 5.3 / 4.2 us per step at 1024^2 against 2 us of arithmetic; the gpu-scope fences and the flag polls
//     sat on the step's critical path.)
//   * CG keeps x, r and d resident, exchanges d's boundary rows the same way, and sums its two dot products over
//     per-CTA LL slots that every CTA reads back and adds in slot order -- every CTA holds bit-identical Krylov
//     scalars and runs the scalar stage (finalize_stage) redundantly, as k_cg_persistent does.  x_new is streamed
//     to the global ping-pong buffers every iteration, so that on exit they hold the last two iterates exactly
//     like the fused kernels leave them (the loser is the reference's VARo).
// Arithmetic: the star engine's operation order (star_cells_eq, FLAT), one rounding per reference operation ->
// Euler is bit-identical to the streaming path; CG differs only in the summation order of the dot products
// (as between any two kernel variants).
// Preconditions (host): 2-D mesh, constant-coefficient star operators, every face Dirichlet (static shell: the
// boundary cells never change after the first BC application), single GPU, cooperative launch available
// (co-residency is what makes waiting on another CTA legal).
#pragma once
#include "kernels_tma_pw.cuh"

namespace pa {

constexpr int kResThreads = 512;
constexpr int kResMaxCtas = 148;
constexpr int kResSmemMax = 227 * 1024;
constexpr int kResSlotLines = 2;  // all-reduce values per CTA and epoch

// ---- LL lines: {data, seq} pairs ---------------------------------------------------------------------
template <typename T>
struct LLOf;
template <>
struct LLOf<double> {
  typedef uint4 line;  // {lo, seq, hi, seq}
};
template <>
struct LLOf<float> {
  typedef uint2 line;  // {bits, seq}
};
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned seq) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(seq),
               "r"((unsigned)__double2hiint(v)), "r"(seq)
               : "memory");
}
__device__ __forceinline__ void ll_store(uint2* p, float v, unsigned seq) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4* p, unsigned seq, double& v) {
  unsigned a, b, c, d;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  v = __hiloint2double((int)c, (int)a);
  return b == seq && d == seq;
}
__device__ __forceinline__ bool ll_try(const uint2* p, unsigned seq, float& v) {
  unsigned a, b;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
  v = __uint_as_float(a);
  return b == seq;
}
// N consecutive lines: all loads are issued before the first check
template <typename T, int N>
__device__ __forceinline__ void ll_load(const typename LLOf<T>::line* p, unsigned seq, T (&v)[N]) {
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int e = 0; e < N; ++e) ok &= ll_try(p + e, seq, v[e]);
  } while (!ok);
}
template <typename T, int N>
__device__ __forceinline__ void ll_store_vec(typename LLOf<T>::line* p, unsigned seq, const T (&v)[N]) {
#pragma unroll
  for (int e = 0; e < N; ++e) ll_store(p + e, v[e], seq);
}

template <typename T>
__device__ __forceinline__ void sts_vec(T* p, const T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  V q;
  T* s = reinterpret_cast<T*>(&q);
#pragma unroll
  for (int e = 0; e < VecOf<T>::N; ++e) s[e] = v[e];
  *reinterpret_cast<V*>(p) = q;
}
template <typename T>
__device__ __forceinline__ void ldcg_vec(const T* p, T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  V q = __ldcg(reinterpret_cast<const V*>(p));
  const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
  for (int e = 0; e < VecOf<T>::N; ++e) v[e] = s[e];
}

// this CTA's rows and the thread's walk over their 16-byte vectors
struct ResCtx {
  int n2, nv;       // row length in cells / in vectors
  int row0, rows;   // first owned global row, owned rows
  int dq, dr;       // kResThreads / nv, kResThreads % nv  (incremental (row, vector) walk)
};
template <typename T>
__device__ __forceinline__ void res_ctx_init(ResCtx& c, const GridDev& g, int R) {
  c.n2 = g.n[2];
  c.nv = g.n[2] / VecOf<T>::N;
  c.row0 = blockIdx.x * R;
  c.rows = min(R, g.n[0] - c.row0);
  c.dq = kResThreads / c.nv;
  c.dr = kResThreads % c.nv;
}

// exchange buffer of one launch: [parity 2][cta][side 2][n2] lines; side 0 = the CTA's first row, 1 = its last
template <typename T>
struct ResLL {
  typedef typename LLOf<T>::line line;
  line* base;
  int n2, ctas;
  __device__ __forceinline__ line* row(unsigned parity, int cta, int side) const {
    return base + (((long long)(parity & 1u) * ctas + cta) * 2 + side) * n2;
  }
};

// the operator sum at one cell of a 2-D grid: star_cells_eq (kernels_tma_pw.cuh) with K::FLAT, same operation
// order; the CG form (one operator + the implicit-Euler shift, star_cells of kernels_tma.cuh) is the same code
// with nops == 1
template <typename T, bool LEAN, int NOPS>
__device__ __forceinline__ T res_star(const EqDev<T>& eq, const OpScale<T> (&sc)[NOPS > 0 ? NOPS : kMaxOps], int clx,
                                      int cz, T v0, T xp, T xm, T zp, T zm) {
  constexpr int MAXO = NOPS > 0 ? NOPS : kMaxOps;
  const int nops = NOPS > 0 ? NOPS : eq.nops;
  const int ix = LEAN ? 0 : clx, iz = LEAN ? 0 : cz;
  T res = (T)0;
#pragma unroll
  for (int q = 0; q < MAXO; ++q) {
    if (q >= nops) break;
    const OpDev<T>& o = eq.op[q];
    T s = o.coef[0][ix][0] * xp;
    s = s + o.coef[0][ix][1] * v0;
    s = s + o.coef[0][ix][2] * xm;
    T acc = s;
    T s2 = o.coef[2][iz][0] * zp;
    s2 = s2 + o.coef[2][iz][1] * v0;
    s2 = s2 + o.coef[2][iz][2] * zm;
    acc = acc + s2;
    if (sc[q].use) acc = acc * sc[q].scale;
    res = res + acc;
    if (o.has_shift) {
      const T m = o.shift * v0;
      res = m + res;
    }
  }
  return res;
}

// A(phi) on one vector of the thread's cells: `c` points at the vector inside its resident row (for the columns
// left and right of it), vm / vp are the same columns of the rows above / below.  emit(e, in_region, value).
template <typename T, int NOPS, typename F>
__device__ __forceinline__ void res_apply_vec(const GridDev& g, const EqDev<T>& eq,
                                              const OpScale<T> (&sc)[NOPS > 0 ? NOPS : kMaxOps], const T* c, int n2,
                                              int grow, int col, const T (&v0)[VecOf<T>::N],
                                              const T (&vm)[VecOf<T>::N], const T (&vp)[VecOf<T>::N], F emit) {
  constexpr int VEC = VecOf<T>::N;
  const T zl = col > 0 ? c[-1] : (T)0;
  const T zr = col + VEC < n2 ? c[VEC] : (T)0;
  const int clx = coef_class(g, 0, grow);
  const int lo2 = g.lo[2] > 2 ? g.lo[2] : 2, hi2 = g.hi[2] < n2 - 2 ? g.hi[2] : n2 - 2;
  if (clx == 0 && col >= lo2 && col + VEC <= hi2) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
      emit(e, true, res_star<T, true, NOPS>(eq, sc, 0, 0, v0[e], vp[e], vm[e], zp, zm));
    }
  } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int z = col + e;
      const bool in = z >= g.lo[2] && z < g.hi[2];
      const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
      T a = (T)0;
      if (in) a = res_star<T, false, NOPS>(eq, sc, clx, coef_class(g, 2, z), v0[e], vp[e], vm[e], zp, zm);
      emit(e, in, a);
    }
  }
}

// the same columns of the row above (dir -1) / below (+1) local row lr of a resident buffer `buf` (rows n2 apart):
// from shared memory, or -- across the CTA's first / last row -- from the neighbour's LL row with sequence number
// seq (seq == 0: from the global array `g0` instead, the state before the first exchange); rows outside the grid
// read as zeros (they only reach shell cells, which are never computed)
template <typename T>
__device__ __forceinline__ void res_neighbour_row(const ResCtx& c, const ResLL<T>& ll, const T* buf, int lr, int col,
                                                  int dir, unsigned parity, unsigned seq, const T* g0,
                                                  T (&v)[VecOf<T>::N]) {
  constexpr int VEC = VecOf<T>::N;
  const int nr = lr + dir;
  if (nr >= 0 && nr < c.rows) {
    lds_vec<T>(buf + (long long)nr * c.n2 + col, v);
    return;
  }
  const int ncta = (int)blockIdx.x + dir;
  if (ncta < 0 || ncta >= (int)gridDim.x) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = (T)0;
    return;
  }
  if (seq == 0u) {
    ldcg_vec<T>(g0 + (long long)(c.row0 + nr) * c.n2 + col, v);
    return;
  }
  ll_load<T, VEC>(ll.row(parity, ncta, dir < 0 ? 1 : 0) + col, seq, v);
}

// walk the vectors of local rows [ra, rb):  f(local row, column of the vector's first cell)
template <typename F>
__device__ __forceinline__ void res_rows(const ResCtx& c, int ra, int rb, F f) {
  if (rb <= ra) return;
  int lr = ra + (int)threadIdx.x / c.nv, cv = (int)threadIdx.x % c.nv;
  while (lr < rb) {
    f(lr, cv);
    lr += c.dq;
    cv += c.dr;
    if (cv >= c.nv) {
      cv -= c.nv;
      ++lr;
    }
  }
}
// Interior rows as a MARCH: the thread keeps one column block and walks down a contiguous run of rows with the rows
// above / at / below in registers (one 16-byte shared-memory load per row instead of three, no index arithmetic
// per vector).  The general walk above costs more integer and address instructions per vector than the stencil has
// fp64 instructions (measured: 0.55 us per 1024-cell row against 0.24 us of fp64 issue), so it is kept for the
// boundary rows and for row lengths that do not map onto the 512 threads.
// Mapping: nv >= 512: every thread takes columns tid, tid + 512, ... and all interior rows; nv < 512: G = 512 / nv
// thread groups share the interior rows in contiguous runs.  `use`: the mapping keeps >= 85 % of the threads busy.
struct ResMarch {
  bool use;
  int cv, cstep;  // first column block (in vectors) and the step to the next one
  int ra, rb;     // local rows [ra, rb)
};
__device__ __forceinline__ ResMarch res_march(const ResCtx& c) {
  ResMarch m;
  const int nint = c.rows - 2;
  m.cv = m.cstep = m.ra = m.rb = 0;
  m.use = false;
  if (nint < 1) return m;
  if (c.nv >= kResThreads) {
    const int passes = (c.nv + kResThreads - 1) / kResThreads;
    m.use = c.nv * 20 >= passes * kResThreads * 17;
    m.cv = threadIdx.x;
    m.cstep = kResThreads;
    m.ra = 1;
    m.rb = c.rows - 1;
  } else {
    int G = kResThreads / c.nv;
    if (G > nint) G = nint;
    const int chunk = (nint + G - 1) / G;
    m.use = (long long)nint * c.nv * 20 >= (long long)chunk * kResThreads * 17;
    const int grp = threadIdx.x / c.nv;
    m.cv = threadIdx.x - grp * c.nv;
    m.cstep = c.nv;  // one column block per thread
    m.ra = 1 + grp * chunk;
    m.rb = min(m.ra + chunk, c.rows - 1);
    if (grp >= G) m.rb = m.ra;
  }
  return m;
}
// f(local row, column, pointer to the vector in `buf`, v0, vm, vp) for the thread's share of the interior rows
template <typename T, typename F>
__device__ __forceinline__ void res_march_rows(const ResCtx& c, const ResMarch& m, const T* buf, F f) {
  constexpr int VEC = VecOf<T>::N;
  if (m.rb <= m.ra) return;
  for (int cv = m.cv; cv < c.nv; cv += m.cstep) {
    const int col = cv * VEC;
    const T* p = buf + (long long)m.ra * c.n2 + col;
    T vm[VEC], v0[VEC], vp[VEC];
    lds_vec<T>(p - c.n2, vm);
    lds_vec<T>(p, v0);
    for (int lr = m.ra; lr < m.rb; ++lr) {
      lds_vec<T>(p + c.n2, vp);
      f(lr, col, p, v0, vm, vp);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        vm[e] = v0[e];
        v0[e] = vp[e];
      }
      p += c.n2;
    }
  }
}

// the same share of the interior rows without the register window:  f(local row, column)
template <typename T, typename F>
__device__ __forceinline__ void res_march_plain(const ResCtx& c, const ResMarch& m, F f) {
  if (m.rb <= m.ra) return;
  for (int cv = m.cv; cv < c.nv; cv += m.cstep)
    for (int lr = m.ra; lr < m.rb; ++lr) f(lr, cv * VecOf<T>::N);
}

// the first and the last owned row (one row if the CTA owns a single one)
template <typename F>
__device__ __forceinline__ void res_boundary_rows(const ResCtx& c, F f) {
  const int nb = c.rows > 1 ? 2 : 1;
  for (int i = threadIdx.x; i < nb * c.nv; i += kResThreads) {
    const int k = i >= c.nv ? 1 : 0;
    f(k ? c.rows - 1 : 0, i - k * c.nv);
  }
}
// a vector of a boundary row goes to the neighbour(s) that read it
template <typename T>
__device__ __forceinline__ void res_send(const ResCtx& c, const ResLL<T>& ll, int lr, int col, unsigned parity,
                                         unsigned seq, const T (&v)[VecOf<T>::N]) {
  if (lr == 0 && blockIdx.x > 0) ll_store_vec<T, VecOf<T>::N>(ll.row(parity, blockIdx.x, 0) + col, seq, v);
  if (lr == c.rows - 1 && blockIdx.x + 1 < gridDim.x)
    ll_store_vec<T, VecOf<T>::N>(ll.row(parity, blockIdx.x, 1) + col, seq, v);
}

// =========================================================================================
// explicit Euler: nsteps steps of  phi <- phi + dt (rhs - A(phi))  on the region
// b0 holds phi on entry; step s reads buffer s&1 and writes (s+1)&1; the last TWO steps store every row to
// global memory, so that on exit b[nsteps&1] is the result and the other array the step before it (VARo).
// Sequence numbers seq0+1 .. seq0+nsteps are this launch's (never reused on the same exchange buffer).
// =========================================================================================
template <typename T, int NOPS, bool HAS_RHS>
__global__ void __launch_bounds__(kResThreads, 1)
k_euler_resident(GridDev g, EqDev<T> eq, T* __restrict__ b0, T* __restrict__ b1, const T* __restrict__ rhs, T dt,
                 int nsteps, int R, typename LLOf<T>::line* llbase, unsigned seq0) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int MAXO = NOPS > 0 ? NOPS : kMaxOps;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  ResCtx c;
  res_ctx_init<T>(c, g, R);
  const ResLL<T> ll{llbase, c.n2, (int)gridDim.x};
  const long long rowlen = c.n2;
  T* const sb0 = reinterpret_cast<T*>(base);
  T* const sb1 = sb0 + (long long)R * rowlen;

  for (int i = threadIdx.x; i < c.rows * c.nv; i += kResThreads) {  // resident copy of the owned rows
    const int lr = i / c.nv, cv = i - lr * c.nv;
    T v[VEC];
    ldcg_vec<T>(b0 + (long long)(c.row0 + lr) * rowlen + cv * VEC, v);
    sts_vec<T>(sb0 + (long long)lr * rowlen + cv * VEC, v);
  }
  __syncthreads();

  OpScale<T> sc[MAXO];
#pragma unroll
  for (int q = 0; q < MAXO; ++q)
    if (q < (NOPS > 0 ? NOPS : eq.nops)) sc[q] = op_scale<T>(eq.op[q]);

  const ResMarch march = res_march(c);
  for (int s = 0; s < nsteps; ++s) {
    const T* cur = (s & 1) ? sb1 : sb0;
    T* nxt = (s & 1) ? sb0 : sb1;
    T* gout = (s & 1) ? b0 : b1;
    const bool all_rows = s + 2 >= nsteps;
    const unsigned seq_in = s == 0 ? 0u : seq0 + (unsigned)s, seq_out = seq0 + (unsigned)s + 1u;
    auto finish = [&](int lr, int col, const T* p, const T (&v0)[VEC], const T (&vm)[VEC], const T (&vp)[VEC],
                      bool boundary) {
      const int grow = c.row0 + lr;
      T o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = v0[e];
      if (grow >= g.lo[0] && grow < g.hi[0]) {
        T av[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) av[e] = (T)0;
        if (HAS_RHS) {
          typedef typename VecOf<T>::type V;
          V q = __ldg(reinterpret_cast<const V*>(rhs + (long long)grow * rowlen + col));
          const T* qs = reinterpret_cast<const T*>(&q);
#pragma unroll
          for (int e = 0; e < VEC; ++e) av[e] = qs[e];
        }
        res_apply_vec<T, NOPS>(g, eq, sc, p, c.n2, grow, col, v0, vm, vp, [&](int e, bool in, T a) {
          if (in) {
            const T res = av[e] - a;
            o[e] = v0[e] + dt * res;
          }
        });
      }
      sts_vec<T>(nxt + (long long)lr * rowlen + col, o);
      if (boundary && s + 1 < nsteps) res_send<T>(c, ll, lr, col, (unsigned)(s + 1), seq_out, o);
      if (all_rows) sts_vec<T>(gout + (long long)grow * rowlen + col, o);
    };
    auto cell_row = [&](int lr, int cv, bool boundary) {
      const int col = cv * VEC;
      const T* p = cur + (long long)lr * rowlen + col;
      T v0[VEC], vm[VEC], vp[VEC];
      lds_vec<T>(p, v0);
      res_neighbour_row<T>(c, ll, cur, lr, col, -1, (unsigned)s, seq_in, b0, vm);
      res_neighbour_row<T>(c, ll, cur, lr, col, +1, (unsigned)s, seq_in, b0, vp);
      finish(lr, col, p, v0, vm, vp, boundary);
    };
    res_boundary_rows(c, [&](int lr, int cv) { cell_row(lr, cv, true); });
    if (march.use)
      res_march_rows<T>(c, march, cur, [&](int lr, int col, const T* p, const T (&v0)[VEC], const T (&vm)[VEC],
                                           const T (&vp)[VEC]) { finish(lr, col, p, v0, vm, vp, false); });
    else
      res_rows(c, 1, c.rows - 1, [&](int lr, int cv) { cell_row(lr, cv, false); });
    __syncthreads();  // end of step s: `cur` is free, `nxt` complete
  }
}

// =========================================================================================
// CG, the whole solve.  On entry: xa = x after the BC application (xb holds the same shell), r = rhs - A(x) on the
// region and 0 elsewhere, d = r (PW_RESID writes both), st = k_state_init + ST_CG_INIT (rr in sum[R_RR]).
// Sequence numbers: d rows of iteration k carry seq0 + k + 1, the all-reduce of epoch e carries seq0 + e.
// =========================================================================================
template <int NS>
__device__ __forceinline__ void res_allsum(double (&v)[NS], uint4* slots, unsigned epoch, unsigned seq, double* red,
                                           double* bc) {
  block_sum<NS>(v, red);
  if (threadIdx.x == 0) {
    uint4* mine = slots + ((long long)(epoch & 1u) * kResMaxCtas + blockIdx.x) * kResSlotLines;
#pragma unroll
    for (int s = 0; s < NS; ++s) ll_store(mine + s, v[s], seq);
  }
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = 0.0;
  if (threadIdx.x < gridDim.x)
    ll_load<double, NS>(slots + ((long long)(epoch & 1u) * kResMaxCtas + threadIdx.x) * kResSlotLines, seq, acc);
  block_sum<NS>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) bc[s] = acc[s];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NS; ++s) v[s] = bc[s];
}

template <typename T>
__global__ void __launch_bounds__(kResThreads, 1)
k_cg_resident(GridDev g, EqDev<T> eq, T* __restrict__ xa, T* __restrict__ xb, const T* __restrict__ r_in,
              const T* __restrict__ d_in, SolverState* st, int R, typename LLOf<T>::line* llbase, uint4* slots,
              unsigned seq0) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ SolverState ls;  // this CTA's copy of the solver state (bit-identical in every CTA)
  __shared__ double red[2 * 32];
  __shared__ double bc[2];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  ResCtx c;
  res_ctx_init<T>(c, g, R);
  const ResLL<T> ll{llbase, c.n2, (int)gridDim.x};
  const long long rowlen = c.n2;
  T* sd = reinterpret_cast<T*>(base);
  T* sr = sd + (long long)R * rowlen;
  T* sx = sr + (long long)R * rowlen;

  if (threadIdx.x == 0) ls = *st;
  for (int i = threadIdx.x; i < c.rows * c.nv; i += kResThreads) {
    const int lr = i / c.nv, cv = i - lr * c.nv;
    const long long go = (long long)(c.row0 + lr) * rowlen + cv * VEC, so = (long long)lr * rowlen + cv * VEC;
    T v[VEC];
    ldcg_vec<T>(d_in + go, v);
    sts_vec<T>(sd + so, v);
    ldcg_vec<T>(r_in + go, v);
    sts_vec<T>(sr + so, v);
    ldcg_vec<T>(xa + go, v);
    sts_vec<T>(sx + so, v);
  }
  __syncthreads();

  OpScale<T> sc[1];
  sc[0] = op_scale<T>(eq.op[0]);
  const ResMarch march = res_march(c);
  unsigned epoch = 1u;
  unsigned it = 0;
  while (!ls.done) {
    const unsigned seq_d = seq0 + it + 1u;
    // ---- phase A: d = r + beta d on the region (linalg.py:141), boundary rows first ---------------------
    const T beta = (T)ls.scal[S_BETA];
    auto dupd = [&](int lr, int col, bool boundary) {
      const int grow = c.row0 + lr;
      T* dp = sd + (long long)lr * rowlen + col;
      T dv[VEC], rv[VEC];
      lds_vec<T>(dp, dv);
      if (grow >= g.lo[0] && grow < g.hi[0]) {
        lds_vec<T>(sr + (long long)lr * rowlen + col, rv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int z = col + e;
          if (z >= g.lo[2] && z < g.hi[2]) dv[e] = rv[e] + beta * dv[e];
        }
        sts_vec<T>(dp, dv);
      }
      if (boundary) res_send<T>(c, ll, lr, col, it, seq_d, dv);
    };
    double qa[1] = {0.0};
    auto dad_fin = [&](int lr, int col, const T* dp, const T (&dv)[VEC], const T (&vm)[VEC], const T (&vp)[VEC]) {
      const int grow = c.row0 + lr;
      if (grow < g.lo[0] || grow >= g.hi[0]) return;
      res_apply_vec<T, 1>(g, eq, sc, dp, c.n2, grow, col, dv, vm, vp, [&](int e, bool in, T a) {
        if (in) {
          const T q = dv[e] * a;
          qa[0] += (double)q;
        }
      });
    };
    // general walk: the rows above / below from shared memory or, across the CTA's edge, from the neighbour's LL row
    auto with_rows = [&](int lr, int cv, auto fin) {
      const int col = cv * VEC;
      const T* dp = sd + (long long)lr * rowlen + col;
      T dv[VEC], vm[VEC], vp[VEC];
      lds_vec<T>(dp, dv);
      res_neighbour_row<T>(c, ll, sd, lr, col, -1, it, seq_d, (const T*)nullptr, vm);
      res_neighbour_row<T>(c, ll, sd, lr, col, +1, it, seq_d, (const T*)nullptr, vp);
      fin(lr, col, dp, dv, vm, vp);
    };
    res_boundary_rows(c, [&](int lr, int cv) { dupd(lr, cv * VEC, true); });
    if (march.use)
      res_march_plain<T>(c, march, [&](int lr, int col) { dupd(lr, col, false); });
    else
      res_rows(c, 1, c.rows - 1, [&](int lr, int cv) { dupd(lr, cv * VEC, false); });
    __syncthreads();  // every own row of d is updated
    // the interior rows first: the neighbours' rows are still in flight
    if (march.use)
      res_march_rows<T>(c, march, sd, dad_fin);
    else
      res_rows(c, 1, c.rows - 1, [&](int lr, int cv) { with_rows(lr, cv, dad_fin); });
    res_boundary_rows(c, [&](int lr, int cv) { with_rows(lr, cv, dad_fin); });
    res_allsum<1>(qa, slots, epoch, seq0 + epoch, red, bc);
    ++epoch;
    if (threadIdx.x == 0) {
      ls.sum[R_A] = qa[0];
      finalize_stage<T>(ST_CG_DAD, &ls);  // alpha = rr / dAd (linalg.py:114-120)
    }
    __syncthreads();
    // ---- phase B: x_new = x + alpha d ; r -= alpha A(d) ; sums |r|^2, |dx|^2 (linalg.py:122-137) ------
    const T alpha = (T)ls.scal[S_ALPHA];
    T* gxn = ((it + 1u) & 1u) ? xb : xa;
    double qb[2] = {0.0, 0.0};
    auto upd_fin = [&](int lr, int col, const T* dp, const T (&dv)[VEC], const T (&vm)[VEC], const T (&vp)[VEC]) {
      const int grow = c.row0 + lr;
      if (grow < g.lo[0] || grow >= g.hi[0]) return;
      T* rp = sr + (long long)lr * rowlen + col;
      T* xp = sx + (long long)lr * rowlen + col;
      T rv[VEC], xv[VEC];
      lds_vec<T>(rp, rv);
      lds_vec<T>(xp, xv);
      const bool rshell = grow == 0 || grow == g.n[0] - 1;
      res_apply_vec<T, 1>(g, eq, sc, dp, c.n2, grow, col, dv, vm, vp, [&](int e, bool in, T a) {
        if (in) {
          const T xo = xv[e];
          const T xn = xo + alpha * dv[e];
          const T rn = rv[e] - alpha * a;
          xv[e] = xn;
          rv[e] = rn;
          const T q = rn * rn;
          qb[0] += (double)q;
          const int z = col + e;
          if (!rshell && z != 0 && z != c.n2 - 1) {
            const T df = xn - xo;
            const T q2 = df * df;
            qb[1] += (double)q2;
          }
        }
      });
      sts_vec<T>(rp, rv);
      sts_vec<T>(xp, xv);
      sts_vec<T>(gxn + (long long)grow * rowlen + col, xv);
    };
    res_boundary_rows(c, [&](int lr, int cv) { with_rows(lr, cv, upd_fin); });
    if (march.use)
      res_march_rows<T>(c, march, sd, upd_fin);
    else
      res_rows(c, 1, c.rows - 1, [&](int lr, int cv) { with_rows(lr, cv, upd_fin); });
    res_allsum<2>(qb, slots, epoch, seq0 + epoch, red, bc);
    ++epoch;
    if (threadIdx.x == 0) {
      ls.sum[R_A] = qb[0];
      ls.sum[R_B] = qb[1];
      ls.sum[R_SHELL] = 0.0;  // static shell: the boundary cells never change
      finalize_stage<T>(ST_CG_FIN, &ls);
    }
    __syncthreads();
    ++it;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = ls;
}

// ---- host ------------------------------------------------------------------------------------
struct ResPlan {
  int ctas, R;
  size_t smem;
};

// Exchange buffers (all-reduce slots + LL rows): a small per-device pool handed out round-robin (two resident
// launches only run at the same time if they come from different streams and both fit the SMs).  Nothing is zeroed
// per launch: every launch takes a fresh range of sequence numbers, so lines of earlier launches never match.
struct ResBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  unsigned next_seq = 1u;
};
constexpr size_t kResSlotBytes = 16384;  // [2][kResMaxCtas][kResSlotLines] uint4, rounded up

inline ResBuf* res_exchange_buffer(size_t ll_bytes) {
  constexpr int kPool = 4;
  static ResBuf pool[kMaxDevices][kPool];
  static unsigned rr[kMaxDevices] = {};
  const int dev = current_device();
  ResBuf& b = pool[dev][__atomic_fetch_add(&rr[dev], 1u, __ATOMIC_RELAXED) % kPool];
  const size_t need = kResSlotBytes + ll_bytes;
  if (b.bytes < need) {
    if (b.ptr) cudaFree(b.ptr);  // (synchronises the device: no launch still uses it)
    b.ptr = nullptr;
    b.bytes = 0;
    // cudaMemset of device memory is asynchronous: without the synchronisation a launch on a non-blocking stream
    // can start first, and the late memset then erases lines the kernel is waiting for (observed: a hang)
    if (cudaMalloc(&b.ptr, need) != cudaSuccess || cudaMemset(b.ptr, 0, need) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
      cudaGetLastError();
      if (b.ptr) cudaFree(b.ptr);
      b.ptr = nullptr;
      return nullptr;
    }
    b.bytes = need;  // (next_seq keeps counting: recycled memory may hold this pool entry's old lines)
  }
  return &b;
}
// the first of `count` fresh sequence numbers is seq0 + 1; on wrap-around the buffer is cleared on the stream
inline bool res_take_seq(ResBuf& b, unsigned count, cudaStream_t s, unsigned* seq0) {
  if (b.next_seq > 0xffffffffu - count - 16u) {
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemset(b.ptr, 0, b.bytes) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess)
      return false;
    b.next_seq = 1u;
  }
  *seq0 = b.next_seq;
  b.next_seq += count + 1u;
  return true;
}

// resident rows per CTA: Euler 2R (two buffers), CG 3R (d, r, x)
template <typename T>
inline bool res_plan(const GridDev& g, bool cg, ResPlan& p) {
  constexpr int VEC = VecOf<T>::N;
  if (!g.act[0] || g.act[1] || !g.act[2]) return false;  // 2-D meshes: kernel axes (0, 2)
  if (g.n[2] % VEC != 0 || g.n[2] < 2 * VEC || g.n[0] < 3) return false;
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop || sms < 1) return false;
  if (sms > kResMaxCtas) sms = kResMaxCtas;
  const int want = g.n[0] < sms ? g.n[0] : sms;
  p.R = (g.n[0] + want - 1) / want;
  p.ctas = (g.n[0] + p.R - 1) / p.R;
  const long long rows = cg ? 3LL * p.R : 2LL * p.R;
  const long long bytes = rows * g.n[2] * (long long)sizeof(T) + 128;
  if (bytes > kResSmemMax - 4096) return false;  // (static shared memory of the kernels: < 2 KB)
  p.smem = (size_t)bytes;
  return true;
}

template <typename T>
inline size_t res_ll_bytes(const GridDev& g, const ResPlan& p) {
  return (size_t)2 * p.ctas * 2 * g.n[2] * sizeof(typename LLOf<T>::line);
}

template <typename T>
inline bool res_eq_ok(const pa_equation& eq) {
  if (eq.nops < 1 || eq.nops > PA_MAX_OPS) return false;
  for (int k = 0; k < eq.nops; ++k) {
    const pa_op& o = eq.ops[k];
    if (o.kind != PA_OP_STAR || o.param_field != nullptr || o.edge != 0 || o.coef_tab[0] || o.coef_tab[1] ||
        o.coef_tab[2])
      return false;
  }
  return true;
}

template <typename T, int NOPS, bool HAS_RHS>
static bool launch_euler_resident_n(cudaStream_t s, const ResPlan& p, const GridDev& g, const EqDev<T>& eq, T* b0,
                                    T* b1, const T* rhs, T dt, int nsteps) {
  if (cudaFuncSetAttribute(k_euler_resident<T, NOPS, HAS_RHS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)p.smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ResBuf* buf = res_exchange_buffer(res_ll_bytes<T>(g, p));
  unsigned seq0 = 0;
  if (!buf || !res_take_seq(*buf, (unsigned)nsteps + 1u, s, &seq0)) return false;
  typename LLOf<T>::line* ll = (typename LLOf<T>::line*)((char*)buf->ptr + kResSlotBytes);
  int R = p.R;
  GridDev gg = g;
  EqDev<T> e = eq;
  void* args[] = {(void*)&gg, (void*)&e, (void*)&b0, (void*)&b1, (void*)&rhs, (void*)&dt, (void*)&nsteps, (void*)&R,
                  (void*)&ll, (void*)&seq0};
  if (cudaLaunchCooperativeKernel((void*)k_euler_resident<T, NOPS, HAS_RHS>, dim3(p.ctas), dim3(kResThreads), args,
                                  p.smem, s) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// nsteps explicit Euler steps in one launch; false: not launched (the caller takes the streaming path)
template <typename T>
bool launch_euler_resident(cudaStream_t s, const GridDev& g, const pa_equation& peq, const EqDev<T>& eq, T* b0, T* b1,
                           const T* rhs, T dt, int nsteps) {
  ResPlan p;
  if (!res_eq_ok<T>(peq) || !res_plan<T>(g, false, p)) return false;
  if (rhs) {
    if (eq.nops == 1) return launch_euler_resident_n<T, 1, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
    if (eq.nops == 2) return launch_euler_resident_n<T, 2, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
    return launch_euler_resident_n<T, 0, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
  }
  if (eq.nops == 1) return launch_euler_resident_n<T, 1, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
  if (eq.nops == 2) return launch_euler_resident_n<T, 2, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
  return launch_euler_resident_n<T, 0, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps);
}

// the whole CG solve in one launch (eq: ONE star operator, has_shift for the implicit-Euler term); at most
// max_it + 1 iterations run (finalize_stage ST_CG_FIN)
template <typename T>
bool launch_cg_resident(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, T* xa, T* xb, const T* r, const T* d,
                        SolverState* st, int max_it) {
  ResPlan p;
  if (eq.nops != 1 || max_it < 0 || max_it > 1000000000 || !res_plan<T>(g, true, p)) return false;
  if (cudaFuncSetAttribute(k_cg_resident<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ResBuf* buf = res_exchange_buffer(res_ll_bytes<T>(g, p));
  unsigned seq0 = 0;
  if (!buf || !res_take_seq(*buf, 2u * ((unsigned)max_it + 3u) + 2u, s, &seq0)) return false;
  uint4* slots = (uint4*)buf->ptr;
  typename LLOf<T>::line* ll = (typename LLOf<T>::line*)((char*)buf->ptr + kResSlotBytes);
  int R = p.R;
  GridDev gg = g;
  EqDev<T> e = eq;
  void* args[] = {(void*)&gg, (void*)&e, (void*)&xa, (void*)&xb, (void*)&r, (void*)&d, (void*)&st, (void*)&R,
                  (void*)&ll, (void*)&slots, (void*)&seq0};
  if (cudaLaunchCooperativeKernel((void*)k_cg_resident<T>, dim3(p.ctas), dim3(kResThreads), args, p.smem, s) !=
      cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

}  // namespace pa
