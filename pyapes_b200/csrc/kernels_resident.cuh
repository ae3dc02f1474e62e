// Shared-memory-RESIDENT kernels for 2-D grids that fit the SMs' shared memory (1024^2 fp64 = 8 MB of the
// 33 MB that 148 SMs hold): explicit Euler stepping and the whole CG solve as ONE cooperative launch.
//
// Why: on such grids a time step / a CG phase of the streaming TMA kernels costs 8-18 us although its data
// would move in 2-3 us -- launch, pipeline fill and the row-by-row march of a 2-D "plane" dominate
// (BASELINE config 3 at 1024^2: 0.30 of the HBM-equivalent roofline, CG 0.27).  Here the field never leaves
// the SM between steps:
//   * CTA b (one per SM) owns R consecutive rows of the (n0, n2) grid and keeps them, with one halo row above and
//     below, in shared memory for the whole launch;
//   * the rows a neighbour needs -- the CTA's FIRST and LAST row -- are stored into a global exchange buffer as LL
//     lines: every 8-byte word carries 4 bytes of data and a 4-byte sequence number (the scheme of NCCL's LL
//     protocol; aligned 8-byte accesses are single transactions).  The consumer -- the neighbour's thread that
//     computes the same column of ITS boundary row in the next step -- loads the line into its halo row and retries
//     until both sequence numbers are the ones it expects.  No flag word, no fence, no polling lane, no barrier
//     between CTAs: the dependency is per column.  Lines are requested at the start of a step without waiting and
//     looked at when the boundary rows are due (after 2/5 of the interior), so that they travel while both CTAs
//     compute.  (First version: red.release per warp / a publishing lane + bulk async copies into the halo rows --
//     5.3 / 4.2 us per step at 1024^2 against 2 us of arithmetic; the gpu-scope fences and the flag polls sat on the
//     step's critical path.)
//   * CG keeps x, r and d resident, exchanges d's boundary rows the same way, and sums its two dot products by an LL
//     all-gather into per-CTA inboxes (res_allsum), added in slot order -- every CTA holds bit-identical Krylov
//     scalars and runs the scalar stage (finalize_stage) redundantly, as k_cg_persistent does.  x_new is streamed
//     to the global ping-pong buffers every iteration, so that on exit they hold the last two iterates exactly
//     like the fused kernels leave them (the loser is the reference's VARo).
//   * two code paths per kernel: the uniform-coefficient fast path (ResFast: every face Dirichlet => the coefficient
//     classes are bitwise equal => every region cell is the class-0 stencil; thread-fixed columns, two rows per trip)
//     and the general item loop (any coefficients, any row length).
// Arithmetic: the star engine's operation order (star_cells_eq, FLAT), one rounding per reference operation ->
// Euler is bit-identical to the streaming path; CG differs only in the summation order of the dot products
// (as between any two kernel variants).
// Preconditions (host): 2-D mesh, constant-coefficient star operators, every face Dirichlet (static shell: the
// boundary cells never change after the first BC application), single GPU, cooperative launch available
// (co-residency is what makes waiting on another CTA legal).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
// Watchdog of the waits: a line that has not arrived after kResTimeoutNs (a CTA that died, a bug) must not leave the
// GPU spinning for ever.  The waiter raises g_res_abort, every other wait sees it within ~100 us and gives up as
// well, the kernels leave their loops, and the host reports PA_ERR_CUDA (res_check_abort).  The word lives in the
// translation unit that instantiates the kernels (tu_resident.cu); one resident launch runs at a time (ResBuf::last).
constexpr unsigned long long kResTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
static __device__ unsigned int g_res_abort;
__device__ __forceinline__ bool res_aborted() { return *(volatile unsigned int*)&g_res_abort != 0u; }

// N consecutive lines: all loads are issued before the first check
template <typename T, int N>
__device__ __forceinline__ void ll_load(const typename LLOf<T>::line* p, unsigned seq, T (&v)[N]) {
  unsigned ns = 32u, spins = 0u;
  unsigned long long t0 = 0ull;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int e = 0; e < N; ++e) ok &= ll_try(p + e, seq, v[e]);
    if (ok) break;
    // back off (32 .. 256 ns): hundreds of waiting threads re-reading at full rate only load the L2 slices the
    // awaited stores have to get through
    __nanosleep(ns);
    if (ns < 256u) ns <<= 1;
    if ((++spins & 255u) == 0u) {  // every ~65 us of waiting
      if (res_aborted()) break;
      const unsigned long long now = global_timer_ns();
      if (t0 == 0ull) {
        t0 = now;
      } else if (now - t0 > kResTimeoutNs) {
        *(volatile unsigned int*)&g_res_abort = 1u;
        __threadfence();
        break;
      }
    }
  }
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
  int dq, dr;       // kResThreads / nv, kResThreads % nv  (incremental (position, vector) walk)
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
    return base + (((int)(parity & 1u) * ctas + cta) * 2 + side) * n2;  // (< 2^31 lines, res_plan)
  }
};

// ---- the thread's work items of one pass --------------------------------------------------------------
// ONE loop body serves every row of a pass -- the kernels' code must stay small: a first version with separate
// code for boundary rows, a two-row register march for the interior and its remainders was > 64 KB of SASS per
// kernel, every piece of it executed once per step, and ran at ~1 us per 1024-cell row (instruction fetch) against
// 0.24 us of fp64 issue.
// A pass visits the CTA's rows in positions 0 .. rows-1; the two boundary rows (first and last owned row) sit at
// positions rs-1 and rs: rs = 1 puts them first, rs = rows-1 last, anything between in the middle of the pass.
// Item i = tid + k * 512 is vector (i % nv) of the row at position (i / nv).
struct ResIt {
  int j, cv;
};
__device__ __forceinline__ ResIt res_it_first(const ResCtx& c) {
  ResIt a;
  a.j = (int)threadIdx.x / c.nv;
  a.cv = (int)threadIdx.x - a.j * c.nv;
  return a;
}
__device__ __forceinline__ ResIt res_it_next(const ResCtx& c, ResIt a) {
  a.j += c.dq;
  a.cv += c.dr;
  if (a.cv >= c.nv) {
    a.cv -= c.nv;
    ++a.j;
  }
  return a;
}
__device__ __forceinline__ int res_row_at(const ResCtx& c, int rs, int j) {
  if (c.rows < 2) return 0;
  return j < rs - 1 ? j + 1 : (j == rs - 1 ? 0 : (j == rs ? c.rows - 1 : j - 1));
}
// neighbour rows of local row `row` that live in another CTA: bit 0 = the row above, bit 1 = the row below
__device__ __forceinline__ unsigned res_halo_dirs(const ResCtx& c, int row) {
  unsigned m = 0u;
  if (row == 0 && blockIdx.x > 0) m |= 1u;
  if (row == c.rows - 1 && blockIdx.x + 1 < gridDim.x) m |= 2u;
  return m;
}
// One halo vector: from the neighbour's LL row (sequence number seq; seq == 0: from the global array g0, the state
// before the first exchange) into the halo row of the resident buffer (`buf0` = its row 0; halo rows at -1 and at
// c.rows).  Non-blocking calls return false when the line has not arrived yet.  The halo cell (side, column) is
// private to the thread that computes the boundary vector of that column: no barrier between this store and the
// thread's own later loads.
template <typename T>
__device__ __forceinline__ bool res_halo_get(const ResCtx& c, const ResLL<T>& ll, T* buf0, int col, int below,
                                             unsigned parity, unsigned seq, const T* g0, bool blocking) {
  constexpr int VEC = VecOf<T>::N;
  T v[VEC];
  T* dst = buf0 + (below ? c.rows : -1) * c.n2 + col;
  if (seq == 0u) {
    ldcg_vec<T>(g0 + (c.row0 + (below ? c.rows : -1)) * c.n2 + col, v);
  } else {
    const typename LLOf<T>::line* src = ll.row(parity, (int)blockIdx.x + (below ? 1 : -1), below ? 0 : 1) + col;
    if (blocking) {
      ll_load<T, VEC>(src, seq, v);
    } else {
      bool ok = true;
#pragma unroll
      for (int e = 0; e < VEC; ++e) ok &= ll_try(src + e, seq, v[e]);
      if (!ok) return false;
    }
  }
  sts_vec<T>(dst, v);
  return true;
}
// Request the halo vectors of the thread's boundary items early, without waiting (`got`: one bit per halo vector in
// item order).  What had not arrived is fetched again, blocking, when the item is computed (res_halo_need).
template <typename T>
__device__ __forceinline__ unsigned res_halo_prefetch(const ResCtx& c, const ResLL<T>& ll, T* buf0, int rs,
                                                      unsigned parity, unsigned seq, const T* g0) {
  unsigned got = 0u;
  int bit = 0;
  // the thread's items at the two boundary positions rs-1, rs (one position if the CTA owns a single row), in item
  // order -- the order in which the pass meets them
  const int lo = c.rows < 2 ? 0 : (rs - 1) * c.nv, hi = c.rows < 2 ? c.nv : (rs + 1) * c.nv;
  for (int i = lo + (((int)threadIdx.x - lo) & (kResThreads - 1)); i < hi; i += kResThreads) {
    const int j = c.rows < 2 ? 0 : (i < rs * c.nv ? rs - 1 : rs);
    const int cv = i - j * c.nv;
    const unsigned dirs = res_halo_dirs(c, res_row_at(c, rs, j));
#pragma unroll
    for (int b = 0; b < 2; ++b)
      if ((dirs >> b) & 1u) {
        if (res_halo_get<T>(c, ll, buf0, cv * VecOf<T>::N, b, parity, seq, g0, false)) got |= 1u << (bit & 31);
        ++bit;
      }
  }
  return got;
}
template <typename T>
__device__ __forceinline__ void res_halo_need(const ResCtx& c, const ResLL<T>& ll, T* buf0, int row, int col,
                                              unsigned parity, unsigned seq, const T* g0, unsigned got, int& bit) {
  if (row != 0 && row != c.rows - 1) return;
  const unsigned dirs = res_halo_dirs(c, row);
#pragma unroll
  for (int b = 0; b < 2; ++b)
    if ((dirs >> b) & 1u) {
      if (!((got >> (bit & 31)) & 1u)) res_halo_get<T>(c, ll, buf0, col, b, parity, seq, g0, true);
      ++bit;
    }
}
// a vector of a boundary row goes to the neighbour(s) that read it
template <typename T>
__device__ __forceinline__ void res_send(const ResCtx& c, const ResLL<T>& ll, int row, int col, unsigned parity,
                                         unsigned seq, const T (&v)[VecOf<T>::N]) {
  if (row == 0 && blockIdx.x > 0) ll_store_vec<T, VecOf<T>::N>(ll.row(parity, blockIdx.x, 0) + col, seq, v);
  if (row == c.rows - 1 && blockIdx.x + 1 < gridDim.x)
    ll_store_vec<T, VecOf<T>::N>(ll.row(parity, blockIdx.x, 1) + col, seq, v);
}

// ---- the operator ---------------------------------------------------------------------------------------
// The operator sum at one cell of a 2-D grid: star_cells_eq (kernels_tma_pw.cuh) with K::FLAT, same operation
// order; the CG form (one operator + the implicit-Euler shift, star_cells of kernels_tma.cuh) is the same code
// with nops == 1.
template <typename T, int NOPS>
struct ResScales {
  OpScale<T> sc[NOPS > 0 ? NOPS : kMaxOps];
  __device__ __forceinline__ explicit ResScales(const EqDev<T>& eq) {
#pragma unroll
    for (int q = 0; q < (NOPS > 0 ? NOPS : kMaxOps); ++q)
      if (q < (NOPS > 0 ? NOPS : eq.nops)) sc[q] = op_scale<T>(eq.op[q]);
  }
};
// SHIFT = false: the equation has no implicit-Euler shift (explicit Euler), the test is compiled out.
// The scale factor is applied unconditionally: when it is not "used" it is exactly 1 and x * 1 == x bit for bit, and a
// multiplication is cheaper than the two selects per value a run-time `use` costs (ncu: 139 FSEL per warp and step).
template <typename T, bool LEAN, int NOPS, bool SHIFT = true>
__device__ __forceinline__ T res_star(const EqDev<T>& eq, const ResScales<T, NOPS>& scs, int clx, int cz, T v0, T xp,
                                      T xm, T zp, T zm) {
  constexpr int MAXO = NOPS > 0 ? NOPS : kMaxOps;
  const int nops = NOPS > 0 ? NOPS : eq.nops;
  const int ix = LEAN ? 0 : clx, iz = LEAN ? 0 : cz;
  T res = (T)0;
#pragma unroll
  for (int q = 0; q < MAXO; ++q) {
    if (q >= nops) break;
    const OpDev<T>& o = eq.op[q];
    const OpScale<T>& sc = scs.sc[q];
    T s = o.coef[0][ix][0] * xp;
    s = s + o.coef[0][ix][1] * v0;
    s = s + o.coef[0][ix][2] * xm;
    T acc = s;
    T s2 = o.coef[2][iz][0] * zp;
    s2 = s2 + o.coef[2][iz][1] * v0;
    s2 = s2 + o.coef[2][iz][2] * zm;
    acc = acc + s2;
    acc = acc * sc.scale;
    res = res + acc;
    if (SHIFT && o.has_shift) {
      const T m = o.shift * v0;
      res = m + res;
    }
  }
  return res;
}
// boundary-adjacent cells (coefficient classes != 0, cells outside the region): two rows and two columns of the grid.
// (Inline: as a real call it forced the vectors of EVERY item through local memory, 3 STL per item on the hot path;
// cold code costs no instruction fetches.)
template <typename T, int NOPS>
__device__ __forceinline__ void res_apply_general(const GridDev& g, const EqDev<T>& eq, const ResScales<T, NOPS>& scs,
                                                  int clx, int col, const T* v0,
                                               const T* vm, const T* vp, T zl, T zr, T* ad, unsigned* inmask) {
  constexpr int VEC = VecOf<T>::N;
  unsigned m = 0u;
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const int z = col + e;
    const bool in = z >= g.lo[2] && z < g.hi[2];
    const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
    const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
    T a = (T)0;
    if (in) {
      a = res_star<T, false, NOPS>(eq, scs, clx, coef_class(g, 2, z), v0[e], vp[e], vm[e], zp, zm);
      m |= 1u << e;
    }
    ad[e] = a;
  }
  *inmask = m;
}
// A(phi) on one vector whose row is inside the region: ad[e] and the mask of region cells.  `p` points at the vector
// inside a resident buffer (rows n2 apart, the rows above and below in place -- halo rows included).
template <typename T, int NOPS>
__device__ __forceinline__ unsigned res_apply_vec(const GridDev& g, const EqDev<T>& eq, const ResScales<T, NOPS>& scs,
                                                  const T* p, int n2, int grow, int col,
                                                  const T (&v0)[VecOf<T>::N], T (&ad)[VecOf<T>::N]) {
  constexpr int VEC = VecOf<T>::N;
  T vm[VEC], vp[VEC];
  lds_vec<T>(p - n2, vm);
  lds_vec<T>(p + n2, vp);
  const T zl = col > 0 ? p[-1] : (T)0;
  const T zr = col + VEC < n2 ? p[VEC] : (T)0;
  const int clx = coef_class(g, 0, grow);
  const int lo2 = g.lo[2] > 2 ? g.lo[2] : 2, hi2 = g.hi[2] < n2 - 2 ? g.hi[2] : n2 - 2;
  if (clx == 0 && col >= lo2 && col + VEC <= hi2) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
      ad[e] = res_star<T, true, NOPS>(eq, scs, 0, 0, v0[e], vp[e], vm[e], zp, zm);
    }
    return (1u << VEC) - 1u;
  }
  unsigned m;
  res_apply_general<T, NOPS>(g, eq, scs, clx, col, v0, vm, vp, zl, zr, ad, &m);
  return m;
}

// ---- the uniform-coefficient fast path ----------------------------------------------------------------
// ncu on the item loop alone (1024^2): 2084 instructions per warp and step for 7 items of 58 fp64 instructions each,
// 60 % issue-active, fp64 pipe 28 % -- the walk, the row mapping, the region / class tests and the addressing of the
// general item cost four times the arithmetic, and one warp alone issues only every 4-6 cycles: a step is as slow as
// its longest warp.  (A march over the class-0 interior only made it worse: the two warps that own the columns next to
// the walls kept all their rows on the item loop and everybody waited for them.)
// When the three coefficient classes of every operator hold the same numbers on both axes (UNI: no Neumann / Symmetry
// face -- always the case here, every face is Dirichlet) no cell needs a class: a thread keeps ONE column block,
// walks down its rows with the rows above / at / below in registers, two rows per trip, and runs nothing but loads,
// stencil, a region select for the wall columns, and the store.  The CTA's first and last row are one more such trip,
// fed from the halo rows.  All warps do the same work.
// Mapping: nv a multiple of 512: every thread takes columns tid, tid + 512, ... and all rows; 512 a multiple of nv:
// G = 512 / nv thread groups share the rows in contiguous runs, the last group takes the boundary rows as two of its
// rows.  Other row lengths (and non-uniform coefficients) stay on the item loop.
struct ResFast {
  bool on;
  bool boundary;   // this thread computes the CTA's first and last row (for its columns)
  int ta, tb;      // this thread's interior rows [ta, tb)
  int cv, cstep;   // this thread's first column block (vectors) and the step to its next one
};
template <typename T>
__device__ __forceinline__ ResFast res_fast_init(const ResCtx& c, const GridDev& g, bool uni) {
  ResFast f;
  f.on = false;
  f.boundary = false;
  f.ta = f.tb = f.cv = 0;
  f.cstep = 1;
  // interior rows must all be region rows (every face Dirichlet: the region is the grid minus its shell)
  if (!uni || g.lo[0] > 1 || g.hi[0] < g.n[0] - 1) return f;
  const int nint = c.rows > 2 ? c.rows - 2 : 0;
  if (c.nv % kResThreads == 0) {
    f.on = true;
    f.boundary = true;
    f.cv = threadIdx.x;
    f.cstep = kResThreads;
    f.ta = 1;
    f.tb = 1 + nint;
  } else if (kResThreads % c.nv == 0) {
    f.on = true;
    const int G = kResThreads / c.nv, grp = (int)threadIdx.x / c.nv;
    const int chunk = (nint + 2 + G - 1) / G;  // rows per group, the boundary rows counted as two
    f.cv = (int)threadIdx.x - grp * c.nv;
    f.cstep = c.nv;
    f.ta = min(1 + grp * chunk, 1 + nint);
    f.tb = grp == G - 1 ? 1 + nint : min(f.ta + chunk, 1 + nint);
    f.boundary = grp == G - 1;
  }
  return f;
}
// mask of the vector's cells that lie inside the region along the contiguous axis (all ones but for the two wall columns)
template <typename T>
__device__ __forceinline__ unsigned res_col_mask(const GridDev& g, int col) {
  unsigned m = 0u;
#pragma unroll
  for (int e = 0; e < VecOf<T>::N; ++e)
    if (col + e >= g.lo[2] && col + e < g.hi[2]) m |= 1u << e;
  return m;
}
// rows [ra, rb) of the thread's column blocks, two rows per trip (both rows are computed before either is stored:
// the compiler must assume that `cur` and `nxt` alias):
//   cell(v0, vm, vp, zl, zr, local row, col, column mask, row inside the region) -> o,   put(row, col, o, mask)
//   (mask = the vector's region cells: the column mask, or 0 for a row outside the region)
// (interior rows of a CTA are region rows: res_fast_init)
template <typename T, typename FC, typename FP>
__device__ __forceinline__ void res_fast_march(const ResCtx& c, const ResFast& f, const GridDev& g, const T* cur, int ra,
                                               int rb, FC cell, FP put) {
  constexpr int VEC = VecOf<T>::N;
  if (rb <= ra) return;
  const int n2 = c.n2;
  for (int cv = f.cv; cv < c.nv; cv += f.cstep) {
    const int col = cv * VEC;
    const unsigned cm = res_col_mask<T>(g, col);
    const T* p = cur + ra * n2 + col;
    T vm[VEC], v0[VEC];
    lds_vec<T>(p - n2, vm);
    lds_vec<T>(p, v0);
    int lr = ra;
    for (; lr + 1 < rb; lr += 2) {
      T v1[VEC], v2[VEC], oa[VEC], ob[VEC];
      lds_vec<T>(p + n2, v1);
      lds_vec<T>(p + 2 * n2, v2);
      const T zla = p[-1], zra = p[VEC], zlb = p[n2 - 1], zrb = p[n2 + VEC];
      cell(v0, vm, v1, zla, zra, lr, col, cm, true, oa);
      cell(v1, v0, v2, zlb, zrb, lr + 1, col, cm, true, ob);
      put(lr, col, oa, cm);
      put(lr + 1, col, ob, cm);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        vm[e] = v1[e];
        v0[e] = v2[e];
      }
      p += 2 * n2;
    }
    if (lr < rb) {
      T v1[VEC], oa[VEC];
      lds_vec<T>(p + n2, v1);
      const T zla = p[-1], zra = p[VEC];
      cell(v0, vm, v1, zla, zra, lr, col, cm, true, oa);
      put(lr, col, oa, cm);
    }
  }
}
// the CTA's first and last row (one row if it owns a single one) for the thread's column blocks, as ONE trip; the
// halo rows of `cur` are in place (res_halo_need has run for these items)
template <typename T, typename FC, typename FP>
__device__ __forceinline__ void res_fast_boundary(const ResCtx& c, const GridDev& g, const T* cur, int col, FC cell,
                                                  FP put) {
  constexpr int VEC = VecOf<T>::N;
  const int n2 = c.n2;
  const unsigned cm = res_col_mask<T>(g, col);
  const bool ra_in = c.row0 >= g.lo[0] && c.row0 < g.hi[0];
  const bool rb_in = c.row0 + c.rows - 1 >= g.lo[0] && c.row0 + c.rows - 1 < g.hi[0];
  const T* pa = cur + col;
  T a0[VEC], am[VEC], ap[VEC], oa[VEC];
  lds_vec<T>(pa, a0);
  lds_vec<T>(pa - n2, am);
  lds_vec<T>(pa + n2, ap);
  const T zla = pa[-1], zra = pa[VEC];
  if (c.rows < 2) {
    cell(a0, am, ap, zla, zra, 0, col, cm, ra_in, oa);
    put(0, col, oa, ra_in ? cm : 0u);
    return;
  }
  const T* pb = cur + (c.rows - 1) * n2 + col;
  T b0[VEC], bm[VEC], bp[VEC], ob[VEC];
  lds_vec<T>(pb, b0);
  lds_vec<T>(pb - n2, bm);
  lds_vec<T>(pb + n2, bp);
  const T zlb = pb[-1], zrb = pb[VEC];
  cell(a0, am, ap, zla, zra, 0, col, cm, ra_in, oa);
  cell(b0, bm, bp, zlb, zrb, c.rows - 1, col, cm, rb_in, ob);
  put(0, col, oa, ra_in ? cm : 0u);
  put(c.rows - 1, col, ob, rb_in ? cm : 0u);
}

// =========================================================================================
// explicit Euler: nsteps steps of  phi <- phi + dt (rhs - A(phi))  on the region
// b0 holds phi on entry; step s reads buffer s&1 and writes (s+1)&1; the last TWO steps store every row to
// global memory, so that on exit b[nsteps&1] is the result and the other array the step before it (VARo).
// Sequence numbers seq0+1 .. seq0+nsteps are this launch's (never reused on the same exchange buffer).
// The boundary rows sit in the MIDDLE of a step: the neighbours' lines (stored in the middle of their previous
// step) are requested first, travel while the first part of the interior is computed, and this step's boundary
// rows are on their way while the second part is.
// =========================================================================================
template <typename T, int NOPS, bool HAS_RHS>
__global__ void __launch_bounds__(kResThreads, 1)
k_euler_resident(GridDev g, EqDev<T> eq, T* __restrict__ b0,
                 T* __restrict__ b1, const T* __restrict__ rhs, T dt, int nsteps, int R,
                 typename LLOf<T>::line* llbase, unsigned seq0, int uni, unsigned long long* dbg) {
  constexpr int VEC = VecOf<T>::N;
  // PA_RES_DEBUG: time stamps of one step, CTA gridDim/2, thread 0 (null in normal runs)
  auto stamp = [&](int s, int k) {
    if (dbg != nullptr && s == 8 && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) dbg[k] = global_timer_ns();
  };
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  ResCtx c;
  res_ctx_init<T>(c, g, R);
  const ResLL<T> ll{llbase, c.n2, (int)gridDim.x};
  const int rowlen = c.n2;  // (a resident grid has < 2^31 cells: 32-bit offsets everywhere)
  // two buffers of R + 2 rows; sb* point at row 0, the halo rows are rows -1 and c.rows
  T* const sb0 = reinterpret_cast<T*>(base) + rowlen;
  T* const sb1 = sb0 + (R + 2) * rowlen;

  for (int i = threadIdx.x; i < (c.rows + 2) * c.nv; i += kResThreads) {  // resident copy, halo rows zeroed
    const int lr = i / c.nv - 1, cv = i - (lr + 1) * c.nv;
    T v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = (T)0;
    sts_vec<T>(sb1 + lr * rowlen + cv * VEC, v);
    if (lr >= 0 && lr < c.rows) ldcg_vec<T>(b0 + (c.row0 + lr) * rowlen + cv * VEC, v);
    sts_vec<T>(sb0 + lr * rowlen + cv * VEC, v);
  }
  __syncthreads();

  const ResScales<T, NOPS> scs(eq);
  const ResFast fast = res_fast_init<T>(c, g, uni != 0);
  // item loop: the boundary rows in the middle of the pass; fast path: the thread's first 2/5 of its rows, its boundary
  // trip, the rest of its rows
  const int rs = c.rows < 2 ? 1 : 1 + ((c.rows - 2) * 2) / 5;
  const int fsplit = fast.ta + ((fast.tb - fast.ta) * 2) / 5;
  for (int s = 0; s < nsteps; ++s) {
    // (watchdog: the host reports the failure.  Looked at every 256 steps only -- the load is an L2 round trip on the
    //  step's critical path; once the word is raised every wait gives up within ~65 us anyway)
    if ((s & 255) == 255 && res_aborted()) break;
    T* cur = (s & 1) ? sb1 : sb0;
    T* nxt = (s & 1) ? sb0 : sb1;
    T* gout = (s & 1) ? b0 : b1;
    const bool all_rows = s + 2 >= nsteps, send = s + 1 < nsteps;
    const unsigned seq_in = s == 0 ? 0u : seq0 + (unsigned)s, seq_out = seq0 + (unsigned)s + 1u;
    stamp(s, 0);
    if (fast.on) {
      // ---- uniform coefficients: every cell of the region is the class-0 stencil ------------------------
      unsigned got = 0u;
      if (fast.boundary) {  // request the neighbours' lines for my columns (bit 2k: above, 2k+1: below, block k)
        int k = 0;
        for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
          if (blockIdx.x > 0 && res_halo_get<T>(c, ll, cur, cv * VEC, 0, (unsigned)s, seq_in, b0, false))
            got |= 1u << (2 * k);
          if (blockIdx.x + 1 < gridDim.x && res_halo_get<T>(c, ll, cur, cv * VEC, 1, (unsigned)s, seq_in, b0, false))
            got |= 2u << (2 * k);
        }
      }
      auto cell = [&](const T (&v0)[VEC], const T (&vm)[VEC], const T (&vp)[VEC], T zl, T zr, int row, int col,
                      unsigned cm, bool rin, T (&o)[VEC]) {
        T av[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) av[e] = (T)0;
        if (HAS_RHS) {
          typedef typename VecOf<T>::type V;
          V q = __ldg(reinterpret_cast<const V*>(rhs + (c.row0 + row) * rowlen + col));
          const T* qs = reinterpret_cast<const T*>(&q);
#pragma unroll
          for (int e = 0; e < VEC; ++e) av[e] = qs[e];
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
          const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
          const T a = res_star<T, true, NOPS, false>(eq, scs, 0, 0, v0[e], vp[e], vm[e], zp, zm);
          const T res = av[e] - a;
          o[e] = v0[e] + dt * res;
        }
        // the wall columns / rows keep their values: a branch that all but two warps of a CTA never take
        if (!rin || cm != (1u << VEC) - 1u) {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (!rin || !((cm >> e) & 1u)) o[e] = v0[e];
        }
      };
      auto put = [&](int row, int col, const T (&o)[VEC], unsigned) {
        sts_vec<T>(nxt + row * rowlen + col, o);
        if (all_rows) sts_vec<T>(gout + (c.row0 + row) * rowlen + col, o);
      };
      auto put_boundary = [&](int row, int col, const T (&o)[VEC], unsigned m) {
        put(row, col, o, m);
        if (send) res_send<T>(c, ll, row, col, (unsigned)(s + 1), seq_out, o);
      };
      for (int seg = 0; seg < 2; ++seg) {  // ONE instance of the march
        res_fast_march<T>(c, fast, g, cur, seg == 0 ? fast.ta : fsplit, seg == 0 ? fsplit : fast.tb, cell, put);
        if (seg == 0 && fast.boundary) {
          int k = 0;
          for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
            const int col = cv * VEC;
            if (blockIdx.x > 0 && !((got >> (2 * k)) & 1u))
              res_halo_get<T>(c, ll, cur, col, 0, (unsigned)s, seq_in, b0, true);
            if (blockIdx.x + 1 < gridDim.x && !((got >> (2 * k + 1)) & 1u))
              res_halo_get<T>(c, ll, cur, col, 1, (unsigned)s, seq_in, b0, true);
            res_fast_boundary<T>(c, g, cur, col, cell, put_boundary);
          }
        }
      }
    } else {
      // ---- any coefficients, any row length: the item loop ------------------------------------------------
      const unsigned got = res_halo_prefetch<T>(c, ll, cur, rs, (unsigned)s, seq_in, b0);
      int bit = 0;
      // one item: the new values of a vector
      auto compute = [&](ResIt a, int& row, int& col, T (&o)[VEC]) {
        row = res_row_at(c, rs, a.j);
        col = a.cv * VEC;
        res_halo_need<T>(c, ll, cur, row, col, (unsigned)s, seq_in, b0, got, bit);
        const int grow = c.row0 + row;
        const T* p = cur + row * rowlen + col;
        T v0[VEC];
        lds_vec<T>(p, v0);
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = v0[e];
        if (grow >= g.lo[0] && grow < g.hi[0]) {
          T av[VEC], ad[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) av[e] = (T)0;
          if (HAS_RHS) {
            typedef typename VecOf<T>::type V;
            V q = __ldg(reinterpret_cast<const V*>(rhs + grow * rowlen + col));
            const T* qs = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int e = 0; e < VEC; ++e) av[e] = qs[e];
          }
          const unsigned in = res_apply_vec<T, NOPS>(g, eq, scs, p, c.n2, grow, col, v0, ad);
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if ((in >> e) & 1u) {
              const T res = av[e] - ad[e];
              o[e] = v0[e] + dt * res;
            }
        }
      };
      auto emit = [&](int row, int col, const T (&o)[VEC]) {
        sts_vec<T>(nxt + row * rowlen + col, o);
        if (send && (row == 0 || row == c.rows - 1)) res_send<T>(c, ll, row, col, (unsigned)(s + 1), seq_out, o);
        if (all_rows) sts_vec<T>(gout + (c.row0 + row) * rowlen + col, o);
      };
      // two items per trip: both are computed before either is stored
      for (ResIt a = res_it_first(c); a.j < c.rows;) {
        const ResIt b = res_it_next(c, a);
        int ra, ca, rb = 0, cb = 0;
        T oa[VEC], ob[VEC];
        compute(a, ra, ca, oa);
        const bool two = b.j < c.rows;
        if (two) compute(b, rb, cb, ob);
        emit(ra, ca, oa);
        if (two) emit(rb, cb, ob);
        a = res_it_next(c, b);
      }
    }
    stamp(s, 1);
    __syncthreads();  // end of step s: `cur` is free, `nxt` complete (but for its halo rows: next step's fetches)
    stamp(s, 2);
    if (dbg != nullptr && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0 && (s == 8 || s == nsteps - 1))
      dbg[s == 8 ? 3 : 4] = global_timer_ns();
  }
}

// =========================================================================================
// CG, the whole solve.  On entry: xa = x after the BC application (xb holds the same shell), r = rhs - A(x) on the
// region and 0 elsewhere, d = r (PW_RESID writes both), st = k_state_init + ST_CG_INIT (rr in sum[R_RR]).
// Sequence numbers: d rows of iteration k carry seq0 + k + 1, the all-reduce of epoch e carries seq0 + e.
// =========================================================================================
// All-CTA sum as an LL all-gather into private inboxes: CTA b stores its partial sums into line [parity][dst][b] of
// EVERY CTA dst and polls only its own inbox, so that a line has exactly one writer and one reader.  (First version:
// one slot per CTA, read back by all 148 CTAs -- 148 x 148 polling loads on 37 cache lines; measured 3.4-8.2 us per
// all-reduce at 256^2, now 2.1-2.5.)  Every CTA adds the contributions in the same order: bit-identical sums.
template <int NS>
__device__ __forceinline__ void res_allsum(double (&v)[NS], uint4* inbox, unsigned epoch, unsigned seq, double* red,
                                           double* bc) {
  block_sum<NS>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) bc[s] = v[s];
  }
  __syncthreads();
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = 0.0;
  if (threadIdx.x < gridDim.x) {
    uint4* par = inbox + (long long)(epoch & 1u) * kResMaxCtas * kResMaxCtas * kResSlotLines;
    uint4* dst = par + ((long long)threadIdx.x * kResMaxCtas + blockIdx.x) * kResSlotLines;
#pragma unroll
    for (int s = 0; s < NS; ++s) ll_store(dst + s, bc[s], seq);
    ll_load<double, NS>(par + ((long long)blockIdx.x * kResMaxCtas + threadIdx.x) * kResSlotLines, seq, acc);
  }
  __syncthreads();  // bc has been read by every sender
  block_sum<NS>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) bc[s] = acc[s];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NS; ++s) v[s] = bc[s];
  __syncthreads();  // bc is free for the next reduction
}

template <typename T>
__global__ void __launch_bounds__(kResThreads, 1)
k_cg_resident(GridDev g, EqDev<T> eq, T* __restrict__ xa,
              T* __restrict__ xb, const T* __restrict__ r_in, const T* __restrict__ d_in, SolverState* st, int R,
              typename LLOf<T>::line* llbase, uint4* inbox, unsigned seq0, int uni, unsigned long long* dbg) {
  constexpr int VEC = VecOf<T>::N;
  auto stamp = [&](unsigned i, int k) {
    if (dbg != nullptr && i == 8u && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) dbg[k] = global_timer_ns();
  };
  extern __shared__ unsigned char smem_dyn[];
  __shared__ SolverState ls;  // this CTA's copy of the solver state (bit-identical in every CTA)
  __shared__ double red[2 * 32];
  __shared__ double bc[2];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  ResCtx c;
  res_ctx_init<T>(c, g, R);
  const ResLL<T> ll{llbase, c.n2, (int)gridDim.x};
  const int rowlen = c.n2;
  T* sd = reinterpret_cast<T*>(base) + rowlen;  // d: rows -1 .. R (halo rows at -1 and c.rows)
  T* sr = sd + (R + 1) * rowlen;     // r: rows 0 .. R-1
  T* sx = sr + R * rowlen;           // x: rows 0 .. R-1

  if (threadIdx.x == 0) ls = *st;
  for (int i = threadIdx.x; i < (c.rows + 2) * c.nv; i += kResThreads) {
    const int lr = i / c.nv - 1, cv = i - (lr + 1) * c.nv;
    const long long go = (c.row0 + lr) * rowlen + cv * VEC, so = lr * rowlen + cv * VEC;
    T v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = (T)0;
    if (lr < 0 || lr >= c.rows) {
      sts_vec<T>(sd + so, v);
      continue;
    }
    ldcg_vec<T>(d_in + go, v);
    sts_vec<T>(sd + so, v);
    ldcg_vec<T>(r_in + go, v);
    sts_vec<T>(sr + so, v);
    ldcg_vec<T>(xa + go, v);
    sts_vec<T>(sx + so, v);
  }
  __syncthreads();

  const ResScales<T, 1> scs(eq);
  const ResFast fast = res_fast_init<T>(c, g, uni != 0);
  unsigned epoch = 1u;
  unsigned it = 0;
  while (!ls.done) {
    if ((it & 63u) == 63u && res_aborted()) break;  // (watchdog, every 64 iterations: see k_euler_resident)
    const unsigned seq_d = seq0 + it + 1u;
    if (dbg != nullptr && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0 && (it == 8u || it == 908u))
      dbg[it == 8u ? 10 : 11] = global_timer_ns();
    // ---- phase A: d = r + beta d on the region (linalg.py:141), boundary rows first ---------------------
    const T beta = (T)ls.scal[S_BETA];
    stamp(it, 0);
    auto dupd = [&](int row, int col) {
      const int grow = c.row0 + row;
      T* dp = sd + row * rowlen + col;
      T dv[VEC], rv[VEC];
      lds_vec<T>(dp, dv);
      if (grow >= g.lo[0] && grow < g.hi[0]) {
        lds_vec<T>(sr + row * rowlen + col, rv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int z = col + e;
          if (z >= g.lo[2] && z < g.hi[2]) dv[e] = rv[e] + beta * dv[e];
        }
        sts_vec<T>(dp, dv);
      }
      if (row == 0 || row == c.rows - 1) res_send<T>(c, ll, row, col, it, seq_d, dv);
    };
    if (fast.on) {
      constexpr unsigned FULL = (1u << VEC) - 1u;
      for (int cv = fast.cv; cv < c.nv; cv += fast.cstep) {
        const int col = cv * VEC;
        if (fast.boundary) {
          dupd(0, col);
          if (c.rows > 1) dupd(c.rows - 1, col);
        }
        // interior rows are region rows (res_fast_init): only the column mask is left to look at
        const unsigned cm = res_col_mask<T>(g, col);
        for (int row = fast.ta; row < fast.tb; ++row) {
          T* dp = sd + row * rowlen + col;
          T dv[VEC], rv[VEC];
          lds_vec<T>(dp, dv);
          lds_vec<T>(sr + row * rowlen + col, rv);
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (cm == FULL || ((cm >> e) & 1u)) dv[e] = rv[e] + beta * dv[e];
          sts_vec<T>(dp, dv);
        }
      }
    } else {
      for (ResIt a = res_it_first(c); a.j < c.rows; a = res_it_next(c, a)) dupd(res_row_at(c, 1, a.j), a.cv * VEC);
    }
    __syncthreads();  // every own row of d is updated
    stamp(it, 1);
    // A(d) and d.A(d): the interior rows first -- the neighbours' rows, requested here, are still in flight
    double qa[1] = {0.0};
    // A(d) on one item (0 outside the region)
    auto apply_d = [&](int row, int col, T (&dv)[VEC], T (&ad)[VEC]) -> unsigned {
      const int grow = c.row0 + row;
      const T* dp = sd + row * rowlen + col;
      lds_vec<T>(dp, dv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) ad[e] = (T)0;
      if (grow < g.lo[0] || grow >= g.hi[0]) return 0u;
      return res_apply_vec<T, 1>(g, eq, scs, dp, c.n2, grow, col, dv, ad);
    };
    // uniform coefficients: the class-0 stencil on a vector, 0 where the cell is outside the region
    auto fast_cell = [&](const T (&v0)[VEC], const T (&vm)[VEC], const T (&vp)[VEC], T zl, T zr, int row, int col,
                         unsigned cm, bool rin, T (&ad)[VEC]) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
        const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
        ad[e] = res_star<T, true, 1>(eq, scs, 0, 0, v0[e], vp[e], vm[e], zp, zm);
      }
      if (!rin || cm != (1u << VEC) - 1u) {
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (!rin || !((cm >> e) & 1u)) ad[e] = (T)0;
      }
    };
    auto dad_put = [&](int row, int col, const T (&ad)[VEC], unsigned m) {
      if (m == 0u) return;
      T dv[VEC];
      lds_vec<T>(sd + row * rowlen + col, dv);
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (m == (1u << VEC) - 1u || ((m >> e) & 1u)) {
          const T q = dv[e] * ad[e];
          qa[0] += (double)q;
        }
    };
    if (fast.on) {
      unsigned got = 0u;
      if (fast.boundary) {
        int k = 0;
        for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
          if (blockIdx.x > 0 && res_halo_get<T>(c, ll, sd, cv * VEC, 0, it, seq_d, (const T*)nullptr, false))
            got |= 1u << (2 * k);
          if (blockIdx.x + 1 < gridDim.x && res_halo_get<T>(c, ll, sd, cv * VEC, 1, it, seq_d, (const T*)nullptr, false))
            got |= 2u << (2 * k);
        }
      }
      res_fast_march<T>(c, fast, g, sd, fast.ta, fast.tb, fast_cell, dad_put);
      if (fast.boundary) {
        int k = 0;
        for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
          const int col = cv * VEC;
          if (blockIdx.x > 0 && !((got >> (2 * k)) & 1u))
            res_halo_get<T>(c, ll, sd, col, 0, it, seq_d, (const T*)nullptr, true);
          if (blockIdx.x + 1 < gridDim.x && !((got >> (2 * k + 1)) & 1u))
            res_halo_get<T>(c, ll, sd, col, 1, it, seq_d, (const T*)nullptr, true);
          res_fast_boundary<T>(c, g, sd, col, fast_cell, dad_put);
        }
      }
    } else {
      const int rs_last = c.rows < 2 ? 1 : c.rows - 1;
      const unsigned got = res_halo_prefetch<T>(c, ll, sd, rs_last, it, seq_d, (const T*)nullptr);
      int bit = 0;
      for (ResIt a = res_it_first(c); a.j < c.rows;) {
        const ResIt b = res_it_next(c, a);
        const bool two = b.j < c.rows;
        const int ra = res_row_at(c, rs_last, a.j), ca = a.cv * VEC;
        const int rb = two ? res_row_at(c, rs_last, b.j) : 0, cb = b.cv * VEC;
        res_halo_need<T>(c, ll, sd, ra, ca, it, seq_d, (const T*)nullptr, got, bit);
        if (two) res_halo_need<T>(c, ll, sd, rb, cb, it, seq_d, (const T*)nullptr, got, bit);
        T da[VEC], db[VEC], aa[VEC], ab[VEC];
        const unsigned ma = apply_d(ra, ca, da, aa);
        const unsigned mb = two ? apply_d(rb, cb, db, ab) : 0u;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if ((ma >> e) & 1u) {
            const T q = da[e] * aa[e];
            qa[0] += (double)q;
          }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if ((mb >> e) & 1u) {
            const T q = db[e] * ab[e];
            qa[0] += (double)q;
          }
        a = res_it_next(c, b);
      }
    }
    stamp(it, 2);
    res_allsum<1>(qa, inbox, epoch, seq0 + epoch, red, bc);
    stamp(it, 3);
    ++epoch;
    if (threadIdx.x == 0) {
      ls.sum[R_A] = qa[0];
      finalize_stage<T>(ST_CG_DAD, &ls);  // alpha = rr / dAd (linalg.py:114-120)
    }
    __syncthreads();
    stamp(it, 4);
    // ---- phase B: x_new = x + alpha d ; r -= alpha A(d) ; sums |r|^2, |dx|^2 (linalg.py:122-137) ------
    // (A(d) is recomputed -- the halo rows of d are still in place -- instead of kept in 14 registers per thread)
    const T alpha = (T)ls.scal[S_ALPHA];
    T* gxn = ((it + 1u) & 1u) ? xb : xa;
    double qb[2] = {0.0, 0.0};
    auto update = [&](int row, int col, unsigned m, const T (&dv)[VEC], const T (&ad)[VEC]) {
      if (m == 0u) return;
      const int grow = c.row0 + row;
      T* rp = sr + row * rowlen + col;
      T* xp = sx + row * rowlen + col;
      T rv[VEC], xv[VEC];
      lds_vec<T>(rp, rv);
      lds_vec<T>(xp, xv);
      const bool rshell = grow == 0 || grow == g.n[0] - 1;
      // (a vector whose cells are all region cells and that touches no wall: the common case, no per-cell tests)
      const bool plain = m == (1u << VEC) - 1u && !rshell && col > 0 && col + VEC < c.n2;
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (plain || ((m >> e) & 1u)) {
          const T xo = xv[e];
          const T xn = xo + alpha * dv[e];
          const T rn = rv[e] - alpha * ad[e];
          xv[e] = xn;
          rv[e] = rn;
          const T q = rn * rn;
          qb[0] += (double)q;
          const int z = col + e;
          if (plain || (!rshell && z != 0 && z != c.n2 - 1)) {
            const T df = xn - xo;
            const T q2 = df * df;
            qb[1] += (double)q2;
          }
        }
      sts_vec<T>(rp, rv);
      sts_vec<T>(xp, xv);
      sts_vec<T>(gxn + grow * rowlen + col, xv);
    };
    if (fast.on) {
      auto upd_put = [&](int row, int col, const T (&ad)[VEC], unsigned m) {
        if (m == 0u) return;
        T dv[VEC];
        lds_vec<T>(sd + row * rowlen + col, dv);
        update(row, col, m, dv, ad);
      };
      if (fast.boundary)
        for (int cv = fast.cv; cv < c.nv; cv += fast.cstep) res_fast_boundary<T>(c, g, sd, cv * VEC, fast_cell, upd_put);
      res_fast_march<T>(c, fast, g, sd, fast.ta, fast.tb, fast_cell, upd_put);
    } else {
      for (ResIt a = res_it_first(c); a.j < c.rows;) {
        const ResIt b = res_it_next(c, a);
        const bool two = b.j < c.rows;
        const int ra = a.j, ca = a.cv * VEC, rb = two ? b.j : 0, cb = b.cv * VEC;
        T da[VEC], db[VEC], aa[VEC], ab[VEC];
        const unsigned ma = apply_d(ra, ca, da, aa);
        const unsigned mb = two ? apply_d(rb, cb, db, ab) : 0u;
        update(ra, ca, ma, da, aa);
        update(rb, cb, mb, db, ab);
        a = res_it_next(c, b);
      }
    }
    stamp(it, 5);
    res_allsum<2>(qb, inbox, epoch, seq0 + epoch, red, bc);
    stamp(it, 6);
    ++epoch;
    if (threadIdx.x == 0) {
      ls.sum[R_A] = qb[0];
      ls.sum[R_B] = qb[1];
      ls.sum[R_SHELL] = 0.0;  // static shell: the boundary cells never change
      finalize_stage<T>(ST_CG_FIN, &ls);
    }
    __syncthreads();
    stamp(it, 7);
    ++it;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = ls;
}

// =========================================================================================
// Jacobi, the whole solve (uniform-coefficient fast path only; the host falls back to the streaming sweeps otherwise).
// x lives in two resident buffers like the Euler field; a sweep is an Euler-like pass
//     x_new = x + (rhs - A(x)) / diag     on the region            (oracle jacobi)
// followed by ONE all-reduce of |x_new - x|^2 and the scalar stage ST_JA_FIN in every CTA.  x_new is streamed to the
// global ping-pong buffers every sweep (the loser is VARo).  The quotient is div_rcp, the streaming sweep's: the
// correctly rounded one.  Sequence numbers: rows of sweep k carry seq0 + k + 1, the all-reduce of sweep k the same.
// =========================================================================================
template <typename T, int NOPS>
__global__ void __launch_bounds__(kResThreads, 1)
k_jacobi_resident(GridDev g, EqDev<T> eq, T* __restrict__ xa, T* __restrict__ xb, const T* __restrict__ rhs,
                  SolverState* st, int R, typename LLOf<T>::line* llbase, uint4* inbox, unsigned seq0) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ SolverState ls;
  __shared__ double red[32];
  __shared__ double bc[1];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  ResCtx c;
  res_ctx_init<T>(c, g, R);
  const ResLL<T> ll{llbase, c.n2, (int)gridDim.x};
  const int rowlen = c.n2;
  T* const sb0 = reinterpret_cast<T*>(base) + rowlen;
  T* const sb1 = sb0 + (R + 2) * rowlen;
  if (threadIdx.x == 0) ls = *st;
  for (int i = threadIdx.x; i < (c.rows + 2) * c.nv; i += kResThreads) {
    const int lr = i / c.nv - 1, cv = i - (lr + 1) * c.nv;
    T v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = (T)0;
    sts_vec<T>(sb1 + lr * rowlen + cv * VEC, v);
    if (lr >= 0 && lr < c.rows) ldcg_vec<T>(xa + (c.row0 + lr) * rowlen + cv * VEC, v);
    sts_vec<T>(sb0 + lr * rowlen + cv * VEC, v);
  }
  __syncthreads();

  const ResScales<T, NOPS> scs(eq);
  const ResFast fast = res_fast_init<T>(c, g, true);  // (the host launches this kernel for uniform coefficients only)
  const int fsplit = fast.ta + ((fast.tb - fast.ta) * 2) / 5;
  const T dgl = star_diag<T, KFlat, NOPS>(eq, 0, 0, 0);
  const T rcl = (T)1 / dgl;
  const bool den_ok = exp_window(dgl, -DivWin<T>::DEN, DivWin<T>::DEN);
  unsigned s = 0;
  while (!ls.done) {
    if ((s & 63u) == 63u && res_aborted()) break;  // (watchdog, every 64 sweeps: see k_euler_resident)
    T* cur = (s & 1u) ? sb1 : sb0;
    T* nxt = (s & 1u) ? sb0 : sb1;
    T* gout = ((s + 1u) & 1u) ? xb : xa;
    const unsigned seq_in = s == 0u ? 0u : seq0 + s, seq_out = seq0 + s + 1u;
    double qs[1] = {0.0};
    unsigned got = 0u;
    if (fast.boundary) {
      int k = 0;
      for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
        if (blockIdx.x > 0 && res_halo_get<T>(c, ll, cur, cv * VEC, 0, s, seq_in, xa, false)) got |= 1u << (2 * k);
        if (blockIdx.x + 1 < gridDim.x && res_halo_get<T>(c, ll, cur, cv * VEC, 1, s, seq_in, xa, false))
          got |= 2u << (2 * k);
      }
    }
    auto cell = [&](const T (&v0)[VEC], const T (&vm)[VEC], const T (&vp)[VEC], T zl, T zr, int row, int col,
                    unsigned cm, bool rin, T (&o)[VEC]) {
      typedef typename VecOf<T>::type V;
      V q = __ldg(reinterpret_cast<const V*>(rhs + (c.row0 + row) * rowlen + col));
      const T* av = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
        const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
        const T a = res_star<T, true, NOPS>(eq, scs, 0, 0, v0[e], vp[e], vm[e], zp, zm);
        const T res = av[e] - a;
        const bool in = rin && ((cm >> e) & 1u);
        T xn = v0[e];
        if (in) {
          xn = v0[e] + div_rcp<T>(res, dgl, rcl, den_ok);
          const T df = xn - v0[e];  // (every face Dirichlet: the region is the grid minus its shell)
          const T q2 = df * df;
          qs[0] += (double)q2;
        }
        o[e] = xn;
      }
    };
    auto put = [&](int row, int col, const T (&o)[VEC], unsigned) {
      sts_vec<T>(nxt + row * rowlen + col, o);
      sts_vec<T>(gout + (c.row0 + row) * rowlen + col, o);
    };
    auto put_boundary = [&](int row, int col, const T (&o)[VEC], unsigned m) {
      put(row, col, o, m);
      res_send<T>(c, ll, row, col, s + 1u, seq_out, o);
    };
    for (int seg = 0; seg < 2; ++seg) {
      res_fast_march<T>(c, fast, g, cur, seg == 0 ? fast.ta : fsplit, seg == 0 ? fsplit : fast.tb, cell, put);
      if (seg == 0 && fast.boundary) {
        int k = 0;
        for (int cv = fast.cv; cv < c.nv; cv += fast.cstep, ++k) {
          const int col = cv * VEC;
          if (blockIdx.x > 0 && !((got >> (2 * k)) & 1u)) res_halo_get<T>(c, ll, cur, col, 0, s, seq_in, xa, true);
          if (blockIdx.x + 1 < gridDim.x && !((got >> (2 * k + 1)) & 1u))
            res_halo_get<T>(c, ll, cur, col, 1, s, seq_in, xa, true);
          res_fast_boundary<T>(c, g, cur, col, cell, put_boundary);
        }
      }
    }
    res_allsum<1>(qs, inbox, s + 1u, seq_out, red, bc);
    if (threadIdx.x == 0) {
      ls.sum[R_B] = qs[0];
      ls.sum[R_SHELL] = 0.0;  // static shell
      finalize_stage<T>(ST_JA_FIN, &ls);
    }
    __syncthreads();
    ++s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = ls;
}

// ---- host ------------------------------------------------------------------------------------
struct ResPlan {
  int ctas, R;
  size_t smem;
};

// Exchange buffer (all-reduce inboxes + LL rows): ONE per device, grow-only.  Nothing is zeroed per launch: every
// launch takes a fresh range of sequence numbers, so lines of earlier launches never match.  Resident launches of one
// process on one device are serialised through an event (two of them would share the lines).
struct ResBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  unsigned next_seq = 1u;
  cudaEvent_t last = nullptr;
};
constexpr size_t kResSlotBytes = (size_t)2 * kResMaxCtas * kResMaxCtas * kResSlotLines * 16;  // the inboxes, 1.4 MB

inline ResBuf* res_exchange_buffer(size_t ll_bytes, cudaStream_t s) {
  static ResBuf pool[kMaxDevices];
  ResBuf& b = pool[current_device()];
  const size_t need = kResSlotBytes + ll_bytes;
  if (b.bytes < need) {
    // first allocation: 32 MiB (covers 1024^2 fp64 with room to spare), so that a process that starts on small grids
    // does not walk through a series of re-allocations; beyond that geometric growth
    size_t want = need > 2 * b.bytes ? need : 2 * b.bytes;
    if (want < ((size_t)32 << 20)) want = (size_t)32 << 20;
    if (b.ptr) cudaFree(b.ptr);  // (synchronises the device: no launch still uses it)
    b.ptr = nullptr;
    b.bytes = 0;
    // cudaMemset of device memory is asynchronous: without the synchronisation a launch on a non-blocking stream
    // can start first, and the late memset then erases lines the kernel is waiting for (observed: a hang)
    if (cudaMalloc(&b.ptr, want) != cudaSuccess || cudaMemset(b.ptr, 0, want) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
      cudaGetLastError();
      if (b.ptr) cudaFree(b.ptr);
      b.ptr = nullptr;
      return nullptr;
    }
    b.bytes = want;  // (next_seq keeps counting: recycled memory may hold old lines)
  }
  if (b.last == nullptr && cudaEventCreateWithFlags(&b.last, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (cudaStreamWaitEvent(s, b.last, 0) != cudaSuccess) cudaGetLastError();  // (never recorded yet: a no-op)
  return &b;
}
// watchdog word: cleared on the launch stream in front of every resident launch, read back after the caller's
// synchronisation (res_check_abort; only meaningful inside the translation unit that owns g_res_abort)
inline bool res_clear_abort(cudaStream_t s) {
  const unsigned int zero = 0u;
  return cudaMemcpyToSymbolAsync(g_res_abort, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, s) == cudaSuccess;
}
inline bool res_abort_raised() {
  unsigned int v = 0u;
  if (cudaMemcpyFromSymbol(&v, g_res_abort, sizeof(v), 0, cudaMemcpyDeviceToHost) != cudaSuccess) return true;
  return v != 0u;
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

// PA_RES_DEBUG=1: device buffer for the kernels' phase time stamps, printed by the launcher after a synchronisation
inline unsigned long long* res_debug_buffer() {
  static unsigned long long* buf = nullptr;
  if (getenv("PA_RES_DEBUG") == nullptr) return nullptr;
  if (!buf && cudaMalloc(&buf, 16 * sizeof(unsigned long long)) != cudaSuccess) buf = nullptr;
  if (buf) cudaMemset(buf, 0, 16 * sizeof(unsigned long long));
  return buf;
}
inline void res_debug_print(const char* what, unsigned long long* dbg, int n, cudaStream_t s) {
  if (!dbg) return;
  unsigned long long h[16];
  cudaStreamSynchronize(s);
  cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
  fprintf(stderr, "[PA_RES_DEBUG] %s: phase stamps (ns since the first):", what);
  for (int k = 1; k < n; ++k) fprintf(stderr, " %lld", (long long)(h[k] - h[0]));
  if (h[11] > h[10]) fprintf(stderr, "  | 900 iterations: %.3f us each", (double)(h[11] - h[10]) / 900e3);
  fprintf(stderr, "\n");
}

// resident rows per CTA: Euler 2 (R + 2) (two buffers with halo rows), CG 3R + 2 (d with halo rows, r, x)
template <typename T>
inline bool res_plan(const GridDev& g, bool cg, ResPlan& p) {
  constexpr int VEC = VecOf<T>::N;
  if (!g.act[0] || g.act[1] || !g.act[2]) return false;  // 2-D meshes: kernel axes (0, 2)
  if (g.n[2] % VEC != 0 || g.n[2] < 2 * VEC || g.n[0] < 3) return false;
  if (g.n[2] / VEC > 6 * kResThreads) return false;  // <= 16 halo vectors per thread (res_halo_prefetch's bit mask)
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop || sms < 1) return false;
  if (sms > kResMaxCtas) sms = kResMaxCtas;
  const int want = g.n[0] < sms ? g.n[0] : sms;
  p.R = (g.n[0] + want - 1) / want;
  p.ctas = (g.n[0] + p.R - 1) / p.R;
  const long long rows = cg ? 3LL * p.R + 2 : 2LL * (p.R + 2);
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

// =========================================================================================
// Explicit operators on small / medium 2-D grids: one thread per 16-byte vector, neighbours straight from global
// memory.  The TMA star engine walks a 2-D grid row by row through a producer / consumer pipeline; a single
// application of an operator on 1024^2 then costs ~10 us of pipeline latency for 3 us of traffic (0.25 of the
// roofline, launch- and latency-bound).  Here there is no pipeline to fill: 512 K independent threads, every row is
// read three times but from L1 / L2.  Same semantics as PW_APPLY / PW_GRAD: defined on EVERY cell with torch.roll's
// wrap-around on both axes (fdc.py:171-200), class-indexed coefficients next to the walls, same operation order.
// =========================================================================================
template <typename T, int NOPS, bool GRAD>
__global__ void __launch_bounds__(256)
k_apply_direct2d(GridDev g, EqDev<T> eq, const T* __restrict__ phi, T* __restrict__ out) {
  constexpr int VEC = VecOf<T>::N;
  const int n0 = g.n[0], n2 = g.n[2], nv = n2 / VEC;
  const ResScales<T, NOPS> scs(eq);
  const int total = n0 * nv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / nv, col = (i - row * nv) * VEC;
    const int rm = row == 0 ? n0 - 1 : row - 1, rp = row == n0 - 1 ? 0 : row + 1;
    const T* p = phi + row * n2 + col;
    T v0[VEC], vm[VEC], vp[VEC];
    lds_vec<T>(p, v0);
    lds_vec<T>(phi + rm * n2 + col, vm);
    lds_vec<T>(phi + rp * n2 + col, vp);
    const T zl = col > 0 ? p[-1] : phi[row * n2 + n2 - 1];
    const T zr = col + VEC < n2 ? p[VEC] : phi[row * n2];
    const int clx = coef_class(g, 0, row);
    const bool lean = clx == 0 && col >= 2 && col + VEC <= n2 - 2;
    T o0[VEC], o2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const T zp = (e == VEC - 1) ? zr : v0[e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl : v0[e > 0 ? e - 1 : 0];
      if (GRAD) {
        // central gradient, one output array per mesh axis (fdc.py:80-87), op order of k_grad / PW_GRAD
        const OpDev<T>& o = eq.op[0];
        const int cz = lean ? 0 : coef_class(g, 2, col + e);
        T s = o.coef[0][clx][0] * vp[e];
        s = s + o.coef[0][clx][1] * v0[e];
        s = s + o.coef[0][clx][2] * vm[e];
        if (o.has_param) s = s * o.param;
        o0[e] = s;
        s = o.coef[2][cz][0] * zp;
        s = s + o.coef[2][cz][1] * v0[e];
        s = s + o.coef[2][cz][2] * zm;
        if (o.has_param) s = s * o.param;
        o2[e] = s;
      } else if (lean) {
        o0[e] = res_star<T, true, NOPS, false>(eq, scs, 0, 0, v0[e], vp[e], vm[e], zp, zm);
      } else {
        o0[e] = res_star<T, false, NOPS, false>(eq, scs, clx, coef_class(g, 2, col + e), v0[e], vp[e], vm[e], zp, zm);
      }
    }
    typedef typename VecOf<T>::type V;
    V q;
    T* qs = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int e = 0; e < VEC; ++e) qs[e] = o0[e];
    __stcs(reinterpret_cast<V*>(out + row * n2 + col), q);
    if (GRAD) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) qs[e] = o2[e];
      __stcs(reinterpret_cast<V*>(out + total * VEC + row * n2 + col), q);
    }
  }
}

// 2-D meshes up to kDirect2dCells cells (32-bit offsets; beyond, the TMA engine's streaming wins anyway)
constexpr long long kDirect2dCells = 1LL << 23;
template <typename T>
inline bool direct2d_eligible(const GridDev& g, const pa_equation& eq) {
  constexpr int VEC = VecOf<T>::N;
  if (!g.act[0] || g.act[1] || !g.act[2]) return false;
  if (g.n[2] % VEC != 0 || g.n[2] < 2 * VEC || g.n[0] < 3) return false;
  if (g.cells > kDirect2dCells) return false;
  return res_eq_ok<T>(eq);
}
template <typename T, bool GRAD>
bool launch_apply_direct2d(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const T* phi, T* out) {
  const long long vecs = g.cells / VecOf<T>::N;
  long long blocks = (vecs + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (GRAD)
    k_apply_direct2d<T, 1, true><<<(int)blocks, 256, 0, s>>>(g, eq, phi, out);
  else if (eq.nops == 1)
    k_apply_direct2d<T, 1, false><<<(int)blocks, 256, 0, s>>>(g, eq, phi, out);
  else if (eq.nops == 2)
    k_apply_direct2d<T, 2, false><<<(int)blocks, 256, 0, s>>>(g, eq, phi, out);
  else
    k_apply_direct2d<T, 0, false><<<(int)blocks, 256, 0, s>>>(g, eq, phi, out);
  return cudaGetLastError() == cudaSuccess;
}

// the three coefficient classes of every operator hold the same numbers on both axes of the 2-D grid (no Neumann /
// Symmetry face; compared bit for bit, like TmaPlan::coef_uniform)
inline bool res_coef_uniform(const pa_equation& eq) {
  for (int k = 0; k < eq.nops; ++k)
    for (int a = 0; a < 3; a += 2)
      for (int cls = 1; cls < 3; ++cls)
        for (int q = 0; q < 3; ++q)
          if (std::memcmp(&eq.ops[k].coef[a][cls][q], &eq.ops[k].coef[a][0][q], sizeof(double)) != 0) return false;
  return true;
}

template <typename T, int NOPS, bool HAS_RHS>
static bool launch_euler_resident_n(cudaStream_t s, const ResPlan& p, const GridDev& g, const EqDev<T>& eq, T* b0,
                                    T* b1, const T* rhs, T dt, int nsteps, int uni) {
  if (cudaFuncSetAttribute(k_euler_resident<T, NOPS, HAS_RHS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)p.smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ResBuf* buf = res_exchange_buffer(res_ll_bytes<T>(g, p), s);
  unsigned seq0 = 0;
  if (!buf || !res_take_seq(*buf, (unsigned)nsteps + 1u, s, &seq0) || !res_clear_abort(s)) return false;
  typename LLOf<T>::line* ll = (typename LLOf<T>::line*)((char*)buf->ptr + kResSlotBytes);
  int R = p.R;
  GridDev gg = g;
  EqDev<T> e = eq;
  unsigned long long* dbg = res_debug_buffer();
  void* args[] = {(void*)&gg, (void*)&e, (void*)&b0, (void*)&b1, (void*)&rhs, (void*)&dt, (void*)&nsteps, (void*)&R,
                  (void*)&ll, (void*)&seq0, (void*)&uni, (void*)&dbg};
  if (cudaLaunchCooperativeKernel((void*)k_euler_resident<T, NOPS, HAS_RHS>, dim3(p.ctas), dim3(kResThreads), args,
                                  p.smem, s) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaEventRecord(buf->last, s);
  res_debug_print("euler step 8: items | barrier | (3) | end of the last step", dbg, 5, s);
  return true;
}

// after the stream of a resident launch has been synchronised: true if its watchdog fired (tu_resident.cu)
bool res_check_abort();

// nsteps explicit Euler steps in one launch; false: not launched (the caller takes the streaming path)
template <typename T>
bool launch_euler_resident(cudaStream_t s, const GridDev& g, const pa_equation& peq, const EqDev<T>& eq, T* b0, T* b1,
                           const T* rhs, T dt, int nsteps) {
  ResPlan p;
  if (!res_eq_ok<T>(peq) || !res_plan<T>(g, false, p)) return false;
  const char* ev = getenv("PA_RES_PATH");  // "items": keep the general item loop (A/B runs, tests)
  const int uni = (res_coef_uniform(peq) && !(ev != nullptr && strcmp(ev, "items") == 0)) ? 1 : 0;
  if (rhs) {
    if (eq.nops == 1) return launch_euler_resident_n<T, 1, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
    if (eq.nops == 2) return launch_euler_resident_n<T, 2, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
    return launch_euler_resident_n<T, 0, true>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
  }
  if (eq.nops == 1) return launch_euler_resident_n<T, 1, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
  if (eq.nops == 2) return launch_euler_resident_n<T, 2, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
  return launch_euler_resident_n<T, 0, false>(s, p, g, eq, b0, b1, rhs, dt, nsteps, uni);
}

// the whole CG solve in one launch (eq: ONE star operator, has_shift for the implicit-Euler term); at most
// max_it + 1 iterations run (finalize_stage ST_CG_FIN)
template <typename T>
bool launch_cg_resident(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, T* xa, T* xb, const T* r, const T* d,
                        SolverState* st, int max_it, int uni) {
  ResPlan p;
  if (eq.nops != 1 || max_it < 0 || max_it > 1000000000 || !res_plan<T>(g, true, p)) return false;
  if (cudaFuncSetAttribute(k_cg_resident<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ResBuf* buf = res_exchange_buffer(res_ll_bytes<T>(g, p), s);
  unsigned seq0 = 0;
  if (!buf || !res_take_seq(*buf, 2u * ((unsigned)max_it + 3u) + 2u, s, &seq0) || !res_clear_abort(s)) return false;
  uint4* inbox = (uint4*)buf->ptr;
  typename LLOf<T>::line* ll = (typename LLOf<T>::line*)((char*)buf->ptr + kResSlotBytes);
  int R = p.R;
  GridDev gg = g;
  EqDev<T> e = eq;
  unsigned long long* dbg = res_debug_buffer();
  const char* ev = getenv("PA_RES_PATH");  // "items": keep the general item loop (A/B runs, tests)
  if (ev != nullptr && strcmp(ev, "items") == 0) uni = 0;
  void* args[] = {(void*)&gg, (void*)&e, (void*)&xa, (void*)&xb, (void*)&r, (void*)&d, (void*)&st, (void*)&R,
                  (void*)&ll, (void*)&inbox, (void*)&seq0, (void*)&uni, (void*)&dbg};
  if (cudaLaunchCooperativeKernel((void*)k_cg_resident<T>, dim3(p.ctas), dim3(kResThreads), args, p.smem, s) !=
      cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaEventRecord(buf->last, s);
  res_debug_print("cg iteration 8: d update+barrier | d.A(d) | all-reduce | alpha | x,r update | all-reduce | beta", dbg,
                  8, s);
  return true;
}

// the whole Jacobi solve in one launch: 2-D mesh, every face Dirichlet, uniform coefficients, a thread mapping with
// fixed columns (row length a multiple or a divisor of 512 vectors); false: not launched (streaming sweeps instead)
template <typename T, int NOPS>
static bool launch_jacobi_resident_n(cudaStream_t s, const ResPlan& p, const GridDev& g, const EqDev<T>& eq, T* xa, T* xb,
                                     const T* rhs, SolverState* st, int max_it) {
  if (cudaFuncSetAttribute(k_jacobi_resident<T, NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  ResBuf* buf = res_exchange_buffer(res_ll_bytes<T>(g, p), s);
  unsigned seq0 = 0;
  if (!buf || !res_take_seq(*buf, (unsigned)max_it + 4u, s, &seq0) || !res_clear_abort(s)) return false;
  uint4* inbox = (uint4*)buf->ptr;
  typename LLOf<T>::line* ll = (typename LLOf<T>::line*)((char*)buf->ptr + kResSlotBytes);
  int R = p.R;
  GridDev gg = g;
  EqDev<T> e = eq;
  void* args[] = {(void*)&gg, (void*)&e, (void*)&xa, (void*)&xb, (void*)&rhs, (void*)&st, (void*)&R, (void*)&ll,
                  (void*)&inbox, (void*)&seq0};
  if (cudaLaunchCooperativeKernel((void*)k_jacobi_resident<T, NOPS>, dim3(p.ctas), dim3(kResThreads), args, p.smem, s) !=
      cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaEventRecord(buf->last, s);
  return true;
}
template <typename T>
bool launch_jacobi_resident(cudaStream_t s, const GridDev& g, const pa_equation& peq, const EqDev<T>& eq, T* xa, T* xb,
                            const T* rhs, SolverState* st, int max_it) {
  ResPlan p;
  if (max_it < 0 || max_it > 1000000000 || !res_eq_ok<T>(peq) || !res_coef_uniform(peq) || !res_plan<T>(g, false, p))
    return false;
  const int nv = g.n[2] / VecOf<T>::N;
  if (nv % kResThreads != 0 && kResThreads % nv != 0) return false;  // the fast path's thread mapping (res_fast_init)
  if (g.lo[0] > 1 || g.hi[0] < g.n[0] - 1) return false;
  const char* ev = getenv("PA_RES_PATH");
  if (ev != nullptr && strcmp(ev, "items") == 0) return false;
  if (eq.nops == 1) return launch_jacobi_resident_n<T, 1>(s, p, g, eq, xa, xb, rhs, st, max_it);
  if (eq.nops == 2) return launch_jacobi_resident_n<T, 2>(s, p, g, eq, xa, xb, rhs, st, max_it);
  return false;
}

}  // namespace pa
