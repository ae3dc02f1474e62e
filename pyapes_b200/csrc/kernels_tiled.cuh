// Tiled fast path for the hot configuration: ONE constant-coefficient star operator
// (Laplacian with any BC mix / const-u Div), 2-D or 3-D, fp64 or fp32.
//
// 2.5-D register blocking.  A CTA owns a (TY x TZ) tile of the (axis1, axis2) plane and
// marches along axis 0 over a chunk of planes.  A thread owns RY CONSECUTIVE rows x VEC
// contiguous columns (one 16-byte vector per row) and keeps planes x-1, x, x+1 of its cells in
// registers, so
//   * axis-0 neighbours come from registers (previous / next plane),
//   * axis-1 neighbours come from registers for the inner rows; only each thread's first and
//     last row go through (triple-buffered) shared memory, one __syncthreads per plane,
//   * axis-2 neighbours come from the adjacent lane by warp shuffle,
//   * the one-cell ring around the tile is loaded by the edge warps / edge lanes themselves
//     (vector loads for the rows above/below, scalar loads for the two side columns).
// Every field value is therefore read from L2/HBM once per tile (+ring), with 16-byte
// coalesced accesses.  All index arithmetic, wrap-around and predicates are hoisted out of the
// plane loop; tiles that touch neither the array edge nor a boundary-adjacent coefficient
// class take a predicate-free LEAN instantiation (CTA-uniform branch).
//
// CG fusion (SURVEY.md §8d canonical variant, 8 words / cell / iteration):
//   phase A:  d_new = r + beta*d   (to the other d buffer: neighbours still need d_old on
//             their halos), dAd = sum d_new * A(d_new) -> alpha          [R r, R d, W d]
//   phase B:  x_new = x + alpha*d, r -= alpha*A(d)  (A(d) recomputed, never stored),
//             sums r.r and |x_new-x|^2 (non-shell)                 [R x, R d, R r, W x, W r]
// Arithmetic order per cell is identical to eval_equation() in common.cuh (bit-exact).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "kernels_generic.cuh"

#ifndef PA_MINB2
#define PA_MINB2 2
#endif
#ifndef PA_MINB4
#define PA_MINB4 1
#endif

namespace pa {

struct TilePlan {
  int tiles_y, tiles_z, chunks, cx;  // grid geometry; cx = planes per chunk
  int vec_ok;                        // 16-byte vector access legal (n2 % VEC == 0)
  int fuse_fin;                      // phase B also finalizes the iteration (static shell)
  int ry;                            // rows per thread of the instantiation to launch (2 or 4)
  int dist;                          // multi-GPU: leave raw sums for the NCCL all-reduce
  // blockIdx.z -> chunk (halo-exchange overlap on slabs): the first b_lo launch slots are the lowest
  // chunks, the next b_hi the highest ones, the rest continue from chunk0 -- so the chunks that hold the
  // first / last owned plane can run first (tile_chunk()).  signal_halo: those boundary CTAs count
  // themselves in SolverState::halo_count when their planes are written (k_wait_halo).
  // accum = 1 adds this launch's sums to the ones an earlier sub-launch stored
  int chunk0, b_lo, b_hi, signal_halo, accum;
  // TMA kernels, periodic faces on kernel axes 1/2: the boxes cannot wrap (out-of-bounds halo cells
  // arrive as zeros), so boundary tiles read the wrapped halo row / column straight from the global
  // arrays behind the halo tensor maps
  int wrap;
  const void* src0;
  const void* src1;
  // multi-GPU, TMA CG kernels: all-reduce over peer memory inside the kernel (common.cuh P2PDev);
  // peers == nullptr -> raw sums are left for the NCCL all-reduce + k_finalize pair
  P2PDev p2p;
  // slabs, TMA CG kernels: halo exchange through the neighbours' landing zones (common.cuh HaloDev).
  // ghost_last (phase A): the two chunks that read a ghost plane are scheduled LAST, which gives the
  // neighbour's phase B the longest time to deliver it.
  HaloDev halo = {{nullptr, nullptr}, 0, {nullptr, nullptr}, nullptr, 0, 0, 0};
  int ghost_last = 0;
  // TMA kernels are PERSISTENT: a launch has at most 2 CTAs per SM and every CTA walks the work items
  // (tile, chunk) id = blockIdx.x, blockIdx.x + gridDim.x, ... with ONE mbarrier pipeline that keeps running
  // across item boundaries (the producer lane prefetches the next item while the consumers finish this one).
  // nz = chunk slots of this launch (== chunks unless a sub-launch takes a subset).
  int nz = 0;
  int work_slot = 0;  // entry of the launch in the work-counter pool (kernels_tma.cuh)
  int first_static = 0;  // 1: a CTA's first item is its block index, the counter serves the rest
  // star engine: bit a = the three coefficient classes of every operator hold the same numbers on kernel axis a (no
  // Neumann / Symmetry face there); tiles whose wall-adjacent cells only lie on such axes take class 0 like the LEAN
  // path (kernels_tma_pw.cuh, UNI)
  int uni = 0;
};

__device__ __forceinline__ int tile_chunk(const TilePlan& p, int z) {
  if (p.ghost_last) return z < p.chunks - 2 ? z + 1 : (z == p.chunks - 2 ? 0 : z);
  if (z < p.b_lo) return z;
  if (z < p.b_lo + p.b_hi) return p.chunks - p.b_hi + (z - p.b_lo);
  return p.chunk0 + (z - p.b_lo - p.b_hi);
}

template <typename T>
struct VecOf;
template <>
struct VecOf<double> {
  static constexpr int N = 2;
  typedef double2 type;
};
template <>
struct VecOf<float> {
  static constexpr int N = 4;
  typedef float4 type;
};

template <typename T, int RY_>
struct TileCfg {
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int TXT = 32;            // lanes along axis 2: a warp spans one row segment
  static constexpr int TYT = 8;             // warps per CTA, stacked along axis 1
  static constexpr int RY = RY_;            // consecutive rows per thread
  static constexpr int TY = TYT * RY;
  static constexpr int TZ = TXT * VEC;
  static constexpr int THREADS = TXT * TYT;
  static constexpr int NBUF = 3;
  // per plane buffer: bot[TYT+1][TZ] (bot[0] = ring row above the tile; bot[w+1] = last row
  // of warp w) and top[TYT+1][TZ] (top[w] = first row of warp w; top[TYT] = ring row below)
  static constexpr int HALF = (TYT + 1) * TZ;
  static constexpr int PLANE = 2 * HALF;
  static constexpr size_t SMEM = sizeof(T) * NBUF * PLANE;
};

static inline int tile_default_ry() {
  static int ry = 0;
  if (!ry) {
    const char* e = getenv("PA_TILE_RY");
    ry = (e && atoi(e) == 4) ? 4 : 2;
  }
  return ry;
}

template <typename T, int RY>
inline bool plan_tiles_ry(const GridDev& g, const pa_equation& eq, TilePlan& p, int slots);

template <typename T>
inline bool plan_tiles(const GridDev& g, const pa_equation& eq, TilePlan& p) {
  p.ry = tile_default_ry();
  return p.ry == 2 ? plan_tiles_ry<T, 2>(g, eq, p, kNumSMs * PA_MINB2)
                   : plan_tiles_ry<T, 4>(g, eq, p, kNumSMs * PA_MINB4);
}

template <typename T, int RY>
inline bool plan_tiles_ry(const GridDev& g, const pa_equation& eq, TilePlan& p, int slots) {
  typedef TileCfg<T, RY> C;
  if (eq.nops != 1 || eq.ops[0].kind != PA_OP_STAR || eq.ops[0].param_field != nullptr || eq.ops[0].edge != 0 || eq.ops[0].coef_tab[0] || eq.ops[0].coef_tab[1] || eq.ops[0].coef_tab[2]) return false;
  if (!g.act[1] || !g.act[2]) return false;  // 1-D meshes stay on the generic kernels
  if (g.n[1] < 4 || g.n[2] < 2 * C::VEC) return false;
  p.tiles_y = (g.n[1] + C::TY - 1) / C::TY;
  p.tiles_z = (g.n[2] + C::TZ - 1) / C::TZ;
  const int tiles = p.tiles_y * p.tiles_z;
  // Wave-aware chunking along axis 0: maximise (CTA slots used in the last wave) x (useful
  // planes / planes incl. the two halo planes a chunk re-reads).
  int best_c = 1;
  double best = -1.0;
  const int maxc = g.n[0] >= 8 ? g.n[0] / 4 : 1;
  for (int c = 1; c <= maxc; ++c) {
    const int cx = (g.n[0] + c - 1) / c;
    const int cc = (g.n[0] + cx - 1) / cx;
    const long long items = (long long)cc * tiles;
    if (items > kMaxPartials) break;
    const long long waves = (items + slots - 1) / slots;
    const double quant = (double)items / (double)(waves * slots);
    const double halo = g.act[0] ? (double)cx / (double)(cx + 2) : 1.0;
    // mild preference for >= 2 waves worth of CTAs so the tail of one wave overlaps the next
    const double score = quant * halo * (items >= slots ? 1.0 : (double)items / slots);
    if (score > best + 1e-9) {
      best = score;
      best_c = cc;
    }
  }
  p.cx = (g.n[0] + best_c - 1) / best_c;
  p.chunks = (g.n[0] + p.cx - 1) / p.cx;
  p.vec_ok = (g.n[2] % C::VEC == 0) ? 1 : 0;
  p.fuse_fin = 0;
  p.dist = 0;
  p.chunk0 = 0;
  p.b_lo = p.b_hi = p.signal_halo = 0;
  p.halo = HaloDev{{nullptr, nullptr}, 0, {nullptr, nullptr}, nullptr, 0, 0, 0};
  p.ghost_last = 0;
  p.wrap = 0;
  p.src0 = p.src1 = nullptr;
  p.p2p = P2PDev{nullptr, 0, 0, 0, 0};
  p.accum = 0;
  return true;
}

// ---- small helpers ------------------------------------------------------------------------
__device__ __forceinline__ int wrapi(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

template <typename T>
__device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T>
__device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// Per-thread geometry, computed once per CTA.
template <typename T, int RY>
struct TileCtx {
  typedef TileCfg<T, RY> C;
  int lane, warp;
  int row_off[RY];   // element offset of (row k, first column) within a plane, wrapped
  int ring_off;      // ring row: above the tile (warp 0) / below it (warp TYT-1)
  int col_delta;     // row_off[k] + col_delta = this lane's ring-column cell (lane 0 / 31)
  int zc[C::VEC];         // wrapped column of element e (scalar path)
  bool vec;               // 16-B accesses legal for this thread
  // GENERAL path only
  unsigned valid, inreg, nonshell;  // bit k*VEC+e
  int cly[RY], clz[C::VEC];
};

template <typename T, int RY>
__device__ __forceinline__ void tile_setup(const GridDev& g, const TilePlan& p, TileCtx<T, RY>& c,
                                           int y0, int z0) {
  typedef TileCfg<T, RY> C;
  c.lane = threadIdx.x & 31;
  c.warp = threadIdx.x >> 5;
  const int zg = z0 + c.lane * C::VEC;
  const int zw = wrapi(zg, g.n[2]);  // n2 >= TZ is not required: wrap once is enough (n2 >= 2*VEC..)
  c.vec = p.vec_ok != 0;
#pragma unroll
  for (int e = 0; e < C::VEC; ++e) c.zc[e] = (zg + e) % g.n[2];
  const int yb = y0 + c.warp * RY;
#pragma unroll
  for (int k = 0; k < RY; ++k) {
    const int yw = (yb + k) % g.n[1];
    c.row_off[k] = yw * g.n[2] + (c.vec ? (zg % g.n[2]) : 0);
  }
  {
    // ring column: lane 0 -> z0-1, lane 31 -> z0+TZ (wrapped)
    const int zr = (c.lane == 0) ? wrapi(z0 - 1, g.n[2]) : (z0 + C::TZ) % g.n[2];
    c.col_delta = zr - (c.vec ? (zg % g.n[2]) : 0);
    const int yr = (c.warp == 0) ? wrapi(y0 - 1, g.n[1]) : (y0 + C::TY) % g.n[1];
    c.ring_off = yr * g.n[2] + (c.vec ? (zg % g.n[2]) : 0);
  }
  (void)zw;
  c.valid = c.inreg = c.nonshell = 0u;
#pragma unroll
  for (int k = 0; k < RY; ++k) {
    const int y = yb + k;
    c.cly[k] = (y < g.n[1]) ? coef_class(g, 1, y) : 0;
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) {
      const int z = zg + e;
      const bool v = (y < g.n[1]) && (z < g.n[2]);
      const bool rg = v && y >= g.lo[1] && y < g.hi[1] && z >= g.lo[2] && z < g.hi[2];
      const bool ns = v && y != 0 && y != g.n[1] - 1 && z != 0 && z != g.n[2] - 1;
      const unsigned bit = 1u << (k * C::VEC + e);
      if (v) c.valid |= bit;
      if (rg) c.inreg |= bit;
      if (ns) c.nonshell |= bit;
    }
  }
#pragma unroll
  for (int e = 0; e < C::VEC; ++e) c.clz[e] = (zg + e < g.n[2]) ? coef_class(g, 2, zg + e) : 0;
}

template <typename T, int RY, bool LEAN = false>
__device__ __forceinline__ void ld_row(const T* __restrict__ base, int off,
                                       const TileCtx<T, RY>& c, T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  constexpr int N = VecOf<T>::N;
  if (LEAN || c.vec) {
    V q = *reinterpret_cast<const V*>(base + off);
    const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = s[e];
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = base[off + c.zc[e]];
  }
}

template <typename T, int RY, bool LEAN>
__device__ __forceinline__ void st_row(T* __restrict__ base, int off, const TileCtx<T, RY>& c,
                                       int k, const T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  constexpr int N = VecOf<T>::N;
  if (LEAN || (c.vec && ((c.valid >> (k * N)) & ((1u << N) - 1u)) == ((1u << N) - 1u))) {
    V q;
    T* s = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int e = 0; e < N; ++e) s[e] = v[e];
    *reinterpret_cast<V*>(base + off) = q;
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e)
      if ((c.valid >> (k * N + e)) & 1u) base[off + (c.vec ? e : c.zc[e])] = v[e];
  }
}

// The star operator on this thread's cells of the centre plane.
//   vm / vc / vp : planes x-1 / x / x+1 (registers);  up / dn : rows above the thread's first
//   row / below its last row (from smem);  zl / zr : columns left of element 0 / right of
//   element VEC-1 for every row (shuffle or ring).
template <typename T, int RY, bool LEAN, typename F>
__device__ __forceinline__ void star_rows(const GridDev& g, const OpDev<T>& o,
                                          const TileCtx<T, RY>& c, const T (&cx)[3], bool actx,
                                          const T (&vm)[RY][VecOf<T>::N],
                                          const T (&vc)[RY][VecOf<T>::N],
                                          const T (&vp)[RY][VecOf<T>::N],
                                          const T (&up)[VecOf<T>::N], const T (&dn)[VecOf<T>::N],
                                          const T (&zl)[RY], const T (&zr)[RY], F emit) {
  constexpr int VEC = VecOf<T>::N;
#pragma unroll
  for (int k = 0; k < RY; ++k) {
    T yap, yac, yam;
    if (LEAN) {
      yap = o.coef[1][0][0];
      yac = o.coef[1][0][1];
      yam = o.coef[1][0][2];
    } else {
      yap = o.coef[1][c.cly[k]][0];
      yac = o.coef[1][c.cly[k]][1];
      yam = o.coef[1][c.cly[k]][2];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      T zap, zac, zam;
      if (LEAN) {
        zap = o.coef[2][0][0];
        zac = o.coef[2][0][1];
        zam = o.coef[2][0][2];
      } else {
        zap = o.coef[2][c.clz[e]][0];
        zac = o.coef[2][c.clz[e]][1];
        zam = o.coef[2][c.clz[e]][2];
      }
      const T v0 = vc[k][e];
      const T yp = (k == RY - 1) ? dn[e] : vc[k + 1 < RY ? k + 1 : k][e];
      const T ym = (k == 0) ? up[e] : vc[k > 0 ? k - 1 : 0][e];
      const T zp = (e == VEC - 1) ? zr[k] : vc[k][e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl[k] : vc[k][e > 0 ? e - 1 : 0];
      T acc = (T)0;
      if (actx) {
        T s = cx[0] * vp[k][e];
        s = s + cx[1] * v0;
        s = s + cx[2] * vm[k][e];
        acc = acc + s;
      }
      {
        T s = yap * yp;
        s = s + yac * v0;
        s = s + yam * ym;
        acc = acc + s;
      }
      {
        T s = zap * zp;
        s = s + zac * v0;
        s = s + zam * zm;
        acc = acc + s;
      }
      if (o.has_param) acc = acc * o.param;
      acc = acc * o.sign;
      const T res = (T)0 + acc;
      emit(k, e, res);
    }
  }
}

// Exchange of one plane's tile-internal edges through shared memory.
//   publish: every warp stores its first row (top[w]) and last row (bot[w+1]); warp 0 also the
//            ring row above (bot[0]), warp TYT-1 the ring row below (top[TYT]).
//   gather : a warp reads bot[w] (row above its first row) and top[w+1] (row below its last).
template <typename T, int RY>
__device__ __forceinline__ void publish(T* sm, const TileCtx<T, RY>& c,
                                        const T (&v)[RY][VecOf<T>::N], const T (&ringrow)[VecOf<T>::N]) {
  typedef TileCfg<T, RY> C;
  typedef typename VecOf<T>::type V;
  auto put = [&](T* dst, const T (&x)[C::VEC]) {
    V q;
    T* s = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) s[e] = x[e];
    *reinterpret_cast<V*>(dst + c.lane * C::VEC) = q;
  };
  put(sm + C::HALF + c.warp * C::TZ, v[0]);           // top[w]
  put(sm + (c.warp + 1) * C::TZ, v[RY - 1]);          // bot[w+1]
  if (c.warp == 0) put(sm, ringrow);                  // bot[0]
  if (c.warp == C::TYT - 1) put(sm + C::HALF + C::TYT * C::TZ, ringrow);  // top[TYT]
}

template <typename T, int RY>
__device__ __forceinline__ void gather(const T* sm, const TileCtx<T, RY>& c, T (&up)[VecOf<T>::N],
                                       T (&dn)[VecOf<T>::N]) {
  typedef TileCfg<T, RY> C;
  typedef typename VecOf<T>::type V;
  V a = *reinterpret_cast<const V*>(sm + c.warp * C::TZ + c.lane * C::VEC);               // bot[w]
  V b = *reinterpret_cast<const V*>(sm + C::HALF + (c.warp + 1) * C::TZ + c.lane * C::VEC);  // top[w+1]
  const T* pa_ = reinterpret_cast<const T*>(&a);
  const T* pb_ = reinterpret_cast<const T*>(&b);
#pragma unroll
  for (int e = 0; e < C::VEC; ++e) {
    up[e] = pa_[e];
    dn[e] = pb_[e];
  }
}

// z-neighbours of the centre plane: adjacent lanes, ring column on lanes 0 / 31
template <typename T, int RY>
__device__ __forceinline__ void z_neighbours(const TileCtx<T, RY>& c, const T (&vc)[RY][VecOf<T>::N],
                                             const T (&ringcol)[RY], T (&zl)[RY], T (&zr)[RY]) {
  constexpr int VEC = VecOf<T>::N;
#pragma unroll
  for (int k = 0; k < RY; ++k) {
    T l = shfl_up1<T>(vc[k][VEC - 1]);
    T r = shfl_dn1<T>(vc[k][0]);
    zl[k] = (c.lane == 0) ? ringcol[k] : l;
    zr[k] = (c.lane == 31) ? ringcol[k] : r;
  }
}

// =========================================================================================
// CG phase A
// =========================================================================================
template <typename T, int RY, bool LEAN>
__device__ __forceinline__ void phaseA_body(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                            const T* __restrict__ r, const T* __restrict__ d_old,
                                            T* __restrict__ d_new, T beta, T* smem, int y0, int z0,
                                            double& acc_out) {
  typedef TileCfg<T, RY> C;
  constexpr int VEC = C::VEC;
  TileCtx<T, RY> c;
  tile_setup<T, RY>(g, p, c, y0, z0);
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  const long long n12 = (long long)g.n[1] * g.n[2];
  const bool actx = g.act[0] != 0;
  const bool edge_lane = (c.lane == 0) || (c.lane == 31);
  const bool ring_up = c.warp == 0, ring_dn = c.warp == C::TYT - 1;

  T vm[RY][VEC], vc[RY][VEC], vp[RY][VEC];
  T rr[RY][VEC], dd[RY][VEC];        // raw prefetch of the next plane
  T rrow_r[VEC], rrow_d[VEC];        // raw ring row (warp 0: above, warp TYT-1: below)
  T rcol_r[RY], rcol_d[RY];          // raw ring column (lane 0: left, lane 31: right)
  T ringcol_c[RY], ringcol_p[RY], ringrow[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) rrow_r[e] = rrow_d[e] = ringrow[e] = (T)0;
#pragma unroll
  for (int k = 0; k < RY; ++k) rcol_r[k] = rcol_d[k] = ringcol_c[k] = ringcol_p[k] = (T)0;

  auto fetch = [&](int xp, bool rings) {
    const T* rp = r + (long long)xp * n12;
    const T* dp = d_old + (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < RY; ++k) {
      ld_row<T, RY, LEAN>(rp, c.row_off[k], c, rr[k]);
      ld_row<T, RY, LEAN>(dp, c.row_off[k], c, dd[k]);
    }
    if (rings) {
      if (ring_up || ring_dn) {
        ld_row<T, RY, LEAN>(rp, c.ring_off, c, rrow_r);
        ld_row<T, RY, LEAN>(dp, c.ring_off, c, rrow_d);
      }
      if (edge_lane) {
#pragma unroll
        for (int k = 0; k < RY; ++k) {
          rcol_r[k] = rp[c.row_off[k] + c.col_delta];
          rcol_d[k] = dp[c.row_off[k] + c.col_delta];
        }
      }
    }
  };
  auto combine = [&](T (&out)[RY][VEC]) {  // d_new = r + beta*d           (linalg.py:141)
#pragma unroll
    for (int k = 0; k < RY; ++k)
#pragma unroll
      for (int e = 0; e < VEC; ++e) out[k][e] = rr[k][e] + beta * dd[k][e];
  };
  auto combine_rings = [&](T (&col)[RY]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) ringrow[e] = rrow_r[e] + beta * rrow_d[e];
#pragma unroll
    for (int k = 0; k < RY; ++k) col[k] = rcol_r[k] + beta * rcol_d[k];
  };
  auto write_d = [&](int xp, const T (&v)[RY][VEC]) {
    T* dp = d_new + (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < RY; ++k) st_row<T, RY, LEAN>(dp, c.row_off[k], c, k, v[k]);
  };

  // prologue
  if (actx) {
    fetch(wrapi(x0 - 1, g.n[0]), false);
    combine(vm);
  }
  fetch(x0, true);
  combine(vc);
  combine_rings(ringcol_c);
  publish<T, RY>(smem + (x0 % C::NBUF) * C::PLANE, c, vc, ringrow);
  write_d(x0, vc);
  if (actx) fetch(wrapi(x0 + 1, g.n[0]), x0 + 1 < x1);

  double acc = 0.0;
  for (int x = x0; x < x1; ++x) {
    const bool more = x + 1 < x1;
    if (actx) {
      combine(vp);
      if (more) {
        combine_rings(ringcol_p);
        publish<T, RY>(smem + ((x + 1) % C::NBUF) * C::PLANE, c, vp, ringrow);
        write_d(x + 1, vp);
        fetch(wrapi(x + 2, g.n[0]), x + 2 < x1);  // in flight during this plane's stencil
      }
    }
    __syncthreads();
    const bool xin = x >= g.lo[0] && x < g.hi[0] && x >= g.olo0 && x < g.ohi0;
    if (xin) {
      T up[VEC], dn[VEC], zl[RY], zr[RY];
      gather<T, RY>(smem + (x % C::NBUF) * C::PLANE, c, up, dn);
      z_neighbours<T, RY>(c, vc, ringcol_c, zl, zr);
      const int clx = actx ? coef_class(g, 0, x) : 0;
      const T cx[3] = {o.coef[0][clx][0], o.coef[0][clx][1], o.coef[0][clx][2]};
      // d is identically 0 outside the solver region, so d*Ad needs no region mask; only
      // cells past the array end (partial tiles) must be dropped
      star_rows<T, RY, LEAN>(g, o, c, cx, actx, vm, vc, vp, up, dn, zl, zr, [&](int k, int e, T ad) {
        if (LEAN || ((c.valid >> (k * VEC + e)) & 1u)) {
          T q = vc[k][e] * ad;
          acc += (double)q;
        }
      });
    }
#pragma unroll
    for (int k = 0; k < RY; ++k) {
      ringcol_c[k] = ringcol_p[k];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        vm[k][e] = vc[k][e];
        vc[k][e] = vp[k][e];
      }
    }
  }
  acc_out = acc;
}

template <typename T, int RY>
__global__ void __launch_bounds__(TileCfg<T, RY>::THREADS, RY == 2 ? PA_MINB2 : PA_MINB4)
k_cg_phaseA(TilePlan p, GridDev g, OpDev<T> o, const T* __restrict__ r, const T* __restrict__ d_old,
            T* __restrict__ d_new, SolverState* st, double* partials) {
  typedef TileCfg<T, RY> C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  if (st->done) return;
  const T beta = (T)st->scal[S_BETA];
  const int y0 = blockIdx.y * C::TY, z0 = blockIdx.x * C::TZ;
  const bool full = p.vec_ok && (y0 + C::TY <= g.n[1]) && (z0 + C::TZ <= g.n[2]);
  const bool edge = (y0 < 2) || (y0 + C::TY > g.n[1] - 2) || (z0 < 2) || (z0 + C::TZ > g.n[2] - 2);
  double acc[1] = {0.0};
  if (full && !edge)
    phaseA_body<T, RY, true>(p, g, o, r, d_old, d_new, beta, smem, y0, z0, acc[0]);
  else
    phaseA_body<T, RY, false>(p, g, o, r, d_old, d_new, beta, smem, y0, z0, acc[0]);
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  grid_reduce<1>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 1>{st, R_A, p.dist ? ST_NONE : ST_CG_DAD});
}

// =========================================================================================
// CG phase B
// =========================================================================================
template <typename T, int RY, bool LEAN>
__device__ __forceinline__ void phaseB_body(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                            const T* __restrict__ x_old, T* __restrict__ x_new,
                                            const T* __restrict__ d, T* __restrict__ r, T alpha,
                                            T* smem, int y0, int z0, double (&acc_out)[2]) {
  typedef TileCfg<T, RY> C;
  constexpr int VEC = C::VEC;
  TileCtx<T, RY> c;
  tile_setup<T, RY>(g, p, c, y0, z0);
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  const long long n12 = (long long)g.n[1] * g.n[2];
  const bool actx = g.act[0] != 0;
  const bool edge_lane = (c.lane == 0) || (c.lane == 31);
  const bool ring_up = c.warp == 0, ring_dn = c.warp == C::TYT - 1;

  T vm[RY][VEC], vc[RY][VEC], vp[RY][VEC];
  T ringcol_c[RY], ringcol_p[RY], ringrow[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) ringrow[e] = (T)0;
#pragma unroll
  for (int k = 0; k < RY; ++k) ringcol_c[k] = ringcol_p[k] = (T)0;

  auto fetch_d = [&](int xp, T (&out)[RY][VEC], T (&col)[RY], bool rings) {
    const T* dp = d + (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < RY; ++k) ld_row<T, RY, LEAN>(dp, c.row_off[k], c, out[k]);
    if (rings) {
      if (ring_up || ring_dn) ld_row<T, RY, LEAN>(dp, c.ring_off, c, ringrow);
      if (edge_lane) {
#pragma unroll
        for (int k = 0; k < RY; ++k) col[k] = dp[c.row_off[k] + c.col_delta];
      }
    }
  };

  if (actx) fetch_d(wrapi(x0 - 1, g.n[0]), vm, ringcol_c, false);
  fetch_d(x0, vc, ringcol_c, true);
  publish<T, RY>(smem + (x0 % C::NBUF) * C::PLANE, c, vc, ringrow);
  if (actx) fetch_d(wrapi(x0 + 1, g.n[0]), vp, ringcol_p, x0 + 1 < x1);

  double a0 = 0.0, a1 = 0.0;
  for (int x = x0; x < x1; ++x) {
    const bool more = x + 1 < x1;
    // x and r of this plane: needed only after the stencil, so their latency hides behind it
    T xv[RY][VEC], rv[RY][VEC];
    {
      const T* xp_ = x_old + (long long)x * n12;
      const T* rp_ = r + (long long)x * n12;
#pragma unroll
      for (int k = 0; k < RY; ++k) {
        ld_row<T, RY, LEAN>(xp_, c.row_off[k], c, xv[k]);
        ld_row<T, RY, LEAN>(rp_, c.row_off[k], c, rv[k]);
      }
    }
    __syncthreads();
    const bool xreg = x >= g.lo[0] && x < g.hi[0];
    const bool xown = x >= g.olo0 && x < g.ohi0;
    const int gx = x + g.goff0;
    const bool xshell = actx && (gx == 0 || gx == g.gn0 - 1);
    T xn[RY][VEC];
    T* xo_ = x_new + (long long)x * n12;
    if (xreg) {
      T up[VEC], dn[VEC], zl[RY], zr[RY], rn[RY][VEC];
      gather<T, RY>(smem + (x % C::NBUF) * C::PLANE, c, up, dn);
      z_neighbours<T, RY>(c, vc, ringcol_c, zl, zr);
      const int clx = actx ? coef_class(g, 0, x) : 0;
      const T cx[3] = {o.coef[0][clx][0], o.coef[0][clx][1], o.coef[0][clx][2]};
      star_rows<T, RY, LEAN>(g, o, c, cx, actx, vm, vc, vp, up, dn, zl, zr, [&](int k, int e, T ad) {
        const bool in = LEAN || ((c.inreg >> (k * VEC + e)) & 1u);
        // outside the region d == 0 and r == 0: x + alpha*0 == x exactly; r must stay untouched
        xn[k][e] = xv[k][e] + alpha * vc[k][e];          // linalg.py:122
        const T t = rv[k][e] - alpha * ad;               // linalg.py:131
        rn[k][e] = in ? t : rv[k][e];
        if (xown) {
          if (in) {  // r == 0 outside the region; wrapped duplicates of a partial tile are not `in`
            const T q = t * t;
            a0 += (double)q;
          }
          if (!xshell && (LEAN || ((c.nonshell >> (k * VEC + e)) & 1u))) {
            const T df = xn[k][e] - xv[k][e];
            const T q2 = df * df;
            a1 += (double)q2;
          }
        }
      });
      T* ro_ = r + (long long)x * n12;
#pragma unroll
      for (int k = 0; k < RY; ++k) {
        st_row<T, RY, LEAN>(xo_, c.row_off[k], c, k, xn[k]);
        st_row<T, RY, LEAN>(ro_, c.row_off[k], c, k, rn[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < RY; ++k) st_row<T, RY, LEAN>(xo_, c.row_off[k], c, k, xv[k]);
    }
    // plane x+1 becomes the centre: publish its edges now (its loads have landed: the stencil
    // above needed them), visible after the next iteration's barrier
    if (actx && more) publish<T, RY>(smem + ((x + 1) % C::NBUF) * C::PLANE, c, vp, ringrow);
#pragma unroll
    for (int k = 0; k < RY; ++k) {
      ringcol_c[k] = ringcol_p[k];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        vm[k][e] = vc[k][e];
        vc[k][e] = vp[k][e];
      }
    }
    if (actx && more) fetch_d(wrapi(x + 2, g.n[0]), vp, ringcol_p, x + 2 < x1);
  }
  acc_out[0] = a0;
  acc_out[1] = a1;
}

template <typename T, int RY>
__global__ void __launch_bounds__(TileCfg<T, RY>::THREADS, RY == 2 ? PA_MINB2 : PA_MINB4)
k_cg_phaseB(TilePlan p, GridDev g, OpDev<T> o, const T* __restrict__ x_old, T* __restrict__ x_new,
            const T* __restrict__ d, T* __restrict__ r, SolverState* st, double* partials) {
  typedef TileCfg<T, RY> C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA];
  const int y0 = blockIdx.y * C::TY, z0 = blockIdx.x * C::TZ;
  const bool full = p.vec_ok && (y0 + C::TY <= g.n[1]) && (z0 + C::TZ <= g.n[2]);
  const bool edge = (y0 < 2) || (y0 + C::TY > g.n[1] - 2) || (z0 < 2) || (z0 + C::TZ > g.n[2] - 2);
  double acc[2] = {0.0, 0.0};
  if (full && !edge)
    phaseB_body<T, RY, true>(p, g, o, x_old, x_new, d, r, alpha, smem, y0, z0, acc);
  else
    phaseB_body<T, RY, false>(p, g, o, x_old, x_new, d, r, alpha, smem, y0, z0, acc);
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  // static shell (all faces Dirichlet): shell cells never change after the initial BC
  // application, so the shell part of ||x_new - x_old|| is exactly 0 and the iteration's
  // scalar update can run right here instead of after separate BC / shell-norm launches.
  if (p.fuse_fin) {
    if (threadIdx.x == 0) st->sum[R_SHELL] = 0.0;  // same value from every CTA
    grid_reduce<2>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 2>{st, R_A, ST_CG_FIN});
  } else {
    grid_reduce<2>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 2>{st, R_A, ST_NONE});
  }
}

template <typename T, int RY>
inline void launch_cg_phaseA_ry(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                                const T* r, const T* d_old, T* d_new, SolverState* st, double* partials) {
  typedef TileCfg<T, RY> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseA<T, RY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(p.tiles_z, p.tiles_y, p.chunks);
  k_cg_phaseA<T, RY><<<grid, C::THREADS, C::SMEM, s>>>(p, g, eq.op[0], r, d_old, d_new, st, partials);
}

template <typename T, int RY>
inline void launch_cg_phaseB_ry(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                                const T* x_old, T* x_new, const T* d, T* r, SolverState* st,
                                double* partials) {
  typedef TileCfg<T, RY> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseB<T, RY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(p.tiles_z, p.tiles_y, p.chunks);
  k_cg_phaseB<T, RY><<<grid, C::THREADS, C::SMEM, s>>>(p, g, eq.op[0], x_old, x_new, d, r, st, partials);
}

template <typename T>
void launch_cg_phaseA(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                             const T* r, const T* d_old, T* d_new, SolverState* st, double* partials) {
  if (p.ry == 2)
    launch_cg_phaseA_ry<T, 2>(s, p, g, eq, r, d_old, d_new, st, partials);
  else
    launch_cg_phaseA_ry<T, 4>(s, p, g, eq, r, d_old, d_new, st, partials);
}

template <typename T>
void launch_cg_phaseB(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                             const T* x_old, T* x_new, const T* d, T* r, SolverState* st,
                             double* partials) {
  if (p.ry == 2)
    launch_cg_phaseB_ry<T, 2>(s, p, g, eq, x_old, x_new, d, r, st, partials);
  else
    launch_cg_phaseB_ry<T, 4>(s, p, g, eq, x_old, x_new, d, r, st, partials);
}

}  // namespace pa
