// Tiled fast path for the hot configuration: ONE constant-coefficient star operator
// (Laplacian with any BC mix / const-u Div / 1-D Grad), 2-D or 3-D, fp64 or fp32.
//
// 2.5-D blocking: a CTA owns a (TY x TZ) tile of the (axis1, axis2) plane and marches along
// axis 0 over a chunk of planes.  Each thread keeps its own cells of planes x-1, x, x+1 in
// registers; the centre plane (plus a one-cell ring) goes through triple-buffered shared
// memory for the axis-1/axis-2 neighbours, so every field value is read from L2/HBM once per
// tile (+ ring) instead of seven times.  16-byte vector loads/stores along the contiguous axis.
//
// CG fusion (SURVEY.md §8d canonical variant, 8 words / cell / iteration):
//   phase A:  d_new = r + beta*d   (written to the other d buffer),
//             dAd   = sum d_new * A(d_new)          -> alpha       [R r, R d, W d]
//   phase B:  x_new = x + alpha*d, r -= alpha*A(d)  (A(d) recomputed, never stored),
//             sums r.r and |x_new-x|^2 (non-shell)                 [R x, R d, R r, W x, W r]
// Arithmetic order per cell is identical to eval_equation() in common.cuh.
#pragma once
#include "common.cuh"
#include "kernels_generic.cuh"

namespace pa {

struct TilePlan {
  int tiles_y, tiles_z, chunks, cx;  // grid geometry; cx = planes per chunk
  int vec_ok;                        // 16-byte vector access legal (n2 % VEC == 0)
};

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

// compile-time tile shape
template <typename T>
struct TileCfg {
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int TXT = 32;             // threads along z (one warp per row segment)
  static constexpr int TYT = 8;              // thread rows
  static constexpr int RY = 4;               // rows per thread
  static constexpr int TY = TYT * RY;        // 32 rows
  static constexpr int TZ = TXT * VEC;       // 64 (fp64) / 128 (fp32) columns
  static constexpr int PITCH = TZ + 2 * VEC; // interior starts VEC elements in (16-B aligned)
  static constexpr int PLANE = (TY + 2) * PITCH;
  static constexpr int NBUF = 3;
  static constexpr int THREADS = TXT * TYT;
  static constexpr int RING = 2 * TZ + 2 * TY;  // ring cells (no corners: star stencil)
  static constexpr size_t SMEM = sizeof(T) * NBUF * PLANE;
};

template <typename T>
inline bool plan_tiles(const GridDev& g, const pa_equation& eq, TilePlan& p) {
  typedef TileCfg<T> C;
  if (eq.nops != 1 || eq.ops[0].kind != PA_OP_STAR) return false;
  if (!g.act[1] || !g.act[2]) return false;  // 1-D meshes stay on the generic kernels
  if (g.n[1] < 3 || g.n[2] < 3) return false;
  p.tiles_y = (g.n[1] + C::TY - 1) / C::TY;
  p.tiles_z = (g.n[2] + C::TZ - 1) / C::TZ;
  int tiles = p.tiles_y * p.tiles_z;
  // enough CTAs for >= ~4 per SM, chunks of at least 8 planes, at most kMaxPartials CTAs
  int want = (kNumSMs * 4 + tiles - 1) / tiles;
  int maxc = g.n[0] / 8 > 0 ? g.n[0] / 8 : 1;
  int chunks = want < maxc ? want : maxc;
  if (chunks < 1) chunks = 1;
  while ((long long)chunks * tiles > kMaxPartials && chunks > 1) --chunks;
  if ((long long)chunks * tiles > kMaxPartials) return false;
  p.cx = (g.n[0] + chunks - 1) / chunks;
  p.chunks = (g.n[0] + p.cx - 1) / p.cx;
  p.vec_ok = (g.n[2] % C::VEC == 0) ? 1 : 0;
  return true;
}

// ---- per-thread tile bookkeeping ----------------------------------------------------------
template <typename T>
struct TileCtx {
  typedef TileCfg<T> C;
  int tx, ty;          // thread coordinates
  int y0, z0;          // tile origin
  int zg;              // global z of this thread's first element
  int yg[C::RY];       // global y of this thread's rows
  bool zin[C::VEC];    // element inside the array
  bool yin[C::RY];
  bool vec;            // this thread may use 16-B accesses
  // ring cell owned by this thread (threads < RING): global (y,z) with wrap, smem offset
  bool has_ring;
  int ring_y, ring_z, ring_s;
  bool ring_valid;
};

template <typename T>
__device__ __forceinline__ void tile_setup(const GridDev& g, const TilePlan& p, TileCtx<T>& c) {
  typedef TileCfg<T> C;
  c.tx = threadIdx.x % C::TXT;
  c.ty = threadIdx.x / C::TXT;
  c.y0 = blockIdx.y * C::TY;
  c.z0 = blockIdx.x * C::TZ;
  c.zg = c.z0 + c.tx * C::VEC;
#pragma unroll
  for (int e = 0; e < C::VEC; ++e) c.zin[e] = (c.zg + e) < g.n[2];
#pragma unroll
  for (int k = 0; k < C::RY; ++k) {
    c.yg[k] = c.y0 + c.ty + k * C::TYT;
    c.yin[k] = c.yg[k] < g.n[1];
  }
  c.vec = p.vec_ok && (c.zg + C::VEC <= g.n[2]);
  // ring: [0,TZ) row above, [TZ,2TZ) row below, [2TZ,2TZ+TY) left column, rest right column
  int t = threadIdx.x;
  c.has_ring = t < C::RING;
  int ly, lz;  // local coords in [-1, TY] x [-1, TZ]
  if (t < C::TZ) {
    ly = -1;
    lz = t;
  } else if (t < 2 * C::TZ) {
    ly = C::TY;
    lz = t - C::TZ;
  } else if (t < 2 * C::TZ + C::TY) {
    ly = t - 2 * C::TZ;
    lz = -1;
  } else {
    ly = t - 2 * C::TZ - C::TY;
    lz = C::TZ;
  }
  // rows/columns past the array end: the "below"/"right" ring sits right after the last
  // valid row/column of a partial tile
  int rows = min(C::TY, g.n[1] - c.y0), cols = min(C::TZ, g.n[2] - c.z0);
  if (ly == C::TY) ly = rows;
  if (lz == C::TZ) lz = cols;
  c.ring_valid = c.has_ring && ly <= rows && lz <= cols && (ly < rows || lz < cols) &&
                 !((ly == -1 || ly == rows) && lz >= cols) && !((lz == -1 || lz == cols) && ly >= rows);
  int gy = c.y0 + ly, gz = c.z0 + lz;
  if (gy < 0) gy += g.n[1];
  if (gy >= g.n[1]) gy -= g.n[1];
  if (gz < 0) gz += g.n[2];
  if (gz >= g.n[2]) gz -= g.n[2];
  c.ring_y = gy;
  c.ring_z = gz;
  c.ring_s = (ly + 1) * C::PITCH + (lz + C::VEC);
}

template <typename T>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, bool vec, const bool* zin,
                                         T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  constexpr int N = VecOf<T>::N;
  if (vec) {
    V q = *reinterpret_cast<const V*>(p);
    const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = s[e];
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = zin[e] ? p[e] : (T)0;
  }
}

template <typename T>
__device__ __forceinline__ void store_vec(T* __restrict__ p, bool vec, const bool* zin,
                                          const T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  constexpr int N = VecOf<T>::N;
  if (vec) {
    V q;
    T* s = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int e = 0; e < N; ++e) s[e] = v[e];
    *reinterpret_cast<V*>(p) = q;
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e)
      if (zin[e]) p[e] = v[e];
  }
}

// own cells of one plane -> shared memory (16-B aligned interior).  Rows / elements past
// the array end are NOT written: their slots hold the ring of a partial tile.
template <typename T>
__device__ __forceinline__ void smem_put(T* sm, const TileCtx<T>& c,
                                         const T (&v)[TileCfg<T>::RY][VecOf<T>::N]) {
  typedef TileCfg<T> C;
  typedef typename VecOf<T>::type V;
#pragma unroll
  for (int k = 0; k < C::RY; ++k) {
    if (!c.yin[k]) continue;
    int ly = c.ty + k * C::TYT;
    T* dst = &sm[(ly + 1) * C::PITCH + C::VEC + c.tx * C::VEC];
    if (c.zin[C::VEC - 1]) {
      V q;
      T* s = reinterpret_cast<T*>(&q);
#pragma unroll
      for (int e = 0; e < C::VEC; ++e) s[e] = v[k][e];
      *reinterpret_cast<V*>(dst) = q;
    } else {
#pragma unroll
      for (int e = 0; e < C::VEC; ++e)
        if (c.zin[e]) dst[e] = v[k][e];
    }
  }
}

// the star operator on this thread's cells of plane x, given the three planes
//   vm / vp: planes x-1 / x+1 (registers), vc: plane x (registers), sm: plane x with ring
template <typename T, typename F>
__device__ __forceinline__ void star_plane(const GridDev& g, const OpDev<T>& o, const TileCtx<T>& c,
                                           int x, const T* sm,
                                           const T (&vm)[TileCfg<T>::RY][VecOf<T>::N],
                                           const T (&vc)[TileCfg<T>::RY][VecOf<T>::N],
                                           const T (&vp)[TileCfg<T>::RY][VecOf<T>::N], F emit) {
  typedef TileCfg<T> C;
  typedef typename VecOf<T>::type V;
  const int clx = g.act[0] ? coef_class(g, 0, x) : 0;
  const T xap = o.coef[0][clx][0], xac = o.coef[0][clx][1], xam = o.coef[0][clx][2];
#pragma unroll
  for (int k = 0; k < C::RY; ++k) {
    if (!c.yin[k]) continue;
    const int ly = c.ty + k * C::TYT;
    const int cly = coef_class(g, 1, c.yg[k]);
    const T yap = o.coef[1][cly][0], yac = o.coef[1][cly][1], yam = o.coef[1][cly][2];
    const T* row = &sm[(ly + 1) * C::PITCH + C::VEC + c.tx * C::VEC];
    V up4 = *reinterpret_cast<const V*>(row + C::PITCH);   // y+1
    V dn4 = *reinterpret_cast<const V*>(row - C::PITCH);   // y-1
    const T* upv = reinterpret_cast<const T*>(&up4);
    const T* dnv = reinterpret_cast<const T*>(&dn4);
    const T zl = row[-1], zr = row[C::VEC];
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) {
      if (!c.zin[e]) continue;
      const int zgl = c.zg + e;
      const int clz = coef_class(g, 2, zgl);
      const T v0 = vc[k][e];
      const T zm = (e == 0) ? zl : vc[k][e - 1];
      // next element in registers, unless it is past the array end (then the ring holds it)
      T zp;
      if (e == C::VEC - 1) zp = zr;
      else zp = c.zin[e + 1 < C::VEC ? e + 1 : e] ? vc[k][e + 1 < C::VEC ? e + 1 : e] : row[e + 1];
      T acc = (T)0;
      if (g.act[0]) {
        T s = xap * vp[k][e];
        s = s + xac * v0;
        s = s + xam * vm[k][e];
        acc = acc + s;
      }
      {
        T s = yap * upv[e];
        s = s + yac * v0;
        s = s + yam * dnv[e];
        acc = acc + s;
      }
      {
        T s = o.coef[2][clz][0] * zp;
        s = s + o.coef[2][clz][1] * v0;
        s = s + o.coef[2][clz][2] * zm;
        acc = acc + s;
      }
      if (o.has_param) acc = acc * o.param;
      acc = acc * o.sign;
      T res = (T)0 + acc;
      emit(k, e, zgl, res);
    }
  }
}

__device__ __forceinline__ int wrap_plane(int x, int n) { return x < 0 ? x + n : (x >= n ? x - n : x); }

// =========================================================================================
// CG phase A
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(TileCfg<T>::THREADS, 2)
k_cg_phaseA(TilePlan p, GridDev g, OpDev<T> o, const T* __restrict__ r, const T* __restrict__ d_old,
            T* __restrict__ d_new, SolverState* st, double* partials) {
  typedef TileCfg<T> C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  if (st->done) return;
  const T beta = (T)st->scal[S_BETA];
  TileCtx<T> c;
  tile_setup<T>(g, p, c);
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  const long long n12 = (long long)g.n[1] * g.n[2];

  T vm[C::RY][C::VEC], vc[C::RY][C::VEC], vp[C::RY][C::VEC];
  T rr[C::RY][C::VEC], dd[C::RY][C::VEC];  // raw prefetch
  T ring_r = (T)0, ring_d = (T)0;
#pragma unroll
  for (int k = 0; k < C::RY; ++k)
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) rr[k][e] = dd[k][e] = (T)0;

  auto fetch = [&](int xp) {  // raw r, d of plane xp (own cells + ring cell)
    const long long base = (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < C::RY; ++k) {
      if (c.yin[k]) {
        const long long off = base + (long long)c.yg[k] * g.n[2] + c.zg;
        load_vec<T>(r + off, c.vec, c.zin, rr[k]);
        load_vec<T>(d_old + off, c.vec, c.zin, dd[k]);
      }
    }
    if (c.ring_valid) {
      const long long off = base + (long long)c.ring_y * g.n[2] + c.ring_z;
      ring_r = r[off];
      ring_d = d_old[off];
    }
  };
  auto combine = [&](T (&out)[C::RY][C::VEC], T* sm, bool with_ring) {  // d_new = r + beta d
#pragma unroll
    for (int k = 0; k < C::RY; ++k)
#pragma unroll
      for (int e = 0; e < C::VEC; ++e) out[k][e] = rr[k][e] + beta * dd[k][e];
    if (sm != nullptr) {
      smem_put<T>(sm, c, out);
      if (with_ring && c.ring_valid) sm[c.ring_s] = ring_r + beta * ring_d;
    }
  };
  auto write_d = [&](int xp, const T (&v)[C::RY][C::VEC]) {
    const long long base = (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < C::RY; ++k)
      if (c.yin[k])
        store_vec<T>(d_new + base + (long long)c.yg[k] * g.n[2] + c.zg, c.vec, c.zin, v[k]);
  };

  // prologue: plane x0-1 (registers only), plane x0 (registers + smem), prefetch x0+1
  if (g.act[0]) {
    fetch(wrap_plane(x0 - 1, g.n[0]));
    combine(vm, nullptr, false);
  }
  fetch(x0);
  combine(vc, smem + (x0 % C::NBUF) * C::PLANE, true);
  write_d(x0, vc);
  if (g.act[0]) fetch(wrap_plane(x0 + 1, g.n[0]));

  double acc[1] = {0.0};
  for (int x = x0; x < x1; ++x) {
    if (g.act[0]) {
      combine(vp, smem + ((x + 1) % C::NBUF) * C::PLANE, x + 1 < x1);
      if (x + 1 < x1) {
        write_d(x + 1, vp);
        fetch(wrap_plane(x + 2, g.n[0]));  // in flight during this plane's stencil
      }
    }
    __syncthreads();
    const bool xin = x >= g.lo[0] && x < g.hi[0] && x >= g.olo0 && x < g.ohi0;
    if (xin) {
      star_plane<T>(g, o, c, x, smem + (x % C::NBUF) * C::PLANE, vm, vc, vp,
                    [&](int k, int e, int zgl, T ad) {
                      if (c.yg[k] >= g.lo[1] && c.yg[k] < g.hi[1] && zgl >= g.lo[2] && zgl < g.hi[2]) {
                        T q = vc[k][e] * ad;
                        acc[0] += (double)q;
                      }
                    });
    }
#pragma unroll
    for (int k = 0; k < C::RY; ++k)
#pragma unroll
      for (int e = 0; e < C::VEC; ++e) {
        vm[k][e] = vc[k][e];
        vc[k][e] = vp[k][e];
      }
  }
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  grid_reduce<1>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 1>{st, R_A, ST_CG_DAD});
}

// =========================================================================================
// CG phase B
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(TileCfg<T>::THREADS, 2)
k_cg_phaseB(TilePlan p, GridDev g, OpDev<T> o, const T* __restrict__ x_old, T* __restrict__ x_new,
            const T* __restrict__ d, T* __restrict__ r, SolverState* st, double* partials) {
  typedef TileCfg<T> C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA];
  TileCtx<T> c;
  tile_setup<T>(g, p, c);
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  const long long n12 = (long long)g.n[1] * g.n[2];

  T vm[C::RY][C::VEC], vc[C::RY][C::VEC], vp[C::RY][C::VEC];
  T xv[C::RY][C::VEC], rv[C::RY][C::VEC];
  T ring_d = (T)0;

  auto fetch_d = [&](int xp, T (&out)[C::RY][C::VEC], bool ring) {
    const long long base = (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < C::RY; ++k) {
      if (c.yin[k])
        load_vec<T>(d + base + (long long)c.yg[k] * g.n[2] + c.zg, c.vec, c.zin, out[k]);
      else {
#pragma unroll
        for (int e = 0; e < C::VEC; ++e) out[k][e] = (T)0;
      }
    }
    if (ring && c.ring_valid) ring_d = d[base + (long long)c.ring_y * g.n[2] + c.ring_z];
  };
  auto fetch_xr = [&](int xp) {
    const long long base = (long long)xp * n12;
#pragma unroll
    for (int k = 0; k < C::RY; ++k) {
      if (c.yin[k]) {
        const long long off = base + (long long)c.yg[k] * g.n[2] + c.zg;
        load_vec<T>(x_old + off, c.vec, c.zin, xv[k]);
        load_vec<T>(r + off, c.vec, c.zin, rv[k]);
      }
    }
  };
  auto publish = [&](int xp, const T (&v)[C::RY][C::VEC], bool ring) {
    T* sm = smem + (xp % C::NBUF) * C::PLANE;
    smem_put<T>(sm, c, v);
    if (ring && c.ring_valid) sm[c.ring_s] = ring_d;
  };

  if (g.act[0]) fetch_d(wrap_plane(x0 - 1, g.n[0]), vm, false);
  fetch_d(x0, vc, true);
  publish(x0, vc, true);
  if (g.act[0]) fetch_d(wrap_plane(x0 + 1, g.n[0]), vp, x0 + 1 < x1);
  fetch_xr(x0);

  double acc[2] = {0.0, 0.0};
  for (int x = x0; x < x1; ++x) {
    if (g.act[0] && x + 1 < x1) publish(x + 1, vp, true);
    __syncthreads();
    const bool xreg = x >= g.lo[0] && x < g.hi[0];
    const bool xown = x >= g.olo0 && x < g.ohi0;
    const int gx = x + g.goff0;
    const bool xshell = g.act[0] && (gx == 0 || gx == g.gn0 - 1);
    T xn[C::RY][C::VEC], rn[C::RY][C::VEC];
#pragma unroll
    for (int k = 0; k < C::RY; ++k)
#pragma unroll
      for (int e = 0; e < C::VEC; ++e) {
        xn[k][e] = xv[k][e];
        rn[k][e] = rv[k][e];
      }
    if (xreg) {
      star_plane<T>(g, o, c, x, smem + (x % C::NBUF) * C::PLANE, vm, vc, vp,
                    [&](int k, int e, int zgl, T ad) {
                      if (c.yg[k] >= g.lo[1] && c.yg[k] < g.hi[1] && zgl >= g.lo[2] && zgl < g.hi[2]) {
                        xn[k][e] = xv[k][e] + alpha * vc[k][e];   // linalg.py:122
                        T t = rv[k][e] - alpha * ad;              // linalg.py:131
                        rn[k][e] = t;
                        if (xown) {
                          T q = t * t;
                          acc[0] += (double)q;
                        }
                      }
                    });
    }
    // store, and |x_new - x_old|^2 over owned non-shell cells
    {
      const long long base = (long long)x * n12;
#pragma unroll
      for (int k = 0; k < C::RY; ++k) {
        if (!c.yin[k]) continue;
        const long long off = base + (long long)c.yg[k] * g.n[2] + c.zg;
        store_vec<T>(x_new + off, c.vec, c.zin, xn[k]);
        if (xreg) store_vec<T>(r + off, c.vec, c.zin, rn[k]);
        const bool yshell = c.yg[k] == 0 || c.yg[k] == g.n[1] - 1;
        if (xown && !xshell && !yshell) {
#pragma unroll
          for (int e = 0; e < C::VEC; ++e) {
            const int zgl = c.zg + e;
            if (c.zin[e] && zgl != 0 && zgl != g.n[2] - 1) {
              T df = xn[k][e] - xv[k][e];
              T q = df * df;
              acc[1] += (double)q;
            }
          }
        }
      }
    }
    // rotate and prefetch
#pragma unroll
    for (int k = 0; k < C::RY; ++k)
#pragma unroll
      for (int e = 0; e < C::VEC; ++e) {
        vm[k][e] = vc[k][e];
        vc[k][e] = vp[k][e];
      }
    if (x + 1 < x1) {
      if (g.act[0]) fetch_d(wrap_plane(x + 2, g.n[0]), vp, x + 2 < x1);
      fetch_xr(x + 1);
    }
  }
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  grid_reduce<2>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 2>{st, R_A, ST_NONE});
}

template <typename T>
inline void launch_cg_phaseA(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                             const T* r, const T* d_old, T* d_new, SolverState* st, double* partials) {
  typedef TileCfg<T> C;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseA<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(p.tiles_z, p.tiles_y, p.chunks);
  k_cg_phaseA<T><<<grid, C::THREADS, C::SMEM, s>>>(p, g, eq.op[0], r, d_old, d_new, st, partials);
}

template <typename T>
inline void launch_cg_phaseB(cudaStream_t s, const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                             const T* x_old, T* x_new, const T* d, T* r, SolverState* st,
                             double* partials) {
  typedef TileCfg<T> C;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseB<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(p.tiles_z, p.tiles_y, p.chunks);
  k_cg_phaseB<T><<<grid, C::THREADS, C::SMEM, s>>>(p, g, eq.op[0], x_old, x_new, d, r, st, partials);
}

}  // namespace pa
