// TMA-staged star engine for everything that is "one stencil application + a pointwise
// epilogue": residual init, Jacobi sweep, explicit Euler step, BiCGSTAB's two operator
// applications.  Same producer/consumer mbarrier pipeline as kernels_tma.cuh, but
//   * the equation may have up to PA_MAX_OPS constant-coefficient STAR operators (e.g. upwind
//     Div + Laplacian of the advection-diffusion problems), summed in the reference's order;
//   * one optional auxiliary own-tile input (rhs or r0) rides in the same pipeline stage.
// Algorithmic traffic: R in (+halo), [R aux], W out  = 2-3 words per cell.
#pragma once
#include <type_traits>

#include "kernels_tma.cuh"

namespace pa {

enum PwMode {
  PW_RESID = 0,    // out = rhs - A(in) on the region, 0 elsewhere; out2 = out; sum out^2
  PW_JACOBI = 1,   // out = in + (rhs - A(in)) / diag   on the region, in elsewhere; sum |d|^2
  PW_EULER = 2,    // out = in + dt*(rhs - A(in))       on the region, in elsewhere
  PW_APPLY_V = 3,  // out = A(in) on the region, 0 elsewhere; sum r0*out
  PW_APPLY_T = 4,  // out = A(in) ...; sums out*in, out*out, r0*out; skipped when finished_flag
  // the explicit operators (ops._Aop, FDC().laplacian/.div/.grad): defined on EVERY cell with
  // torch.roll's wrap-around on every axis (fdc.py:171-200), no region, no sums.  R in, W out(s).
  PW_APPLY = 5,    // out = sum_k sign*param*Op_k(in)                        2 words per cell
  PW_GRAD = 6,     // out[a] = param * d(in)/dx_a, one array per mesh axis   1 + d words per cell
};

// MODE decides the pipeline: the explicit operators (PW_APPLY / PW_GRAD) stream ONE input and nothing
// else, so a stage is only the halo tile (9.8 KB) and 2 planes in flight per CTA (S = 4) are not enough
// bytes in flight to cover the DRAM latency (ncu, 512^3 Laplacian: 53 % of the stall samples were
// consumers waiting on `full`); they take S = 8.
template <typename T, typename K, int MODE = 0>
struct PwCfg : TmaCfg<T, K> {
  typedef TmaCfg<T, K> B;
  static constexpr bool NO_AUX = (MODE == PW_APPLY || MODE == PW_GRAD);
  static constexpr int S = NO_AUX ? 8 : B::S;
  static constexpr int STAGE = B::HALO_SLOT + (NO_AUX ? 0 : B::OWN_SLOT);  // in halo, aux own
  static constexpr int BAR_BYTES = 2 * S * 8;
  static constexpr size_t SMEM = (size_t)S * STAGE + BAR_BYTES + 128;
};

struct PwPlan {
  TilePlan tile;
  CUtensorMap in_halo[2];  // the two ping-pong buffers of the stencilled field (or one)
  CUtensorMap aux_own;     // rhs / r0
  int has_aux;
};

// ---- exact shortcuts ------------------------------------------------------------------------
// Every shortcut below returns the SAME bits as the reference's operation sequence:
//  * the per-operator accumulator starts from the first axis' sum instead of 0 + sum: the two only
//    differ in the sign of an all-zero sum, which the `0 + acc` of the operator sum erases;
//  * (acc * param) * sign == acc * (param * sign) because sign is +-1 (rounding is symmetric), and a
//    factor of exactly 1 is skipped;
//  * x / d with d constant over a plane: q = x * RN(1/d) followed by two FMA residual corrections
//    is the correctly rounded quotient (Markstein: a correctly rounded reciprocal + a faithful
//    quotient estimate), as long as nothing under/overflows -- operands outside a safe exponent
//    window (and 0, Inf, NaN) take the plain division.
template <typename T>
struct OpScale {
  T scale;
  bool use;
};

template <typename T>
__device__ __forceinline__ OpScale<T> op_scale(const OpDev<T>& o) {
  OpScale<T> r;
  r.scale = o.has_param ? o.param * o.sign : o.sign;
  r.use = r.scale != (T)1;
  return r;
}

__device__ __forceinline__ bool exp_window(double v, int lo, int hi) {  // 2^lo <= |v| < 2^hi
  const unsigned e = ((unsigned)__double2hiint(v) >> 20) & 0x7ffu;
  return (e - (unsigned)(1023 + lo)) < (unsigned)(hi - lo);
}
__device__ __forceinline__ bool exp_window(float v, int lo, int hi) {
  const unsigned e = ((unsigned)__float_as_int(v) >> 23) & 0xffu;
  return (e - (unsigned)(127 + lo)) < (unsigned)(hi - lo);
}
template <typename T>
struct DivWin;
template <>
struct DivWin<double> {
  static constexpr int NUM = 600, DEN = 300;
};
template <>
struct DivWin<float> {
  static constexpr int NUM = 60, DEN = 30;
};

// a / d given rcp = RN(1/d); `den_ok` = d inside its exponent window (uniform per plane)
template <typename T>
__device__ __forceinline__ T div_rcp(T a, T d, T rcp, bool den_ok) {
  if (den_ok && exp_window(a, -DivWin<T>::NUM, DivWin<T>::NUM)) {
    T q = a * rcp;
    T e = fma(-d, q, a);
    q = fma(e, rcp, q);
    e = fma(-d, q, a);
    return fma(e, rcp, q);
  }
  return a / d;
}

// the same quotient once the exponent windows have been checked (for a whole warp at a time, see pw_consumer)
template <typename T>
__device__ __forceinline__ T div_rcp_checked(T a, T d, T rcp) {
  T q = a * rcp;
  T e = fma(-d, q, a);
  q = fma(e, rcp, q);
  e = fma(-d, q, a);
  return fma(e, rcp, q);
}

// diagonal of the operator sum for one coefficient-class triple (Jacobi)
template <typename T, typename K, int NOPS>
__device__ __forceinline__ T star_diag(const EqDev<T>& eq, int clx, int cy, int cz) {
  T diag = (T)0;
  const int nops = NOPS > 0 ? NOPS : eq.nops;
#pragma unroll
  for (int q = 0; q < nops; ++q) {
    const OpDev<T>& o = eq.op[q];
    T dacc = (T)0;
    dacc = dacc + o.coef[0][clx][1];
    if (!K::FLAT) dacc = dacc + o.coef[1][cy][1];
    dacc = dacc + o.coef[2][cz][1];
    if (o.has_param) dacc = dacc * o.param;
    dacc = dacc * o.sign;
    diag = diag + dacc;
  }
  return diag;
}

// sum over operators of sign*param*(star), reference order (ops.py:130-149, fdc.py:103-108).
// Kernel axis 0 is always active here (pw_eligible: 3-D meshes, and 2-D meshes as FLAT tiles).
// UNI: the coefficient classes are bitwise equal (TilePlan::uni): class 0 everywhere, same bits, no per-cell selects.
template <typename T, typename K, bool LEAN, int NOPS, bool UNI, typename F>
__device__ __forceinline__ void star_cells_eq(const EqDev<T>& eq, const ConsCtx<T, K>& c, int clx,
                                              const T (&vm)[K::RY][VecOf<T>::N],
                                              const T (&vc)[K::RY][VecOf<T>::N],
                                              const T (&vp)[K::RY][VecOf<T>::N], const T (&up)[VecOf<T>::N],
                                              const T (&dn)[VecOf<T>::N], const T (&zl)[K::RY],
                                              const T (&zr)[K::RY], F emit) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int MAXO = NOPS > 0 ? NOPS : kMaxOps;
  const int nops = NOPS > 0 ? NOPS : eq.nops;  // compile-time count: unrolled, constant operands
  OpScale<T> sc[MAXO];
#pragma unroll
  for (int q = 0; q < MAXO; ++q)
    if (q < nops) sc[q] = op_scale<T>(eq.op[q]);
#pragma unroll
  for (int k = 0; k < K::RY; ++k) {
    const int cy = (LEAN || UNI) ? 0 : c.cly[k];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int cz = (LEAN || UNI) ? 0 : c.clz[e];
      const T v0 = vc[k][e];
      const T yp = (k == K::RY - 1) ? dn[e] : vc[k + 1 < K::RY ? k + 1 : k][e];
      const T ym = (k == 0) ? up[e] : vc[k > 0 ? k - 1 : 0][e];
      const T zp = (e == VEC - 1) ? zr[k] : vc[k][e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl[k] : vc[k][e > 0 ? e - 1 : 0];
      T res = (T)0;
#pragma unroll
      for (int q = 0; q < MAXO; ++q) {
        if (q >= nops) break;
        const OpDev<T>& o = eq.op[q];
        T acc;
        {
          T s = o.coef[0][clx][0] * vp[k][e];
          s = s + o.coef[0][clx][1] * v0;
          s = s + o.coef[0][clx][2] * vm[k][e];
          acc = s;
        }
        if (!K::FLAT) {
          T s = o.coef[1][cy][0] * yp;
          s = s + o.coef[1][cy][1] * v0;
          s = s + o.coef[1][cy][2] * ym;
          acc = acc + s;
        }
        {
          T s = o.coef[2][cz][0] * zp;
          s = s + o.coef[2][cz][1] * v0;
          s = s + o.coef[2][cz][2] * zm;
          acc = acc + s;
        }
        if (sc[q].use) acc = acc * sc[q].scale;
        res = res + acc;
      }
      emit(k, e, res);
    }
  }
}

// UNI (general path only; the kernel dispatches on the CTA-uniform TilePlan::uni): coefficient class 0 on every axis
// and the LEAN Jacobi quotient -- the edge tiles (56 % of the tiles of a 256^2 plane) then differ from the LEAN ones
// only by their region / validity masks.
template <typename T, typename K, int MODE, bool LEAN, int NOPS, bool UNI = false>
__device__ __forceinline__ void pw_consumer(const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                                            T* __restrict__ out, T* __restrict__ out2, T dt, bool has_aux,
                                            unsigned char* stages, uint64_t* full, uint64_t* empty,
                                            int y0, int z0, int x0, int x1, double (&acc_out)[3]) {
  typedef PwCfg<T, K, MODE> C;
  constexpr int VEC = C::VEC;
  constexpr bool ALL = (MODE == PW_APPLY || MODE == PW_GRAD);  // every cell, wrap-around on every axis
  ConsCtx<T, K> c;
  cons_setup<T, K>(g, c, y0, z0);
  const bool actx = g.act[0] != 0;
  const long long n12 = (long long)g.n[1] * g.n[2];
  T* op_ = out + (long long)x0 * n12 + c.goff;
  T* op2_ = out2 ? out2 + (long long)x0 * n12 + c.goff : nullptr;

  auto halo = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * C::STAGE); };
  auto aux = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * C::STAGE + C::HALO_SLOT); };
  auto load_own = [&](int s, T (&v)[K::RY][VEC]) {
    const T* h = halo(s) + c.hoff;
#pragma unroll
    for (int k = 0; k < K::RY; ++k) lds_vec<T>(h + k * C::BOXZ, v[k]);
  };
  auto release = [&](int s) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&empty[s]);
  };

  T A[K::RY][VEC], B[K::RY][VEC], Cc[K::RY][VEC];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (actx) {
    mbar_wait(&full[0], 0);
    load_own(0, A);
    release(0);
  }
  mbar_wait(&full[1 % C::S], 0);
  load_own(1 % C::S, B);

  auto step = [&](T (&vm)[K::RY][VEC], T (&vc)[K::RY][VEC], T (&vp)[K::RY][VEC], int x, int i) {
    const int sn = i & (C::S - 1), sc = (i - 1) & (C::S - 1);
    if (actx) {
      mbar_wait(&full[sn], (i / C::S) & 1);
      load_own(sn, vp);
    }
    const T* h = halo(sc) + c.hoff;
    const bool xreg = ALL || (x >= g.lo[0] && x < g.hi[0]);
    const bool xown = x >= g.olo0 && x < g.ohi0;
    const int gx = x + g.goff0;
    const bool xshell = actx && (gx == 0 || gx == g.gn0 - 1);
    T ax[K::RY][VEC];
    StepSum<T> s0(a0), s1(a1), s2(a2);
    T dgl = (T)1, rcl = (T)1;  // LEAN Jacobi: one diagonal (and its reciprocal) per plane
    bool den_ok = false;
    if (xreg) {
      T up[VEC], dn[VEC], zl[K::RY], zr[K::RY];
      if (!K::FLAT) {
        lds_vec<T>(h - C::BOXZ, up);
        lds_vec<T>(h + K::RY * C::BOXZ, dn);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) up[e] = dn[e] = (T)0;
      }
#pragma unroll
      for (int k = 0; k < K::RY; ++k) {
        zl[k] = h[k * C::BOXZ - 1];
        zr[k] = h[k * C::BOXZ + VEC];
      }
      if (!LEAN && p.wrap) {
        const T* ing = static_cast<const T*>(p.src0);
        wrap_halo<T, K, ALL>(g, c, x, up, dn, zl, zr, [&](long long i) { return ing[i]; });
        if (ALL && !K::FLAT && c.zg < g.n[2]) {
          // partial tile along axis 1: the row below the last valid row is row 0 (it sits in one of
          // the thread's own -- invalid, never stored -- rows)
#pragma unroll
          for (int k = 0; k + 1 < K::RY; ++k)
            if (c.yb + k + 1 == g.n[1]) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) vc[k + 1][e] = ing[(long long)x * n12 + c.zg + e];
            }
        }
      }
      const int clx = coef_class(g, 0, x);  // (plane-uniform: a dynamic index costs nothing, LEAN and UNI included)
      if (MODE == PW_GRAD) {
        // central gradient, one output array per mesh axis (fdc.py:80-87); op order of k_grad
        const OpDev<T>& o = eq.op[0];
#pragma unroll
        for (int k = 0; k < K::RY; ++k) {
          const int cy = (LEAN || UNI) ? 0 : c.cly[k];
          T g0[VEC], g1[VEC], g2[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int cz = (LEAN || UNI) ? 0 : c.clz[e];
            const T v0 = vc[k][e];
            const T yp = (k == K::RY - 1) ? dn[e] : vc[k + 1 < K::RY ? k + 1 : k][e];
            const T ym = (k == 0) ? up[e] : vc[k > 0 ? k - 1 : 0][e];
            const T zp = (e == VEC - 1) ? zr[k] : vc[k][e + 1 < VEC ? e + 1 : e];
            const T zm = (e == 0) ? zl[k] : vc[k][e > 0 ? e - 1 : 0];
            T s = o.coef[0][clx][0] * vp[k][e];
            s = s + o.coef[0][clx][1] * v0;
            s = s + o.coef[0][clx][2] * vm[k][e];
            if (o.has_param) s = s * o.param;
            g0[e] = s;
            if (!K::FLAT) {
              s = o.coef[1][cy][0] * yp;
              s = s + o.coef[1][cy][1] * v0;
              s = s + o.coef[1][cy][2] * ym;
              if (o.has_param) s = s * o.param;
              g1[e] = s;
            }
            s = o.coef[2][cz][0] * zp;
            s = s + o.coef[2][cz][1] * v0;
            s = s + o.coef[2][cz][2] * zm;
            if (o.has_param) s = s * o.param;
            g2[e] = s;
          }
          T* row = op_ + (long long)k * g.n[2];
          stg_row<T, K, LEAN, true>(row, c, k, g0);
          if (!K::FLAT) {
            stg_row<T, K, LEAN, true>(row + g.cells, c, k, g1);
            stg_row<T, K, LEAN, true>(row + 2 * g.cells, c, k, g2);
          } else {
            stg_row<T, K, LEAN, true>(row + g.cells, c, k, g2);
          }
        }
        op_ += n12;
        release(sc);
        return;
      }
      star_cells_eq<T, K, LEAN, NOPS, UNI>(eq, c, clx, vm, vc, vp, up, dn, zl, zr,
                                           [&](int k, int e, T v) { ax[k][e] = v; });
      if (MODE == PW_JACOBI && (LEAN || UNI)) {
        dgl = star_diag<T, K, NOPS>(eq, clx, 0, 0);
        rcl = (T)1 / dgl;
        den_ok = exp_window(dgl, -DivWin<T>::DEN, DivWin<T>::DEN);
      }
    }
#pragma unroll
    for (int k = 0; k < K::RY; ++k) {
      T av[VEC], o[VEC];
      if (!C::NO_AUX && has_aux) {
        lds_vec<T>(aux(sc) + c.ooff + k * C::OBOXZ, av);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) av[e] = (T)0;
      }
      // Jacobi with one diagonal per plane: the exponent windows of div_rcp are checked for the WARP's cells of this row
      // in one vote -- per cell they were a compare chain and two branches, 80 of the ~290 instructions of a LEAN plane
      // step (ncu source page) -- and the per-cell fallback only runs for a warp that holds a zero / tiny / huge residual
      bool jac_fast = false;
      if (MODE == PW_JACOBI && (LEAN || UNI)) {
        bool okw = den_ok;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const bool in = xreg && (LEAN || ((c.inreg >> (k * VEC + e)) & 1u));
          const T res = av[e] - ax[k][e];
          okw = okw && (!in || exp_window(res, -DivWin<T>::NUM, DivWin<T>::NUM));
        }
        jac_fast = __all_sync(0xffffffffu, okw) != 0;
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const bool in = xreg && (LEAN || ((c.inreg >> (k * VEC + e)) & 1u));
        const T xc = vc[k][e];
        if (MODE == PW_RESID) {
          const T res = in ? av[e] - ax[k][e] : (T)0;
          o[e] = res;
          if (in && xown) {
            const T q = res * res;
            s0.add(q);
          }
        } else if (MODE == PW_JACOBI) {
          T xn = xc;
          if (in) {
            const T res = av[e] - ax[k][e];
            if (LEAN || UNI)
              xn = xc + (jac_fast ? div_rcp_checked<T>(res, dgl, rcl) : div_rcp<T>(res, dgl, rcl, den_ok));
            else
              xn = xc + res / star_diag<T, K, NOPS>(eq, coef_class(g, 0, x), c.cly[k], c.clz[e]);
          }
          o[e] = xn;
          if (xown && !xshell && (LEAN || ((c.nonshell >> (k * VEC + e)) & 1u))) {
            const T df = xn - xc;
            const T q = df * df;
            s1.add(q);
          }
        } else if (MODE == PW_EULER) {
          T xn = xc;
          if (in) {
            const T res = av[e] - ax[k][e];
            xn = xc + dt * res;
          }
          o[e] = xn;
        } else if (MODE == PW_APPLY) {
          o[e] = ax[k][e];  // every valid cell (stg_row masks the others)
        } else {  // PW_APPLY_V / PW_APPLY_T
          const T a = in ? ax[k][e] : (T)0;
          o[e] = a;
          if (in && xown) {
            if (MODE == PW_APPLY_V) {
              const T q = av[e] * a;
              s0.add(q);
            } else {
              const T q0 = a * xc, q1 = a * a, q2 = av[e] * a;
              s0.add(q0);
              s1.add(q1);
              s2.add(q2);
            }
          }
        }
      }
      stg_row<T, K, LEAN, MODE == PW_APPLY>(op_ + (long long)k * g.n[2], c, k, o);
      if (MODE == PW_RESID && op2_) stg_row<T, K, LEAN>(op2_ + (long long)k * g.n[2], c, k, o);
    }
    s0.flush();
    s1.flush();
    s2.flush();
    op_ += n12;
    if (op2_) op2_ += n12;
    release(sc);
  };

  int x = x0, i = 2;
  if (LEAN) {  // 3x unrolled: no register rotation
    while (true) {
      step(A, B, Cc, x, i);
      if (++x >= x1) break;
      ++i;
      step(B, Cc, A, x, i);
      if (++x >= x1) break;
      ++i;
      step(Cc, A, B, x, i);
      if (++x >= x1) break;
      ++i;
    }
  } else {  // boundary tiles: one copy of the body (instruction-cache footprint), rotate registers
    for (; x < x1; ++x, ++i) {
      step(A, B, Cc, x, i);
#pragma unroll
      for (int k = 0; k < K::RY; ++k)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          A[k][e] = B[k][e];
          B[k][e] = Cc[k][e];
        }
    }
  }
  acc_out[0] = a0;
  acc_out[1] = a1;
  acc_out[2] = a2;
}

template <typename T, typename K, int MODE, int NOPS>
__global__ void __launch_bounds__(PwCfg<T, K, MODE>::THREADS, 2)
k_star_tma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_aux,
           TilePlan p, GridDev g, EqDev<T> eq, T* __restrict__ out, T* __restrict__ out2, T dt,
           int has_aux, SolverState* st, double* partials, int stage) {
  typedef PwCfg<T, K, MODE> C;
  extern __shared__ unsigned char smem_dyn[];
  if (st != nullptr && st->done) return;
  if (MODE == PW_APPLY_T && st->finished_flag) return;
  // 128-byte aligned start, derived by pointer arithmetic so the compiler keeps the shared
  // address space (LDS instead of generic LD)
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  unsigned char* stages = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)C::S * C::STAGE);
  uint64_t* empty = full + C::S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y0 = blockIdx.y * C::TY, z0 = blockIdx.x * C::TZ;
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  const bool actx = g.act[0] != 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C::CWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  if (warp == C::CWARPS) {
    if (lane == 0) {
      const int n = x1 - x0 + 2;
      for (int i = 0; i < n; ++i) {
        if (!actx && i != 1) continue;
        const int pl = x0 - 1 + i;
        const int s = i & (C::S - 1);
        if (i >= C::S) mbar_wait(&empty[s], ((i / C::S) - 1) & 1);
        const bool inner = !C::NO_AUX && has_aux && (pl >= x0 && pl < x1);
        mbar_expect_tx(&full[s], (uint32_t)(C::HALO_BYTES + (inner ? C::OWN_BYTES : 0)));
        unsigned char* sb = stages + (size_t)s * C::STAGE;
        const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
        for (int b = 0; b < C::NB; ++b) {
          const int zb = z0 + b * C::OBOXZ;
          tma_load_3d(sb + b * C::HBOX_SLOT, &tm_in, zb - C::HZ, C::FLAT ? 0 : y0 - 1, xw, &full[s]);
          if (inner) tma_load_3d(sb + C::HALO_SLOT + b * C::OBOX_SLOT, &tm_aux, zb, y0, xw, &full[s]);
        }
      }
    }
  } else {
    const bool full_tile = (C::FLAT || y0 + C::TY <= g.n[1]) && (z0 + C::TZ <= g.n[2]);
    const bool edge_y = !C::FLAT && ((y0 < 2) || (y0 + C::TY > g.n[1] - 2));
    const bool edge_z = (z0 < 2) || (z0 + C::TZ > g.n[2] - 2);
    // class-free tile: on each tiled axis either the classes hold the same numbers (TilePlan::uni bit) or the tile has
    // no cell next to a wall (its classes are 0 anyway) -- with Neumann / Symmetry faces on ONE axis (config 4) only
    // the two tile rows along those walls keep the per-cell class selects and the true division of Jacobi
    const bool uni_tile = (!edge_y || (p.uni & 2)) && (!edge_z || (p.uni & 4));
    if (full_tile && !edge_y && !edge_z)
      pw_consumer<T, K, MODE, true, NOPS>(p, g, eq, out, out2, dt, has_aux != 0, stages, full, empty, y0, z0, x0,
                                           x1, acc);
    else if (uni_tile)
      pw_consumer<T, K, MODE, false, NOPS, true>(p, g, eq, out, out2, dt, has_aux != 0, stages, full, empty, y0, z0,
                                                  x0, x1, acc);
    else
      pw_consumer<T, K, MODE, false, NOPS>(p, g, eq, out, out2, dt, has_aux != 0, stages, full, empty, y0, z0, x0,
                                            x1, acc);
  }
  if (MODE == PW_EULER || MODE == PW_APPLY || MODE == PW_GRAD) return;  // no reductions
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  P2PDev pp = p.p2p;  // slabs: the host sets p.p2p when this launch sums over the ranks itself
  pp.slot0 = R_A;
  pp.count = 3;
  grid_reduce<3>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 3>{st, R_A, stage, 0, pp});
}

// =========================================================================================
// BiCGSTAB second half step, fused:  s = r - alpha v  (never stored unless `s_out`),  t = A(s),
// sums |s|^2, t.s, t.t, r0.t  -- two halo inputs (r, v) combined on the fly like CG phase A, r0 as
// the own-tile input.  Traffic: R r, R v, R r0, W t = 4 words per cell (s-stage + T-apply: 6).
// =========================================================================================
template <typename T, typename K>
struct Pw2Cfg : TmaCfg<T, K> {
  typedef TmaCfg<T, K> B;
  static constexpr int STAGE = 2 * B::HALO_SLOT + B::OWN_SLOT;  // r halo, v halo, r0 own
  static constexpr size_t SMEM = (size_t)B::S * STAGE + B::BAR_BYTES + 128;
};

template <typename T, typename K, bool LEAN, int NOPS>
__device__ __forceinline__ void pw2_consumer(const TilePlan& p, const GridDev& g, const EqDev<T>& eq,
                                             T* __restrict__ t_out, T* __restrict__ s_out, T alpha,
                                             unsigned char* stages, uint64_t* full, uint64_t* empty, int y0,
                                             int z0, int x0, int x1, double (&acc_out)[4]) {
  typedef Pw2Cfg<T, K> C;
  constexpr int VEC = C::VEC;
  ConsCtx<T, K> c;
  cons_setup<T, K>(g, c, y0, z0);
  const long long n12 = (long long)g.n[1] * g.n[2];
  T* tp_ = t_out + (long long)x0 * n12 + c.goff;
  T* sp_ = s_out ? s_out + (long long)x0 * n12 + c.goff : nullptr;

  auto rt = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * C::STAGE) + c.hoff; };
  auto vt = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * C::STAGE + C::HALO_SLOT) + c.hoff; };
  auto aux = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * C::STAGE + 2 * C::HALO_SLOT); };
  // s = r - alpha v on the thread's own cells of the plane in stage s   (linalg.py:230)
  auto own_s = [&](int s, T (&o)[K::RY][VEC]) {
    const T* rp = rt(s);
    const T* vp = vt(s);
#pragma unroll
    for (int k = 0; k < K::RY; ++k) {
      T a[VEC], b[VEC];
      lds_vec<T>(rp + k * C::BOXZ, a);
      lds_vec<T>(vp + k * C::BOXZ, b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[k][e] = a[e] - alpha * b[e];
    }
  };
  auto release = [&](int s) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&empty[s]);
  };

  T A[K::RY][VEC], B[K::RY][VEC], Cc[K::RY][VEC];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  mbar_wait(&full[0], 0);
  own_s(0, A);
  release(0);
  mbar_wait(&full[1 % C::S], 0);
  own_s(1 % C::S, B);

  auto step = [&](T (&vm)[K::RY][VEC], T (&vc)[K::RY][VEC], T (&vp)[K::RY][VEC], int x, int i) {
    const int sn = i & (C::S - 1), sc = (i - 1) & (C::S - 1);
    mbar_wait(&full[sn], (i / C::S) & 1);
    own_s(sn, vp);
    const bool xreg = x >= g.lo[0] && x < g.hi[0];
    const bool xown = x >= g.olo0 && x < g.ohi0;
    T ax[K::RY][VEC];
    StepSum<T> s0(a0), s1(a1), s2(a2), s3(a3);
    if (xreg) {
      const T* rp = rt(sc);
      const T* vq = vt(sc);
      T up[VEC], dn[VEC], zl[K::RY], zr[K::RY];
      if (!K::FLAT) {
        T a[VEC], b[VEC];
        lds_vec<T>(rp - C::BOXZ, a);
        lds_vec<T>(vq - C::BOXZ, b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) up[e] = a[e] - alpha * b[e];
        lds_vec<T>(rp + K::RY * C::BOXZ, a);
        lds_vec<T>(vq + K::RY * C::BOXZ, b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) dn[e] = a[e] - alpha * b[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) up[e] = dn[e] = (T)0;
      }
#pragma unroll
      for (int k = 0; k < K::RY; ++k) {
        zl[k] = rp[k * C::BOXZ - 1] - alpha * vq[k * C::BOXZ - 1];
        zr[k] = rp[k * C::BOXZ + VEC] - alpha * vq[k * C::BOXZ + VEC];
      }
      if (!LEAN && p.wrap) {
        const T* rg = static_cast<const T*>(p.src0);
        const T* vg = static_cast<const T*>(p.src1);
        wrap_halo<T, K>(g, c, x, up, dn, zl, zr, [&](long long i) { return rg[i] - alpha * vg[i]; });
      }
      star_cells_eq<T, K, LEAN, NOPS, false>(eq, c, coef_class(g, 0, x), vm, vc, vp, up, dn, zl, zr,
                                             [&](int k, int e, T v) { ax[k][e] = v; });
    }
#pragma unroll
    for (int k = 0; k < K::RY; ++k) {
      T r0v[VEC], o[VEC];
      lds_vec<T>(aux(sc) + c.ooff + k * C::OBOXZ, r0v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const bool in = xreg && (LEAN || ((c.inreg >> (k * VEC + e)) & 1u));
        const T sv = vc[k][e];
        const T a = in ? ax[k][e] : (T)0;
        o[e] = a;
        if (xown && (LEAN || ((c.valid >> (k * VEC + e)) & 1u))) {  // s == 0 outside the region
          const T q = sv * sv;
          s0.add(q);
        }
        if (in && xown) {
          const T q1 = a * sv, q2 = a * a, q3 = r0v[e] * a;
          s1.add(q1);
          s2.add(q2);
          s3.add(q3);
        }
      }
      stg_row<T, K, LEAN>(tp_ + (long long)k * g.n[2], c, k, o);
      if (sp_) stg_row<T, K, LEAN>(sp_ + (long long)k * g.n[2], c, k, vc[k]);
    }
    s0.flush();
    s1.flush();
    s2.flush();
    s3.flush();
    tp_ += n12;
    if (sp_) sp_ += n12;
    release(sc);
  };

  int x = x0, i = 2;
  if (LEAN) {
    while (true) {
      step(A, B, Cc, x, i);
      if (++x >= x1) break;
      ++i;
      step(B, Cc, A, x, i);
      if (++x >= x1) break;
      ++i;
      step(Cc, A, B, x, i);
      if (++x >= x1) break;
      ++i;
    }
  } else {
    for (; x < x1; ++x, ++i) {
      step(A, B, Cc, x, i);
#pragma unroll
      for (int k = 0; k < K::RY; ++k)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          A[k][e] = B[k][e];
          B[k][e] = Cc[k][e];
        }
    }
  }
  acc_out[0] = a0;
  acc_out[1] = a1;
  acc_out[2] = a2;
  acc_out[3] = a3;
}

template <typename T, typename K, int NOPS>
__global__ void __launch_bounds__(Pw2Cfg<T, K>::THREADS, 2)
k_bi_st_tma(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_v,
            const __grid_constant__ CUtensorMap tm_r0, TilePlan p, GridDev g, EqDev<T> eq, T* __restrict__ t_out,
            T* __restrict__ s_out, SolverState* st, double* partials, int stage) {
  typedef Pw2Cfg<T, K> C;
  extern __shared__ unsigned char smem_dyn[];
  if (st->done) return;
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  unsigned char* stages = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)C::S * C::STAGE);
  uint64_t* empty = full + C::S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y0 = blockIdx.y * C::TY, z0 = blockIdx.x * C::TZ;
  const int x0 = blockIdx.z * p.cx, x1 = min(x0 + p.cx, g.n[0]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C::CWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (warp == C::CWARPS) {
    if (lane == 0) {
      const int n = x1 - x0 + 2;
      for (int i = 0; i < n; ++i) {
        const int pl = x0 - 1 + i;
        const int s = i & (C::S - 1);
        if (i >= C::S) mbar_wait(&empty[s], ((i / C::S) - 1) & 1);
        const bool inner = (pl >= x0 && pl < x1);
        mbar_expect_tx(&full[s], (uint32_t)(2 * C::HALO_BYTES + (inner ? C::OWN_BYTES : 0)));
        unsigned char* sb = stages + (size_t)s * C::STAGE;
        const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
        for (int b = 0; b < C::NB; ++b) {
          const int zb = z0 + b * C::OBOXZ;
          tma_load_3d(sb + b * C::HBOX_SLOT, &tm_r, zb - C::HZ, C::FLAT ? 0 : y0 - 1, xw, &full[s]);
          tma_load_3d(sb + C::HALO_SLOT + b * C::HBOX_SLOT, &tm_v, zb - C::HZ, C::FLAT ? 0 : y0 - 1, xw, &full[s]);
          if (inner) tma_load_3d(sb + 2 * C::HALO_SLOT + b * C::OBOX_SLOT, &tm_r0, zb, y0, xw, &full[s]);
        }
      }
    }
  } else {
    const T alpha = (T)st->scal[S_ALPHA];
    const bool full_tile = (C::FLAT || y0 + C::TY <= g.n[1]) && (z0 + C::TZ <= g.n[2]);
    const bool edge = (!C::FLAT && ((y0 < 2) || (y0 + C::TY > g.n[1] - 2))) || (z0 < 2) || (z0 + C::TZ > g.n[2] - 2);
    if (full_tile && !edge)
      pw2_consumer<T, K, true, NOPS>(p, g, eq, t_out, s_out, alpha, stages, full, empty, y0, z0, x0, x1, acc);
    else
      pw2_consumer<T, K, false, NOPS>(p, g, eq, t_out, s_out, alpha, stages, full, empty, y0, z0, x0, x1, acc);
  }
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  P2PDev pp = p.p2p;
  pp.slot0 = R_A;
  pp.count = 4;
  grid_reduce<4>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 4>{st, R_A, stage, 0, pp});
}

// ---- host ------------------------------------------------------------------------------------
// TilePlan::uni: bit a set = the three coefficient classes of every operator are bitwise equal on kernel axis a
template <typename T>
inline int eq_uniform(const GridDev& g, const EqDev<T>& eq) {
  int mask = 0;
  for (int a = 0; a < 3; ++a) {
    bool same = true;
    for (int k = 0; k < eq.nops; ++k)
      for (int cls = 1; cls < 3; ++cls)
        if (std::memcmp(&eq.op[k].coef[a][cls][0], &eq.op[k].coef[a][0][0], 3 * sizeof(T)) != 0) same = false;
    if (same || !g.act[a]) mask |= 1 << a;
  }
  return mask;
}

template <typename T>
inline bool pw_eligible(const GridDev& g, const pa_equation& eq, int nfaces, const pa_face_bc* faces) {
  constexpr int VEC = VecOf<T>::N;
  if (eq.nops < 1 || eq.nops > PA_MAX_OPS) return false;
  for (int k = 0; k < eq.nops; ++k) {
    const pa_op& o = eq.ops[k];
    if (o.kind != PA_OP_STAR || o.param_field != nullptr || o.edge != 0 || o.coef_tab[0] || o.coef_tab[1] ||
        o.coef_tab[2])
      return false;
  }
  if (!g.act[2] || (!g.act[1] && !g.act[0])) return false;
  if (g.n[2] % VEC != 0 || g.n[2] < 2 * VEC) return false;
  if (g.act[1] && g.n[1] < 4) return false;
  int wrap = 0;
  if (!tma_wrap_ok<T>(g, nfaces, faces, &wrap)) return false;
  return encode_tiled_fn() != nullptr;
}

template <typename T>
inline void pw_tile_plan(const GridDev& g, TilePlan& p, int nfaces = 0, const pa_face_bc* faces = nullptr) {
  const bool flat = tma_flat(g);
  const int ty = flat ? 1 : PwCfg<T, KStd>::TY;
  const int tz = flat ? PwCfg<T, KFlat>::TZ : PwCfg<T, KStd>::TZ;
  p.ry = flat ? 1 : KStd::RY;
  p.tiles_y = flat ? 1 : (g.n[1] + ty - 1) / ty;
  p.tiles_z = (g.n[2] + tz - 1) / tz;
  tma_chunks(g, p.tiles_y * p.tiles_z, p);
  tma_wrap_ok<T>(g, nfaces, faces, &p.wrap);
}

// One launch of the engine.  `in` is the stencilled field, `aux` rhs / r0 (may be null).
template <typename T, typename K, int MODE, int NOPS>
static void launch_star_tma_n(cudaStream_t s, const CUtensorMap& tm_in, const CUtensorMap& tm_aux,
                              const GridDev& g, const EqDev<T>& eq, const TilePlan& tile, bool has_aux, T* out,
                              T* out2, T dt, SolverState* st, double* partials, int stage) {
  typedef PwCfg<T, K, MODE> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_star_tma<T, K, MODE, NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(tile.tiles_z, tile.tiles_y, tile.chunks);
  k_star_tma<T, K, MODE, NOPS><<<grid, C::THREADS, C::SMEM, s>>>(tm_in, tm_aux, tile, g, eq, out, out2, dt,
                                                                has_aux ? 1 : 0, st, partials, stage);
}

template <typename T, typename K, int MODE>
static bool launch_star_tma_k(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile_in,
                              const T* in, const T* aux, T* out, T* out2, T dt, SolverState* st,
                              double* partials, int stage) {
  typedef PwCfg<T, K> C;
  TilePlan tile = tile_in;
  tile.src0 = in;  // wrap-around reads of periodic axes 1/2
  tile.uni = getenv("PA_STAR_NO_UNI") == nullptr ? eq_uniform<T>(g, eq) : 0;
  CUtensorMap tm_in, tm_aux;
  if (!make_map<T>(&tm_in, in, g, C::BOXZ, C::BOXY)) return false;
  if (!make_map<T>(&tm_aux, aux ? aux : in, g, C::OBOXZ, C::TY)) return false;
  // operator count known at compile time for the common equations (1: Poisson, 2: adv-diff)
  if (eq.nops == 1)
    launch_star_tma_n<T, K, MODE, 1>(s, tm_in, tm_aux, g, eq, tile, aux != nullptr, out, out2, dt, st, partials, stage);
  else if (eq.nops == 2)
    launch_star_tma_n<T, K, MODE, 2>(s, tm_in, tm_aux, g, eq, tile, aux != nullptr, out, out2, dt, st, partials, stage);
  else
    launch_star_tma_n<T, K, MODE, 0>(s, tm_in, tm_aux, g, eq, tile, aux != nullptr, out, out2, dt, st, partials, stage);
  return true;
}

template <typename T, typename K, int NOPS>
static void launch_bi_st_n(cudaStream_t s, const CUtensorMap& tm_r, const CUtensorMap& tm_v, const CUtensorMap& tm_r0,
                           const GridDev& g, const EqDev<T>& eq, const TilePlan& tile, T* t_out, T* s_out,
                           SolverState* st, double* partials, int stage) {
  typedef Pw2Cfg<T, K> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_bi_st_tma<T, K, NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attr = true;
  }
  dim3 grid(tile.tiles_z, tile.tiles_y, tile.chunks);
  k_bi_st_tma<T, K, NOPS><<<grid, C::THREADS, C::SMEM, s>>>(tm_r, tm_v, tm_r0, tile, g, eq, t_out, s_out, st, partials,
                                                           stage);
}

template <typename T, typename K>
static bool launch_bi_st_k(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile_in, const T* r,
                           const T* v, const T* r0, T* t_out, T* s_out, SolverState* st, double* partials,
                           int stage) {
  typedef Pw2Cfg<T, K> C;
  TilePlan tile = tile_in;
  tile.src0 = r;
  tile.src1 = v;
  CUtensorMap tm_r, tm_v, tm_r0;
  if (!make_map<T>(&tm_r, r, g, C::BOXZ, C::BOXY) || !make_map<T>(&tm_v, v, g, C::BOXZ, C::BOXY) ||
      !make_map<T>(&tm_r0, r0, g, C::OBOXZ, C::TY))
    return false;
  if (eq.nops == 1)
    launch_bi_st_n<T, K, 1>(s, tm_r, tm_v, tm_r0, g, eq, tile, t_out, s_out, st, partials, stage);
  else if (eq.nops == 2)
    launch_bi_st_n<T, K, 2>(s, tm_r, tm_v, tm_r0, g, eq, tile, t_out, s_out, st, partials, stage);
  else
    launch_bi_st_n<T, K, 0>(s, tm_r, tm_v, tm_r0, g, eq, tile, t_out, s_out, st, partials, stage);
  return true;
}

// BiCGSTAB s/t half step (k_bi_st_tma); s_out may be null (s is then recomputed by the x update)
template <typename T>
bool launch_bi_st_tma(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile, const T* r,
                      const T* v, const T* r0, T* t_out, T* s_out, SolverState* st, double* partials, int stage) {
  return tma_flat(g) ? launch_bi_st_k<T, KFlat>(s, g, eq, tile, r, v, r0, t_out, s_out, st, partials, stage)
                     : launch_bi_st_k<T, KStd>(s, g, eq, tile, r, v, r0, t_out, s_out, st, partials, stage);
}

// ---- explicit operators on every cell (PW_APPLY / PW_GRAD) ------------------------------------------
// Constant-coefficient star operators on 2-D / 3-D grids whose contiguous axis is a whole number of
// 16-byte vectors; everything else (field advection, Tensor coefficient, rz tables, 1-D, odd n2) stays
// on k_apply / k_grad.  edge=True is NOT a reason to leave: the one-sided face values are written by the
// shell kernels (kernels_generic.cuh k_apply_shell / k_grad_shell) after the main pass.
template <typename T>
inline bool apply_eligible(const GridDev& g, const pa_equation& eq) {
  constexpr int VEC = VecOf<T>::N;
  if (eq.nops < 1 || eq.nops > PA_MAX_OPS) return false;
  for (int k = 0; k < eq.nops; ++k) {
    const pa_op& o = eq.ops[k];
    if (o.kind != PA_OP_STAR || o.param_field != nullptr || o.coef_tab[0] || o.coef_tab[1] || o.coef_tab[2])
      return false;
  }
  if (!g.act[2] || (!g.act[1] && !g.act[0])) return false;
  if (g.n[2] % VEC != 0 || g.n[2] < 2 * VEC) return false;
  if (g.act[1] && g.n[1] < 4) return false;
  if (g.act[0] && g.n[0] < 3) return false;
  return encode_tiled_fn() != nullptr;
}

template <typename T>
inline void apply_tile_plan(const GridDev& g, TilePlan& p) {
  pw_tile_plan<T>(g, p);
  p.wrap = 1;  // edge tiles always fetch the wrapped halo row / column (wrap_halo<ALL>)
}

template <typename T, typename K>
static bool launch_star_grad_k(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile_in,
                               const T* in, T* out) {
  typedef PwCfg<T, K> C;
  TilePlan tile = tile_in;
  tile.src0 = in;
  tile.uni = getenv("PA_STAR_NO_UNI") == nullptr ? eq_uniform<T>(g, eq) : 0;
  CUtensorMap tm_in;
  if (!make_map<T>(&tm_in, in, g, C::BOXZ, C::BOXY)) return false;
  launch_star_tma_n<T, K, PW_GRAD, 1>(s, tm_in, tm_in, g, eq, tile, false, out, nullptr, (T)0, nullptr, nullptr, 0);
  return true;
}

// out[(axis), cell] = param * d(in)/dx_axis for every active mesh axis (eq.op[0] is the Grad star)
template <typename T>
bool launch_star_grad(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile, const T* in,
                      T* out) {
  return tma_flat(g) ? launch_star_grad_k<T, KFlat>(s, g, eq, tile, in, out)
                     : launch_star_grad_k<T, KStd>(s, g, eq, tile, in, out);
}

template <typename T, int MODE>
bool launch_star_tma(cudaStream_t s, const GridDev& g, const EqDev<T>& eq, const TilePlan& tile,
                     const T* in, const T* aux, T* out, T* out2, T dt, SolverState* st,
                     double* partials, int stage) {
  return tma_flat(g) ? launch_star_tma_k<T, KFlat, MODE>(s, g, eq, tile, in, aux, out, out2, dt, st, partials, stage)
                     : launch_star_tma_k<T, KStd, MODE>(s, g, eq, tile, in, aux, out, out2, dt, st, partials, stage);
}

}  // namespace pa
