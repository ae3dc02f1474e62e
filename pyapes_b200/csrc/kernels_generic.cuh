// Generic (any equation, any BC mix, 1-D/2-D/3-D) kernels: one thread per cell, neighbours
// read straight from global memory.  They define the semantics; the tiled kernels in
// kernels_tiled.cuh are the fast path for the single-STAR-operator case and must agree
// with these bit for bit.
#pragma once
#include "common.cuh"

namespace pa {

// stages of the device-side scalar recurrences
enum Stage {
  ST_NONE = 0,
  ST_CG_INIT,
  ST_CG_DAD,
  ST_CG_UPD,   // leaves raw sums for ST_CG_FIN
  ST_CG_FIN,
  ST_BI_INIT,
  ST_BI_V,
  ST_BI_S,
  ST_BI_T,
  ST_BI_ST,    // fused s/t half step: sums |s|^2, t.s, t.t, r0.t
  ST_BI_X,     // leaves raw sums for ST_BI_FIN
  ST_BI_FIN,
  ST_JA_UPD,
  ST_JA_FIN,
};

// raw-sum slots in SolverState::sum
enum { R_A = 0, R_B = 1, R_C = 2, R_SHELL = 3, R_RR = 4 };
// scalar slots in SolverState::scal
enum { S_ALPHA = 0, S_BETA = 1, S_OMEGA = 2, S_RHO = 3, S_RHO_NEXT = 4 };

template <typename T>
__device__ __forceinline__ void check_exit(SolverState* st, T tol) {
  // linalg.py:334-336 then 144-150 / 109
  if (isnan(tol) || isinf(tol)) {
    st->done = 1;
    st->status = PA_BAD_TOL;
  }
}

// Finalize a stage from the (already grid- and rank-reduced) raw sums in st->sum.
// Every scalar is rounded to the field dtype T where torch would hold a T tensor.
template <typename T>
__device__ void finalize_stage(int stage, SolverState* st) {
  double* sum = st->sum;
  double* sc = st->scal;
  switch (stage) {
    case ST_CG_INIT:  // rr = sum(r*r)
      sum[R_RR] = (double)(T)sum[R_A];
      break;
    case ST_CG_DAD: {  // alpha = nan_to_num(sum(r*r) / sum(d*Ad))      linalg.py:118-120
      T a = (T)sum[R_RR] / (T)sum[R_A];
      sc[S_ALPHA] = (double)nan_to_num0<T>(a);
    } break;
    case ST_CG_FIN: {
      // tol = ||x_new - x_old||_2 over the whole array                  linalg.py:134
      T tol = (T)sqrt(sum[R_B] + sum[R_SHELL]);
      st->tol = (double)tol;
      if (isnan(tol) || isinf(tol)) {
        st->done = 1;
        st->status = PA_BAD_TOL;
        break;
      }
      T rr_new = (T)sum[R_A];
      sc[S_BETA] = (double)(rr_new / (T)sum[R_RR]);  // linalg.py:137
      sum[R_RR] = (double)rr_new;
      st->itr += 1;  // linalg.py:144
      if (st->itr > st->max_it) {  // linalg.py:146-150
        st->done = 1;
        st->status = PA_MAXIT;
      } else if (!((double)tol > st->tolerance)) {  // linalg.py:109
        st->done = 1;
        st->status = PA_CONVERGED;
      }
    } break;
    case ST_BI_INIT: {  // linalg.py:201-206, then the head of the first iteration 212-214
      T rho_next = (T)sum[R_A];
      st->tol = (double)(T)sqrt((double)rho_next);
      sc[S_RHO] = 1.0;
      sc[S_ALPHA] = 1.0;
      sc[S_OMEGA] = 1.0;
      T beta = rho_next / (T)1 * (T)1 / (T)1;
      sc[S_BETA] = (double)beta;
      sc[S_RHO] = (double)rho_next;
    } break;
    case ST_BI_V: {  // linalg.py:222-225
      st->itr += 1;
      T a = (T)sc[S_RHO] / (T)sum[R_A];
      sc[S_ALPHA] = (double)nan_to_num0<T>(a);
    } break;
    case ST_BI_S: {  // tol = ||r - alpha v||                              linalg.py:233-240
      T tol = (T)sqrt(sum[R_A]);
      st->tol = (double)tol;
      if (isnan(tol) || isinf(tol)) {
        st->done = 1;
        st->status = PA_BAD_TOL;
        break;
      }
      st->finished_flag = ((double)tol <= st->tolerance) ? 1 : 0;
    } break;
    case ST_BI_ST: {  // ST_BI_S then ST_BI_T on the sums of the fused kernel (k_bi_st_tma)
      T tol = (T)sqrt(sum[R_A]);
      st->tol = (double)tol;
      if (isnan(tol) || isinf(tol)) {
        st->done = 1;
        st->status = PA_BAD_TOL;
        break;
      }
      st->finished_flag = ((double)tol <= st->tolerance) ? 1 : 0;
      if (st->finished_flag) break;
      T w = (T)sum[R_B] / (T)sum[R_C];
      w = nan_to_num0<T>(w);
      sc[S_OMEGA] = (double)w;
      sc[S_RHO_NEXT] = (double)((-w) * (T)sum[R_SHELL]);
      T beta = (T)sc[S_RHO_NEXT] / (T)sc[S_RHO] * (T)sc[S_ALPHA] / (T)sc[S_OMEGA];
      sc[S_BETA] = (double)beta;
      sc[S_RHO] = sc[S_RHO_NEXT];
    } break;
    case ST_BI_T: {  // omega, rho_next                                     linalg.py:246-250
      if (st->finished_flag) break;  // early exit taken after the first half step (linalg.py:235-240)
      T w = (T)sum[R_A] / (T)sum[R_B];
      w = nan_to_num0<T>(w);
      sc[S_OMEGA] = (double)w;
      sc[S_RHO_NEXT] = (double)((-w) * (T)sum[R_C]);
      // head of the next iteration (linalg.py:212-214): every input is known here already, and the
      // fused x/r/p update of this iteration needs beta for p = r_new + beta (p - omega v)
      T beta = (T)sc[S_RHO_NEXT] / (T)sc[S_RHO] * (T)sc[S_ALPHA] / (T)sc[S_OMEGA];
      sc[S_BETA] = (double)beta;
      sc[S_RHO] = sc[S_RHO_NEXT];
    } break;
    case ST_BI_FIN: {
      st->swaps += 1;  // this stage runs right after the x update of the iteration (never once `done` is set)
      if (st->finished_flag) {  // early exit path: `finished = True; continue`
        st->done = 1;
        st->status = PA_CONVERGED;
        break;
      }
      T tol = (T)sqrt(sum[R_A]);  // ||s - omega t||                       linalg.py:262
      st->tol = (double)tol;
      if (isnan(tol) || isinf(tol)) {
        st->done = 1;
        st->status = PA_BAD_TOL;
        break;
      }
      bool fin = ((double)tol <= st->tolerance);
      if (st->itr >= st->max_it) {  // linalg.py:268-271
        st->done = 1;
        st->status = PA_MAXIT;
      } else if (fin) {
        st->done = 1;
        st->status = PA_CONVERGED;
      }
    } break;
    case ST_JA_FIN: {
      T tol = (T)sqrt(sum[R_B] + sum[R_SHELL]);
      st->tol = (double)tol;
      if (isnan(tol) || isinf(tol)) {
        st->done = 1;
        st->status = PA_BAD_TOL;
        break;
      }
      st->itr += 1;
      if (st->itr > st->max_it) {
        st->done = 1;
        st->status = PA_MAXIT;
      } else if (!((double)tol > st->tolerance)) {
        st->done = 1;
        st->status = PA_CONVERGED;
      }
    } break;
    default:
      break;
  }
}

template <typename T>
__global__ void k_finalize(int stage, SolverState* st) {
  if (st->done) return;
  finalize_stage<T>(stage, st);
}

// store raw sums; optionally finalize in place (single-GPU: no allreduce in between)
template <typename T, int NS>
struct StoreSums {
  SolverState* st;
  int slot0;
  int stage;  // ST_NONE: just store
  int accum = 0;  // add to what an earlier sub-launch stored
  P2PDev p2p = {nullptr, 0, 0, 0, 0};  // multi-GPU: sum over ranks through peer memory before `stage`
  __device__ void operator()(double (&acc)[NS]) const {
#pragma unroll
    for (int s = 0; s < NS; ++s) st->sum[slot0 + s] = accum ? st->sum[slot0 + s] + acc[s] : acc[s];
    if (p2p.peers != nullptr) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < p2p.count; ++k) v[k] = st->sum[p2p.slot0 + k];
      if (!p2p_allreduce(p2p, st->epoch, v)) {  // a peer never arrived (watchdog, common.cuh)
        st->done = 1;
        st->status = PA_PEER_LOST;
        return;
      }
      for (int k = 0; k < p2p.count; ++k) st->sum[p2p.slot0 + k] = v[k];
      st->epoch += 1ull;
    }
    if (stage != ST_NONE) finalize_stage<T>(stage, st);
  }
};

// ---------------------------------------------------------------------------------------
// operator application on every cell (roll semantics) — ops._Aop
// ---------------------------------------------------------------------------------------
// one-sided first difference along axis a at a face cell: (3/2 v0 - 2 v1 + 1/2 v2)
// (fdc.py:270-284); `dir` = +1 on the lower face (neighbours at +1,+2), -1 on the upper face
template <typename T>
__device__ __forceinline__ T edge_first(const T* __restrict__ phi, long long idx, long long st, int dir) {
  T t = (T)1.5 * phi[idx];
  t = t - (T)2 * phi[idx + dir * st];
  t = t + (T)0.5 * phi[idx + 2 * dir * st];
  return t;
}

// the operator sum at one cell incl. the edge=True face formulas
template <typename T>
__device__ __forceinline__ T apply_cell(const GridDev& g, const EqDev<T>& eq, const T* __restrict__ phi,
                                        const Cell& c) {
  const long long idx = c.idx;
  T val = eval_equation<T>(g, eq, c, [&](long long j) { return phi[j]; });
  const OpDev<T>& o = eq.op[0];
  if (o.edge == 1) {
    // edge=True Laplacian (fdc.py:223-258): replace by the one-sided second derivative along
    // the face's axis; axes in order, so the last axis wins on shared edges
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!g.act[a]) continue;
      // (axis 0 of a slab: the GLOBAL plane index decides -- local plane 0 may be a ghost plane)
      const int i = (a == 0) ? c.i[0] + g.goff0 : c.i[a], n = (a == 0) ? g.gn0 : g.n[a];
      if (i != 0 && i != n - 1) continue;
      const long long st = stride_of(g, a) * (i == 0 ? 1 : -1);
      T t = (T)2 * phi[idx];
      t = t - (T)5 * phi[idx + st];
      t = t + (T)4 * phi[idx + 2 * st];
      t = t - phi[idx + 3 * st];
      val = t / (o.dx[a] * o.dx[a]);
    }
  } else if (o.edge == 2) {
    // edge=True Div on a 1-D mesh (fdc.py:290-348): the only active axis is kernel axis 2
    const int i = c.i[2], n = g.n[2];
    if (i == 0) {
      T t = -edge_first<T>(phi, idx, 1, 1);
      val = t / o.dx[2] * o.adv_const;
    } else if (i == n - 1) {
      T t = edge_first<T>(phi, idx, 1, -1);
      val = t / o.dx[2] * o.adv_const;
    }
  }
  return val;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_apply(GridDev g, EqDev<T> eq, const T* __restrict__ phi,
                                                  T* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    out[idx] = apply_cell<T>(g, eq, phi, c);
  }
}

// explicit gradient at one cell: one STAR operator, component per active axis (fdc.py:80-87)
template <typename T>
__device__ __forceinline__ void grad_cell(const GridDev& g, const OpDev<T>& o, const T* __restrict__ phi,
                                          T* __restrict__ out, const Cell& c) {
  const long long idx = c.idx;
  T vc = phi[idx];
  int comp = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (!g.act[a]) continue;
    long long st = stride_of(g, a);
    int i = c.i[a], n = g.n[a];
    T vp = phi[(i + 1 == n) ? idx - (long long)(n - 1) * st : idx + st];
    T vm = phi[(i == 0) ? idx + (long long)(n - 1) * st : idx - st];
    int cls = coef_class(g, a, i);
    const T* ct = o.coef_tab[a] != nullptr ? o.coef_tab[a] + 3 * i : &o.coef[a][cls][0];
    T s = ct[0] * vp;
    s = s + ct[1] * vc;
    s = s + ct[2] * vm;
    if (o.edge) {  // edge=True Grad (fdc.py:260-288); axis 0 of a slab: the GLOBAL plane index decides
      const int gi = (a == 0) ? i + g.goff0 : i, gn = (a == 0) ? g.gn0 : n;
      if (gi == 0)
        s = -edge_first<T>(phi, idx, st, 1) / o.dx[a];
      else if (gi == gn - 1)
        s = edge_first<T>(phi, idx, st, -1) / o.dx[a];
    }
    if (o.param_field != nullptr)
      s = s * o.param_field[idx];
    else if (o.has_param)
      s = s * o.param;
    out[(long long)comp * g.cells + idx] = s;
    ++comp;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_grad(GridDev g, OpDev<T> o, const T* __restrict__ phi,
                                                 T* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    grad_cell<T>(g, o, phi, out, c);
  }
}

// The outer shell only (every cell with an index 0 or n-1 along an active axis, each once): rewrites the
// face cells after the TMA main pass when edge=True (one-sided formulas, fdc.py:203-366).  blockIdx.y =
// face id (axis*2 + side); a cell belongs to the first axis (in order) on whose boundary it lies.
template <typename T, bool GRAD>
__global__ void __launch_bounds__(kBlock) k_apply_shell(GridDev g, EqDev<T> eq, const T* __restrict__ phi,
                                                        T* __restrict__ out) {
  const int face = blockIdx.y, ax = face >> 1, up = face & 1;
  if (!g.act[ax]) return;
  if (up && g.n[ax] == 1) return;
  const int bb = (ax == 0) ? 1 : 0, cc = (ax == 2) ? 1 : 2;
  const long long ncell = (long long)g.n[bb] * g.n[cc];
  // axis 0 of a slab: the local plane that holds the global face, if this rank has it
  const int plane = (ax == 0) ? (up ? g.gn0 - 1 : 0) - g.goff0 : (up ? g.n[ax] - 1 : 0);
  if (plane < 0 || plane >= g.n[ax]) return;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ncell;
       k += (long long)gridDim.x * blockDim.x) {
    int ib = (int)(k / g.n[cc]), ic = (int)(k - (long long)ib * g.n[cc]);
    Cell c;
    c.i[ax] = plane;
    c.i[bb] = ib;
    c.i[cc] = ic;
    c.idx = ((long long)c.i[0] * g.n[1] + c.i[1]) * g.n[2] + c.i[2];
    bool dup = false;
    for (int e = 0; e < ax; ++e) {
      if (!g.act[e]) continue;
      const int gi = (e == 0) ? c.i[0] + g.goff0 : c.i[e], gn = (e == 0) ? g.gn0 : g.n[e];
      dup |= (gi == 0) | (gi == gn - 1);
    }
    if (dup) continue;
    if (GRAD)
      grad_cell<T>(g, eq.op[0], phi, out, c);
    else
      out[c.idx] = apply_cell<T>(g, eq, phi, c);
  }
}

// ---------------------------------------------------------------------------------------
// boundary conditions: one launch per face, in list order (bcs.py:197-280)
// ---------------------------------------------------------------------------------------
// value of face cell k (column `base` of the two other axes) of face f, from the current phi
template <typename T>
__device__ __forceinline__ T bc_face_value(const FaceDev<T>& f, const T* phi, long long base, long long k, long long sa,
                                           int n) {
  // planes: face, 1 and 2 inward, 1 and 2 "forward" (wrapped), as the rolled masks give
  const int pf = f.side < 0 ? 0 : n - 1;
  const int p1 = ((pf - f.side) % n + n) % n, p2 = ((pf - 2 * f.side) % n + n) % n;
  const int f1 = ((pf + f.side) % n + n) % n, f2 = ((pf + 2 * f.side) % n + n) % n;
  switch (f.kind) {
    case PA_BC_DIRICHLET:
      return f.values ? f.values[k] : f.value;
    case PA_BC_NEUMANN: {
      T cterm = f.values ? f.values[k] : f.value;
      T t = (T)(4.0 / 3.0) * phi[base + p1 * sa];
      t = t - (T)(1.0 / 3.0) * phi[base + p2 * sa];
      return t + cterm;
    }
    case PA_BC_SYMMETRY:
      return phi[base + p1 * sa];
    default:  // periodic
      if (f.side < 0) {
        T t = phi[base + p1 * sa] - phi[base + f1 * sa];
        return t + phi[base + f2 * sa];
      }
      return phi[base + f1 * sa];
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_bc_face(GridDev g, FaceDev<T> f, T* __restrict__ phi,
                                                    const SolverState* st) {
  if (st != nullptr && st->done) return;
  const int a = f.axis;
  const int b = (a == 0) ? 1 : 0, c = (a == 2) ? 1 : 2;  // the two other axes, b outer
  const long long ncell = (long long)g.n[b] * g.n[c];
  const long long sa = stride_of(g, a), sb = stride_of(g, b), sc = stride_of(g, c);
  const int n = g.n[a];
  const int pf = f.side < 0 ? 0 : n - 1;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ncell;
       k += (long long)gridDim.x * blockDim.x) {
    long long ib = k / g.n[c], ic = k - ib * g.n[c];
    long long base = ib * sb + ic * sc;
    phi[base + pf * sa] = bc_face_value<T>(f, phi, base, k, sa, n);
  }
}

// Two consecutive faces of the list that are the two sides of ONE axis, as one launch (api.cu launch_bcs decides when
// that keeps the list-order semantics): non-periodic pairs touch and read disjoint planes (n >= 5), blockIdx.y picks
// the face; a periodic pair (lower, then upper) is done column by column by one thread -- the upper face is the new
// lower face value (bcs.py:262-280).
template <typename T>
__global__ void __launch_bounds__(kBlock) k_bc_face_pair(GridDev g, FaceDev<T> f0, FaceDev<T> f1, int periodic_pair,
                                                         T* phi, const SolverState* st) {
  if (st != nullptr && st->done) return;
  const int a = f0.axis;
  const int b = (a == 0) ? 1 : 0, c = (a == 2) ? 1 : 2;
  const long long ncell = (long long)g.n[b] * g.n[c];
  const long long sa = stride_of(g, a), sb = stride_of(g, b), sc = stride_of(g, c);
  const int n = g.n[a];
  if (periodic_pair && blockIdx.y != 0) return;
  const FaceDev<T>& f = (blockIdx.y == 0) ? f0 : f1;
  const int pf = f.side < 0 ? 0 : n - 1;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ncell;
       k += (long long)gridDim.x * blockDim.x) {
    long long ib = k / g.n[c], ic = k - ib * g.n[c];
    long long base = ib * sb + ic * sc;
    const T v = bc_face_value<T>(f, phi, base, k, sa, n);
    phi[base + pf * sa] = v;
    if (periodic_pair) phi[base + (long long)(f1.side < 0 ? 0 : n - 1) * sa] = v;  // upper = the new lower value
  }
}

// ---------------------------------------------------------------------------------------
// shell norm: sum over cells on the outer shell of (a - b)^2, each cell once, then finalize.
// blockIdx.y = face id (axis*2 + side).  A cell belongs to the first axis (in order) on
// whose boundary it lies.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock) k_shell_norm(GridDev g, const T* __restrict__ a,
                                                       const T* __restrict__ b, SolverState* st,
                                                       double* partials, int stage, P2PDev p2p) {
  if (st->done) return;
  double v[1] = {0.0};
  const int face = blockIdx.y, ax = face >> 1, up = face & 1;
  if (g.act[ax]) {
    const int bb = (ax == 0) ? 1 : 0, cc = (ax == 2) ? 1 : 2;
    const long long ncell = (long long)g.n[bb] * g.n[cc];
    int plane;  // local index of the face plane, -1 if this rank does not own it
    if (ax == 0) {
      int gi = up ? g.gn0 - 1 : 0;
      plane = gi - g.goff0;
      if (plane < g.olo0 || plane >= g.ohi0) plane = -1;
      if (up && g.gn0 == 1) plane = -1;
    } else {
      plane = up ? g.n[ax] - 1 : 0;
      if (up && g.n[ax] == 1) plane = -1;
    }
    if (plane >= 0) {
      for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ncell;
           k += (long long)gridDim.x * blockDim.x) {
        int ib = (int)(k / g.n[cc]), ic = (int)(k - (long long)ib * g.n[cc]);
        Cell c;
        c.i[ax] = plane;
        c.i[bb] = ib;
        c.i[cc] = ic;
        c.idx = ((long long)c.i[0] * g.n[1] + c.i[1]) * g.n[2] + c.i[2];
        if (!owned(g, c)) continue;
        // skip cells already counted by an earlier axis' faces
        bool dup = false;
        for (int e = 0; e < ax; ++e) {
          if (!g.act[e]) continue;
          int gi = (e == 0) ? c.i[0] + g.goff0 : c.i[e];
          int gn = (e == 0) ? g.gn0 : g.n[e];
          dup |= (gi == 0) | (gi == gn - 1);
        }
        // lower and upper face of a 1-cell-thick... cannot happen (n >= 3, linalg.py:44-45)
        if (dup) continue;
        T d = a[c.idx] - b[c.idx];
        T q = d * d;
        v[0] += (double)q;
      }
    }
  }
  int nblocks = gridDim.x * gridDim.y;
  int bid = blockIdx.y * gridDim.x + blockIdx.x;
  grid_reduce<1>(v, partials, nblocks, bid, &st->ticket[3], StoreSums<T, 1>{st, R_SHELL, stage, 0, p2p});
}

// ---------------------------------------------------------------------------------------
// CG (linalg.py:74-159)
// ---------------------------------------------------------------------------------------
// r = rhs - A(x) on the solver region, 0 elsewhere; d = r; rr = sum r*r
template <typename T>
__global__ void __launch_bounds__(kBlock) k_residual_init(GridDev g, EqDev<T> eq,
                                                          const T* __restrict__ x,
                                                          const T* __restrict__ rhs,
                                                          T* __restrict__ r, T* __restrict__ d,
                                                          SolverState* st, double* partials,
                                                          int stage) {
  double v[1] = {0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    T res = (T)0;
    if (in_region(g, c)) {
      T ax = eval_equation<T>(g, eq, c, [&](long long j) { return x[j]; });
      res = rhs[idx] - ax;
      if (owned(g, c)) {
        T q = res * res;
        v[0] += (double)q;
      }
    }
    r[idx] = res;
    if (d != nullptr) d[idx] = res;
  }
  grid_reduce<1>(v, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 1>{st, R_A, stage});
}

// d = r + beta*d on the region (linalg.py:141)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_cg_dupdate(GridDev g, const T* __restrict__ r,
                                                       T* __restrict__ d, const SolverState* st) {
  if (st->done) return;
  const T beta = (T)st->scal[S_BETA];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    // ghost planes of a slab too: r's ghosts are exchanged, so d's ghosts follow bit for bit
    if (in_region(g, c) || !owned(g, c)) d[idx] = r[idx] + beta * d[idx];
  }
}

// dAd = sum d * A(d) over the region  (linalg.py:114-120); Ad is not stored
template <typename T>
__global__ void __launch_bounds__(kBlock) k_cg_dAd(GridDev g, EqDev<T> eq, const T* __restrict__ d,
                                                   SolverState* st, double* partials, int stage) {
  if (st->done) return;
  double v[1] = {0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    if (in_region(g, c) && owned(g, c)) {
      T ad = eval_equation<T>(g, eq, c, [&](long long j) { return d[j]; });
      T q = d[idx] * ad;
      v[0] += (double)q;
    }
  }
  grid_reduce<1>(v, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 1>{st, R_A, stage});
}

// x_new = x + alpha d ; r -= alpha Ad (Ad recomputed) ; sums rr_new and interior |dx|^2
template <typename T>
__global__ void __launch_bounds__(kBlock) k_cg_update(GridDev g, EqDev<T> eq,
                                                      const T* __restrict__ x,
                                                      T* __restrict__ x_new,
                                                      const T* __restrict__ d, T* __restrict__ r,
                                                      SolverState* st, double* partials) {
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA];
  double v[2] = {0.0, 0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    T xo = x[idx];
    T xn = xo;
    if (in_region(g, c)) {
      T ad = eval_equation<T>(g, eq, c, [&](long long j) { return d[j]; });
      xn = xo + alpha * d[idx];        // linalg.py:122
      T rn = r[idx] - alpha * ad;      // linalg.py:131
      r[idx] = rn;
      if (owned(g, c)) {
        T q = rn * rn;
        v[0] += (double)q;
      }
    }
    x_new[idx] = xn;
    if (owned(g, c) && !on_shell(g, c)) {
      T df = xn - xo;
      T q = df * df;
      v[1] += (double)q;
    }
  }
  grid_reduce<2>(v, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 2>{st, R_A, ST_NONE});
}

// ---------------------------------------------------------------------------------------
// BiCGSTAB (linalg.py:162-279), stored s and t
// ---------------------------------------------------------------------------------------
// p = r + beta (p - omega v)   (linalg.py:217)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_bi_p(GridDev g, const T* __restrict__ r,
                                                 T* __restrict__ p, const T* __restrict__ v,
                                                 const SolverState* st) {
  if (st->done) return;
  const T beta = (T)st->scal[S_BETA], omega = (T)st->scal[S_OMEGA];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    if (in_region(g, c)) {
      T t = p[idx] - omega * v[idx];
      p[idx] = r[idx] + beta * t;
    }
  }
}

// out = A(in) on the region; up to three dot products with stored vectors
//   mode 0 (v-stage): out=v, sums: r0.v
//   mode 1 (t-stage): out=t, sums: t.s, t.t, r0.t   (skipped when finished_flag)
template <typename T, int MODE>
__global__ void __launch_bounds__(kBlock) k_bi_apply(GridDev g, EqDev<T> eq,
                                                     const T* __restrict__ in, T* __restrict__ out,
                                                     const T* __restrict__ r0, SolverState* st,
                                                     double* partials, int stage) {
  if (st->done) return;
  if (MODE == 1 && st->finished_flag) return;
  double v[3] = {0.0, 0.0, 0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    if (in_region(g, c)) {
      T a = eval_equation<T>(g, eq, c, [&](long long j) { return in[j]; });
      out[idx] = a;
      if (owned(g, c)) {
        if (MODE == 0) {
          T q = r0[idx] * a;
          v[0] += (double)q;
        } else {
          T q0 = a * in[idx], q1 = a * a, q2 = r0[idx] * a;
          v[0] += (double)q0;
          v[1] += (double)q1;
          v[2] += (double)q2;
        }
      }
    }
  }
  grid_reduce<3>(v, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 3>{st, R_A, stage});
}

// s = r - alpha v ; ||s||^2   (linalg.py:230-233)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_bi_s(GridDev g, const T* __restrict__ r,
                                                 const T* __restrict__ v, T* __restrict__ s,
                                                 SolverState* st, double* partials, int stage) {
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA];
  double acc[1] = {0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    if (in_region(g, c)) {
      T sv = r[idx] - alpha * v[idx];
      s[idx] = sv;
      if (owned(g, c)) {
        T q = sv * sv;
        acc[0] += (double)q;
      }
    }
  }
  grid_reduce<1>(acc, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 1>{st, R_A, stage});
}

// x_new = x + alpha p (+ s omega) ; r = s - omega t ; ||r||^2   (linalg.py:236, 253-262)
// and, fused, the head of the NEXT iteration: p = r_new + beta (p - omega v)   (linalg.py:217)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_bi_x(GridDev g, const T* __restrict__ x,
                                                 T* __restrict__ x_new, T* __restrict__ p,
                                                 const T* __restrict__ s, const T* __restrict__ t,
                                                 const T* __restrict__ v, T* __restrict__ r, SolverState* st,
                                                 double* partials) {
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA], omega = (T)st->scal[S_OMEGA], beta = (T)st->scal[S_BETA];
  const bool early = st->finished_flag != 0;
  double acc[1] = {0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    T xn = x[idx];
    if (in_region(g, c)) {
      const T pv = p[idx];
      xn = xn + alpha * pv;
      if (!early) {
        T sv = s[idx];
        xn = xn + sv * omega;
        T rn = sv - omega * t[idx];
        r[idx] = rn;
        T tt = pv - omega * v[idx];
        p[idx] = rn + beta * tt;
        if (owned(g, c)) {
          T q = rn * rn;
          acc[0] += (double)q;
        }
      }
    }
    x_new[idx] = xn;
  }
  grid_reduce<1>(acc, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                 StoreSums<T, 1>{st, R_A, ST_NONE});
}

// Streaming (index-free) versions of the three axpy stages.  Outside the solver region r, p, v,
// s, t are identically 0, so the updates can run over every cell: 0 + beta*(0 - omega*0) == 0 and
// x + alpha*0 == x exactly.  16-byte vector accesses, grid-stride.  Single-GPU only (all cells
// owned).  `n` must be a multiple of the vector width; the caller falls back otherwise.
template <typename T>
struct StreamVec {
  static constexpr int N = 16 / sizeof(T);
  struct alignas(16) type {
    T v[16 / sizeof(T)];
  };
};

template <typename T>
__global__ void __launch_bounds__(kBlock) k_bi_p_stream(long long nvec, const T* __restrict__ r,
                                                        T* __restrict__ p, const T* __restrict__ v,
                                                        const SolverState* st) {
  typedef typename StreamVec<T>::type V;
  if (st->done) return;
  const T beta = (T)st->scal[S_BETA], omega = (T)st->scal[S_OMEGA];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    V a = reinterpret_cast<const V*>(r)[i], b = reinterpret_cast<const V*>(p)[i],
      c = reinterpret_cast<const V*>(v)[i], o;
#pragma unroll
    for (int e = 0; e < StreamVec<T>::N; ++e) {
      T t = b.v[e] - omega * c.v[e];
      o.v[e] = a.v[e] + beta * t;
    }
    reinterpret_cast<V*>(p)[i] = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_bi_s_stream(long long nvec, const T* __restrict__ r,
                                                        const T* __restrict__ v, T* __restrict__ s,
                                                        SolverState* st, double* partials, int stage) {
  typedef typename StreamVec<T>::type V;
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA];
  double acc[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    V a = reinterpret_cast<const V*>(r)[i], c = reinterpret_cast<const V*>(v)[i], o;
    T part = (T)0;  // one 16-byte vector: summed in T, added once (fp64: 2 values; fp32: 4, one conversion)
#pragma unroll
    for (int e = 0; e < StreamVec<T>::N; ++e) {
      T sv = a.v[e] - alpha * c.v[e];
      o.v[e] = sv;
      T q = sv * sv;
      part += q;
    }
    acc[0] += (double)part;
    reinterpret_cast<V*>(s)[i] = o;
  }
  grid_reduce<1>(acc, partials, gridDim.x, blockIdx.x, &st->ticket[0], StoreSums<T, 1>{st, R_A, stage});
}

// RS: s is not stored -- it is recomputed as r - alpha v from the (old) r, with the same operation
// as the s-stage, so the bits are the same (the fused k_bi_st_tma path, 15 words per iteration)
template <typename T, bool RS>
__global__ void __launch_bounds__(kBlock) k_bi_x_stream(long long nvec, const T* __restrict__ x,
                                                        T* __restrict__ x_new, T* __restrict__ p,
                                                        const T* __restrict__ s, const T* __restrict__ t,
                                                        const T* __restrict__ v, T* __restrict__ r,
                                                        SolverState* st, double* partials,
                                                        P2PDev p2p = P2PDev{nullptr, 0, 0, 0, 0}) {
  typedef typename StreamVec<T>::type V;
  if (st->done) return;
  const T alpha = (T)st->scal[S_ALPHA], omega = (T)st->scal[S_OMEGA], beta = (T)st->scal[S_BETA];
  const bool early = st->finished_flag != 0;
  double acc[1] = {0.0};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    V xv = reinterpret_cast<const V*>(x)[i], pv = reinterpret_cast<const V*>(p)[i], xo;
    if (early) {
#pragma unroll
      for (int e = 0; e < StreamVec<T>::N; ++e) xo.v[e] = xv.v[e] + alpha * pv.v[e];
    } else {
      V sv = reinterpret_cast<const V*>(RS ? r : s)[i], tv = reinterpret_cast<const V*>(t)[i],
        vv = reinterpret_cast<const V*>(v)[i], ro, po;
      if (RS) {
#pragma unroll
        for (int e = 0; e < StreamVec<T>::N; ++e) sv.v[e] = sv.v[e] - alpha * vv.v[e];  // linalg.py:230
      }
      T part = (T)0;
#pragma unroll
      for (int e = 0; e < StreamVec<T>::N; ++e) {
        T xn = xv.v[e] + alpha * pv.v[e];
        xo.v[e] = xn + sv.v[e] * omega;
        T rn = sv.v[e] - omega * tv.v[e];
        ro.v[e] = rn;
        T tt = pv.v[e] - omega * vv.v[e];
        po.v[e] = rn + beta * tt;  // next iteration's p (linalg.py:217), beta from ST_BI_T
        T q = rn * rn;
        part += q;
      }
      acc[0] += (double)part;
      reinterpret_cast<V*>(r)[i] = ro;
      reinterpret_cast<V*>(p)[i] = po;
    }
    reinterpret_cast<V*>(x_new)[i] = xo;
  }
  p2p.slot0 = R_A;
  p2p.count = 1;
  grid_reduce<1>(acc, partials, gridDim.x, blockIdx.x, &st->ticket[0], StoreSums<T, 1>{st, R_A, ST_NONE, 0, p2p});
}

// ---------------------------------------------------------------------------------------
// Jacobi and explicit Euler (not in the reference)
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T eq_diag(const GridDev& g, const EqDev<T>& eq, const Cell& c) {
  T res = (T)0;
  for (int k = 0; k < eq.nops; ++k) {
    const OpDev<T>& o = eq.op[k];
    T acc = (T)0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!g.act[a]) continue;
      T Ac;
      if (o.kind == PA_OP_STAR && o.coef_tab[a] != nullptr) {
        Ac = o.coef_tab[a][3 * c.i[a] + 1];
      } else if (o.kind == PA_OP_STAR) {
        Ac = o.coef[a][coef_class(g, a, c.i[a])][1];
      } else if (o.kind == PA_OP_DIV_UPWINDFD_FIELD) {
        T u = o.adv[c.idx];
        T up = u > (T)0 ? u : (T)0, um = u < (T)0 ? u : (T)0;
        Ac = (up - um) / o.dx[a];
      } else {
        Ac = (T)0;
      }
      acc = acc + Ac;
    }
    if (o.param_field != nullptr)
      acc = acc * o.param_field[c.idx];
    else if (o.has_param)
      acc = acc * o.param;
    acc = acc * o.sign;
    res = res + acc;
  }
  return res;
}

// MODE 0: Jacobi  x_new = x + (rhs - A x)/diag ; MODE 1: Euler  x_new = x + dt (rhs - A x)
template <typename T, int MODE>
__global__ void __launch_bounds__(kBlock) k_pointwise_update(GridDev g, EqDev<T> eq,
                                                             const T* __restrict__ x,
                                                             T* __restrict__ x_new,
                                                             const T* __restrict__ rhs, T dt,
                                                             SolverState* st, double* partials) {
  if (st != nullptr && st->done) return;
  double v[2] = {0.0, 0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < g.cells;
       idx += (long long)gridDim.x * blockDim.x) {
    Cell c = decode(g, idx);
    T xo = x[idx];
    T xn = xo;
    if (in_region(g, c)) {
      T ax = eval_equation<T>(g, eq, c, [&](long long j) { return x[j]; });
      T src = rhs ? rhs[idx] : (T)0;
      T res = src - ax;
      if (MODE == 0)
        xn = xo + res / eq_diag<T>(g, eq, c);
      else
        xn = xo + dt * res;
    }
    x_new[idx] = xn;
    if (MODE == 0 && owned(g, c) && !on_shell(g, c)) {
      T df = xn - xo;
      T q = df * df;
      v[1] += (double)q;
    }
  }
  if (MODE == 0)
    grid_reduce<2>(v, partials, gridDim.x, blockIdx.x, &st->ticket[0],
                   StoreSums<T, 2>{st, R_A, ST_NONE});
}

}  // namespace pa
