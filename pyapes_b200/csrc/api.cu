// C-ABI entry points + host-side iteration drivers.  See include/pyapes_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels_generic.cuh"
#include "kernels_tiled.cuh"
#include "kernels_tma.cuh"
#include "kernels_tma_pw.cuh"
#include "kernels_small.cuh"
#include "kernels_resident.cuh"
#include "dist.cuh"

namespace pa {

// The tiled / TMA kernels are instantiated in their own translation units (tu_*.cu), compiled
// in parallel by __graft_entry__.build(); here they are only declared.
#define PA_EXTERN(T)                                                                                           \
  extern template void launch_cg_phaseA<T>(cudaStream_t, const TilePlan&, const GridDev&, const EqDev<T>&,     \
                                           const T*, const T*, T*, SolverState*, double*);                     \
  extern template void launch_cg_phaseB<T>(cudaStream_t, const TilePlan&, const GridDev&, const EqDev<T>&,     \
                                           const T*, T*, const T*, T*, SolverState*, double*);                 \
  extern template void launch_cg_phaseA_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&,  \
                                               int, T*, SolverState*, double*);                                \
  extern template void launch_cg_phaseB_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&,  \
                                               int, T*, T*, SolverState*, double*, int);                       \
  extern template bool launch_cg_coop_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&, T*, \
                                             T*, T*, T*, T*, SolverState*, double*);                           \
  extern template bool launch_star_tma<T, PW_RESID>(cudaStream_t, const GridDev&, const EqDev<T>&,             \
                                                    const TilePlan&, const T*, const T*, T*, T*, T,            \
                                                    SolverState*, double*, int);                               \
  extern template bool launch_star_tma<T, PW_JACOBI>(cudaStream_t, const GridDev&, const EqDev<T>&,            \
                                                     const TilePlan&, const T*, const T*, T*, T*, T,           \
                                                     SolverState*, double*, int);                              \
  extern template bool launch_star_tma<T, PW_EULER>(cudaStream_t, const GridDev&, const EqDev<T>&,             \
                                                    const TilePlan&, const T*, const T*, T*, T*, T,            \
                                                    SolverState*, double*, int);                               \
  extern template bool launch_star_tma<T, PW_APPLY_V>(cudaStream_t, const GridDev&, const EqDev<T>&,           \
                                                      const TilePlan&, const T*, const T*, T*, T*, T,          \
                                                      SolverState*, double*, int);                             \
  extern template bool launch_star_tma<T, PW_APPLY_T>(cudaStream_t, const GridDev&, const EqDev<T>&,           \
                                                      const TilePlan&, const T*, const T*, T*, T*, T,          \
                                                      SolverState*, double*, int);
PA_EXTERN(double)
PA_EXTERN(float)
#define PA_EXTERN_APPLY(T)                                                                                     \
  extern template bool launch_star_tma<T, PW_APPLY>(cudaStream_t, const GridDev&, const EqDev<T>&,             \
                                                    const TilePlan&, const T*, const T*, T*, T*, T,            \
                                                    SolverState*, double*, int);                               \
  extern template bool launch_star_grad<T>(cudaStream_t, const GridDev&, const EqDev<T>&, const TilePlan&,     \
                                           const T*, T*);
PA_EXTERN_APPLY(double)
PA_EXTERN_APPLY(float)
#undef PA_EXTERN_APPLY
#define PA_EXTERN_RES(T)                                                                                           \
  extern template bool launch_euler_resident<T>(cudaStream_t, const GridDev&, const pa_equation&, const EqDev<T>&, \
                                                T*, T*, const T*, T, int);                                         \
  extern template bool launch_apply_direct2d<T, false>(cudaStream_t, const GridDev&, const EqDev<T>&, const T*, T*); \
  extern template bool launch_apply_direct2d<T, true>(cudaStream_t, const GridDev&, const EqDev<T>&, const T*, T*);  \
  extern template bool launch_jacobi_resident<T>(cudaStream_t, const GridDev&, const pa_equation&, const EqDev<T>&,  \
                                                 T*, T*, const T*, SolverState*, int);                               \
  extern template bool launch_cg_resident<T>(cudaStream_t, const GridDev&, const EqDev<T>&, T*, T*, const T*,      \
                                             const T*, SolverState*, int, int);
PA_EXTERN_RES(double)
PA_EXTERN_RES(float)
#undef PA_EXTERN_RES
extern template bool launch_bi_st_tma<double>(cudaStream_t, const GridDev&, const EqDev<double>&, const TilePlan&,
                                              const double*, const double*, const double*, double*, double*,
                                              SolverState*, double*, int);
extern template bool launch_bi_st_tma<float>(cudaStream_t, const GridDev&, const EqDev<float>&, const TilePlan&,
                                             const float*, const float*, const float*, float*, float*, SolverState*,
                                             double*, int);
#undef PA_EXTERN

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PA_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(PA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

static int check_grid(const pa_grid* g) {
  if (!g) return fail(PA_ERR_ARG, "grid is null");
  if (g->ndim < 1 || g->ndim > 3) return fail(PA_ERR_ARG, "grid.ndim must be 1..3");
  for (int a = 0; a < 3; ++a) {
    if (g->n[a] < 1) return fail(PA_ERR_ARG, "grid.n must be >= 1");
    const bool active = (a == 2) || (a == 0 && g->ndim >= 2) || (a == 1 && g->ndim == 3);
    if (!active && g->n[a] != 1) return fail(PA_ERR_ARG, "inactive kernel axes must have extent 1");
    if (g->lo[a] < 0 || g->hi[a] > g->n[a] || g->lo[a] > g->hi[a])
      return fail(PA_ERR_ARG, "grid.lo/hi out of range");
  }
  if (g->olo0 < 0 || g->ohi0 > g->n[0] || g->olo0 > g->ohi0)
    return fail(PA_ERR_ARG, "grid.olo0/ohi0 out of range");
  if ((long long)g->n[0] * g->n[1] * g->n[2] >= (1LL << 40))
    return fail(PA_ERR_ARG, "grid too large");
  return PA_OK;
}

static int check_eq(const pa_equation* eq) {
  if (!eq) return fail(PA_ERR_ARG, "equation is null");
  if (eq->nops < 1 || eq->nops > PA_MAX_OPS)
    return fail(PA_ERR_UNSUPPORTED, "equation must have 1..PA_MAX_OPS operators");
  for (int k = 0; k < eq->nops; ++k) {
    int kind = eq->ops[k].kind;
    if (kind < PA_OP_STAR || kind > PA_OP_DIV_UPWINDFD_FIELD)
      return fail(PA_ERR_ARG, "unknown operator kind");
    if (kind != PA_OP_STAR && eq->ops[k].adv == nullptr)
      return fail(PA_ERR_ARG, "field-advection operator without adv pointer");
    if (eq->ops[k].edge != 0 && eq->nops != 1)
      return fail(PA_ERR_ARG, "edge=True is an option of the explicit single-operator calls");
  }
  return PA_OK;
}

static int check_faces(int nfaces, const pa_face_bc* faces) {
  if (nfaces < 0 || nfaces > PA_MAX_FACES) return fail(PA_ERR_ARG, "nfaces must be 0..6");
  if (nfaces > 0 && !faces) return fail(PA_ERR_ARG, "faces is null");
  for (int f = 0; f < nfaces; ++f) {
    if (faces[f].axis < 0 || faces[f].axis > 2) return fail(PA_ERR_ARG, "face axis");
    if (faces[f].side != -1 && faces[f].side != 1) return fail(PA_ERR_ARG, "face side");
    if (faces[f].kind < PA_BC_DIRICHLET || faces[f].kind > PA_BC_PERIODIC)
      return fail(PA_ERR_ARG, "face kind");
  }
  return PA_OK;
}

// auto-selection threshold of the persistent small-grid CG kernel (cells); see DESIGN.md §4
// measured (tools/bench_small.py): 9.7-10.9 us per iteration against 16.2 us for the fused kernels up to
// 256^2 / 32^3, break-even near 110 k cells
constexpr long long kSmallCgCells = 80000;
// auto-selection window of the cooperative whole-solve TMA kernel (tools/bench_small.py, B200): 17-18 us per
// iteration against 19-20 for separate launches at 32^3 .. 64^3 and 36 against 38 at 1024^2, equal from 128^3 on
// (the iteration is bound by the latency of the plane-by-plane march inside a phase, not by the launches)
constexpr long long kCoopCgCells = 1500000;

static inline int grid_blocks(long long cells) {
  long long b = (cells + kBlock - 1) / kBlock;
  long long cap = (long long)kNumSMs * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// All faces Dirichlet (or none): after the initial BC application the shell of x never changes
// (the solver region excludes it and phase B copies it), so per-iteration BC launches and the
// shell norm are exact no-ops and can be skipped.
static int static_shell(int nfaces, const pa_face_bc* faces) {
  for (int f = 0; f < nfaces; ++f)
    if (faces[f].kind != PA_BC_DIRICHLET) return 0;
  return 1;
}

struct Launcher {
  cudaStream_t s;
  int count = 0;
  bool ok = true;  // false once a TMA launch could not be issued (tensor-map encode failed)
};

// device scratch of the slab-periodic boundary condition (two planes), grown on demand — never
// while a stream capture is active (run_solver / euler_impl reserve it before capturing)
static void* bc_scratch(size_t bytes) {
  static void* p_dev[kMaxDevices] = {};
  static size_t cap_dev[kMaxDevices] = {};
  void*& p = p_dev[current_device()];
  size_t& cap = cap_dev[current_device()];
  if (bytes > cap) {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) == cudaSuccess) cap = bytes;
  }
  return p;
}

// phi[face] = (phi[1 inward] - phi[N-1]) + phi[N-2]  with the two far planes in `far` (N-2 first)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_bc_periodic_lo_slab(long long n, T* __restrict__ face,
                                                                const T* __restrict__ inner,
                                                                const T* __restrict__ far, const SolverState* st) {
  if (st != nullptr && st->done) return;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    T t = inner[k] - far[n + k];
    face[k] = t + far[k];
  }
}

template <typename T>
static void launch_bcs(Launcher& L, const GridDev& g, int nfaces, const pa_face_bc* faces, T* phi,
                       const SolverState* st, const Dist* dist = nullptr) {
  for (int f = 0; f < nfaces; ++f) {
    FaceDev<T> fd;
    fd.axis = faces[f].axis;
    fd.side = faces[f].side;
    fd.kind = faces[f].kind;
    fd.value = (T)faces[f].value;
    fd.values = (const T*)faces[f].values;
    if (!g.act[fd.axis]) continue;
    if (fd.axis == 0 && fd.kind == PA_BC_PERIODIC && dist && dist->ring) {
      // periodic faces along the slab axis (bcs.py:262-280) need the other end of the ring, at
      // this position of the face list:
      //   lower: phi[0] = phi[1] - phi[N-1] + phi[N-2]   rank P-1 sends its last two planes to rank 0
      //   upper: phi[N-1] = phi[0]                        rank 0 sends plane 0 to rank P-1
      const long long plane = (long long)g.n[1] * g.n[2];
      const int last = dist->nranks - 1;
      if (fd.side < 0) {
        if (dist->rank == last) dist_send<T>(*dist, phi + (long long)(g.ohi0 - 2) * plane, 2 * plane, 0, L.s);
        if (dist->rank == 0) {
          T* far = (T*)bc_scratch((size_t)(2 * plane) * sizeof(T));
          if (!far) {
            L.ok = false;
            continue;
          }
          dist_recv<T>(*dist, far, 2 * plane, last, L.s);
          int blocks = (int)((plane + kBlock - 1) / kBlock);
          if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
          k_bc_periodic_lo_slab<T><<<blocks, kBlock, 0, L.s>>>(plane, phi + (long long)g.olo0 * plane,
                                                               phi + (long long)(g.olo0 + 1) * plane, far, st);
        }
      } else {
        if (dist->rank == 0) dist_send<T>(*dist, phi + (long long)g.olo0 * plane, plane, last, L.s);
        if (dist->rank == last) dist_recv<T>(*dist, phi + (long long)(g.ohi0 - 1) * plane, plane, 0, L.s);
      }
      ++L.count;
      continue;
    }
    // along the slab axis only the rank that owns the face plane applies it
    if (fd.axis == 0) {
      int gi = fd.side < 0 ? 0 : g.gn0 - 1;
      int pl = gi - g.goff0;
      if (pl < g.olo0 || pl >= g.ohi0) continue;
      // slab-decomposed: the rank that owns a global x face has no ghost plane on that side,
      // so the face plane is local plane 0 / n[0]-1 and k_bc_face's local indexing holds
      if (pl != (fd.side < 0 ? 0 : g.n[0] - 1)) continue;
    }
    int b = (fd.axis == 0) ? 1 : 0, c = (fd.axis == 2) ? 1 : 2;
    long long ncell = (long long)g.n[b] * g.n[c];
    int blocks = (int)((ncell + kBlock - 1) / kBlock);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    if (blocks < 1) blocks = 1;
    // The next face of the list is the other side of this axis: one launch for both when that keeps the list-order
    // semantics -- two non-periodic faces (they touch and read disjoint planes once n >= 5), or the periodic pair
    // lower-then-upper (the upper face is the new lower value, column by column).  Single GPU only (on slabs an x
    // face may belong to another rank).  6 -> 3 boundary launches per sweep for a full 3-D face list.
    if (!dist && f + 1 < nfaces && faces[f + 1].axis == faces[f].axis && faces[f + 1].side == -faces[f].side &&
        g.n[fd.axis] >= 5) {
      const bool p0 = faces[f].kind == PA_BC_PERIODIC, p1 = faces[f + 1].kind == PA_BC_PERIODIC;
      const bool pair_np = !p0 && !p1;
      const bool pair_per = p0 && p1 && faces[f].side < 0 && faces[f].values == nullptr && faces[f + 1].values == nullptr;
      if (pair_np || pair_per) {
        FaceDev<T> fe;
        fe.axis = faces[f + 1].axis;
        fe.side = faces[f + 1].side;
        fe.kind = faces[f + 1].kind;
        fe.value = (T)faces[f + 1].value;
        fe.values = (const T*)faces[f + 1].values;
        k_bc_face_pair<T><<<dim3(blocks, pair_per ? 1 : 2), kBlock, 0, L.s>>>(g, fd, fe, pair_per ? 1 : 0, phi, st);
        ++L.count;
        ++f;
        continue;
      }
    }
    k_bc_face<T><<<blocks, kBlock, 0, L.s>>>(g, fd, phi, st);
    ++L.count;
  }
}

template <typename T>
static void launch_shell(Launcher& L, const GridDev& g, const T* a, const T* b, SolverState* st,
                         double* partials, int stage, P2PDev p2p = P2PDev{nullptr, 0, 0, 0, 0}) {
  long long m = 1;
  for (int ax = 0; ax < 3; ++ax) {
    int bb = (ax == 0) ? 1 : 0, cc = (ax == 2) ? 1 : 2;
    long long nc = (long long)g.n[bb] * g.n[cc];
    if (g.act[ax] && nc > m) m = nc;
  }
  int bx = (int)((m + kBlock - 1) / kBlock);
  if (bx > 128) bx = 128;
  dim3 grid(bx, 6);
  k_shell_norm<T><<<grid, kBlock, 0, L.s>>>(g, a, b, st, partials, stage, p2p);
  ++L.count;
}

__global__ void k_state_init(SolverState* st, double tolerance, int max_it, unsigned long long epoch = 1ull) {
  st->epoch = epoch;
  st->halo_count = 0u;
  st->halo_target = 0u;
  st->halo_cnt[0] = st->halo_cnt[1] = 0u;
  for (int i = 0; i < 8; ++i) {
    st->sum[i] = 0.0;
    st->scal[i] = 0.0;
    st->ticket[i] = 0u;
  }
  st->tol = 1.0;
  st->tolerance = tolerance;
  st->itr = 0;
  st->max_it = max_it;
  st->done = 0;
  st->status = PA_RUNNING;
  st->finished_flag = 0;
  st->swaps = 0;
}

// ---- workspace carving -------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  SolverState* st;
  double* partials;
  char* vec[8];
};

static size_t ws_bytes(long long cells, size_t esz, int nvec) {
  size_t s = align_up(sizeof(SolverState), 256);
  s += align_up(sizeof(double) * kNumSums * kMaxPartials, 256);
  s += (size_t)nvec * align_up((size_t)cells * esz, 256);
  return s;
}

static void carve(void* ws, long long cells, size_t esz, int nvec, Workspace& w) {
  char* p = (char*)ws;
  w.st = (SolverState*)p;
  p += align_up(sizeof(SolverState), 256);
  w.partials = (double*)p;
  p += align_up(sizeof(double) * kNumSums * kMaxPartials, 256);
  for (int i = 0; i < nvec; ++i) {
    w.vec[i] = p;
    p += align_up((size_t)cells * esz, 256);
  }
}

static int method_nvec(int method) {
  switch (method) {
    case PA_METHOD_CG: return 3;
    case PA_METHOD_BICGSTAB: return 6;
    default: return 0;
  }
}

// Mailbox for polling the device-side state: MAPPED pinned memory that a one-thread kernel writes straight from the
// device.  (It used to be a 200-byte cudaMemcpyAsync: a device-to-host copy queues on the copy engine behind whatever
// bulk download the caller has in flight on another stream, and a solver that polls every 32 iterations then idles
// the GPU until that download has finished.  A kernel's stores do not go through a copy engine.)
static SolverState* host_mailbox() {
  static SolverState* p = nullptr;
  if (!p) {
    if (cudaHostAlloc((void**)&p, sizeof(SolverState), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
  }
  return p;
}

__global__ void k_state_mirror(const SolverState* dev, SolverState* host) {
  *host = *dev;
  __threadfence_system();
}

static int poll_state(cudaStream_t s, const SolverState* dev, SolverState** out) {
  SolverState* h = host_mailbox();
  if (!h) return fail(PA_ERR_CUDA, "cudaHostAlloc failed");
  SolverState* hd = nullptr;
  if (cudaHostGetDevicePointer((void**)&hd, h, 0) == cudaSuccess && hd != nullptr) {
    k_state_mirror<<<1, 1, 0, s>>>(dev, hd);
  } else {  // no mapped memory on this platform: the copy engine it is
    cudaGetLastError();
    PA_CUDA(cudaMemcpyAsync(h, dev, sizeof(SolverState), cudaMemcpyDeviceToHost, s));
  }
  PA_CUDA(cudaStreamSynchronize(s));
  *out = h;
  return PA_OK;
}

static void fill_report(pa_report* rep, const SolverState* h, int launches) {
  rep->itr = h->itr;
  rep->status = h->status;
  rep->tol = h->tol;
  rep->result_in_alt = 0;
  rep->launches = launches;
  rep->swaps = 0;
  rep->reserved = 0;
}

static void set_swaps(pa_report* rep, int swaps) {
  rep->swaps = swaps;
  rep->result_in_alt = swaps & 1;
}

// The iteration loop runs on a private non-blocking stream (stream capture is not allowed
// on the legacy default stream torch hands us); it is ordered after the caller's stream by
// an event, and the call returns only after that stream has drained (poll_state).
static int solver_stream(cudaStream_t caller, cudaStream_t* out) {
  static cudaStream_t s_dev[kMaxDevices] = {};
  static cudaEvent_t ev_dev[kMaxDevices] = {};
  cudaStream_t& s = s_dev[current_device()];
  cudaEvent_t& ev = ev_dev[current_device()];
  if (!s) {
    PA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    PA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  PA_CUDA(cudaEventRecord(ev, caller));
  PA_CUDA(cudaStreamWaitEvent(s, ev, 0));
  *out = s;
  return PA_OK;
}

// second private stream for the halo exchange, with the two events that fork/join it
struct CommLane {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static CommLane* comm_lane() {
  static CommLane lane_dev[kMaxDevices];
  CommLane& lane = lane_dev[current_device()];
  if (!lane.s) {
    int lo = 0, hi = 0;  // highest priority: its few CTAs must not queue behind the compute kernel's
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&lane.s, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&lane.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&lane.join, cudaEventDisableTiming);
  }
  return &lane;
}

// ---- peer-memory mailboxes (common.cuh P2PDev) ------------------------------------------------
// Process-wide: one communicator per process.  `epoch` is the number of peer all-reduces completed so
// far; it is identical on every rank because all ranks run the same solves with bitwise identical
// scalars (and therefore the same iteration counts).
struct P2PCtx {
  unsigned long long* local = nullptr;       // this rank's IPC allocation: mailbox | halo flags | landing zone
  unsigned long long** peers_dev = nullptr;  // device array [nranks] of mapped mailboxes
  char* peer_base[16] = {};                  // the same mappings, host side (halo landing zones of the peers)
  int rank = 0, nranks = 0;
  unsigned long long epoch = 0;
  bool ready = false;
  size_t halo_cap = 0;                       // bytes per landing plane (0: no landing zone)
};
static P2PCtx g_p2p;
constexpr size_t kMailboxBytes = 2 * 16 * 8 * sizeof(unsigned long long);
// layout of the IPC allocation (common.cuh HaloDev): [0, 2048) mailbox; [2048, 4096) two flag words
// (lower ghost ready, upper ghost ready); from 4096: landing planes [slot 2][side 2], halo_cap bytes each
constexpr size_t kHaloFlagOff = kMailboxBytes;
constexpr size_t kHaloLandOff = 4096;
static_assert(kMailboxBytes == 2048, "layout");

// bytes per landing plane: PA_HALO_MIB (default 8 MiB = a 1024^2 fp64 plane); 0 disables the landing zone
static size_t halo_cap_bytes() {
  const char* e = getenv("PA_HALO_MIB");
  long mib = e ? atol(e) : 8;
  if (mib < 0) mib = 0;
  return (size_t)mib << 20;
}

static P2PDev p2p_dev(const Dist* dist) {
  P2PDev d{nullptr, 0, 0, 0, 0};
  if (dist && g_p2p.ready && g_p2p.nranks == dist->nranks && g_p2p.rank == dist->rank && dist->nranks <= 16) {
    d.peers = g_p2p.peers_dev;
    d.me = g_p2p.rank;
    d.nranks = g_p2p.nranks;
  }
  return d;
}

// Peer-memory halo exchange of the TMA CG kernels (common.cuh HaloDev): possible when the mailboxes are up,
// a plane fits a landing slot and axes 1/2 are not periodic (the wrapped halo reads go to the r array itself).
template <typename T>
static HaloDev halo_dev(const Dist* dist, const GridDev& g, const TilePlan& tile) {
  HaloDev h{{nullptr, nullptr}, 0, {nullptr, nullptr}, nullptr, 0, 0, 0};
  const size_t plane = (size_t)g.n[1] * g.n[2] * sizeof(T);
  if (!dist || p2p_dev(dist).peers == nullptr || g_p2p.halo_cap == 0 || plane > g_p2p.halo_cap || tile.wrap) return h;
  if (getenv("PA_NO_PEER_HALO")) return h;
  const int P = dist->nranks, me = dist->rank;
  const int lower = me > 0 ? me - 1 : (dist->ring ? P - 1 : -1);
  const int upper = me < P - 1 ? me + 1 : (dist->ring ? 0 : -1);
  const size_t cap = g_p2p.halo_cap;
  h.slot_bytes = (long long)(2 * cap);
  if (lower >= 0) {  // my first owned plane is the lower neighbour's UPPER ghost (side 1)
    h.dst[0] = g_p2p.peer_base[lower] + kHaloLandOff + cap;
    h.flag_dst[0] = (unsigned long long*)(g_p2p.peer_base[lower] + kHaloFlagOff) + 1;
  }
  if (upper >= 0) {  // my last owned plane is the upper neighbour's LOWER ghost (side 0)
    h.dst[1] = g_p2p.peer_base[upper] + kHaloLandOff;
    h.flag_dst[1] = (unsigned long long*)(g_p2p.peer_base[upper] + kHaloFlagOff);
  }
  h.flag_src = (const unsigned long long*)((char*)g_p2p.local + kHaloFlagOff);
  h.tiles = tile.tiles_y * tile.tiles_z;
  h.on = 1;
  return h;
}

__global__ void k_halo_flags_init(unsigned long long* flags, unsigned long long v) {
  flags[0] = v;
  flags[1] = v;
}

// ---- CG -----------------------------------------------------------------------------------
// One iteration = [d update] -> d.Ad -> x/r update -> BC faces -> shell norm + scalars.
// Tiled variant fuses the first two and recomputes Ad in the third (8 words / cell).
template <typename T>
static void cg_iteration(Launcher& L, const GridDev& g, const EqDev<T>& eq, int nfaces,
                         const pa_face_bc* faces, const Workspace& w, T* cur, T* nxt, bool tiled,
                         const TilePlan& plan, int parity, cudaEvent_t* marks = nullptr,
                         const TmaPlan* tma = nullptr, const Dist* dist = nullptr) {
  auto mark = [&](int i) {
    if (marks) cudaEventRecord(marks[i], L.s);
  };
  bool overlapped = false;
  mark(0);
  T* r = (T*)w.vec[0];
  T* d = (T*)w.vec[1];
  int nb = grid_blocks(g.cells);
  // slabs + peer mailboxes: the all-reduces and scalar stages run inside the fused kernels
  const bool p2p = tma != nullptr && tma->tile.p2p.peers != nullptr;
  if (tma) {
    T* d_new = (T*)w.vec[2 - parity];
    launch_cg_phaseA_tma<T>(L.s, *tma, g, eq, parity, d_new, w.st, w.partials);
    ++L.count;
    if (dist && !p2p) {
      dist_allreduce(*dist, &w.st->sum[R_A], 1, L.s);
      k_finalize<T><<<1, 1, 0, L.s>>>(ST_CG_DAD, w.st);
      L.count += 2;
    }
    mark(1);
    const bool peer_halo = tma->tile.halo.on != 0;
    CommLane* lane = (dist && !peer_halo && tma_interior_chunks(*tma, g) >= 4) ? comm_lane() : nullptr;  // enough interior work to hide it
    if (peer_halo) {
      // r's boundary planes travel INSIDE phase B (stores into the neighbours' landing zones); boundary
      // chunks first so that they arrive early.  No NCCL call, no second stream, no waiting kernel.
      launch_cg_phaseB_tma<T>(L.s, *tma, g, eq, parity, nxt, r, w.st, w.partials, tma_interior_chunks(*tma, g) >= 1 ? 4 : 0);
      ++L.count;
      overlapped = true;
    } else if (lane) {
      // overlap: ONE phase B launch whose boundary chunks are scheduled first and count themselves
      // in; on the (high-priority) communication stream k_wait_halo waits for that count, then their r
      // planes go out while the interior chunks are still running.  (A split into two launches cost
      // a wave: 1.7 + 5.2 -> 2 + 6 waves instead of 6.9 -> 7.)
      // (phase B is enqueued BEFORE the waiting kernel: outside graph capture the host may block
      //  inside the NCCL enqueue, and a spinning k_wait_halo must never wait for a kernel that the
      //  host has not been able to submit yet)
      cudaEventRecord(lane->fork, L.s);
      launch_cg_phaseB_tma<T>(L.s, *tma, g, eq, parity, nxt, r, w.st, w.partials, 3);
      cudaStreamWaitEvent(lane->s, lane->fork, 0);
      k_wait_halo<<<1, 1, 0, lane->s>>>(w.st, tma_boundary_ctas(*tma, g));
      dist_halo_exchange<T>(*dist, r, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, lane->s);
      cudaEventRecord(lane->join, lane->s);
      L.count += 3;
      overlapped = true;
    } else {
      launch_cg_phaseB_tma<T>(L.s, *tma, g, eq, parity, nxt, r, w.st, w.partials);
      ++L.count;
    }
  } else if (tiled) {
    // the fused d-update recomputes d_new on tile halos, so d is double-buffered: neighbours
    // must still see d_old there (kernels_tiled.cuh)
    T* d_old = (T*)w.vec[1 + parity];
    T* d_new = (T*)w.vec[2 - parity];
    launch_cg_phaseA<T>(L.s, plan, g, eq, r, d_old, d_new, w.st, w.partials);
    ++L.count;
    if (dist) {
      dist_allreduce(*dist, &w.st->sum[R_A], 1, L.s);
      k_finalize<T><<<1, 1, 0, L.s>>>(ST_CG_DAD, w.st);
      L.count += 2;
    }
    mark(1);
    launch_cg_phaseB<T>(L.s, plan, g, eq, cur, nxt, d_new, r, w.st, w.partials);
    ++L.count;
  } else {
    k_cg_dupdate<T><<<nb, kBlock, 0, L.s>>>(g, r, d, w.st);
    k_cg_dAd<T><<<nb, kBlock, 0, L.s>>>(g, eq, d, w.st, w.partials, dist ? ST_NONE : ST_CG_DAD);
    if (dist) {
      dist_allreduce(*dist, &w.st->sum[R_A], 1, L.s);
      k_finalize<T><<<1, 1, 0, L.s>>>(ST_CG_DAD, w.st);
      L.count += 2;
    }
    mark(1);
    k_cg_update<T><<<nb, kBlock, 0, L.s>>>(g, eq, cur, nxt, d, r, w.st, w.partials);
    L.count += 3;
  }
  mark(2);
  if (!dist) {
    if (!((tiled && plan.fuse_fin) || (tma && tma->tile.fuse_fin))) {
      launch_bcs<T>(L, g, nfaces, faces, nxt, w.st);
      launch_shell<T>(L, g, nxt, cur, w.st, w.partials, ST_CG_FIN);
    }
  } else {
    // multi-GPU: boundary planes of the new r to the neighbours' ghosts, then the three sums
    const bool stat = static_shell(nfaces, faces) != 0;
    if (stat) {
      if (!p2p) cudaMemsetAsync(&w.st->sum[R_SHELL], 0, sizeof(double), L.s);
    } else {
      launch_bcs<T>(L, g, nfaces, faces, nxt, w.st, dist);
      P2PDev pp = p2p ? tma->tile.p2p : P2PDev{nullptr, 0, 0, 0, 0};
      pp.slot0 = R_A;
      pp.count = 4;
      launch_shell<T>(L, g, nxt, cur, w.st, w.partials, p2p ? ST_CG_FIN : ST_NONE, pp);
    }
    if (!overlapped) dist_halo_exchange<T>(*dist, r, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, L.s);
    if (!p2p) {  // (p2p, static shell: phase B has already reduced and finalized)
      dist_allreduce(*dist, &w.st->sum[R_A], 4, L.s);
      k_finalize<T><<<1, 1, 0, L.s>>>(ST_CG_FIN, w.st);
      L.count += 2;
    }
    const bool peer_halo = tma != nullptr && tma->tile.halo.on != 0;
    if (overlapped && !peer_halo) cudaStreamWaitEvent(L.s, comm_lane()->join, 0);  // r ghosts before the next phase A
    if (!peer_halo) ++L.count;
  }
  mark(3);
}

// Instrumented CG pass for bench.py: `iters` iterations, no graph, every section bracketed by
// CUDA events on the launching stream.  out_ms = {phase A, phase B, BC faces + shell norm,
// whole loop, launches per iteration} (averages per iteration).
template <typename T>
static int profile_cg(const pa_grid* pg, const pa_equation* peq, int nfaces, const pa_face_bc* faces,
                      T* x, T* x_alt, const T* rhs, int iters, int variant, void* ws, size_t ws_size,
                      double* out_ms, cudaStream_t caller_stream) {
  cudaStream_t stream;
  int rcs = solver_stream(caller_stream, &stream);
  if (rcs != PA_OK) return rcs;
  GridDev g = make_grid(*pg);
  EqDev<T> eq = make_eq<T>(*peq);
  if (ws_size < ws_bytes(g.cells, sizeof(T), method_nvec(PA_METHOD_CG)))
    return fail(PA_ERR_ARG, "workspace too small");
  Workspace w;
  carve(ws, g.cells, sizeof(T), method_nvec(PA_METHOD_CG), w);
  Launcher L{stream};
  TilePlan plan;
  TmaPlan tmap;
  const bool contract = (variant & 0x100) != 0;
  variant &= 0xff;
  bool use_tma = variant == 0 && plan_tma<T>(g, *peq, nfaces, faces, x, x_alt, (T*)w.vec[0], (T*)w.vec[1],
                                             (T*)w.vec[2], tmap);
  if (use_tma) tmap.tile.fuse_fin = static_shell(nfaces, faces);
  if (use_tma) tmap.contract = (contract && !tmap.tile.wrap) ? 1 : 0;
  bool tiled = !use_tma && (variant == 0 || variant == 2) && plan_tiles<T>(g, *peq, plan);
  if (tiled) plan.fuse_fin = static_shell(nfaces, faces);
  k_state_init<<<1, 1, 0, stream>>>(w.st, 1e-300, iters + 10);
  launch_bcs<T>(L, g, nfaces, faces, x, nullptr);
  k_residual_init<T><<<grid_blocks(g.cells), kBlock, 0, stream>>>(g, eq, x, rhs, (T*)w.vec[0],
                                                                 (T*)w.vec[1], w.st, w.partials,
                                                                 ST_CG_INIT);
  std::vector<cudaEvent_t> ev(4 * (size_t)iters);
  for (auto& e : ev) PA_CUDA(cudaEventCreate(&e));
  int before = L.count;
  for (int it = 0; it < iters; ++it) {
    T* cur = (it & 1) ? x_alt : x;
    T* nxt = (it & 1) ? x : x_alt;
    cg_iteration<T>(L, g, eq, nfaces, faces, w, cur, nxt, tiled, plan, it & 1, &ev[4 * (size_t)it],
                    use_tma ? &tmap : nullptr);
  }
  PA_CUDA(cudaStreamSynchronize(stream));
  double a = 0, b = 0, c = 0;
  for (int it = 0; it < iters; ++it) {
    float ms;
    cudaEventElapsedTime(&ms, ev[4 * it], ev[4 * it + 1]);
    a += ms;
    cudaEventElapsedTime(&ms, ev[4 * it + 1], ev[4 * it + 2]);
    b += ms;
    cudaEventElapsedTime(&ms, ev[4 * it + 2], ev[4 * it + 3]);
    c += ms;
  }
  float tot;
  cudaEventElapsedTime(&tot, ev[0], ev[4 * (size_t)iters - 1]);
  out_ms[0] = a / iters;
  out_ms[1] = b / iters;
  out_ms[2] = c / iters;
  out_ms[3] = (double)tot / iters;
  out_ms[4] = (double)(L.count - before) / iters;
  out_ms[5] = use_tma ? 2.0 : (tiled ? 1.0 : 0.0);
  for (auto& e : ev) cudaEventDestroy(e);
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

// Multi-GPU (slab) variant: the axpy stages run over the OWNED planes only (a contiguous range),
// p and s get their ghost planes by one halo exchange each before the stencil that reads them
// (one exchange per operator application), and each of the four reductions is
// all-reduced before its scalar stage is finalized by k_finalize.
template <typename T>
static void bicgstab_iteration(Launcher& L, const GridDev& g, const EqDev<T>& eq, int nfaces,
                               const pa_face_bc* faces, const Workspace& w, T* cur, T* nxt,
                               const TilePlan* pw = nullptr, const Dist* dist = nullptr) {
  T* r0 = (T*)w.vec[0];
  T* r = (T*)w.vec[1];
  T* p = (T*)w.vec[2];
  T* v = (T*)w.vec[3];
  T* s = (T*)w.vec[4];
  T* t = (T*)w.vec[5];
  int nb = grid_blocks(g.cells);
  // streaming axpy stages: legal when the owned range starts and ends on a 16-byte vector
  // boundary (the region test is unnecessary: everything is 0 outside it)
  constexpr int SV = StreamVec<T>::N;
  const long long plane = (long long)g.n[1] * g.n[2];
  const long long off = (long long)g.olo0 * plane;
  const long long nown = (long long)(g.ohi0 - g.olo0) * plane;
  const bool stream_ok = (nown % SV == 0) && (off % SV == 0);
  const long long nvec = nown / SV;
  // (the choice must be the same on every rank: plane % SV is, stream_ok alone is not -- the owned
  //  range starts at a different plane on rank 0)
  const bool p2p_on = dist && pw != nullptr && (plane % SV == 0) && p2p_dev(dist).peers != nullptr;
  auto reduce = [&](int count, int stage) {  // rank-sum of the raw sums, then the scalar stage
    if (!dist || p2p_on) return;
    dist_allreduce(*dist, &w.st->sum[R_A], count, L.s);
    k_finalize<T><<<1, 1, 0, L.s>>>(stage, w.st);
    L.count += 2;
  };
  // slabs with peer mailboxes, TMA + streaming path: every reduction is summed over the ranks inside
  // the kernel that produces it (common.cuh p2p_allreduce)
  const P2PDev P = p2p_on ? p2p_dev(dist) : P2PDev{nullptr, 0, 0, 0, 0};
  const bool p2p = P.peers != nullptr;
  TilePlan pwr;
  if (pw) {
    pwr = *pw;
    pwr.p2p = P;
    pw = &pwr;
  }
  const bool nccl = dist && !p2p;
  const int ST_V = nccl ? ST_NONE : ST_BI_V, ST_S = nccl ? ST_NONE : ST_BI_S, ST_T = nccl ? ST_NONE : ST_BI_T;
  // (p = r + beta (p - omega v) of this iteration was produced by the previous iteration's fused
  //  x/r/p update, or by bicgstab_first_p before the loop; 17 words per cell and iteration)
  if (pw)
    L.ok &= launch_star_tma<T, PW_APPLY_V>(L.s, g, eq, *pw, p, r0, v, nullptr, (T)0, w.st, w.partials, ST_V);
  else
    k_bi_apply<T, 0><<<nb, kBlock, 0, L.s>>>(g, eq, p, v, r0, w.st, w.partials, ST_V);
  reduce(1, ST_BI_V);
  // second half step.  TMA engine + streaming x update: s = r - alpha v is never stored -- the fused
  // kernel forms it on the fly (with halos) for t = A(s) and the x update recomputes it
  // (R r,v,r0 W t | R x,p,r,t,v W x,r,p: with v = A(p) 15 words per iteration).
  const bool fuse_st = pw != nullptr && stream_ok;
  if (fuse_st) {
    if (dist) {  // the stencil of s needs the neighbours' boundary planes of r and v
      dist_halo_exchange<T>(*dist, v, plane, g.olo0, g.ohi0, L.s);
      ++L.count;
    }
    L.ok &= launch_bi_st_tma<T>(L.s, g, eq, *pw, r, v, r0, t, nullptr, w.st, w.partials, nccl ? ST_NONE : ST_BI_ST);
    reduce(4, ST_BI_ST);
    k_bi_x_stream<T, true><<<nb, kBlock, 0, L.s>>>(nvec, cur + off, nxt + off, p + off, s + off, t + off, v + off,
                                                   r + off, w.st, w.partials, P);
    L.count += 3;
    if (dist) {
      dist_halo_exchange<T>(*dist, r, plane, g.olo0, g.ohi0, L.s);
      ++L.count;
    }
  } else {
    if (stream_ok)
      k_bi_s_stream<T><<<nb, kBlock, 0, L.s>>>(nvec, r + off, v + off, s + off, w.st, w.partials, ST_S);
    else
      k_bi_s<T><<<nb, kBlock, 0, L.s>>>(g, r, v, s, w.st, w.partials, ST_S);
    reduce(1, ST_BI_S);
    if (dist) {
      dist_halo_exchange<T>(*dist, s, plane, g.olo0, g.ohi0, L.s);
      ++L.count;
    }
    if (pw)
      L.ok &= launch_star_tma<T, PW_APPLY_T>(L.s, g, eq, *pw, s, r0, t, nullptr, (T)0, w.st, w.partials, ST_T);
    else
      k_bi_apply<T, 1><<<nb, kBlock, 0, L.s>>>(g, eq, s, t, r0, w.st, w.partials, ST_T);
    reduce(3, ST_BI_T);
    if (stream_ok)
      k_bi_x_stream<T, false><<<nb, kBlock, 0, L.s>>>(nvec, cur + off, nxt + off, p + off, s + off, t + off, v + off,
                                                      r + off, w.st, w.partials);
    else
      k_bi_x<T><<<nb, kBlock, 0, L.s>>>(g, cur, nxt, p, s, t, v, r, w.st, w.partials);
    L.count += 4;
  }
  if (dist) {  // ghost planes of the new p for the next iteration's v = A(p)
    dist_halo_exchange<T>(*dist, p, plane, g.olo0, g.ohi0, L.s);
    ++L.count;
  }
  launch_bcs<T>(L, g, nfaces, faces, nxt, w.st, dist);
  if (nccl) {  // (peer mailboxes: |r|^2 was summed over the ranks inside the x update)
    dist_allreduce(*dist, &w.st->sum[R_A], 1, L.s);
    ++L.count;
  }
  k_finalize<T><<<1, 1, 0, L.s>>>(ST_BI_FIN, w.st);
  ++L.count;
}

// p of the first iteration: p = r + beta (p - omega v) with p = v = 0 (linalg.py:217)
template <typename T>
static void bicgstab_first_p(Launcher& L, const GridDev& g, const Workspace& w, const Dist* dist) {
  T* r = (T*)w.vec[1];
  T* p = (T*)w.vec[2];
  T* v = (T*)w.vec[3];
  constexpr int SV = StreamVec<T>::N;
  const long long plane = (long long)g.n[1] * g.n[2];
  const long long off = (long long)g.olo0 * plane;
  const long long nown = (long long)(g.ohi0 - g.olo0) * plane;
  const int nb = grid_blocks(g.cells);
  if ((nown % SV == 0) && (off % SV == 0))
    k_bi_p_stream<T><<<nb, kBlock, 0, L.s>>>(nown / SV, r + off, p + off, v + off, w.st);
  else
    k_bi_p<T><<<nb, kBlock, 0, L.s>>>(g, r, p, v, w.st);
  ++L.count;
  if (dist) {
    dist_halo_exchange<T>(*dist, p, plane, g.olo0, g.ohi0, L.s);
    ++L.count;
  }
}

// Multi-GPU: the sweep reads x with its ghost planes (valid on entry), then the new iterate's
// boundary planes go to the neighbours and {|dx|^2 interior, |dx|^2 shell} are all-reduced.
template <typename T>
static void jacobi_iteration(Launcher& L, const GridDev& g, const EqDev<T>& eq, int nfaces,
                             const pa_face_bc* faces, const Workspace& w, T* cur, T* nxt,
                             const T* rhs, const TilePlan* pw = nullptr, const Dist* dist = nullptr) {
  int nb = grid_blocks(g.cells);
  const bool stat = static_shell(nfaces, faces) != 0;
  const P2PDev P = (dist && pw != nullptr) ? p2p_dev(dist) : P2PDev{nullptr, 0, 0, 0, 0};
  const bool p2p = P.peers != nullptr;
  if (pw) {
    // static shell: the sweep also finalizes the iteration (shell part of the norm is 0); on slabs
    // with peer mailboxes it first sums |dx|^2 over the ranks, inside the kernel
    const bool fuse = stat && (!dist || p2p);
    TilePlan pwr = *pw;
    if (fuse) pwr.p2p = P;
    L.ok &= launch_star_tma<T, PW_JACOBI>(L.s, g, eq, pwr, cur, rhs, nxt, nullptr, (T)0, w.st, w.partials,
                                          fuse ? ST_JA_FIN : ST_NONE);
    ++L.count;
    if (fuse) {
      if (dist) {
        dist_halo_exchange<T>(*dist, nxt, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, L.s);
        ++L.count;
      }
      return;
    }
  } else {
    k_pointwise_update<T, 0><<<nb, kBlock, 0, L.s>>>(g, eq, cur, nxt, rhs, (T)0, w.st, w.partials);
    ++L.count;
  }
  if (dist && stat) {
    cudaMemsetAsync(&w.st->sum[R_SHELL], 0, sizeof(double), L.s);
  } else {
    launch_bcs<T>(L, g, nfaces, faces, nxt, w.st, dist);
    P2PDev pp = P;  // non-static shell on slabs: the shell norm is the last sum -> reduce all four here
    pp.slot0 = R_A;
    pp.count = 4;
    launch_shell<T>(L, g, nxt, cur, w.st, w.partials, (dist && !p2p) ? ST_NONE : ST_JA_FIN, pp);
  }
  if (dist) {
    dist_halo_exchange<T>(*dist, nxt, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, L.s);
    ++L.count;
    if (!p2p || stat) {  // (stat && p2p returned above; this is the NCCL path)
      dist_allreduce(*dist, &w.st->sum[R_A], 4, L.s);
      k_finalize<T><<<1, 1, 0, L.s>>>(ST_JA_FIN, w.st);
      L.count += 2;
    }
  }
}

// Shared driver: init, then iterate in pairs (x -> x_alt -> x) so that a pair can be
// captured once in a CUDA graph and replayed; poll the device-side done flag every
// `check_every` iterations.  Launches after `done` are no-ops (each kernel tests st->done).
template <typename T>
static int run_solver(int method, const pa_grid* pg, const pa_equation* peq, int nfaces,
                      const pa_face_bc* faces, T* x, T* x_alt, const T* rhs,
                      const pa_solver_cfg* cfg, void* ws, size_t ws_size, pa_report* rep,
                      cudaStream_t caller_stream, const Dist* dist = nullptr) {
  cudaStream_t stream;
  {
    int rcs = solver_stream(caller_stream, &stream);
    if (rcs != PA_OK) return rcs;
  }
  GridDev g = make_grid(*pg);
  // nonlinear advection (pa_op.adv_is_iterate): the operator reads its advection speed from the
  // current iterate, which ping-pongs between x and x_alt -> one equation per buffer
  EqDev<T> eq = make_eq<T>(*peq);
  EqDev<T> eq_alt = eq;
  bool nonlinear = false;
  for (int k = 0; k < peq->nops; ++k)
    if (peq->ops[k].kind != PA_OP_STAR && peq->ops[k].adv_is_iterate) {
      eq.op[k].adv = x;
      eq_alt.op[k].adv = x_alt;
      nonlinear = true;
    }
  // (on slabs the iterate's ghost planes are refreshed after every update, see `iteration` below: the central
  //  scheme reads the advection speed one cell up and down the slab axis)
  int nvec = method_nvec(method);
  if (ws_size < ws_bytes(g.cells, sizeof(T), nvec))
    return fail(PA_ERR_ARG, "workspace too small (see pa_solver_workspace_bytes)");
  Workspace w;
  carve(ws, g.cells, sizeof(T), nvec, w);
  Launcher L{stream};

  TilePlan plan;
  bool tiled = false;
  TmaPlan tmap;
  bool use_tma = false;
  // 4: fused kernels, never a whole-solve kernel; 5: the cooperative whole-solve TMA kernel (L2-resident grids)
  const bool auto_fused = cfg->variant == 0 || cfg->variant == 4 || cfg->variant == 5 || cfg->variant == 6;
  if (method == PA_METHOD_CG && auto_fused) {
    use_tma = plan_tma<T>(g, *peq, nfaces, faces, x, x_alt, (T*)w.vec[0], (T*)w.vec[1], (T*)w.vec[2], tmap);
    if (use_tma) tmap.tile.fuse_fin = static_shell(nfaces, faces);
    if (use_tma) tmap.contract = ((cfg->flags & PA_FLAG_CONTRACT) && !tmap.tile.wrap) ? 1 : 0;
  }
  if (method == PA_METHOD_CG && !use_tma && (auto_fused || cfg->variant == 2))
    tiled = plan_tiles<T>(g, *peq, plan);
  if (tiled) plan.fuse_fin = static_shell(nfaces, faces);
  if (dist) {
    if (!nccl_api().ok) return fail(PA_ERR_NCCL, nccl_api().error);
    nccl_first_error() = ncclSuccess;
    plan.dist = tmap.tile.dist = 1;
    plan.fuse_fin = tmap.tile.fuse_fin = 0;
    if (use_tma && cfg->variant != 4) {  // (variant 4 on slabs: keep the NCCL all-reduces, for A/B runs)
      tmap.tile.p2p = p2p_dev(dist);
      if (tmap.tile.p2p.peers != nullptr) {
        tmap.tile.fuse_fin = static_shell(nfaces, faces);
        tmap.tile.halo = halo_dev<T>(dist, g, tmap.tile);
        if (tmap.tile.halo.on) {
          const bool flat = tma_flat(g);
          const int boxz = flat ? TmaCfg<T, KFlat>::BOXZ : TmaCfg<T, KStd>::BOXZ;
          const int boxy = flat ? TmaCfg<T, KFlat>::BOXY : TmaCfg<T, KStd>::BOXY;
          if (!make_map<T>(&tmap.r_land, (const T*)((char*)g_p2p.local + kHaloLandOff), g, boxz, boxy, 4, g_p2p.halo_cap))
            tmap.tile.halo.on = 0;
        }
      }
    }
  }

  TilePlan pwtile;
  const bool pw_ok = cfg->variant != 1 && pw_eligible<T>(g, *peq, nfaces, faces);
  if (pw_ok) pw_tile_plan<T>(g, pwtile, nfaces, faces);
  const TilePlan* pw = pw_ok ? &pwtile : nullptr;

  k_state_init<<<1, 1, 0, stream>>>(w.st, cfg->tol, cfg->max_it, g_p2p.epoch + 1ull);
  ++L.count;
  if (dist && dist->ring) bc_scratch((size_t)2 * g.n[1] * g.n[2] * sizeof(T));  // before any capture
  launch_bcs<T>(L, g, nfaces, faces, x, nullptr, dist);
  int nb = grid_blocks(g.cells);
  size_t vbytes = (size_t)g.cells * sizeof(T);

  if (method == PA_METHOD_CG) {
    T* r = (T*)w.vec[0];
    T* d = (T*)w.vec[1];
    const long long plane = (long long)g.n[1] * g.n[2];
    if (dist) dist_halo_exchange<T>(*dist, x, plane, g.olo0, g.ohi0, stream);  // x ghosts for r0
    if (pw)
      L.ok &= launch_star_tma<T, PW_RESID>(stream, g, eq, *pw, x, rhs, r, d, (T)0, w.st, w.partials,
                                           dist ? ST_NONE : ST_CG_INIT);
    else
      k_residual_init<T><<<nb, kBlock, 0, stream>>>(g, eq, x, rhs, r, d, w.st, w.partials,
                                                    dist ? ST_NONE : ST_CG_INIT);
    ++L.count;
    if (dist) {
      dist_halo_exchange<T>(*dist, r, plane, g.olo0, g.ohi0, stream);
      if (g.olo0 > 0)
        PA_CUDA(cudaMemcpyAsync(d, r, (size_t)g.olo0 * plane * sizeof(T), cudaMemcpyDeviceToDevice, stream));
      if (g.ohi0 < g.n[0])
        PA_CUDA(cudaMemcpyAsync(d + (long long)g.ohi0 * plane, r + (long long)g.ohi0 * plane,
                                (size_t)(g.n[0] - g.ohi0) * plane * sizeof(T), cudaMemcpyDeviceToDevice,
                                stream));
      dist_allreduce(*dist, &w.st->sum[R_A], 1, stream);
      k_finalize<T><<<1, 1, 0, stream>>>(ST_CG_INIT, w.st);
      L.count += 4;
      if (use_tma && tmap.tile.halo.on) {
        // the first phase A (parity 0) reads landing slot 1: seed it with the ghost planes NCCL has just
        // delivered, and raise both flags to the sequence number that phase A will ask for (epoch - 1)
        char* land = (char*)g_p2p.local + kHaloLandOff + 2 * g_p2p.halo_cap;
        const size_t pb = (size_t)plane * sizeof(T);
        if (g.olo0 > 0)
          PA_CUDA(cudaMemcpyAsync(land, r + (long long)(g.olo0 - 1) * plane, pb, cudaMemcpyDeviceToDevice, stream));
        if (g.ohi0 < g.n[0])
          PA_CUDA(cudaMemcpyAsync(land + g_p2p.halo_cap, r + (long long)g.ohi0 * plane, pb, cudaMemcpyDeviceToDevice, stream));
        k_halo_flags_init<<<1, 1, 0, stream>>>((unsigned long long*)((char*)g_p2p.local + kHaloFlagOff), g_p2p.epoch);
        L.count += 1;
      }
    }
  } else if (method == PA_METHOD_BICGSTAB) {
    T* r0 = (T*)w.vec[0];
    T* r = (T*)w.vec[1];
    for (int i = 2; i < 6; ++i) PA_CUDA(cudaMemsetAsync(w.vec[i], 0, vbytes, stream));
    const long long plane = (long long)g.n[1] * g.n[2];
    if (dist) {
      dist_halo_exchange<T>(*dist, x, plane, g.olo0, g.ohi0, stream);  // x ghosts for r0
      // the iteration only writes the owned planes of the iterates: give x_alt the same ghosts
      if (g.olo0 > 0)
        PA_CUDA(cudaMemcpyAsync(x_alt, x, (size_t)g.olo0 * plane * sizeof(T), cudaMemcpyDeviceToDevice, stream));
      if (g.ohi0 < g.n[0])
        PA_CUDA(cudaMemcpyAsync(x_alt + (long long)g.ohi0 * plane, x + (long long)g.ohi0 * plane,
                                (size_t)(g.n[0] - g.ohi0) * plane * sizeof(T), cudaMemcpyDeviceToDevice,
                                stream));
      ++L.count;
    }
    const int st_init = dist ? ST_NONE : ST_BI_INIT;
    if (pw)
      L.ok &= launch_star_tma<T, PW_RESID>(stream, g, eq, *pw, x, rhs, r0, r, (T)0, w.st, w.partials, st_init);
    else
      k_residual_init<T><<<nb, kBlock, 0, stream>>>(g, eq, x, rhs, r0, r, w.st, w.partials, st_init);
    ++L.count;
    if (dist) {
      dist_allreduce(*dist, &w.st->sum[R_A], 1, stream);
      k_finalize<T><<<1, 1, 0, stream>>>(ST_BI_INIT, w.st);
      dist_halo_exchange<T>(*dist, r, plane, g.olo0, g.ohi0, stream);  // the fused s/t kernel reads r with halos
      L.count += 3;
    }
    bicgstab_first_p<T>(L, g, w, dist);
  } else if (dist) {  // Jacobi: the first sweep reads x with its ghost planes
    dist_halo_exchange<T>(*dist, x, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, stream);
    ++L.count;
  }
  PA_CUDA(cudaGetLastError());
  if (!L.ok) return fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled failed (TMA tile descriptor)");

  // while-loop entry condition of cg/jacobi: tol = 1.0 > tolerance (linalg.py:90,109)
  if (method != PA_METHOD_BICGSTAB && !(1.0 > cfg->tol)) {
    PA_CUDA(cudaStreamSynchronize(stream));
    rep->itr = 0;
    rep->status = PA_CONVERGED;
    rep->tol = 1.0;
    rep->result_in_alt = 0;
    rep->launches = L.count;
    rep->swaps = 0;
    rep->reserved = 0;
    return PA_OK;
  }

  // Small grids: the whole CG solve as ONE persistent cooperative kernel (kernels_small.cuh) -- the
  // iteration of the fused kernels is launch-bound there.  variant 3 forces it, auto (0) takes it
  // below kSmallCgCells.
  if (method == PA_METHOD_CG && !dist && !nonlinear && static_shell(nfaces, faces) &&
      (cfg->variant == 3 || (cfg->variant == 0 && g.cells <= kSmallCgCells))) {
    const int maxb = small_cg_max_blocks<T>();
    if (maxb > 0) {
      long long want = (g.cells + kSmallBlock - 1) / kSmallBlock;
      int blocks = (int)(want < maxb ? want : maxb);
      if (blocks > kMaxPartials) blocks = kMaxPartials;
      T* r = (T*)w.vec[0];
      T* d = (T*)w.vec[1];
      double* partA = w.partials;
      double* partB = w.partials + 2 * kMaxPartials;
      PA_CUDA(cudaMemcpyAsync(x_alt, x, vbytes, cudaMemcpyDeviceToDevice, stream));  // the static shell
      void* args[] = {(void*)&g, (void*)&eq, (void*)&x, (void*)&x_alt, (void*)&r, (void*)&d, (void*)&w.st,
                      (void*)&partA, (void*)&partB};
      PA_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_persistent<T>, dim3(blocks), dim3(kSmallBlock), args, 0,
                                          stream));
      L.count += 2;
      SolverState* hs = nullptr;
      int rcp = poll_state(stream, w.st, &hs);
      if (rcp != PA_OK) return rcp;
      PA_CUDA(cudaGetLastError());
      if (!hs->done) return fail(PA_ERR_CUDA, "persistent CG kernel returned without latching `done`");
      fill_report(rep, hs, L.count);
      set_swaps(rep, hs->itr + (hs->status == PA_BAD_TOL ? 1 : 0));
      return PA_OK;
    }
    if (cfg->variant == 3) return fail(PA_ERR_UNSUPPORTED, "cooperative launch is not available on this device");
  }

  // Jacobi on a 2-D grid that fits the SMs' shared memory: the whole solve as ONE cooperative launch
  // (kernels_resident.cuh k_jacobi_resident: uniform coefficients, every face Dirichlet).  variant 6 or auto.
  if (method == PA_METHOD_JACOBI && !dist && !nonlinear && pw_ok && tma_flat(g) && static_shell(nfaces, faces) &&
      (cfg->variant == 6 || (cfg->variant == 0 && getenv("PA_NO_RESIDENT") == nullptr))) {
    PA_CUDA(cudaMemcpyAsync(x_alt, x, vbytes, cudaMemcpyDeviceToDevice, stream));  // the static shell
    if (launch_jacobi_resident<T>(stream, g, *peq, eq, x, x_alt, rhs, w.st, cfg->max_it)) {
      L.count += 2;
      SolverState* hs = nullptr;
      int rcp = poll_state(stream, w.st, &hs);
      if (rcp != PA_OK) return rcp;
      PA_CUDA(cudaGetLastError());
      if (res_check_abort()) return fail(PA_ERR_CUDA, "resident Jacobi kernel: a row / sum exchange timed out (watchdog)");
      if (!hs->done) return fail(PA_ERR_CUDA, "resident Jacobi kernel returned without latching `done`");
      fill_report(rep, hs, L.count);
      set_swaps(rep, hs->itr + (hs->status == PA_BAD_TOL ? 1 : 0));
      return PA_OK;
    }
    cudaGetLastError();
  }

  // fused TMA CG kernels with an implicit-Euler term: operator 0 + diagonal shift (plan_tma)
  EqDev<T> eq_cg = eq;
  if (use_tma && peq->nops == 2) {
    double c = 0.0;
    diagonal_op(peq->ops[1], &c);
    eq_cg.nops = 1;
    eq_cg.op[0].has_shift = 1;
    eq_cg.op[0].shift = (T)c;
  }

  // 2-D grids that fit the SMs' shared memory: the whole solve as ONE cooperative launch with x, r, d resident
  // (kernels_resident.cuh).  variant 6 forces it; auto (0) takes it whenever it fits.
  if (method == PA_METHOD_CG && use_tma && tma_flat(g) && !dist && !nonlinear && static_shell(nfaces, faces) &&
      !tmap.tile.wrap && !tmap.contract &&
      (cfg->variant == 6 || (cfg->variant == 0 && g.cells > kSmallCgCells && getenv("PA_NO_RESIDENT") == nullptr))) {
    PA_CUDA(cudaMemcpyAsync(x_alt, x, vbytes, cudaMemcpyDeviceToDevice, stream));  // the static shell
    if (launch_cg_resident<T>(stream, g, eq_cg, x, x_alt, (const T*)w.vec[0], (const T*)w.vec[1], w.st, cfg->max_it, res_coef_uniform(*peq) ? 1 : 0)) {
      L.count += 2;
      SolverState* hs = nullptr;
      int rcp = poll_state(stream, w.st, &hs);
      if (rcp != PA_OK) return rcp;
      PA_CUDA(cudaGetLastError());
      if (res_check_abort()) return fail(PA_ERR_CUDA, "resident CG kernel: a row / sum exchange timed out (watchdog)");
      if (!hs->done) return fail(PA_ERR_CUDA, "resident CG kernel returned without latching `done`");
      fill_report(rep, hs, L.count);
      set_swaps(rep, hs->itr + (hs->status == PA_BAD_TOL ? 1 : 0));
      return PA_OK;
    }
    cudaGetLastError();
    if (cfg->variant == 6) return fail(PA_ERR_UNSUPPORTED, "resident CG: the grid does not fit the SMs' shared memory (2-D meshes only)");
  }

  // L2-resident grids: the whole solve as ONE cooperative launch of the two TMA phases (kernels_tma.cuh
  // k_cg_coop_tma).  variant 5 forces it; auto (0) takes it between the tiny-grid kernel above and kCoopCgCells.
  if (method == PA_METHOD_CG && use_tma && !dist && !nonlinear && static_shell(nfaces, faces) && !tmap.tile.wrap &&
      !tmap.contract &&
      (cfg->variant == 5 ||
       (cfg->variant == 0 && g.cells > kSmallCgCells && g.cells <= kCoopCgCells && getenv("PA_NO_COOP_CG") == nullptr))) {
    if (launch_cg_coop_tma<T>(stream, tmap, g, eq_cg, x, x_alt, (T*)w.vec[0], (T*)w.vec[1], (T*)w.vec[2], w.st,
                              w.partials)) {
      L.count += 1;
      SolverState* hs = nullptr;
      int rcp = poll_state(stream, w.st, &hs);
      if (rcp != PA_OK) return rcp;
      PA_CUDA(cudaGetLastError());
      if (!hs->done) return fail(PA_ERR_CUDA, "cooperative CG kernel returned without latching `done`");
      fill_report(rep, hs, L.count);
      set_swaps(rep, hs->itr + (hs->status == PA_BAD_TOL ? 1 : 0));
      return PA_OK;
    }
    cudaGetLastError();
    if (cfg->variant == 5) return fail(PA_ERR_UNSUPPORTED, "cooperative launch is not available on this device");
  }
  auto iteration = [&](T* cur, T* nxt) {
    const EqDev<T>& e = use_tma ? eq_cg : ((cur == x) ? eq : eq_alt);
    if (method == PA_METHOD_CG)
      cg_iteration<T>(L, g, e, nfaces, faces, w, cur, nxt, tiled, plan, cur == x ? 0 : 1, nullptr,
                      use_tma ? &tmap : nullptr, dist);
    else if (method == PA_METHOD_BICGSTAB)
      bicgstab_iteration<T>(L, g, e, nfaces, faces, w, cur, nxt, pw, dist);
    else
      jacobi_iteration<T>(L, g, e, nfaces, faces, w, cur, nxt, rhs, pw, dist);
    if (dist && nonlinear && method != PA_METHOD_JACOBI) {  // (Jacobi exchanges its new iterate anyway)
      dist_halo_exchange<T>(*dist, nxt, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, L.s);
      ++L.count;
    }
  };

  const long long max_iters = (method == PA_METHOD_BICGSTAB)
                                  ? (cfg->max_it > 1 ? cfg->max_it : 1)
                                  : (long long)cfg->max_it + 1;
  int check_every = cfg->check_every > 0 ? cfg->check_every : 32;
  if (check_every & 1) ++check_every;  // whole pairs

  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int per_pair = 0;
  if (cfg->use_graph) {
    int before = L.count;
    PA_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    iteration(x, x_alt);
    iteration(x_alt, x);
    cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    per_pair = L.count - before;
    L.count = before;
    if (ce != cudaSuccess) return fail(PA_ERR_CUDA, "graph capture failed");
    PA_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  }

  SolverState* h = nullptr;
  long long it = 0;
  int rc = PA_OK;
  while (true) {
    // never queue more iterations than max_it allows (each queued iteration past `done` is a
    // train of no-op launches)
    long long chunk = check_every;
    const long long left = max_iters + 1 - it;
    if (left < chunk) chunk = left > 2 ? left + (left & 1) : 2;
    for (long long k = 0; k < chunk; k += 2) {
      if (gexec) {
        cudaError_t e = cudaGraphLaunch(gexec, stream);
        if (e != cudaSuccess) {
          rc = fail(PA_ERR_CUDA, std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
          break;
        }
        L.count += per_pair;
      } else {
        iteration(x, x_alt);
        iteration(x_alt, x);
      }
    }
    if (rc != PA_OK) break;
    if (!L.ok) {
      rc = fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled failed (TMA tile descriptor)");
      break;
    }
    it += chunk;
    rc = poll_state(stream, w.st, &h);
    if (rc != PA_OK) break;
    if (h->done) break;
    if (it > max_iters + 2) {
      rc = fail(PA_ERR_CUDA, "solver did not latch `done` (internal error)");
      break;
    }
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  // a distributed solve that ends on an error leaves the host's epoch behind the mailbox contents (and the
  // ranks may have left at different reductions): never trust the mailboxes again, ncclAllReduce takes over
  if (dist && nccl_first_error() != ncclSuccess) {
    g_p2p.ready = false;
    return fail(PA_ERR_NCCL, std::string("NCCL: ") + nccl_api().GetErrorString(nccl_first_error()));
  }
  if (rc != PA_OK) {
    if (dist) g_p2p.ready = false;
    return rc;
  }
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    if (dist) g_p2p.ready = false;
    return fail(PA_ERR_CUDA, cudaGetErrorString(le));
  }
  if (h->status == PA_PEER_LOST) {
    g_p2p.ready = false;  // the mailboxes are in an unknown state: never use them again
    return fail(PA_ERR_NCCL, "multi-GPU: a peer rank did not join a fused all-reduce within 60 s (peer-memory watchdog)");
  }
  if (dist && method == PA_METHOD_BICGSTAB) {
    // the axpy stages only touched owned planes: refresh the ghost planes of both iterates
    const long long plane = (long long)g.n[1] * g.n[2];
    dist_halo_exchange<T>(*dist, x, plane, g.olo0, g.ohi0, stream);
    dist_halo_exchange<T>(*dist, x_alt, plane, g.olo0, g.ohi0, stream);
    PA_CUDA(cudaStreamSynchronize(stream));
  }
  if (dist && g_p2p.ready) g_p2p.epoch = h->epoch - 1ull;  // unchanged if no reduction used the mailboxes
  fill_report(rep, h, L.count);
  // ping-pong swaps == x updates performed.  CG / Jacobi bump itr after the update (an invalid tolerance
  // latches in between: x written, itr not bumped); BiCGSTAB bumps itr at the head of the iteration and can
  // latch PA_BAD_TOL before the update (linalg.py:233-240), so it counts its updates itself.
  int swaps = (method == PA_METHOD_BICGSTAB) ? h->swaps : h->itr + (h->status == PA_BAD_TOL ? 1 : 0);
  set_swaps(rep, swaps);
  return PA_OK;
}

// PA_APPLY_VARIANT=generic forces k_apply / k_grad, =tma the TMA star engine, =direct the direct 2-D kernel wherever
// it applies (A/B runs of the paths; read per call).  Default: 2-D grids up to kDirect2dCells cells take the direct
// kernel (no pipeline to fill: 1024^2 in ~4 us instead of ~10), everything else eligible the TMA engine.
static bool apply_force_generic() {
  const char* e = getenv("PA_APPLY_VARIANT");
  return e != nullptr && strcmp(e, "generic") == 0;
}
static bool apply_use_direct2d(bool eligible) {
  const char* e = getenv("PA_APPLY_VARIANT");
  if (e != nullptr && (strcmp(e, "tma") == 0 || strcmp(e, "generic") == 0)) return false;
  return eligible;
}

static inline dim3 shell_grid(const GridDev& g) {
  long long m = 1;
  for (int ax = 0; ax < 3; ++ax) {
    int bb = (ax == 0) ? 1 : 0, cc = (ax == 2) ? 1 : 2;
    long long nc = (long long)g.n[bb] * g.n[cc];
    if (g.act[ax] && nc > m) m = nc;
  }
  long long bx = (m + kBlock - 1) / kBlock;
  if (bx > kNumSMs * 4) bx = kNumSMs * 4;
  return dim3((unsigned)bx, 6);
}

// ops._Aop / FDC().laplacian / .div: constant-coefficient stars go through the TMA star engine
// (PW_APPLY, 2 words per cell); edge=True adds the shell pass with the one-sided face formulas.
template <typename T>
static int apply_impl(const pa_grid* pg, const pa_equation* peq, const T* phi, T* out,
                      cudaStream_t s) {
  GridDev g = make_grid(*pg);
  EqDev<T> eq = make_eq<T>(*peq);
  if (apply_use_direct2d(direct2d_eligible<T>(g, *peq))) {
    if (!launch_apply_direct2d<T, false>(s, g, eq, phi, out)) return fail(PA_ERR_CUDA, "k_apply_direct2d launch failed");
    if (peq->ops[0].edge != 0) k_apply_shell<T, false><<<shell_grid(g), kBlock, 0, s>>>(g, eq, phi, out);
  } else if (!apply_force_generic() && apply_eligible<T>(g, *peq)) {
    TilePlan tile;
    apply_tile_plan<T>(g, tile);
    if (!launch_star_tma<T, PW_APPLY>(s, g, eq, tile, phi, nullptr, out, nullptr, (T)0, nullptr, nullptr, ST_NONE))
      return fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled failed (TMA tile descriptor)");
    if (peq->ops[0].edge != 0) k_apply_shell<T, false><<<shell_grid(g), kBlock, 0, s>>>(g, eq, phi, out);
  } else {
    k_apply<T><<<grid_blocks(g.cells), kBlock, 0, s>>>(g, eq, phi, out);
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

}  // namespace pa

using namespace pa;

extern "C" {

const char* pa_last_error(void) { return g_err.c_str(); }
int pa_abi_version(void) { return PA_ABI_VERSION; }
int pa_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

#define PA_REQUIRE_DEVICE()                                                               \
  do {                                                                                    \
    if (pa_device_count() < 1)                                                            \
      return fail(PA_ERR_CUDA, "no CUDA device: pyapes_b200 has no CPU path");            \
  } while (0)

// Every entry point runs on the device that OWNS its arrays, whatever the caller's current device is
// (a Field on cuda:1 in a process whose current device is 0): streams, scratch buffers and function
// attributes are cached per device ordinal, and the caller's device is restored on return.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    if (p == nullptr || cudaGetDevice(&prev) != cudaSuccess) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    if (a.type == cudaMemoryTypeDevice && a.device != prev && cudaSetDevice(a.device) == cudaSuccess) switched = true;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

#define PA_DISPATCH(dtype, CALL)                                                          \
  do {                                                                                    \
    if ((dtype) == PA_F64) {                                                              \
      typedef double T;                                                                   \
      return CALL;                                                                        \
    } else if ((dtype) == PA_F32) {                                                       \
      typedef float T;                                                                    \
      return CALL;                                                                        \
    }                                                                                     \
    return fail(PA_ERR_ARG, "dtype must be PA_F32 or PA_F64");                            \
  } while (0)

int pa_stencil_apply(const pa_grid* g, const pa_equation* eq, int dtype, const void* phi,
                     void* out, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq))) return rc;
  if (!phi || !out || phi == out) return fail(PA_ERR_ARG, "phi/out null or aliased");
  PA_DISPATCH(dtype, apply_impl<T>(g, eq, (const T*)phi, (T*)out, (cudaStream_t)stream));
}

}  // extern "C"
template <typename T>
static int grad_impl(const pa_grid* pg, const pa_op* op, const T* phi, T* out, cudaStream_t s) {
  GridDev g = make_grid(*pg);
  pa_equation e;
  memset(&e, 0, sizeof(e));
  e.nops = 1;
  e.ops[0] = *op;
  EqDev<T> eq = make_eq<T>(e);
  if (apply_use_direct2d(direct2d_eligible<T>(g, e))) {
    if (!launch_apply_direct2d<T, true>(s, g, eq, phi, out)) return fail(PA_ERR_CUDA, "k_apply_direct2d launch failed");
    if (op->edge != 0) k_apply_shell<T, true><<<shell_grid(g), kBlock, 0, s>>>(g, eq, phi, out);
  } else if (!apply_force_generic() && apply_eligible<T>(g, e)) {  // PW_GRAD, 1 + d words per cell
    TilePlan tile;
    apply_tile_plan<T>(g, tile);
    if (!launch_star_grad<T>(s, g, eq, tile, phi, out))
      return fail(PA_ERR_CUDA, "cuTensorMapEncodeTiled failed (TMA tile descriptor)");
    if (op->edge != 0) k_apply_shell<T, true><<<shell_grid(g), kBlock, 0, s>>>(g, eq, phi, out);
  } else {
    k_grad<T><<<grid_blocks(g.cells), kBlock, 0, s>>>(g, eq.op[0], phi, out);
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}
extern "C" {

int pa_grad_apply(const pa_grid* g, const pa_op* op, int dtype, const void* phi, void* out,
                  void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g))) return rc;
  if (!op || op->kind != PA_OP_STAR) return fail(PA_ERR_ARG, "grad needs a PA_OP_STAR operator");
  if (!phi || !out || phi == out) return fail(PA_ERR_ARG, "phi/out null or aliased");
  PA_DISPATCH(dtype, grad_impl<T>(g, op, (const T*)phi, (T*)out, (cudaStream_t)stream));
}

}  // extern "C"
template <typename T>
static int bc_impl(const pa_grid* pg, int nfaces, const pa_face_bc* faces, T* phi,
                   cudaStream_t s) {
  GridDev g = make_grid(*pg);
  Launcher L{s};
  launch_bcs<T>(L, g, nfaces, faces, phi, nullptr);
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}
extern "C" {

int pa_bc_apply(const pa_grid* g, int nfaces, const pa_face_bc* faces, int dtype, void* phi,
                void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!phi) return fail(PA_ERR_ARG, "phi is null");
  PA_DISPATCH(dtype, bc_impl<T>(g, nfaces, faces, (T*)phi, (cudaStream_t)stream));
}

size_t pa_solver_workspace_bytes(const pa_grid* g, int dtype, int method) {
  if (!g) return 0;
  long long cells = (long long)g->n[0] * g->n[1] * g->n[2];
  return ws_bytes(cells, dtype == PA_F64 ? 8 : 4, method_nvec(method));
}

// Periodic faces along the slab axis: both ends must be periodic and every rank must carry the
// wrap-around ghost planes (pyapes_b200.parallel.SlabMesh(..., periodic=True)).
static int slab_ring(const pa_grid* g, int nfaces, const pa_face_bc* faces, int nranks, int* ring) {
  int lo = 0, hi = 0;
  for (int f = 0; f < nfaces; ++f)
    if (faces[f].axis == 0 && faces[f].kind == PA_BC_PERIODIC) (faces[f].side < 0 ? lo : hi) = 1;
  *ring = 0;
  if (nranks < 2 || (!lo && !hi)) return PA_OK;
  if (lo != hi) return fail(PA_ERR_UNSUPPORTED, "multi-GPU: a single periodic face along the slab axis");
  if (g->olo0 != 1 || g->ohi0 != g->n[0] - 1)
    return fail(PA_ERR_ARG, "multi-GPU: periodic slab axis needs a ghost plane on both sides of every rank");
  *ring = 1;
  return PA_OK;
}

static int check_solver_args(const pa_grid* g, const pa_equation* eq, int nfaces,
                             const pa_face_bc* faces, void* x, void* x_alt, const void* rhs,
                             const pa_solver_cfg* cfg, void* ws, pa_report* rep) {
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!x || !x_alt || !rhs || !cfg || !ws || !rep) return fail(PA_ERR_ARG, "null argument");
  if (x == x_alt) return fail(PA_ERR_ARG, "x and x_alt must not alias");
  if (cfg->max_it < 0) return fail(PA_ERR_ARG, "max_it must be >= 0");
  return PA_OK;
}

int pa_cg_solve(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg,
                void* ws, size_t ws_bytes_, pa_report* report, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc = check_solver_args(g, eq, nfaces, faces, x, x_alt, rhs, cfg, ws, report);
  if (rc) return rc;
  PA_DISPATCH(dtype, run_solver<T>(PA_METHOD_CG, g, eq, nfaces, faces, (T*)x, (T*)x_alt,
                                   (const T*)rhs, cfg, ws, ws_bytes_, report,
                                   (cudaStream_t)stream));
}

int pa_comm_unique_id(void* out128) {
  NcclApi& a = nccl_api();
  if (!a.ok) return fail(PA_ERR_NCCL, a.error);
  ncclUniqueId id;
  ncclResult_t rc = a.GetUniqueId(&id);
  if (rc != ncclSuccess) return fail(PA_ERR_NCCL, a.GetErrorString(rc));
  memcpy(out128, &id, sizeof(id));
  return PA_OK;
}

int pa_comm_create(const void* id128, int rank, int nranks, void** comm_out) {
  PA_REQUIRE_DEVICE();
  NcclApi& a = nccl_api();
  if (!a.ok) return fail(PA_ERR_NCCL, a.error);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c;
  ncclResult_t rc = a.CommInitRank(&c, nranks, id, rank);
  if (rc != ncclSuccess) return fail(PA_ERR_NCCL, a.GetErrorString(rc));
  *comm_out = (void*)c;
  return PA_OK;
}

int pa_comm_destroy(void* comm) {
  NcclApi& a = nccl_api();
  if (!a.ok || !comm) return PA_OK;
  a.CommDestroy((ncclComm_t)comm);
  return PA_OK;
}

int pa_cg_solve_dist(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                     int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg, void* ws,
                     size_t ws_bytes_, void* comm, int rank, int nranks, pa_report* report,
                     void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc = check_solver_args(g, eq, nfaces, faces, x, x_alt, rhs, cfg, ws, report);
  if (rc) return rc;
  if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(PA_ERR_ARG, "bad communicator");
  Dist d{(ncclComm_t)comm, rank, nranks};
  if ((rc = slab_ring(g, nfaces, faces, nranks, &d.ring))) return rc;
  PA_DISPATCH(dtype, run_solver<T>(PA_METHOD_CG, g, eq, nfaces, faces, (T*)x, (T*)x_alt,
                                   (const T*)rhs, cfg, ws, ws_bytes_, report, (cudaStream_t)stream,
                                   nranks > 1 ? &d : nullptr));
}

int pa_p2p_local_handle(void* out64) {
  PA_REQUIRE_DEVICE();
  if (!out64) return fail(PA_ERR_ARG, "null argument");
  if (!g_p2p.local) {
    // mailbox + halo flags + landing zone in ONE allocation: one IPC handle per rank maps all of it
    size_t cap = halo_cap_bytes();
    size_t bytes = kHaloLandOff + 4 * cap;
    if (cudaMalloc((void**)&g_p2p.local, bytes) != cudaSuccess) {  // no room for the landing zone: mailboxes only
      cudaGetLastError();
      cap = 0;
      bytes = kHaloLandOff;
      PA_CUDA(cudaMalloc((void**)&g_p2p.local, bytes));
    }
    g_p2p.halo_cap = cap;
    PA_CUDA(cudaMemset(g_p2p.local, 0, kHaloLandOff));
    PA_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  PA_CUDA(cudaIpcGetMemHandle(&h, g_p2p.local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, sizeof(h));
  return PA_OK;
}

int pa_p2p_attach(const void* handles, int rank, int nranks) {
  PA_REQUIRE_DEVICE();
  if (!handles || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || !g_p2p.local)
    return fail(PA_ERR_ARG, "bad argument (call pa_p2p_local_handle first; at most 16 ranks)");
  std::vector<unsigned long long*> ptrs((size_t)nranks, nullptr);
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) {
      ptrs[p] = g_p2p.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)p, sizeof(h));
    void* mapped = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(PA_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    ptrs[p] = (unsigned long long*)mapped;
  }
  for (int p = 0; p < nranks; ++p) g_p2p.peer_base[p] = (char*)ptrs[p];
  if (!g_p2p.peers_dev) PA_CUDA(cudaMalloc((void**)&g_p2p.peers_dev, 16 * sizeof(unsigned long long*)));
  PA_CUDA(cudaMemcpy(g_p2p.peers_dev, ptrs.data(), (size_t)nranks * sizeof(unsigned long long*),
                     cudaMemcpyHostToDevice));
  g_p2p.rank = rank;
  g_p2p.nranks = nranks;
  g_p2p.epoch = 0;
  g_p2p.ready = true;
  return PA_OK;
}

int pa_p2p_enabled(void) { return g_p2p.ready ? 1 : 0; }
long long pa_p2p_halo_cap(void) { return (long long)g_p2p.halo_cap; }
int pa_p2p_set_halo_cap(long long bytes) {
  if (bytes < 0 || (size_t)bytes > g_p2p.halo_cap) return fail(PA_ERR_ARG, "halo cap can only be lowered");
  g_p2p.halo_cap = (size_t)bytes;
  return PA_OK;
}
int pa_p2p_disable(void) {
  g_p2p.ready = false;
  return PA_OK;
}

int pa_solve_dist(int method, const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg, void* ws,
                  size_t ws_bytes_, void* comm, int rank, int nranks, pa_report* report, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc = check_solver_args(g, eq, nfaces, faces, x, x_alt, rhs, cfg, ws, report);
  if (rc) return rc;
  if (method != PA_METHOD_CG && method != PA_METHOD_BICGSTAB && method != PA_METHOD_JACOBI)
    return fail(PA_ERR_ARG, "unknown solver method");
  if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(PA_ERR_ARG, "bad communicator");
  Dist d{(ncclComm_t)comm, rank, nranks};
  if ((rc = slab_ring(g, nfaces, faces, nranks, &d.ring))) return rc;
  PA_DISPATCH(dtype, run_solver<T>(method, g, eq, nfaces, faces, (T*)x, (T*)x_alt, (const T*)rhs, cfg, ws,
                                   ws_bytes_, report, (cudaStream_t)stream, nranks > 1 ? &d : nullptr));
}

}  // extern "C"
template <typename T>
static int halo_impl(const pa_grid* pg, T* phi, const Dist& d, cudaStream_t s) {
  if (!nccl_api().ok) return fail(PA_ERR_NCCL, nccl_api().error);
  nccl_first_error() = ncclSuccess;
  dist_halo_exchange<T>(d, phi, (long long)pg->n[1] * pg->n[2], pg->olo0, pg->ohi0, s);
  PA_CUDA(cudaStreamSynchronize(s));
  if (nccl_first_error() != ncclSuccess)
    return fail(PA_ERR_NCCL, std::string("NCCL: ") + nccl_api().GetErrorString(nccl_first_error()));
  return PA_OK;
}
extern "C" {

int pa_halo_exchange(const pa_grid* g, int dtype, void* phi, int ring, void* comm, int rank, int nranks, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g))) return rc;
  if (!phi) return fail(PA_ERR_ARG, "phi is null");
  if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(PA_ERR_ARG, "bad communicator");
  if (nranks == 1) return PA_OK;
  Dist d{(ncclComm_t)comm, rank, nranks};
  d.ring = ring ? 1 : 0;
  PA_DISPATCH(dtype, halo_impl<T>(g, (T*)phi, d, (cudaStream_t)stream));
}

int pa_cg_profile(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, void* x, void* x_alt, const void* rhs, int iters, int variant, void* ws,
                  size_t ws_bytes_, double* out_ms, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!x || !x_alt || !rhs || !ws || !out_ms || iters < 1) return fail(PA_ERR_ARG, "bad argument");
  PA_DISPATCH(dtype, profile_cg<T>(g, eq, nfaces, faces, (T*)x, (T*)x_alt, (const T*)rhs, iters,
                                   variant, ws, ws_bytes_, out_ms, (cudaStream_t)stream));
}

int pa_bicgstab_solve(const pa_grid* g, const pa_equation* eq, int nfaces,
                      const pa_face_bc* faces, int dtype, void* x, void* x_alt,
                      const void* rhs, const pa_solver_cfg* cfg, void* ws, size_t ws_bytes_,
                      pa_report* report, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc = check_solver_args(g, eq, nfaces, faces, x, x_alt, rhs, cfg, ws, report);
  if (rc) return rc;
  PA_DISPATCH(dtype, run_solver<T>(PA_METHOD_BICGSTAB, g, eq, nfaces, faces, (T*)x, (T*)x_alt,
                                   (const T*)rhs, cfg, ws, ws_bytes_, report,
                                   (cudaStream_t)stream));
}

int pa_jacobi_solve(const pa_grid* g, const pa_equation* eq, int nfaces,
                    const pa_face_bc* faces, int dtype, void* x, void* x_alt, const void* rhs,
                    const pa_solver_cfg* cfg, void* ws, size_t ws_bytes_, pa_report* report,
                    void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  int rc = check_solver_args(g, eq, nfaces, faces, x, x_alt, rhs, cfg, ws, report);
  if (rc) return rc;
  PA_DISPATCH(dtype, run_solver<T>(PA_METHOD_JACOBI, g, eq, nfaces, faces, (T*)x, (T*)x_alt,
                                   (const T*)rhs, cfg, ws, ws_bytes_, report,
                                   (cudaStream_t)stream));
}

}  // extern "C"
// out = y + a*x  (one rounding per operation; out may alias y)
template <typename T>
__global__ void __launch_bounds__(kBlock) k_axpy(long long n, T a, const T* __restrict__ x, const T* y, T* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T t = a * x[i];
    out[i] = y[i] + t;
  }
}
template <typename T>
static int axpy_impl(long long n, double a, const T* x, const T* y, T* out, cudaStream_t s) {
  if (n == 0) return PA_OK;
  k_axpy<T><<<grid_blocks(n), kBlock, 0, s>>>(n, (T)a, x, y, out);
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

// n explicit Euler steps, ping-ponging between `a` (holds phi on entry) and `b`.  A pair of
// steps is captured once as a CUDA graph and replayed.  Static shell (all faces Dirichlet):
// after the first step the shell already holds the BC values and later steps copy it, so BC
// launches are only needed in the first step.
// Multi-GPU: every step ends with the halo exchange of the new field's boundary planes.
template <typename T>
static int euler_impl(const pa_grid* pg, const pa_equation* peq, int nfaces,
                      const pa_face_bc* faces, T* a, T* b, const T* rhs, double dt, int nsteps,
                      int* result_in_b, cudaStream_t caller, const Dist* dist = nullptr) {
  cudaStream_t s;
  int rcs = solver_stream(caller, &s);
  if (rcs != PA_OK) return rcs;
  GridDev g = make_grid(*pg);
  EqDev<T> eq_a = make_eq<T>(*peq);
  EqDev<T> eq_b = eq_a;  // nonlinear advection: the speed is the field being advanced (see run_solver)
  for (int k = 0; k < peq->nops; ++k)
    if (peq->ops[k].kind != PA_OP_STAR && peq->ops[k].adv_is_iterate) {
      eq_a.op[k].adv = a;
      eq_b.op[k].adv = b;
    }
  Launcher L{s};
  const bool tma = pw_eligible<T>(g, *peq, nfaces, faces);
  TilePlan tile;
  if (tma) pw_tile_plan<T>(g, tile, nfaces, faces);
  const bool stat = static_shell(nfaces, faces) != 0;
  auto one = [&](T* cur, T* nxt, bool bcs) {
    const EqDev<T>& eq = (cur == a) ? eq_a : eq_b;
    bool done = false;
    if (tma)
      done = launch_star_tma<T, PW_EULER>(s, g, eq, tile, cur, rhs, nxt, nullptr, (T)dt, nullptr, nullptr, ST_NONE);
    if (!done)
      k_pointwise_update<T, 1><<<grid_blocks(g.cells), kBlock, 0, s>>>(g, eq, cur, nxt, rhs, (T)dt, nullptr,
                                                                       nullptr);
    if (bcs) launch_bcs<T>(L, g, nfaces, faces, nxt, nullptr, dist);
    if (dist) dist_halo_exchange<T>(*dist, nxt, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, s);
  };
  if (dist) {
    if (!nccl_api().ok) return fail(PA_ERR_NCCL, nccl_api().error);
    nccl_first_error() = ncclSuccess;
    if (dist->ring) bc_scratch((size_t)2 * g.n[1] * g.n[2] * sizeof(T));  // before any capture
    if (nsteps > 0) dist_halo_exchange<T>(*dist, a, (long long)g.n[1] * g.n[2], g.olo0, g.ohi0, s);
  }
  int done_steps = 0;
  T* cur = a;
  T* nxt = b;
  if (nsteps > 0) {  // the first step always applies the BCs
    one(cur, nxt, true);
    std::swap(cur, nxt);
    ++done_steps;
  }
  int remaining = nsteps - done_steps;
  // 2-D grids that fit the SMs' shared memory: all remaining steps in ONE launch, the field resident in shared
  // memory (kernels_resident.cuh).  PA_EULER_VARIANT=stream keeps the per-step launches (A/B runs; read per call).
  if (remaining >= 2 && !dist && stat && tma) {
    bool linear = true;
    for (int k = 0; k < peq->nops; ++k) linear &= peq->ops[k].kind == PA_OP_STAR;
    const char* ev = getenv("PA_EULER_VARIANT");
    if (linear && !(ev != nullptr && strcmp(ev, "stream") == 0) &&
        launch_euler_resident<T>(s, g, *peq, eq_a, cur, nxt, rhs, (T)dt, remaining)) {
      if (remaining & 1) std::swap(cur, nxt);
      done_steps = nsteps;
      remaining = 0;
      PA_CUDA(cudaStreamSynchronize(s));
      if (res_check_abort()) return fail(PA_ERR_CUDA, "resident Euler kernel: a row exchange timed out (watchdog)");
    }
  }
  if (remaining >= 4) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    PA_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    one(cur, nxt, !stat);
    one(nxt, cur, !stat);
    if (cudaStreamEndCapture(s, &graph) != cudaSuccess) return fail(PA_ERR_CUDA, "graph capture failed");
    PA_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
    for (int k = 0; k + 1 < remaining; k += 2) {
      PA_CUDA(cudaGraphLaunch(gexec, s));
      done_steps += 2;
    }
    PA_CUDA(cudaStreamSynchronize(s));
    cudaGraphExecDestroy(gexec);
    cudaGraphDestroy(graph);
  }
  while (done_steps < nsteps) {
    one(cur, nxt, !stat);
    std::swap(cur, nxt);
    ++done_steps;
  }
  PA_CUDA(cudaStreamSynchronize(s));
  PA_CUDA(cudaGetLastError());
  if (dist && nccl_first_error() != ncclSuccess)
    return fail(PA_ERR_NCCL, std::string("NCCL: ") + nccl_api().GetErrorString(nccl_first_error()));
  if (!L.ok) return fail(PA_ERR_CUDA, "device scratch allocation failed (slab-periodic boundary condition)");
  *result_in_b = (cur == b) ? 1 : 0;
  return PA_OK;
}

extern "C" {

int pa_euler_step(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, const void* phi, void* phi_new, const void* rhs, double dt,
                  void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!phi || !phi_new || phi == phi_new) return fail(PA_ERR_ARG, "phi/phi_new null or aliased");
  int in_b = 0;
  PA_DISPATCH(dtype, euler_impl<T>(g, eq, nfaces, faces, (T*)phi, (T*)phi_new, (const T*)rhs, dt, 1, &in_b,
                                   (cudaStream_t)stream));
}

int pa_euler_steps(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                   int dtype, void* phi, void* phi_alt, const void* rhs, double dt, int nsteps,
                   int* result_in_alt, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!phi || !phi_alt || phi == phi_alt || !result_in_alt || nsteps < 0)
    return fail(PA_ERR_ARG, "bad argument");
  PA_DISPATCH(dtype, euler_impl<T>(g, eq, nfaces, faces, (T*)phi, (T*)phi_alt, (const T*)rhs, dt, nsteps,
                                   result_in_alt, (cudaStream_t)stream));
}

int pa_euler_steps_dist(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                        int dtype, void* phi, void* phi_alt, const void* rhs, double dt, int nsteps,
                        int* result_in_alt, void* comm, int rank, int nranks, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(phi);
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!phi || !phi_alt || phi == phi_alt || !result_in_alt || nsteps < 0)
    return fail(PA_ERR_ARG, "bad argument");
  if (!comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(PA_ERR_ARG, "bad communicator");
  Dist d{(ncclComm_t)comm, rank, nranks};
  if ((rc = slab_ring(g, nfaces, faces, nranks, &d.ring))) return rc;
  PA_DISPATCH(dtype, euler_impl<T>(g, eq, nfaces, faces, (T*)phi, (T*)phi_alt, (const T*)rhs, dt, nsteps,
                                   result_in_alt, (cudaStream_t)stream, nranks > 1 ? &d : nullptr));
}

int pa_axpy(int dtype, long long n, double a, const void* x, const void* y, void* out, void* stream) {
  PA_REQUIRE_DEVICE();
  DeviceGuard guard(x);
  if (n < 0 || !x || !y || !out) return fail(PA_ERR_ARG, "bad argument");
  PA_DISPATCH(dtype, axpy_impl<T>(n, a, (const T*)x, (const T*)y, (T*)out, (cudaStream_t)stream));
}

int pa_cg_solve_host(const pa_grid* g, const pa_equation* eq, int nfaces,
                     const pa_face_bc* faces, int dtype, void* x_host, const void* rhs_host,
                     const pa_solver_cfg* cfg, pa_report* report) {
  PA_REQUIRE_DEVICE();
  int rc;
  if ((rc = check_grid(g)) || (rc = check_eq(eq)) || (rc = check_faces(nfaces, faces))) return rc;
  if (!x_host || !rhs_host || !cfg || !report) return fail(PA_ERR_ARG, "null argument");
  size_t esz = dtype == PA_F64 ? 8 : 4;
  size_t vb = (size_t)g->n[0] * g->n[1] * g->n[2] * esz;
  size_t wsb = pa_solver_workspace_bytes(g, dtype, PA_METHOD_CG);
  char* dev = nullptr;
  PA_CUDA(cudaMalloc((void**)&dev, 3 * align_up(vb, 256) + wsb));
  char* x = dev;
  char* xa = dev + align_up(vb, 256);
  char* rhs = dev + 2 * align_up(vb, 256);
  char* ws = dev + 3 * align_up(vb, 256);
  cudaStream_t s = nullptr;
  cudaError_t e = cudaMemcpyAsync(x, x_host, vb, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rhs, rhs_host, vb, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) {
    cudaFree(dev);
    return fail(PA_ERR_CUDA, cudaGetErrorString(e));
  }
  rc = pa_cg_solve(g, eq, nfaces, faces, dtype, x, xa, rhs, cfg, ws, wsb, report, (void*)s);
  if (rc == PA_OK) {
    e = cudaMemcpyAsync(x_host, report->result_in_alt ? xa : x, vb, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(PA_ERR_CUDA, cudaGetErrorString(e));
    report->result_in_alt = 0;
  }
  cudaFree(dev);
  return rc;
}

}  // extern "C"
