// TMA-staged CG kernels (sm_100a): the fastest path for ONE constant-coefficient star
// operator on 2-D/3-D grids whose axes 1 and 2 are not periodic.
//
// Producer/consumer pipeline per CTA (no CTA-wide barrier in the plane loop):
//   * warp 8 (one elected lane) is the PRODUCER: for every plane of the CTA's chunk it issues
//     `cp.async.bulk.tensor.3d` (TMA) loads of the tile *with its halo* — box
//     (TY+2) x (TZ+2*VEC) — into a ring of S shared-memory stages; completion is signalled on
//     a per-stage `full` mbarrier (expect_tx bytes).  Out-of-bounds halo cells are zero-filled
//     by the TMA unit, which is exactly what a non-periodic edge needs (those values only
//     reach cells outside the solver region, which are masked).
//   * warps 0-7 are CONSUMERS: a thread owns K::RY consecutive rows x one 16-byte vector and keeps
//     planes x-1, x, x+1 of its cells in registers (3x unrolled loop: no register rotation);
//     axis-1/axis-2 neighbours are read straight from the staged halo tile.  When a warp is
//     done with a stage it arrives on that stage's `empty` mbarrier; the producer refills it.
//   The pipeline is S-2 planes deep, so DRAM latency is covered by TMA transactions in flight,
//   not by occupancy or prefetch registers.
//
// CG fusion as in kernels_tiled.cuh (SURVEY.md §8d canonical variant, 8 words/cell/iteration):
//   phase A: d_new = r + beta*d ; dAd = sum d_new*A(d_new)     [R r, R d, W d]
//            (d_new on the halo is recomputed from the staged raw r, d — bit-identical)
//   phase B: x_new = x + alpha*d ; r -= alpha*A(d) ; sums      [R x, R d, R r, W x, W r]
// Per-cell arithmetic order is identical to eval_equation() in common.cuh.
#pragma once
#include <cstring>
#include <cuda.h>

#include "common.cuh"
#include "kernels_generic.cuh"
#include "kernels_tiled.cuh"

namespace pa {

// tile shapes: K::RY consecutive rows per thread; the 8 consumer warps are stacked along axis 1
// (3-D meshes) or, FLAT, laid side by side along the contiguous axis (2-D meshes: n1 == 1, the
// march runs along the first mesh axis and a "plane" is a single row)
struct KStd {
  static constexpr int RY = 2;
  static constexpr bool FLAT = false;
};
struct KFlat {
  static constexpr int RY = 1;
  static constexpr bool FLAT = true;
};

template <typename T, typename K>
struct TmaCfg {
  static constexpr int VEC = VecOf<T>::N;
  static constexpr int RY = K::RY;
  static constexpr bool FLAT = K::FLAT;
  static constexpr int CWARPS = 8;                   // consumer warps
  static constexpr int THREADS = (CWARPS + 1) * 32;  // + producer warp
  static constexpr int WZ = FLAT ? CWARPS : 1;       // warps along axis 2
  static constexpr int TY = FLAT ? 1 : CWARPS * RY;
  static constexpr int TZ = WZ * 32 * VEC;
  static constexpr int HZ = VEC;                     // z halo, keeps own cells 16-B aligned
  // A TMA box dimension is limited to 256 elements, so a FLAT tile (512 cells wide) is staged as
  // NB boxes, one per consumer warp, each with its own z halo.
  static constexpr int NB = FLAT ? CWARPS : 1;
  static constexpr int OBOXZ = TZ / NB;              // own-tile box width
  static constexpr int BOXZ = OBOXZ + 2 * HZ;        // halo box width
  static constexpr int BOXY = FLAT ? 1 : TY + 2;     // no axis-1 halo when axis 1 is inactive
  static constexpr int S = 4;                        // pipeline stages (power of two)
  static constexpr int HBOX_BYTES = BOXY * BOXZ * (int)sizeof(T);
  static constexpr int OBOX_BYTES = TY * OBOXZ * (int)sizeof(T);
  static constexpr int HBOX_SLOT = (HBOX_BYTES + 127) / 128 * 128;
  static constexpr int OBOX_SLOT = (OBOX_BYTES + 127) / 128 * 128;
  static constexpr int HALO_BYTES = NB * HBOX_BYTES;  // bytes one stage's halo tile transfers
  static constexpr int OWN_BYTES = NB * OBOX_BYTES;
  static constexpr int HALO_SLOT = NB * HBOX_SLOT;
  static constexpr int OWN_SLOT = NB * OBOX_SLOT;
  static constexpr int STAGE_A = 2 * HALO_SLOT;              // r halo, d halo
  static constexpr int STAGE_B = HALO_SLOT + 2 * OWN_SLOT;   // d halo, x own, r own
  static constexpr int BAR_BYTES = 128;                      // 2*S mbarriers
  static constexpr size_t SMEM_A = (size_t)S * STAGE_A + BAR_BYTES + 128;
  static constexpr size_t SMEM_B = (size_t)S * STAGE_B + BAR_BYTES + 128;
};


// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the
// hint expires) instead of spinning through try_wait/branch pairs -- the spin was 6.5 branches and
// 4 barrier instructions per cell in phase B (ncu source page), pure issue slots and power.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// ---- host: tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// planes / plane_stride_bytes: a stack of (n1 x n2) planes that is not the field itself (the halo landing zone)
template <typename T>
static inline bool make_map(CUtensorMap* m, const T* base, const GridDev& g, int boxz, int boxy, int planes = 0,
                            size_t plane_stride_bytes = 0) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)g.n[2], (cuuint64_t)g.n[1], (cuuint64_t)(planes > 0 ? planes : g.n[0])};
  cuuint64_t strides[2] = {(cuuint64_t)g.n[2] * sizeof(T),
                           plane_stride_bytes ? (cuuint64_t)plane_stride_bytes : (cuuint64_t)g.n[1] * g.n[2] * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)boxz, (cuuint32_t)boxy, 1u};
  cuuint32_t es[3] = {1u, 1u, 1u};
  CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(m, dt, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct TmaPlan {
  TilePlan tile;            // grid geometry (tiles_y/z, chunks, cx, fuse_fin)
  CUtensorMap x_own[2];     // x, x_alt  (phase B reads the current iterate)
  CUtensorMap r_own, r_halo;
  CUtensorMap d_halo[2];    // the two d buffers
  CUtensorMap r_land;       // this rank's landing zone: [slot 2][side 2] ghost planes of r (HaloDev)
  const void* r_ptr;        // raw arrays behind r_halo / d_halo (wrap-around reads, TilePlan::src*)
  const void* d_ptr[2];
  int coef_uniform = 0;     // coefficient classes of axes 1 and 2 are bitwise equal (star_cells UNI)
  int contract = 0;         // opt-in FMA contraction (PA_FLAG_CONTRACT; never with periodic axes 1/2)
};

// wave-aware chunking along axis 0 (shared by the CG and the star-engine plans)
inline void tma_chunks(const GridDev& g, int tiles, TilePlan& p) {
  const int slots = kNumSMs * 2;
  int best_c = 1;
  double best = -1.0;
  // chunks of >= 8 planes, except on small grids where filling the SMs matters more than the
  // two halo planes a chunk re-reads
  const int maxc = (long long)tiles * (g.n[0] / 8) >= slots ? g.n[0] / 8 : (g.n[0] >= 4 ? g.n[0] / 2 : 1);
  for (int c = 1; c <= maxc; ++c) {
    const int cx = (g.n[0] + c - 1) / c;
    const int cc = (g.n[0] + cx - 1) / cx;
    const long long items = (long long)cc * tiles;
    if (items > kMaxPartials) break;
    const long long waves = (items + slots - 1) / slots;
    const double quant = (double)items / (double)(waves * slots);
    const double halo = g.act[0] ? (double)cx / (double)(cx + 2) : 1.0;
    const double score = quant * halo * (items >= slots ? 1.0 : (double)items / slots);
    if (score > best + 1e-9) {
      best = score;
      best_c = cc;
    }
  }
  p.cx = (g.n[0] + best_c - 1) / best_c;
  p.chunks = (g.n[0] + p.cx - 1) / p.cx;
  p.vec_ok = 1;
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
}

// 3-D meshes: tiles of the (axis1, axis2) plane; 2-D meshes (axis 1 inactive): FLAT row tiles
inline bool tma_flat(const GridDev& g) { return !g.act[1]; }

template <typename T, typename K>
inline bool plan_tma_k(const GridDev& g, const T* x, const T* x_alt, const T* r, const T* d0, const T* d1,
                       TmaPlan& tp) {
  typedef TmaCfg<T, K> C;
  if (g.n[2] % C::VEC != 0 || g.n[2] < 2 * C::VEC) return false;
  if (!C::FLAT && g.n[1] < 4) return false;
  TilePlan& p = tp.tile;
  p.ry = K::RY;
  p.tiles_y = C::FLAT ? 1 : (g.n[1] + C::TY - 1) / C::TY;
  p.tiles_z = (g.n[2] + C::TZ - 1) / C::TZ;
  tma_chunks(g, p.tiles_y * p.tiles_z, p);
  tp.r_ptr = r;
  tp.d_ptr[0] = d0;
  tp.d_ptr[1] = d1;
  return make_map<T>(&tp.x_own[0], x, g, C::OBOXZ, C::TY) && make_map<T>(&tp.x_own[1], x_alt, g, C::OBOXZ, C::TY) &&
         make_map<T>(&tp.r_own, r, g, C::OBOXZ, C::TY) && make_map<T>(&tp.r_halo, r, g, C::BOXZ, C::BOXY) &&
         make_map<T>(&tp.d_halo[0], d0, g, C::BOXZ, C::BOXY) && make_map<T>(&tp.d_halo[1], d1, g, C::BOXZ, C::BOXY);
}

// Periodic faces on kernel axes 1/2: a TMA box cannot wrap around, so boundary tiles fetch the
// wrapped halo row / column from global memory (wrap_halo); that needs whole tiles along the axis.
template <typename T>
inline bool tma_wrap_ok(const GridDev& g, int nfaces, const pa_face_bc* faces, int* wrap) {
  *wrap = 0;
  const bool flat = tma_flat(g);
  for (int f = 0; f < nfaces; ++f) {
    if (faces[f].kind != PA_BC_PERIODIC || faces[f].axis == 0) continue;
    if (faces[f].axis == 1) {
      if (flat || g.n[1] % TmaCfg<T, KStd>::TY != 0) return false;
    } else {
      const int tz = flat ? TmaCfg<T, KFlat>::TZ : TmaCfg<T, KStd>::TZ;
      if (g.n[2] % tz != 0) return false;
    }
    *wrap = 1;
  }
  return true;
}

// a STAR operator that is c*phi and nothing else, with sign +1 and no parameter: the form in
// which the host layer lowers the implicit-Euler term (pyapes_b200/solver/ops.py)
inline bool diagonal_op(const pa_op& o, double* c) {
  if (o.kind != PA_OP_STAR || o.param_field != nullptr || o.edge != 0 || o.has_param || o.sign != 1.0) return false;
  for (int a = 0; a < 3; ++a) {
    if (o.coef_tab[a]) return false;
    for (int cls = 0; cls < 3; ++cls)
      for (int k = 0; k < 3; ++k) {
        const bool centre = (a == 2 && k == 1);
        if (!centre && o.coef[a][cls][k] != 0.0) return false;
        if (centre && o.coef[a][cls][k] != o.coef[2][0][1]) return false;
      }
  }
  if (c) *c = o.coef[2][0][1];
  return true;
}

template <typename T>
inline bool plan_tma(const GridDev& g, const pa_equation& eq, int nfaces, const pa_face_bc* faces,
                     const T* x, const T* x_alt, const T* r, const T* d0, const T* d1, TmaPlan& tp) {
  const pa_op& o = eq.ops[0];
  if (!(eq.nops == 1 || (eq.nops == 2 && diagonal_op(eq.ops[1], nullptr))) || o.kind != PA_OP_STAR ||
      o.param_field != nullptr || o.edge != 0 || o.coef_tab[0] || o.coef_tab[1] || o.coef_tab[2])
    return false;
  if (!g.act[2] || (!g.act[1] && !g.act[0])) return false;  // 1-D meshes stay on the generic kernels
  int wrap = 0;
  if (!tma_wrap_ok<T>(g, nfaces, faces, &wrap)) return false;
  const bool ok = tma_flat(g) ? plan_tma_k<T, KFlat>(g, x, x_alt, r, d0, d1, tp)
                              : plan_tma_k<T, KStd>(g, x, x_alt, r, d0, d1, tp);
  tp.tile.wrap = wrap;
  tp.coef_uniform = 1;
  for (int a = 1; a < 3; ++a)
    for (int cls = 1; cls < 3; ++cls)
      for (int k = 0; k < 3; ++k)
        if (std::memcmp(&o.coef[a][cls][k], &o.coef[a][0][k], sizeof(double)) != 0) tp.coef_uniform = 0;
  return ok;
}

// ---- work items of the persistent kernels --------------------------------------------------------
// item id -> (chunk slot, tile row, tile column), chunk slot slowest: CTAs that run side by side work on
// neighbouring tiles of the same chunk (their halos meet in L2), and on slabs the boundary chunks come first.
struct WorkItem {
  int y0, z0, x0, x1;
  bool lean;  // the tile touches neither the array edge nor a boundary-adjacent coefficient class
};

template <typename C>
__device__ __forceinline__ WorkItem work_item(const TilePlan& p, const GridDev& g, int id) {
  WorkItem w;
  const int tz = id % p.tiles_z;
  const int t = id / p.tiles_z;
  const int ty = t % p.tiles_y;
  const int zc = t / p.tiles_y;
  w.z0 = tz * C::TZ;
  w.y0 = ty * C::TY;
  w.x0 = tile_chunk(p, zc) * p.cx;
  w.x1 = min(w.x0 + p.cx, g.n[0]);
  const bool full_tile = (C::FLAT || w.y0 + C::TY <= g.n[1]) && (w.z0 + C::TZ <= g.n[2]);
  const bool edge = (!C::FLAT && ((w.y0 < 2) || (w.y0 + C::TY > g.n[1] - 2))) || (w.z0 < 2) || (w.z0 + C::TZ > g.n[2] - 2);
  w.lean = full_tile && !edge;
  return w;
}

__device__ __forceinline__ int work_items(const TilePlan& p) { return p.tiles_z * p.tiles_y * p.nz; }

// Dynamic claims: the CTAs take the next item id from a global counter (like the hardware CTA scheduler
// did when every item was a CTA) -- a static round-robin gave some CTAs nothing but edge tiles, which run the
// longer general path, and cost 10 % at 512^3.  The producer lane claims, and hands the id to the consumer
// warps through a small shared-memory ring that is published by the `full` barrier of the item's first plane
// (mbarrier arrive = release, wait = acquire); id -1 ends the walk.  Counters come from a per-module pool, one
// entry per launch (round-robin on the host); the last CTA to leave a launch resets its entry.
struct WorkCounter {
  unsigned int next, done;
};
constexpr int kWorkPool = 64;
constexpr int kItemRing = 8;  // > items in flight: the producer is at most S <= 8 planes (>= 3 per item) ahead
static __device__ WorkCounter g_work_pool[kWorkPool];

static inline int next_work_slot() {
  static unsigned int rr = 0;
  return (int)(__atomic_fetch_add(&rr, 1u, __ATOMIC_RELAXED) % kWorkPool);
}

// producer lane: claim the next item; before publishing it, the stage of its first plane must be free
template <typename C>
__device__ __forceinline__ int claim_item(const TilePlan& p, int* item_ring, unsigned seq, uint64_t* full, uint64_t* empty,
                                          unsigned gi) {
  // the first item of every CTA is its block index (no atomic in front of the launch's first TMA load); the
  // counter hands out the ids from gridDim.x on
  const int id = (seq == 0 && p.first_static) ? (int)blockIdx.x
                                              : (int)(gridDim.x * (unsigned)p.first_static + atomicAdd(&g_work_pool[p.work_slot].next, 1u));
  const int s = gi & (C::S - 1);
  if (gi >= (unsigned)C::S) mbar_wait(&empty[s], ((gi / C::S) - 1) & 1);
  const bool over = id >= work_items(p);
  item_ring[seq & (kItemRing - 1)] = over ? -1 : id;
  if (over) mbar_arrive(&full[s]);  // completes the phase without data: the consumers read the sentinel
  return over ? -1 : id;
}

// consumer warps: the item the producer published for sequence number seq (-1: none left)
template <typename C>
__device__ __forceinline__ int take_item(const int* item_ring, unsigned seq, uint64_t* full, unsigned gi) {
  mbar_wait(&full[gi & (C::S - 1)], (gi / C::S) & 1);
  return *(volatile const int*)&item_ring[seq & (kItemRing - 1)];
}

// a kernel that goes on after the sentinel (several phases per launch): the sentinel used one stage of the ring
// like a plane would have, so the consumer warps hand it back like a plane
template <typename C>
__device__ __forceinline__ void consumer_release_sentinel(uint64_t* empty, unsigned gi) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[gi & (C::S - 1)]);
}

// every CTA, once, when it has no work left: the last one re-arms the launch's counter
__device__ __forceinline__ void work_leave(const TilePlan& p) {
  __syncthreads();
  if (threadIdx.x == 0) {
    WorkCounter* w = &g_work_pool[p.work_slot];
    __threadfence();
    if (atomicAdd(&w->done, 1u) == gridDim.x - 1u) {
      w->next = 0u;
      w->done = 0u;
      __threadfence();
    }
  }
}

// Grid of a launch.  Up to three waves of items the launch is PERSISTENT: 2 CTAs per SM (the occupancy of every TMA
// kernel) walk the items with a running pipeline -- measured +4 % on CG at 256^3, where the fill / drain of one CTA
// per item is a visible share of a short kernel.  Beyond that one CTA per item, handed out by the hardware scheduler
// as before (the CTA takes item blockIdx.x, the counter then answers "none left"): at 512^3 (6.9 waves) 296
// resident CTAs that start together march through the planes in lockstep and phase B measured 0.967 ms in that
// state against 0.857 ms with the scheduler's staggered starts -- the same 13 % a static round-robin lost.
inline int persistent_grid(const TilePlan& p, int nz) {
  const long long items = (long long)p.tiles_z * p.tiles_y * nz;
  const long long slots = (long long)kNumSMs * 2;
  if (items > 3 * slots) return (int)items;
  return (int)(items < slots ? (items > 0 ? items : 1) : slots);
}

// consumer warps only (the producer warp runs ahead in its own loop): named barrier 1
template <int CWARPS>
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(CWARPS * 32) : "memory");
}

// ---- consumer-side geometry -----------------------------------------------------------------------
template <typename T, typename K>
struct ConsCtx {
  int lane, warp;
  int hoff;   // element offset of (own row 0, own element 0) inside a halo tile
  int ooff;   // same inside an own tile
  long long goff;  // element offset of (own row 0, element 0) inside a global plane
  unsigned valid, inreg, nonshell;  // bit k*VEC+e   (GENERAL path)
  int cly[K::RY], clz[VecOf<T>::N];
  int yb, zg;  // global (axis 1, axis 2) index of the thread's first cell
};

template <typename T, typename K>
__device__ __forceinline__ void cons_setup(const GridDev& g, ConsCtx<T, K>& c, int y0, int z0) {
  typedef TmaCfg<T, K> C;
  c.lane = threadIdx.x & 31;
  c.warp = threadIdx.x >> 5;
  const int wy = C::FLAT ? 0 : c.warp, wz = C::FLAT ? c.warp : 0;
  const int col = (wz * 32 + c.lane) * C::VEC;
  if (C::FLAT) {  // one box per warp
    c.hoff = wz * (C::HBOX_SLOT / (int)sizeof(T)) + C::HZ + c.lane * C::VEC;
    c.ooff = wz * (C::OBOX_SLOT / (int)sizeof(T)) + c.lane * C::VEC;
  } else {
    c.hoff = (wy * K::RY + 1) * C::BOXZ + C::HZ + col;
    c.ooff = (wy * K::RY) * C::OBOXZ + col;
  }
  const int yb = y0 + wy * K::RY, zg = z0 + col;
  c.yb = yb;
  c.zg = zg;
  c.goff = (long long)yb * g.n[2] + zg;
  c.valid = c.inreg = c.nonshell = 0u;
#pragma unroll
  for (int k = 0; k < K::RY; ++k) {
    const int y = yb + k;
    c.cly[k] = (!C::FLAT && y < g.n[1]) ? coef_class(g, 1, y) : 0;
#pragma unroll
    for (int e = 0; e < C::VEC; ++e) {
      const int z = zg + e;
      const bool v = (y < g.n[1]) && (z < g.n[2]);
      const bool rg = v && y >= g.lo[1] && y < g.hi[1] && z >= g.lo[2] && z < g.hi[2];
      const bool ns = v && (C::FLAT || (y != 0 && y != g.n[1] - 1)) && z != 0 && z != g.n[2] - 1;
      const unsigned bit = 1u << (k * C::VEC + e);
      if (v) c.valid |= bit;
      if (rg) c.inreg |= bit;
      if (ns) c.nonshell |= bit;
    }
  }
#pragma unroll
  for (int e = 0; e < C::VEC; ++e) c.clz[e] = (zg + e < g.n[2]) ? coef_class(g, 2, zg + e) : 0;
}

// Per-step partial sums.  fp64: added straight into the double accumulator (unchanged order).
// fp32: the products of one plane-step (<= 8 cells) are first summed in float and converted once --
// the F2F conversions (quarter rate) were 2 per cell in phase B; torch's own fp32 sums accumulate in
// float as well, and reductions are not part of the bit-exact contract.
template <typename T>
struct StepSum;
template <>
struct StepSum<double> {
  double& acc;
  __device__ __forceinline__ explicit StepSum(double& a) : acc(a) {}
  __device__ __forceinline__ void add(double q) { acc += q; }
  __device__ __forceinline__ void fma_add(double a, double b) { acc = fma(a, b, acc); }
  __device__ __forceinline__ void flush() {}
};
template <>
struct StepSum<float> {
  double& acc;
  float part = 0.f;
  __device__ __forceinline__ explicit StepSum(double& a) : acc(a) {}
  __device__ __forceinline__ void add(float q) { part += q; }
  __device__ __forceinline__ void fma_add(float a, float b) { part = fmaf(a, b, part); }
  __device__ __forceinline__ void flush() { acc += (double)part; }
};

template <typename T>
__device__ __forceinline__ void lds_vec(const T* p, T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  V q = *reinterpret_cast<const V*>(p);
  const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
  for (int e = 0; e < VecOf<T>::N; ++e) v[e] = s[e];
}

// STREAM: write-once output that nobody re-reads soon (the explicit operators): st.global.cs, so that the
// output lines are the first to leave L2 and the input's halo rows survive until the neighbouring tile reads
// them (ncu, 512^3 Grad with plain stores: 1.48 GB read for 1.07 GB of input)
template <typename T, typename K, bool LEAN, bool STREAM = false>
__device__ __forceinline__ void stg_row(T* p, const ConsCtx<T, K>& c, int k, const T (&v)[VecOf<T>::N]) {
  typedef typename VecOf<T>::type V;
  constexpr int N = VecOf<T>::N;
  const unsigned m = (c.valid >> (k * N)) & ((1u << N) - 1u);
  if (LEAN || m == ((1u << N) - 1u)) {
    V q;
    T* s = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int e = 0; e < N; ++e) s[e] = v[e];
    if (STREAM)
      __stcs(reinterpret_cast<V*>(p), q);
    else
      *reinterpret_cast<V*>(p) = q;
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e)
      if ((m >> e) & 1u) p[e] = v[e];
  }
}

// Wrap-around halos of plane x for periodic kernel axes 1/2 (boundary tiles only; tma_wrap_ok
// guarantees whole tiles).  `val(i)` returns the stencilled quantity at linear index i from global
// memory, computed exactly like the staged one.  A row/column is wrapped when the first/last index of
// the axis is inside the solver region (mesh/tools.py:7-20 opens the slicer on periodic sides).
// ALL: every side wraps, whatever the region -- the explicit operator application, which is defined on
// every cell with torch.roll's wrap-around (fdc.py:171-200); partial tiles allowed (the row below the
// last VALID row is handled by the caller, see pw_consumer).
template <typename T, typename K, bool ALL = false, typename F>
__device__ __forceinline__ void wrap_halo(const GridDev& g, const ConsCtx<T, K>& c, int x, T (&up)[VecOf<T>::N],
                                          T (&dn)[VecOf<T>::N], T (&zl)[K::RY], T (&zr)[K::RY], F val) {
  constexpr int VEC = VecOf<T>::N;
  const long long base = (long long)x * g.n[1] * g.n[2];
  if (ALL && c.zg >= g.n[2]) return;  // a column of a partial tile beyond the array
  if (!K::FLAT) {
    if (c.yb == 0 && (ALL || g.lo[1] == 0)) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) up[e] = val(base + (long long)(g.n[1] - 1) * g.n[2] + c.zg + e);
    }
    if (c.yb + K::RY == g.n[1] && (ALL || g.hi[1] == g.n[1])) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) dn[e] = val(base + c.zg + e);
    }
  }
  if (c.zg == 0 && (ALL || g.lo[2] == 0)) {
#pragma unroll
    for (int k = 0; k < K::RY; ++k)
      if (!ALL || c.yb + k < g.n[1]) zl[k] = val(base + (long long)(c.yb + k) * g.n[2] + g.n[2] - 1);
  }
  if (c.zg + VEC == g.n[2] && (ALL || g.hi[2] == g.n[2])) {
#pragma unroll
    for (int k = 0; k < K::RY; ++k)
      if (!ALL || c.yb + k < g.n[1]) zr[k] = val(base + (long long)(c.yb + k) * g.n[2]);
  }
}

// a*b + c: two roundings (the reference's operation sequence; the library is built with -fmad=false) or, in the
// OPT-IN contraction mode (pa_solver_cfg.flags & PA_FLAG_CONTRACT), one fused multiply-add.  Contraction changes the
// last bits of every stencil value (relative 1e-16 per operation, far inside north_star's 1e-12 per operator) and
// nearly halves the fp64 instructions of the fused CG kernels -- what a power-capped step is short of.
template <bool CONTRACT, typename T>
__device__ __forceinline__ T mad(T a, T b, T c) {
  if (CONTRACT) return fma(a, b, c);
  const T m = a * b;
  return m + c;
}

// the star operator on the thread's cells: same bits as eval_equation, with two exact shortcuts --
// the accumulator starts from the first axis' sum instead of 0 + sum (they differ only in the sign
// of an all-zero sum, erased by the `0 + acc` below), and (acc * param) * sign is one
// multiplication by param*sign (sign = +-1; rounding is symmetric), skipped when it is exactly 1.
// Kernel axis 0 is always active on this path (plan_tma).
// UNI: the three coefficient classes of axes 1 and 2 hold the same numbers (no Neumann / Symmetry face on
// those axes), so the general path takes class 0 like the LEAN one -- same bits, no per-cell selects.
template <typename T, typename K, bool LEAN, typename F, bool UNI = false, bool CONTRACT = false>
__device__ __forceinline__ void star_cells(const OpDev<T>& o, const ConsCtx<T, K>& c, const T (&cx)[3],
                                           bool actx, const T (&vm)[K::RY][VecOf<T>::N],
                                           const T (&vc)[K::RY][VecOf<T>::N], const T (&vp)[K::RY][VecOf<T>::N],
                                           const T (&up)[VecOf<T>::N], const T (&dn)[VecOf<T>::N],
                                           const T (&zl)[K::RY], const T (&zr)[K::RY], F emit) {
  constexpr int VEC = VecOf<T>::N;
  const T scale = o.has_param ? o.param * o.sign : o.sign;
  const bool use_scale = scale != (T)1;
#pragma unroll
  for (int k = 0; k < K::RY; ++k) {
    const int cy = (LEAN || UNI) ? 0 : c.cly[k];
    const T yap = o.coef[1][cy][0], yac = o.coef[1][cy][1], yam = o.coef[1][cy][2];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int cz = (LEAN || UNI) ? 0 : c.clz[e];
      const T zap = o.coef[2][cz][0], zac = o.coef[2][cz][1], zam = o.coef[2][cz][2];
      const T v0 = vc[k][e];
      const T yp = (k == K::RY - 1) ? dn[e] : vc[k + 1 < K::RY ? k + 1 : k][e];
      const T ym = (k == 0) ? up[e] : vc[k > 0 ? k - 1 : 0][e];
      const T zp = (e == VEC - 1) ? zr[k] : vc[k][e + 1 < VEC ? e + 1 : e];
      const T zm = (e == 0) ? zl[k] : vc[k][e > 0 ? e - 1 : 0];
      T acc;
      {
        T s = cx[0] * vp[k][e];
        s = mad<CONTRACT>(cx[1], v0, s);
        s = mad<CONTRACT>(cx[2], vm[k][e], s);
        acc = s;
      }
      if (!K::FLAT) {
        T s = yap * yp;
        s = mad<CONTRACT>(yac, v0, s);
        s = mad<CONTRACT>(yam, ym, s);
        acc = acc + s;
      }
      {
        T s = zap * zp;
        s = mad<CONTRACT>(zac, v0, s);
        s = mad<CONTRACT>(zam, zm, s);
        acc = acc + s;
      }
      if (use_scale) acc = acc * scale;
      T res = CONTRACT ? acc : (T)0 + acc;
      if (o.has_shift) res = mad<CONTRACT>(o.shift, v0, res);  // + the diagonal operator, last (ops.py:151-152)
      emit(k, e, res);
    }
  }
}

// =========================================================================================
// phase B
// =========================================================================================
// WRAP: the launch has periodic faces on kernel axes 1/2 (TilePlan::wrap).  A template parameter of the
// KERNEL, so that the edge tiles of non-periodic problems (30 % of the tiles at 512^2 planes) do not carry
// the wrapped-halo loads, their branches and the registers they pin: as a run-time branch they cost the
// general path 22 % more instructions and 9 spill reloads per plane, and phase B 7 % at 512^3.
// gi0: pipeline counter of the item's first plane (x0-1) -- the CTA's mbarrier ring keeps running across items;
// STRIDE: bytes between stages (a kernel that alternates phases uses the larger of the two layouts)
// HALO: the launch stores its first / last owned plane of r into the neighbours' landing zones as well (slabs with
// the peer-memory halo exchange); a template parameter so that single-GPU launches do not carry those stores
template <typename T, typename K, bool LEAN, bool WRAP = false, bool UNI = false, int STRIDE = TmaCfg<T, K>::STAGE_B,
          bool HALO = false, bool CONTRACT = false>
__device__ __forceinline__ void tmaB_consumer(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                              T* __restrict__ x_new, T* __restrict__ r, T alpha,
                                              unsigned char* stages, uint64_t* full, uint64_t* empty,
                                              int y0, int z0, int x0, int x1, double (&acc_out)[2], unsigned gi0) {
  typedef TmaCfg<T, K> C;
  constexpr int VEC = C::VEC;
  ConsCtx<T, K> c;
  cons_setup<T, K>(g, c, y0, z0);
  constexpr bool actx = true;  // kernel axis 0 is active on every TMA path (plan_tma)
  const long long n12 = (long long)g.n[1] * g.n[2];
  T* xo = x_new + (long long)x0 * n12 + c.goff;
  T* ro = r + (long long)x0 * n12 + c.goff;

  auto halo = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * STRIDE); };
  auto ownx = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * STRIDE + C::HALO_SLOT); };
  auto ownr = [&](int s) {
    return reinterpret_cast<const T*>(stages + (size_t)s * STRIDE + C::HALO_SLOT + C::OWN_SLOT);
  };
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
  double a0 = 0.0, a1 = 0.0;

  // prologue: plane x0-1 (counter gi0) -> A ; plane x0 (counter gi0 + 1) -> B
  {
    const unsigned s0 = gi0 & (C::S - 1), s1 = (gi0 + 1) & (C::S - 1);
    mbar_wait(&full[s0], (gi0 / C::S) & 1);
    load_own(s0, A);
    release(s0);
    mbar_wait(&full[s1], ((gi0 + 1) / C::S) & 1);
    load_own(s1, B);
  }

  auto step = [&](T (&vm)[K::RY][VEC], T (&vc)[K::RY][VEC], T (&vp)[K::RY][VEC], int x, unsigned i) {
    // i = pipeline counter of plane x+1; plane x sits in stage (i-1)%S
    const int sn = i & (C::S - 1), sc = (i - 1) & (C::S - 1);
    if (actx) {
      mbar_wait(&full[sn], (i / C::S) & 1);
      load_own(sn, vp);
    }
    const T* h = halo(sc) + c.hoff;
    const bool xreg = x >= g.lo[0] && x < g.hi[0];
    const bool xown = x >= g.olo0 && x < g.ohi0;
    const int gx = x + g.goff0;
    const bool xshell = actx && (gx == 0 || gx == g.gn0 - 1);
    StepSum<T> s0(a0), s1(a1);
    if (xreg) {
      T ad[K::RY][VEC];
      {
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
        if (WRAP) {
          const T* dg = static_cast<const T*>(p.src0);
          wrap_halo<T, K>(g, c, x, up, dn, zl, zr, [&](long long i) { return dg[i]; });
        }
        const int clx = actx ? coef_class(g, 0, x) : 0;
        const T cx[3] = {o.coef[0][clx][0], o.coef[0][clx][1], o.coef[0][clx][2]};
        auto keep = [&](int k, int e, T v) { ad[k][e] = v; };
        star_cells<T, K, LEAN, decltype(keep), UNI, CONTRACT>(o, c, cx, actx, vm, vc, vp, up, dn, zl, zr, keep);
      }
      // x and r of this plane are only needed now: keep their live range short
#pragma unroll
      for (int k = 0; k < K::RY; ++k) {
        T xv[VEC], rv[VEC], xn[VEC], rn[VEC];
        lds_vec<T>(ownx(sc) + c.ooff + k * C::OBOXZ, xv);
        lds_vec<T>(ownr(sc) + c.ooff + k * C::OBOXZ, rv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const bool in = LEAN || ((c.inreg >> (k * VEC + e)) & 1u);
          xn[e] = mad<CONTRACT>(alpha, vc[k][e], xv[e]);  // linalg.py:122 (d == 0 outside the region)
          const T t = CONTRACT ? fma(-alpha, ad[k][e], rv[e]) : rv[e] - alpha * ad[k][e];  // linalg.py:131
          rn[e] = in ? t : rv[e];
          if (xown) {
            if (in) {
              if (CONTRACT) s0.fma_add(t, t);
              else {
                const T q = t * t;
                s0.add(q);
              }
            }
            if (!xshell && (LEAN || ((c.nonshell >> (k * VEC + e)) & 1u))) {
              const T df = xn[e] - xv[e];
              if (CONTRACT) s1.fma_add(df, df);
              else {
                const T q2 = df * df;
                s1.add(q2);
              }
            }
          }
        }
        stg_row<T, K, LEAN>(xo + (long long)k * g.n[2], c, k, xn);
        stg_row<T, K, LEAN>(ro + (long long)k * g.n[2], c, k, rn);
        if (HALO) {  // first / last owned plane: the same row also goes to the neighbour's landing zone
          const long long so = p.halo.slot * p.halo.slot_bytes;
          if (x == g.olo0 && p.halo.dst[0] != nullptr)
            stg_row<T, K, LEAN>(reinterpret_cast<T*>(static_cast<char*>(p.halo.dst[0]) + so) + c.goff +
                                    (long long)k * g.n[2], c, k, rn);
          if (x == g.ohi0 - 1 && p.halo.dst[1] != nullptr)
            stg_row<T, K, LEAN>(reinterpret_cast<T*>(static_cast<char*>(p.halo.dst[1]) + so) + c.goff +
                                    (long long)k * g.n[2], c, k, rn);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < K::RY; ++k) {
        T xv[VEC];
        lds_vec<T>(ownx(sc) + c.ooff + k * C::OBOXZ, xv);
        stg_row<T, K, LEAN>(xo + (long long)k * g.n[2], c, k, xv);
      }
    }
    s0.flush();
    s1.flush();
    xo += n12;
    ro += n12;
    release(sc);
  };

  int x = x0;
  unsigned i = gi0 + 2;
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
  release(i & (C::S - 1));  // the plane above the chunk: its own cells were read, the ring moves on
  acc_out[0] += a0;
  acc_out[1] += a1;
}

// ---- phase B over this CTA's work items: producer lane / consumer warps ------------------------------
template <typename T, typename K, int STRIDE>
__device__ __forceinline__ unsigned phaseB_produce(const CUtensorMap* tm_d, const CUtensorMap* tm_x,
                                                   const CUtensorMap* tm_r, const TilePlan& p, const GridDev& g,
                                                   unsigned char* stages, uint64_t* full, uint64_t* empty,
                                                   int* item_ring, unsigned gi) {
  typedef TmaCfg<T, K> C;
  for (unsigned seq = 0;; ++seq) {
    const int id = claim_item<C>(p, item_ring, seq, full, empty, gi);
    if (id < 0) break;
    const WorkItem w = work_item<C>(p, g, id);
    const int n = w.x1 - w.x0 + 2;  // planes x0-1 .. x1
    for (int k = 0; k < n; ++k, ++gi) {
      const int pl = w.x0 - 1 + k;
      const int s = gi & (C::S - 1);
      if (k > 0 && gi >= (unsigned)C::S) mbar_wait(&empty[s], ((gi / C::S) - 1) & 1);  // (k == 0: claim_item waited)
      const bool inner = (pl >= w.x0 && pl < w.x1);
      mbar_expect_tx(&full[s], (uint32_t)(C::HALO_BYTES + (inner ? 2 * C::OWN_BYTES : 0)));
      unsigned char* sb = stages + (size_t)s * STRIDE;
      const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
      for (int b = 0; b < C::NB; ++b) {
        const int zb = w.z0 + b * C::OBOXZ;
        tma_load_3d(sb + b * C::HBOX_SLOT, tm_d, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
        if (inner) {
          tma_load_3d(sb + C::HALO_SLOT + b * C::OBOX_SLOT, tm_x, zb, w.y0, xw, &full[s]);
          tma_load_3d(sb + C::HALO_SLOT + C::OWN_SLOT + b * C::OBOX_SLOT, tm_r, zb, w.y0, xw, &full[s]);
        }
      }
    }
  }
  return gi;
}

template <typename T, typename K, bool WRAP, bool UNI, int STRIDE, bool HALO = false, bool CONTRACT = false>
__device__ __forceinline__ unsigned phaseB_consume(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                                   T* __restrict__ x_new, T* __restrict__ r, T alpha, SolverState* st,
                                                   unsigned long long halo_seq, unsigned char* stages, uint64_t* full,
                                                   uint64_t* empty, const int* item_ring, double (&acc)[2],
                                                   unsigned gi) {
  typedef TmaCfg<T, K> C;
  for (unsigned seq = 0;; ++seq) {
    const int id = take_item<C>(item_ring, seq, full, gi);
    if (id < 0) break;
    const WorkItem w = work_item<C>(p, g, id);
    if (w.lean)
      tmaB_consumer<T, K, true, false, false, STRIDE, HALO, CONTRACT>(p, g, o, x_new, r, alpha, stages, full, empty,
                                                                      w.y0, w.z0, w.x0, w.x1, acc, gi);
    else
      tmaB_consumer<T, K, false, WRAP, UNI, STRIDE, HALO, CONTRACT>(p, g, o, x_new, r, alpha, stages, full, empty, w.y0,
                                                                    w.z0, w.x0, w.x1, acc, gi);
    gi += (unsigned)(w.x1 - w.x0 + 2);
    if (HALO) {
      // this item's rows of the first / last owned plane are in the neighbour's landing zone: count it in; the
      // last item of a plane publishes the sequence number in the neighbour's flag word (release at system
      // scope: every thread fences its remote stores before the count, the publisher fences again before the flag)
      const bool has_lo = p.halo.dst[0] != nullptr && w.x0 <= g.olo0 && g.olo0 < w.x1;
      const bool has_hi = p.halo.dst[1] != nullptr && w.x0 <= g.ohi0 - 1 && g.ohi0 - 1 < w.x1;
      if (has_lo || has_hi) {
        __threadfence_system();
        consumer_sync<C::CWARPS>();
        if (threadIdx.x == 0) {
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? has_lo : has_hi)) continue;
            const unsigned int t = atomicAdd(&st->halo_cnt[side], 1u);
            if (t == (unsigned int)p.halo.tiles - 1u) {
              atomicExch(&st->halo_cnt[side], 0u);
              __threadfence_system();
              *(volatile unsigned long long*)p.halo.flag_dst[side] = halo_seq;
            }
          }
        }
      }
    }
    if (p.signal_halo && id / (p.tiles_z * p.tiles_y) < p.b_lo + p.b_hi) {
      // NCCL exchange overlapped with the interior chunks: a boundary item is done (k_wait_halo counts them)
      __threadfence();
      consumer_sync<C::CWARPS>();
      if (threadIdx.x == 0) atomicAdd(&st->halo_count, 1u);
    }
  }
  return gi;
}

template <typename C>
__device__ __forceinline__ void pipe_init(uint64_t* full, uint64_t* empty) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C::CWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
}

template <typename T, typename K, bool WRAP, bool UNI, bool HALO = false, bool CONTRACT = false>
__global__ void __launch_bounds__(TmaCfg<T, K>::THREADS, 2)
k_cg_phaseB_tma(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_r, TilePlan p, GridDev g, OpDev<T> o,
                T* __restrict__ x_new, T* __restrict__ r, SolverState* st, double* partials) {
  typedef TmaCfg<T, K> C;
  extern __shared__ unsigned char smem_dyn[];
  if (st->done) return;
  // sequence number this launch publishes with its boundary planes: the all-reduce epoch at launch (it only
  // moves in the last CTA's reduction, after every CTA has taken its ticket)
  const unsigned long long halo_seq = HALO ? *(volatile unsigned long long*)&st->epoch : 0ull;
  // 128-byte aligned start, derived by pointer arithmetic so the compiler keeps the shared
  // address space (LDS instead of generic LD)
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  unsigned char* stages = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)C::S * C::STAGE_B);
  uint64_t* empty = full + C::S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int item_ring[kItemRing];
  pipe_init<C>(full, empty);
  double acc[2] = {0.0, 0.0};
  if (warp == C::CWARPS) {
    if (lane == 0) phaseB_produce<T, K, C::STAGE_B>(&tm_d, &tm_x, &tm_r, p, g, stages, full, empty, item_ring, 0u);
  } else {
    const T alpha = (T)st->scal[S_ALPHA];
    phaseB_consume<T, K, WRAP, UNI, C::STAGE_B, HALO, CONTRACT>(p, g, o, x_new, r, alpha, st, halo_seq, stages, full,
                                                                empty, item_ring, acc, 0u);
  }
  work_leave(p);
  const int nblocks = gridDim.x;
  const int bid = blockIdx.x;
  if (p.fuse_fin) {
    // static shell: this launch also finalizes the iteration; on slabs the two sums first go
    // through the peer-memory all-reduce (p.p2p), single GPU: p.p2p.peers == nullptr
    if (threadIdx.x == 0) st->sum[R_SHELL] = 0.0;
    P2PDev pp = p.p2p;
    pp.slot0 = R_A;
    pp.count = 2;
    grid_reduce<2>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 2>{st, R_A, ST_CG_FIN, p.accum, pp});
  } else {
    grid_reduce<2>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 2>{st, R_A, ST_NONE, p.accum});
  }
}

// Producer-lane wait for a neighbour's boundary plane: the flag word in this rank's landing zone reaches
// `need`.  Acquire at system scope, then a proxy fence: the plane was written through the generic proxy (by
// another GPU) and is about to be read through the async proxy (TMA).  Same 60 s watchdog as p2p_allreduce.
__device__ __forceinline__ bool halo_wait(const unsigned long long* flag, unsigned long long need) {
  unsigned int spins = 0;
  unsigned long long t0 = 0ull;
  bool ok = true;
  while (true) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= need) break;
    if ((++spins & 0x3ffu) == 0u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0ull) t0 = now;
      else if (now - t0 > kP2PTimeoutNs) {
        ok = false;
        break;
      }
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  return ok;
}

// =========================================================================================
// phase A
// =========================================================================================
template <typename T, typename K, bool LEAN, bool WRAP = false, bool UNI = false, int STRIDE = TmaCfg<T, K>::STAGE_A,
          bool CONTRACT = false>
__device__ __forceinline__ void tmaA_consumer(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                              T* __restrict__ d_new, T beta, unsigned char* stages,
                                              uint64_t* full, uint64_t* empty, int y0, int z0, int x0,
                                              int x1, double& acc_out, unsigned gi0) {
  typedef TmaCfg<T, K> C;
  constexpr int VEC = C::VEC;
  ConsCtx<T, K> c;
  cons_setup<T, K>(g, c, y0, z0);
  constexpr bool actx = true;  // kernel axis 0 is active on every TMA path (plan_tma)
  const long long n12 = (long long)g.n[1] * g.n[2];
  T* dout = d_new + (long long)x0 * n12 + c.goff;

  auto rt = [&](int s) { return reinterpret_cast<const T*>(stages + (size_t)s * STRIDE) + c.hoff; };
  auto dt = [&](int s) {
    return reinterpret_cast<const T*>(stages + (size_t)s * STRIDE + C::HALO_SLOT) + c.hoff;
  };
  // d_new = r + beta*d on the thread's own cells of the plane in stage s   (linalg.py:141)
  auto own_dn = [&](int s, T (&v)[K::RY][VEC]) {
    const T* rp = rt(s);
    const T* dp = dt(s);
#pragma unroll
    for (int k = 0; k < K::RY; ++k) {
      T a[VEC], b[VEC];
      lds_vec<T>(rp + k * C::BOXZ, a);
      lds_vec<T>(dp + k * C::BOXZ, b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[k][e] = mad<CONTRACT>(beta, b[e], a[e]);
    }
  };
  auto write_d = [&](const T (&v)[K::RY][VEC]) {
#pragma unroll
    for (int k = 0; k < K::RY; ++k) stg_row<T, K, LEAN>(dout + (long long)k * g.n[2], c, k, v[k]);
    dout += n12;
  };
  auto release = [&](int s) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&empty[s]);
  };

  T A[K::RY][VEC], B[K::RY][VEC], Cc[K::RY][VEC];
  double acc = 0.0;

  {
    const unsigned s0 = gi0 & (C::S - 1), s1 = (gi0 + 1) & (C::S - 1);
    mbar_wait(&full[s0], (gi0 / C::S) & 1);
    own_dn(s0, A);
    release(s0);
    mbar_wait(&full[s1], ((gi0 + 1) / C::S) & 1);
    own_dn(s1, B);
    write_d(B);
  }

  auto step = [&](T (&vm)[K::RY][VEC], T (&vc)[K::RY][VEC], T (&vp)[K::RY][VEC], int x, unsigned i) {
    const int sn = i & (C::S - 1), sc = (i - 1) & (C::S - 1);
    if (actx) {
      mbar_wait(&full[sn], (i / C::S) & 1);
      own_dn(sn, vp);
      if (x + 1 < x1) write_d(vp);
    }
    const bool xin = x >= g.lo[0] && x < g.hi[0] && x >= g.olo0 && x < g.ohi0;
    StepSum<T> sd(acc);
    if (xin) {
      // halo neighbours of the centre plane: recomputed from the staged raw r, d
      const T* rp = rt(sc);
      const T* dp = dt(sc);
      T up[VEC], dn[VEC], zl[K::RY], zr[K::RY];
      if (!K::FLAT) {
        T a[VEC], b[VEC];
        lds_vec<T>(rp - C::BOXZ, a);
        lds_vec<T>(dp - C::BOXZ, b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) up[e] = mad<CONTRACT>(beta, b[e], a[e]);
        lds_vec<T>(rp + K::RY * C::BOXZ, a);
        lds_vec<T>(dp + K::RY * C::BOXZ, b);
#pragma unroll
        for (int e = 0; e < VEC; ++e) dn[e] = mad<CONTRACT>(beta, b[e], a[e]);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) up[e] = dn[e] = (T)0;
      }
#pragma unroll
      for (int k = 0; k < K::RY; ++k) {
        zl[k] = mad<CONTRACT>(beta, dp[k * C::BOXZ - 1], rp[k * C::BOXZ - 1]);
        zr[k] = mad<CONTRACT>(beta, dp[k * C::BOXZ + VEC], rp[k * C::BOXZ + VEC]);
      }
      if (WRAP) {
        const T* rg = static_cast<const T*>(p.src0);
        const T* dg = static_cast<const T*>(p.src1);
        wrap_halo<T, K>(g, c, x, up, dn, zl, zr, [&](long long i) { return rg[i] + beta * dg[i]; });
      }
      const int clx = actx ? coef_class(g, 0, x) : 0;
      const T cx[3] = {o.coef[0][clx][0], o.coef[0][clx][1], o.coef[0][clx][2]};
      // d == 0 outside the solver region: d*Ad needs no region mask, only array bounds
      auto dot = [&](int k, int e, T ad) {
        if (LEAN || ((c.valid >> (k * VEC + e)) & 1u)) {
          if (CONTRACT) sd.fma_add(vc[k][e], ad);
          else {
            const T q = vc[k][e] * ad;
            sd.add(q);
          }
        }
      };
      star_cells<T, K, LEAN, decltype(dot), UNI, CONTRACT>(o, c, cx, actx, vm, vc, vp, up, dn, zl, zr, dot);
    }
    sd.flush();
    release(sc);
  };

  int x = x0;
  unsigned i = gi0 + 2;
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
  release(i & (C::S - 1));  // the plane above the chunk
  acc_out += acc;
}

// ---- phase A over this CTA's work items ------------------------------------------------------------
template <typename T, typename K, int STRIDE>
__device__ __forceinline__ unsigned phaseA_produce(const CUtensorMap* tm_r, const CUtensorMap* tm_d,
                                                   const CUtensorMap* tm_land, const TilePlan& p, const GridDev& g,
                                                   SolverState* st, unsigned char* stages, uint64_t* full,
                                                   uint64_t* empty, int* item_ring, unsigned gi) {
  typedef TmaCfg<T, K> C;
  // peer-memory halo exchange: the ghost planes of r live in this rank's landing zone (tm_land), written
  // by the neighbours' previous phase B; wait for their flag (sequence number = epoch at launch - 1)
  const bool halo_on = p.halo.on != 0;
  const int gl = (halo_on && g.olo0 > 0) ? g.olo0 - 1 : -2;
  const int gu = (halo_on && g.ohi0 < g.n[0]) ? g.ohi0 : -2;
  const unsigned long long need = halo_on ? *(volatile unsigned long long*)&st->epoch - 1ull : 0ull;
  for (unsigned seq = 0;; ++seq) {
    const int id = claim_item<C>(p, item_ring, seq, full, empty, gi);
    if (id < 0) break;
    const WorkItem w = work_item<C>(p, g, id);
    const int n = w.x1 - w.x0 + 2;
    for (int k = 0; k < n; ++k, ++gi) {
      const int pl = w.x0 - 1 + k;
      const int s = gi & (C::S - 1);
      if (k > 0 && gi >= (unsigned)C::S) mbar_wait(&empty[s], ((gi / C::S) - 1) & 1);  // (k == 0: claim_item waited)
      mbar_expect_tx(&full[s], (uint32_t)(2 * C::HALO_BYTES));
      unsigned char* sb = stages + (size_t)s * STRIDE;
      const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
      const bool ghost = (pl == gl) || (pl == gu);
      if (ghost) {
        const int side = (pl == gl) ? 0 : 1;
        if (!halo_wait(p.halo.flag_src + side, need)) {
          st->done = 1;  // the neighbour never delivered (watchdog): end the solve, the host reports it
          st->status = PA_PEER_LOST;
        }
      }
      for (int b = 0; b < C::NB; ++b) {
        const int zb = w.z0 + b * C::OBOXZ;
        if (ghost)
          tma_load_3d(sb + b * C::HBOX_SLOT, tm_land, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1,
                      p.halo.slot * 2 + ((pl == gl) ? 0 : 1), &full[s]);
        else
          tma_load_3d(sb + b * C::HBOX_SLOT, tm_r, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
        tma_load_3d(sb + C::HALO_SLOT + b * C::HBOX_SLOT, tm_d, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
      }
    }
  }
  return gi;
}

template <typename T, typename K, bool WRAP, bool UNI, int STRIDE, bool CONTRACT = false>
__device__ __forceinline__ unsigned phaseA_consume(const TilePlan& p, const GridDev& g, const OpDev<T>& o,
                                                   T* __restrict__ d_new, T beta, unsigned char* stages,
                                                   uint64_t* full, uint64_t* empty, const int* item_ring, double& acc,
                                                   unsigned gi) {
  typedef TmaCfg<T, K> C;
  for (unsigned seq = 0;; ++seq) {
    const int id = take_item<C>(item_ring, seq, full, gi);
    if (id < 0) break;
    const WorkItem w = work_item<C>(p, g, id);
    if (w.lean)
      tmaA_consumer<T, K, true, false, false, STRIDE, CONTRACT>(p, g, o, d_new, beta, stages, full, empty, w.y0, w.z0,
                                                                w.x0, w.x1, acc, gi);
    else
      tmaA_consumer<T, K, false, WRAP, UNI, STRIDE, CONTRACT>(p, g, o, d_new, beta, stages, full, empty, w.y0, w.z0,
                                                              w.x0, w.x1, acc, gi);
    gi += (unsigned)(w.x1 - w.x0 + 2);
  }
  return gi;
}

template <typename T, typename K, bool WRAP, bool UNI, bool CONTRACT = false>
__global__ void __launch_bounds__(TmaCfg<T, K>::THREADS, 2)
k_cg_phaseA_tma(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_d,
                const __grid_constant__ CUtensorMap tm_land, TilePlan p, GridDev g, OpDev<T> o,
                T* __restrict__ d_new, SolverState* st, double* partials) {
  typedef TmaCfg<T, K> C;
  extern __shared__ unsigned char smem_dyn[];
  if (st->done) return;
  // 128-byte aligned start, derived by pointer arithmetic so the compiler keeps the shared
  // address space (LDS instead of generic LD)
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  unsigned char* stages = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)C::S * C::STAGE_A);
  uint64_t* empty = full + C::S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int item_ring[kItemRing];
  pipe_init<C>(full, empty);
  double acc[1] = {0.0};
  if (warp == C::CWARPS) {
    if (lane == 0)
      phaseA_produce<T, K, C::STAGE_A>(&tm_r, &tm_d, &tm_land, p, g, st, stages, full, empty, item_ring, 0u);
  } else {
    const T beta = (T)st->scal[S_BETA];
    phaseA_consume<T, K, WRAP, UNI, C::STAGE_A, CONTRACT>(p, g, o, d_new, beta, stages, full, empty, item_ring, acc[0],
                                                          0u);
  }
  work_leave(p);
  const int nblocks = gridDim.x;
  const int bid = blockIdx.x;
  {
    P2PDev pp = p.p2p;  // slabs: d.Ad summed over the ranks right here when the peer mailboxes exist
    pp.slot0 = R_A;
    pp.count = 1;
    const int stage = (p.dist && pp.peers == nullptr) ? ST_NONE : ST_CG_DAD;
    grid_reduce<1>(acc, partials, nblocks, bid, &st->ticket[0], StoreSums<T, 1>{st, R_A, stage, 0, pp});
  }
}

// =========================================================================================
// L2-resident grids: the WHOLE CG solve as one cooperative launch of the two phases
// =========================================================================================
// Below a few million cells the six CG vectors stay in the 126 MB L2 and an iteration of the two fused kernels
// is bound by launch + pipeline-fill + reduction latency (1024^2: 16 us per kernel for 3 us of L2 traffic).
// Here the CTAs stay resident for the whole solve: per iteration phase A and phase B run over dynamically
// claimed items with the same producer / consumer code as the stand-alone kernels, the mbarrier ring keeps
// running from phase to phase, and each phase ends in ONE grid-wide step that is reduction, scalar stage and
// barrier at once (last-arriver ticket, then a generation word everybody else spins on).  Preconditions
// (host): single GPU, static shell (every face Dirichlet: no BC work inside the loop), no periodic axis 1/2.
template <int NS, typename Fin>
__device__ __forceinline__ void grid_reduce_sync(double (&v)[NS], double* partials, SolverState* st, unsigned int& gen,
                                                 WorkCounter* work, Fin fin) {
  __shared__ bool last_cta;
  __shared__ double red_smem2[NS * 32];
  block_sum<NS>(v, red_smem2);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) partials[s * kMaxPartials + blockIdx.x] = v[s];
    __threadfence();
    last_cta = atomicAdd(&st->ticket[0], 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  ++gen;
  if (last_cta) {
    __threadfence();
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      acc[s] = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc[s] += __ldcg(&partials[s * kMaxPartials + b]);
    }
    block_sum<NS>(acc, red_smem2);
    if (threadIdx.x == 0) {
      st->ticket[0] = 0u;
      work->next = 0u;  // every CTA has left the phase's item walk
      fin(acc);
      __threadfence();
      *(volatile unsigned int*)&st->ticket[5] = gen;  // release the others
    }
  } else if (threadIdx.x == 0) {
    while (*(volatile unsigned int*)&st->ticket[5] != gen) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T, typename K, bool UNI>
__global__ void __launch_bounds__(TmaCfg<T, K>::THREADS, 2)
k_cg_coop_tma(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1,
              const __grid_constant__ CUtensorMap tm_r_own, const __grid_constant__ CUtensorMap tm_r_halo,
              const __grid_constant__ CUtensorMap tm_d0, const __grid_constant__ CUtensorMap tm_d1, TilePlan p,
              GridDev g, OpDev<T> o, T* __restrict__ xa, T* __restrict__ xb, T* __restrict__ r, T* __restrict__ d0,
              T* __restrict__ d1, SolverState* st, double* partials) {
  typedef TmaCfg<T, K> C;
  constexpr int STRIDE = C::STAGE_A > C::STAGE_B ? C::STAGE_A : C::STAGE_B;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  unsigned char* stages = base;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)C::S * STRIDE);
  uint64_t* empty = full + C::S;
  __shared__ int item_ring[kItemRing];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == C::CWARPS && lane == 0;
  pipe_init<C>(full, empty);
  WorkCounter* work = &g_work_pool[p.work_slot];
  unsigned gi = 0, seq = 0, gen = 0;  // pipeline counter, item sequence number, barrier generation
  int parity = 0;
  while (true) {
    // ---- phase A: d_new = r + beta d_old ; d_new . A(d_new) -> alpha
    {
      const CUtensorMap* tm_d = parity ? &tm_d1 : &tm_d0;
      T* d_new = parity ? d0 : d1;
      double acc[1] = {0.0};
      if (warp == C::CWARPS) {
        if (lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");  // r, d of the previous phase -> TMA reads
          for (;; ++seq) {
            const int id = claim_item<C>(p, item_ring, seq, full, empty, gi);
            if (id < 0) {
              ++seq;
              ++gi;  // the sentinel took one barrier phase
              break;
            }
            const WorkItem w = work_item<C>(p, g, id);
            const int n = w.x1 - w.x0 + 2;
            for (int k = 0; k < n; ++k, ++gi) {
              const int pl = w.x0 - 1 + k;
              const int s = gi & (C::S - 1);
              if (k > 0 && gi >= (unsigned)C::S) mbar_wait(&empty[s], ((gi / C::S) - 1) & 1);
              mbar_expect_tx(&full[s], (uint32_t)(2 * C::HALO_BYTES));
              unsigned char* sb = stages + (size_t)s * STRIDE;
              const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
              for (int b = 0; b < C::NB; ++b) {
                const int zb = w.z0 + b * C::OBOXZ;
                tma_load_3d(sb + b * C::HBOX_SLOT, &tm_r_halo, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
                tma_load_3d(sb + C::HALO_SLOT + b * C::HBOX_SLOT, tm_d, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
              }
            }
          }
        }
      } else {
        const T beta = (T)*(volatile double*)&st->scal[S_BETA];
        for (;; ++seq) {
          const int id = take_item<C>(item_ring, seq, full, gi);
          if (id < 0) {
            ++seq;
            consumer_release_sentinel<C>(empty, gi);
            ++gi;
            break;
          }
          const WorkItem w = work_item<C>(p, g, id);
          if (w.lean)
            tmaA_consumer<T, K, true, false, false, STRIDE>(p, g, o, d_new, beta, stages, full, empty, w.y0, w.z0, w.x0,
                                                            w.x1, acc[0], gi);
          else
            tmaA_consumer<T, K, false, false, UNI, STRIDE>(p, g, o, d_new, beta, stages, full, empty, w.y0, w.z0, w.x0,
                                                           w.x1, acc[0], gi);
          gi += (unsigned)(w.x1 - w.x0 + 2);
        }
      }
      grid_reduce_sync<1>(acc, partials, st, gen, work, StoreSums<T, 1>{st, R_A, ST_CG_DAD, 0});
    }
    // ---- phase B: x_new = x + alpha d ; r -= alpha A(d) ; sums -> beta, tol, itr, done
    {
      const CUtensorMap* tm_d = parity ? &tm_d0 : &tm_d1;  // the d phase A has just written
      const CUtensorMap* tm_x = parity ? &tm_x1 : &tm_x0;
      T* x_new = parity ? xa : xb;
      double acc[2] = {0.0, 0.0};
      if (warp == C::CWARPS) {
        if (lane == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          for (;; ++seq) {
            const int id = claim_item<C>(p, item_ring, seq, full, empty, gi);
            if (id < 0) {
              ++seq;
              ++gi;
              break;
            }
            const WorkItem w = work_item<C>(p, g, id);
            const int n = w.x1 - w.x0 + 2;
            for (int k = 0; k < n; ++k, ++gi) {
              const int pl = w.x0 - 1 + k;
              const int s = gi & (C::S - 1);
              if (k > 0 && gi >= (unsigned)C::S) mbar_wait(&empty[s], ((gi / C::S) - 1) & 1);
              const bool inner = (pl >= w.x0 && pl < w.x1);
              mbar_expect_tx(&full[s], (uint32_t)(C::HALO_BYTES + (inner ? 2 * C::OWN_BYTES : 0)));
              unsigned char* sb = stages + (size_t)s * STRIDE;
              const int xw = pl < 0 ? pl + g.n[0] : (pl >= g.n[0] ? pl - g.n[0] : pl);
              for (int b = 0; b < C::NB; ++b) {
                const int zb = w.z0 + b * C::OBOXZ;
                tma_load_3d(sb + b * C::HBOX_SLOT, tm_d, zb - C::HZ, C::FLAT ? 0 : w.y0 - 1, xw, &full[s]);
                if (inner) {
                  tma_load_3d(sb + C::HALO_SLOT + b * C::OBOX_SLOT, tm_x, zb, w.y0, xw, &full[s]);
                  tma_load_3d(sb + C::HALO_SLOT + C::OWN_SLOT + b * C::OBOX_SLOT, &tm_r_own, zb, w.y0, xw, &full[s]);
                }
              }
            }
          }
        }
      } else {
        const T alpha = (T)*(volatile double*)&st->scal[S_ALPHA];
        for (;; ++seq) {
          const int id = take_item<C>(item_ring, seq, full, gi);
          if (id < 0) {
            ++seq;
            consumer_release_sentinel<C>(empty, gi);
            ++gi;
            break;
          }
          const WorkItem w = work_item<C>(p, g, id);
          if (w.lean)
            tmaB_consumer<T, K, true, false, false, STRIDE>(p, g, o, x_new, r, alpha, stages, full, empty, w.y0, w.z0,
                                                            w.x0, w.x1, acc, gi);
          else
            tmaB_consumer<T, K, false, false, UNI, STRIDE>(p, g, o, x_new, r, alpha, stages, full, empty, w.y0, w.z0,
                                                           w.x0, w.x1, acc, gi);
          gi += (unsigned)(w.x1 - w.x0 + 2);
        }
      }
      if (threadIdx.x == 0 && blockIdx.x == 0) st->sum[R_SHELL] = 0.0;  // static shell
      grid_reduce_sync<2>(acc, partials, st, gen, work, StoreSums<T, 2>{st, R_A, ST_CG_FIN, 0});
    }
    if (*(volatile int*)&st->done) break;
    parity ^= 1;
  }
}

// Communication-stream side of the single-launch overlap: wait until every boundary CTA of the phase B
// that is running on the main stream has written its planes.  One thread; returns at once after `done`.
static __global__ void k_wait_halo(SolverState* st, unsigned int nboundary) {
  if (st->done) return;
  const unsigned int target = st->halo_target + nboundary;
  st->halo_target = target;
  // watchdog like p2p_allreduce's: the count only advances if phase B really runs next to this kernel (CUDA
  // does not promise that for independent graph branches) -- never spin for ever
  unsigned int spins = 0;
  unsigned long long t0 = 0ull;
  while (*(volatile unsigned int*)&st->halo_count < target) {
    if ((++spins & 0xfffu) == 0u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0ull) t0 = now;
      else if (now - t0 > kP2PTimeoutNs) {
        st->done = 1;
        st->status = PA_PEER_LOST;
        return;
      }
    }
  }
  __threadfence();
}

// number of boundary CTAs of a sub == 3 launch
// (work items, now that the kernels are persistent)
inline unsigned int tma_boundary_ctas(const TmaPlan& tp, const GridDev& g) {
  const int c_lo = g.olo0 / tp.tile.cx, c_hi = (g.ohi0 - 1) / tp.tile.cx;
  return (unsigned int)((c_lo + 1 + tp.tile.chunks - c_hi) * tp.tile.tiles_y * tp.tile.tiles_z);
}

// chunks left to the interior sub-launch when phase B is split around the halo exchange
inline int tma_interior_chunks(const TmaPlan& tp, const GridDev& g) {
  return (g.ohi0 - 1) / tp.tile.cx - g.olo0 / tp.tile.cx - 1;
}

// ---- launchers -----------------------------------------------------------------------------------
template <typename T, typename K, bool WRAP, bool UNI, bool CONTRACT = false>
static void launch_cg_phaseA_tma_k(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq,
                                   int parity, T* d_new, SolverState* st, double* partials) {
  typedef TmaCfg<T, K> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseA_tma<T, K, WRAP, UNI, CONTRACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_A);
    attr = true;
  }
  TilePlan tile = tp.tile;
  tile.nz = tile.chunks;
  tile.work_slot = next_work_slot();
  tile.first_static = 1;
  dim3 grid(persistent_grid(tile, tile.nz));
  tile.src0 = tp.r_ptr;
  tile.src1 = tp.d_ptr[parity];
  if (tile.halo.on) {
    tile.halo.slot = 1 - parity;  // what the previous iteration's phase B (parity 1 - p) delivered
    tile.ghost_last = tile.chunks >= 3 ? 1 : 0;
  }
  k_cg_phaseA_tma<T, K, WRAP, UNI, CONTRACT><<<grid, C::THREADS, C::SMEM_A, s>>>(
      tp.r_halo, tp.d_halo[parity], tile.halo.on ? tp.r_land : tp.r_halo, tile, g, eq.op[0], d_new, st, partials);
}

template <typename T>
void launch_cg_phaseA_tma(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq,
                          int parity, T* d_new, SolverState* st, double* partials) {
  // (the uniform-coefficient variant is phase B's only: in phase A it saves 6 % of the general path's
  //  instructions but ptxas gives 4 % back on the LEAN path of the same kernel -- no net gain)
  if (tma_flat(g)) {
    if (tp.tile.wrap)
      launch_cg_phaseA_tma_k<T, KFlat, true, false>(s, tp, g, eq, parity, d_new, st, partials);
    else if (tp.contract)
      launch_cg_phaseA_tma_k<T, KFlat, false, false, true>(s, tp, g, eq, parity, d_new, st, partials);
    else
      launch_cg_phaseA_tma_k<T, KFlat, false, false>(s, tp, g, eq, parity, d_new, st, partials);
  } else {
    if (tp.tile.wrap)
      launch_cg_phaseA_tma_k<T, KStd, true, false>(s, tp, g, eq, parity, d_new, st, partials);
    else if (tp.contract)
      launch_cg_phaseA_tma_k<T, KStd, false, false, true>(s, tp, g, eq, parity, d_new, st, partials);
    else
      launch_cg_phaseA_tma_k<T, KStd, false, false>(s, tp, g, eq, parity, d_new, st, partials);
  }
}

template <typename T, typename K, bool WRAP, bool UNI, bool HALO = false, bool CONTRACT = false>
static void launch_cg_phaseB_tma_k(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq,
                                   int parity, T* x_new, T* r, SolverState* st, double* partials, int sub) {
  typedef TmaCfg<T, K> C;
  static bool attr_dev[kMaxDevices] = {};  // the attribute is per device
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_cg_phaseB_tma<T, K, WRAP, UNI, HALO, CONTRACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_B);
    attr = true;
  }
  // iteration parity p: x_old = x buffer p, d (already updated by phase A) = d buffer 1-p
  TilePlan tile = tp.tile;
  tile.src0 = tp.d_ptr[1 - parity];
  int nz = tile.chunks;
  // the boundary sub-launch must produce the first and the last OWNED plane (they leave with the
  // halo exchange while the interior sub-launch runs): every chunk up to the one holding plane
  // olo0 and from the one holding plane ohi0-1 on
  const int c_lo = g.olo0 / tile.cx, c_hi = (g.ohi0 - 1) / tile.cx;
  if (sub == 1) {
    tile.b_lo = c_lo + 1;
    tile.b_hi = tile.chunks - c_hi;
    nz = tile.b_lo + tile.b_hi;
    tile.fuse_fin = 0;  // the interior sub-launch adds its sums and finalizes
  } else if (sub == 2) {  // interior chunks, sums added to the boundary launch's
    tile.chunk0 = c_lo + 1;
    tile.accum = 1;
    nz = c_hi - c_lo - 1;
  } else if (sub == 3 || sub == 4) {  // ONE launch, boundary chunks first
    tile.b_lo = c_lo + 1;
    tile.b_hi = tile.chunks - c_hi;
    tile.chunk0 = tile.b_lo;
    tile.signal_halo = sub == 3 ? 1 : 0;  // 3: each boundary CTA counts itself in for k_wait_halo (NCCL exchange)
  }
  if (tile.halo.on) tile.halo.slot = parity;
  tile.nz = nz;
  tile.work_slot = next_work_slot();
  tile.first_static = 1;
  dim3 grid(persistent_grid(tile, nz));
  k_cg_phaseB_tma<T, K, WRAP, UNI, HALO, CONTRACT><<<grid, C::THREADS, C::SMEM_B, s>>>(tp.d_halo[1 - parity], tp.x_own[parity], tp.r_own, tile,
                                                           g, eq.op[0], x_new, r, st, partials);
}

template <typename T>
void launch_cg_phaseB_tma(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq,
                          int parity, T* x_new, T* r, SolverState* st, double* partials, int sub = 0) {
  // variants of the general (edge-tile) path: periodic wrap, uniform coefficients (every face of axes 1/2
  // Dirichlet or periodic: 10 % fewer instructions on the edge tiles), neither; and, on slabs with the
  // peer-memory halo exchange (never together with wrap, api.cu halo_dev), the HALO stores
  const bool halo = tp.tile.halo.on != 0;
#define PA_LAUNCH_B(KK)                                                                                            \
  do {                                                                                                             \
    if (tp.tile.wrap)                                                                                              \
      launch_cg_phaseB_tma_k<T, KK, true, false>(s, tp, g, eq, parity, x_new, r, st, partials, sub);               \
    else if (tp.contract && tp.coef_uniform && halo)                                                               \
      launch_cg_phaseB_tma_k<T, KK, false, true, true, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);   \
    else if (tp.contract && tp.coef_uniform)                                                                       \
      launch_cg_phaseB_tma_k<T, KK, false, true, false, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);  \
    else if (tp.contract && halo)                                                                                  \
      launch_cg_phaseB_tma_k<T, KK, false, false, true, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);  \
    else if (tp.contract)                                                                                          \
      launch_cg_phaseB_tma_k<T, KK, false, false, false, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub); \
    else if (tp.coef_uniform && halo)                                                                              \
      launch_cg_phaseB_tma_k<T, KK, false, true, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);         \
    else if (tp.coef_uniform)                                                                                      \
      launch_cg_phaseB_tma_k<T, KK, false, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);               \
    else if (halo)                                                                                                 \
      launch_cg_phaseB_tma_k<T, KK, false, false, true>(s, tp, g, eq, parity, x_new, r, st, partials, sub);        \
    else                                                                                                           \
      launch_cg_phaseB_tma_k<T, KK, false, false>(s, tp, g, eq, parity, x_new, r, st, partials, sub);              \
  } while (0)
  if (tma_flat(g))
    PA_LAUNCH_B(KFlat);
  else
    PA_LAUNCH_B(KStd);
#undef PA_LAUNCH_B
}

// co-resident CTAs of the cooperative CG kernel on this device (0: cooperative launch unavailable)
template <typename T, typename K, bool UNI>
static int cg_coop_max_blocks() {
  typedef TmaCfg<T, K> C;
  constexpr size_t SMEM = (size_t)C::S * (C::STAGE_A > C::STAGE_B ? C::STAGE_A : C::STAGE_B) + C::BAR_BYTES + 128;
  int dev = 0, coop = 0, per_sm = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop) return 0;
  static bool attr_dev[kMaxDevices] = {};
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    cudaFuncSetAttribute(k_cg_coop_tma<T, K, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    attr = true;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_coop_tma<T, K, UNI>, C::THREADS, SMEM) != cudaSuccess)
    return 0;
  return per_sm * sms;
}

template <typename T, typename K, bool UNI>
static bool launch_cg_coop_tma_k(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq, T* xa, T* xb,
                                 T* r, T* d0, T* d1, SolverState* st, double* partials) {
  typedef TmaCfg<T, K> C;
  constexpr size_t SMEM = (size_t)C::S * (C::STAGE_A > C::STAGE_B ? C::STAGE_A : C::STAGE_B) + C::BAR_BYTES + 128;
  const int maxb = cg_coop_max_blocks<T, K, UNI>();
  if (maxb <= 0) return false;
  TilePlan tile = tp.tile;
  tile.nz = tile.chunks;
  tile.work_slot = next_work_slot();
  tile.first_static = 0;  // (the item sequence number runs on over the phases)
  tile.fuse_fin = 1;
  int grid = persistent_grid(tile, tile.nz);
  if (grid > maxb) grid = maxb;
  GridDev gg = g;
  OpDev<T> o = eq.op[0];
  void* args[] = {(void*)&tp.x_own[0], (void*)&tp.x_own[1], (void*)&tp.r_own, (void*)&tp.r_halo, (void*)&tp.d_halo[0],
                  (void*)&tp.d_halo[1], (void*)&tile, (void*)&gg, (void*)&o, (void*)&xa, (void*)&xb, (void*)&r,
                  (void*)&d0, (void*)&d1, (void*)&st, (void*)&partials};
  return cudaLaunchCooperativeKernel((void*)k_cg_coop_tma<T, K, UNI>, dim3(grid), dim3(C::THREADS), args, SMEM, s) ==
         cudaSuccess;
}

// the whole CG solve (after the residual init) as ONE cooperative launch; false: not available, use the loop
template <typename T>
bool launch_cg_coop_tma(cudaStream_t s, const TmaPlan& tp, const GridDev& g, const EqDev<T>& eq, T* xa, T* xb, T* r,
                        T* d0, T* d1, SolverState* st, double* partials) {
  if (tp.tile.wrap) return false;
  if (tma_flat(g))
    return tp.coef_uniform ? launch_cg_coop_tma_k<T, KFlat, true>(s, tp, g, eq, xa, xb, r, d0, d1, st, partials)
                           : launch_cg_coop_tma_k<T, KFlat, false>(s, tp, g, eq, xa, xb, r, d0, d1, st, partials);
  return tp.coef_uniform ? launch_cg_coop_tma_k<T, KStd, true>(s, tp, g, eq, xa, xb, r, d0, d1, st, partials)
                         : launch_cg_coop_tma_k<T, KStd, false>(s, tp, g, eq, xa, xb, r, d0, d1, st, partials);
}

}  // namespace pa
