/*
 * pyapes_b200 — C ABI of the B200-native finite-difference hot path.
 *
 * Drop-in boundary for the three Python seams of the reference (pyapes v0.2.13; file:line
 * are relative to the reference tree):
 *   - operator application  OPStype["Aop"] / ops._Aop        (solver/ops.py:122-154,
 *                                                             solver/fdc.py:67-118,171-200)
 *   - solver                linalg.solve / cg / bicgstab     (solver/linalg.py:33-279)
 *   - BC application        BC.apply / _apply_bc_otf         (variables/bcs.py:197-280,
 *                                                             solver/linalg.py:282-299)
 * plus the two paths BASELINE.json's north_star adds that the reference lacks
 * (Jacobi, explicit Euler Ddt).
 *
 * Conventions
 *   - Plain C: pointers + sizes, no torch types.  All array pointers are DEVICE pointers
 *     unless the function name ends in `_host`.  `stream` is a cudaStream_t passed as void*.
 *   - Fields are contiguous row-major scalar fields (the reference's `(1, *nx)` tensor,
 *     variables/fields.py:52-58).  Kernel coordinates are (n0,n1,n2) with kernel axis 2 the
 *     contiguous one.  3-D meshes map axis j -> kernel axis j; 2-D meshes use kernel axes
 *     (0, 2) (n1 = 1: the kernels march along axis 0); 1-D meshes use kernel axis 2.
 *   - dtype: PA_F64 (reference default) or PA_F32 (backend.py:28-42).
 *   - Every function returns 0 on success or a negative pa_status; pa_last_error() gives
 *     the message.  There is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with PA_ERR_CUDA.
 */
#ifndef PYAPES_B200_H
#define PYAPES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 2
#define PA_MAX_OPS 4
#define PA_MAX_FACES 6

typedef enum { PA_F32 = 0, PA_F64 = 1 } pa_dtype;

typedef enum {
  PA_OK = 0,
  PA_ERR_ARG = -1,     /* bad argument (shape, dtype, null pointer, workspace too small) */
  PA_ERR_CUDA = -2,    /* CUDA runtime error / no device */
  PA_ERR_UNSUPPORTED = -3,
  PA_ERR_NCCL = -4
} pa_status;

/* boundary-condition kinds (variables/bcs.py:197-280) */
typedef enum {
  PA_BC_DIRICHLET = 1,
  PA_BC_NEUMANN = 2,
  PA_BC_SYMMETRY = 3,
  PA_BC_PERIODIC = 4
} pa_bc_kind;

/* One boundary face, in the order the reference applies them (the BC config list,
 * bcs.py:363-440).  `value`:
 *   Dirichlet: the face value (bcs.py:205-211).
 *   Neumann  : the additive term  2/3 * V * (x_face - x_inner) * n_dir  already rounded
 *              in the field dtype by the host (bcs.py:251-253); the kernel computes
 *              4/3*phi[1] - 1/3*phi[2] + value.
 * `values` (optional, device, field dtype): one entry per face cell, row-major over the
 * two other kernel axes — used instead of `value` (callable / tensor bc_val). */
typedef struct {
  int32_t axis; /* kernel axis 0..2 */
  int32_t side; /* -1 lower, +1 upper */
  int32_t kind; /* pa_bc_kind */
  int32_t reserved;
  double value;
  const void* values;
} pa_face_bc;

/* Local grid block.  Single GPU: gn0 == n[0], goff0 == 0.  Slab decomposition along kernel
 * axis 0: n[0] counts the locally allocated planes including ghost planes,
 * goff0 is the global index of local plane 0, gn0 the global extent.
 * [lo,hi) is the reference's boundary_slicer (mesh/tools.py:7-20) in LOCAL indices: the
 * region every solver vector is written in.  [olo0,ohi0) are the locally OWNED planes
 * along axis 0 (norms and dot products are taken over owned cells only). */
typedef struct {
  int32_t n[3];
  int32_t lo[3];
  int32_t hi[3];
  int32_t gn0;
  int32_t goff0;
  int32_t olo0;
  int32_t ohi0;
  int32_t ndim; /* mesh dimension 1..3; active kernel axes: 3 -> 0,1,2; 2 -> 0,2; 1 -> 2 */
  int32_t reserved;
} pa_grid;

/* Operator kinds.  Every constant-coefficient operator of the reference (Laplacian with
 * any BC mix, Grad, Div with a constant advection speed, all limiters) is a 3-point-per-axis
 * star whose coefficients depend only on whether the cell's index along that axis is 1,
 * n-2 or anything else (fdc.py:388-421, 575-609): PA_OP_STAR.  The host computes the three
 * coefficient triples with the reference's own rounding sequence. */
typedef enum {
  PA_OP_STAR = 0,
  PA_OP_DIV_CENTRAL_FIELD = 1, /* limiter "none",  advection from a field (fdc.py:708-743) */
  PA_OP_DIV_UPWIND_FIELD = 2,  /* limiter "upwind", reference formula (fdc.py:746-772) */
  PA_OP_DIV_UPWINDFD_FIELD = 3 /* first-order upwind difference (not in the reference) */
} pa_op_kind;

typedef struct {
  int32_t kind;      /* pa_op_kind */
  int32_t has_param; /* multiply by `param` after the stencil (fdm.py:166-169) */
  double sign;       /* +1 / -1 (fdm.py:95-105, ops.py:140-143) */
  double param;
  /* coef[axis][cls][k]: k = 0:Ap (phi[+1]) 1:Ac 2:Am (phi[-1]);
   * cls = 0 default, 1 index==1 along axis, 2 index==n-2 along axis */
  double coef[3][3][3];
  const void* adv;    /* *_FIELD kinds: advection speed, same shape/dtype as the field */
  double two_dx[3];   /* central field div: divisor 2*dx (fdc.py:607-609) */
  double dx[3];       /* upwind_fd field div */
  int32_t zero_am_lo[3]; /* periodic edits of the central scheme (fdc.py:596-602) */
  int32_t zero_ap_hi[3];
  /* optional per-cell coefficient (Laplacian/Grad called with a Tensor, fdm.py:130,169): the
   * stencil result is multiplied by param_field[cell] instead of `param`.  Same shape/dtype as
   * the field; NULL = use the scalar. */
  const void* param_field;
  /* edge=True of the explicit FDC operators (fdc.py:203-366): one-sided differences on the domain
   * faces.  0 none; 1 Laplacian (the face cell's value is REPLACED by the one-sided second
   * derivative along the face's axis, later axes win); 2 Div on a 1-D mesh (the reference raises
   * IndexError for dim > 1), scaled by adv_const; for pa_grad_apply any non-zero value replaces
   * component a on the faces normal to axis a.  Uses dx[]. */
  int32_t edge;
  /* *_FIELD kinds: 1 = the advection field IS the unknown of the solve (nonlinear advection,
   * fdm.div(var, var), fdm.py:306-312): the solvers read it from the current iterate, which moves
   * every iteration as the reference rebinds var (linalg.py:122-125,253-256); `adv` is then only
   * used by pa_stencil_apply / the initial state.  0 = `adv` is a fixed array. */
  int32_t adv_is_iterate;
  double adv_const;
  /* optional per-index coefficient tables: coef_tab[axis] points to n[axis]*3 device values
   * [Ap, Ac, Am] per index along that kernel axis, used instead of coef[axis][cls][] — the
   * axisymmetric (rz) operators, whose r-axis coefficients carry 1 +- dr/(2r)
   * (tools.py:64-108).  NULL = use the three classes. */
  const void* coef_tab[3];
} pa_op;

/* sum_k sign_k * param_k * Op_k(phi), accumulated in list order (ops.py:130-149) */
typedef struct {
  int32_t nops;
  int32_t reserved;
  pa_op ops[PA_MAX_OPS];
} pa_equation;

/* solver report (linalg.py:22-30) + where the result lives */
typedef enum {
  PA_RUNNING = 0,
  PA_CONVERGED = 1,   /* loop left through its condition */
  PA_MAXIT = 2,       /* "Maximum iteration reached!" RuntimeWarning (linalg.py:146-150,268-271) */
  PA_BAD_TOL = 3,     /* "Invalid tolerance detected!" RuntimeError (linalg.py:334-336) */
  PA_PEER_LOST = 4    /* device-side only (multi-GPU): a peer rank never joined a fused all-reduce within
                         the watchdog time; the entry point returns PA_ERR_NCCL, never this status */
} pa_solve_status;

typedef struct {
  int32_t itr;
  int32_t status;        /* pa_solve_status */
  double tol;
  int32_t result_in_alt; /* 1: the final iterate is in x_alt and the previous one in x */
  int32_t launches;      /* kernels launched by this call */
  int32_t swaps;         /* x updates performed (ping-pong swaps); 0: x_alt was never written */
  int32_t reserved;
} pa_report;

typedef struct {
  double tol;
  int32_t max_it;
  int32_t check_every; /* iterations between host polls of the device-side done flag (0 = default) */
  int32_t use_graph;   /* capture the iteration in a CUDA graph */
  int32_t variant;     /* 0 = auto (small grids: persistent kernel; TMA-staged fused kernels, else
                          register-tiled, else generic), 1 = generic kernels, 2 = register-tiled kernels
                          (no TMA), 3 = CG as one persistent cooperative kernel (small grids, every
                          face Dirichlet), 4 = like 0 but never a whole-solve kernel, 5 = CG as ONE
                          cooperative launch of the two TMA phases (L2-resident grids, every face Dirichlet;
                          auto takes it between 80 k and 1.5 M cells on 3-D meshes), 6 = CG as ONE cooperative
                          launch with x, r, d resident in the SMs' shared memory (2-D meshes up to ~1.2 M fp64
                          cells, every face Dirichlet; auto takes it above 80 k cells whenever it fits) */
  int32_t flags;       /* PA_FLAG_* */
  int32_t reserved;
} pa_solver_cfg;

/* pa_solver_cfg.flags.  PA_FLAG_CONTRACT (opt-in): the fused TMA CG kernels evaluate a*b + c as ONE fused
 * multiply-add instead of the reference's two roundings.  Every stencil value then differs from the reference's in
 * the last bits (relative ~1e-16 per operation; north_star asks 1e-12 per operator), the iteration count of a
 * converged solve may move by one; in exchange the kernels issue ~40 % fewer fp64 instructions, which is what a
 * power-capped step is short of.  Default (0): bit-exact operation order. */
#define PA_FLAG_CONTRACT 1

const char* pa_last_error(void);
int pa_abi_version(void);
int pa_device_count(void);

/* --- operator application: out = sum_k sign*param*Op_k(phi) on EVERY cell, with the
 *     wrap-around semantics of torch.roll (fdc.py:171-200).  Replaces ops._Aop. */
int pa_stencil_apply(const pa_grid* g, const pa_equation* eq, int dtype, const void* phi,
                     void* out, void* stream);

/* --- explicit gradient: out has shape (ndim, n0, n1, n2); op must be PA_OP_STAR.
 *     Replaces fdc.Grad.apply (fdc.py:80-87). */
int pa_grad_apply(const pa_grid* g, const pa_op* op, int dtype, const void* phi, void* out,
                  void* stream);

/* --- boundary conditions, in place, faces in list order.  Replaces _apply_bc_otf. */
int pa_bc_apply(const pa_grid* g, int nfaces, const pa_face_bc* faces, int dtype, void* phi,
                void* stream);

/* --- Krylov / stationary solvers.  x: initial guess in, solution out (or in x_alt, see
 *     pa_report.result_in_alt); x_alt: same size, receives the other of {last, previous}
 *     iterate (the reference's Field.VARo, fields.py:129-140).  rhs must already carry the
 *     reference's adjust_rhs terms (ops.py:63-77).  `ws` is a device workspace of at least
 *     pa_solver_workspace_bytes(). */
size_t pa_solver_workspace_bytes(const pa_grid* g, int dtype, int method);
int pa_cg_solve(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg,
                void* ws, size_t ws_bytes, pa_report* report, void* stream);
int pa_bicgstab_solve(const pa_grid* g, const pa_equation* eq, int nfaces,
                      const pa_face_bc* faces, int dtype, void* x, void* x_alt,
                      const void* rhs, const pa_solver_cfg* cfg, void* ws, size_t ws_bytes,
                      pa_report* report, void* stream);
int pa_jacobi_solve(const pa_grid* g, const pa_equation* eq, int nfaces,
                    const pa_face_bc* faces, int dtype, void* x, void* x_alt, const void* rhs,
                    const pa_solver_cfg* cfg, void* ws, size_t ws_bytes, pa_report* report,
                    void* stream);

/* --- explicit Euler:  phi_new[slicer] = phi + dt*(rhs - A(phi)), then BCs on phi_new.
 *     rhs may be NULL (zero source).  phi and phi_new must not alias. */
int pa_euler_step(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, const void* phi, void* phi_new, const void* rhs, double dt,
                  void* stream);

/* --- `nsteps` explicit Euler steps ping-ponging between phi (input) and phi_alt; the step pair is
 *     replayed as a CUDA graph.  *result_in_alt = 1 if the final field is in phi_alt. */
int pa_euler_steps(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                   int dtype, void* phi, void* phi_alt, const void* rhs, double dt, int nsteps,
                   int* result_in_alt, void* stream);

/* --- out[i] = y[i] + a*x[i], i < n (a rounded to dtype first; out may alias y).  Builds the
 *     right-hand side rhs + phi_old/dt of an implicit Euler step (fdm.Ddt with a Krylov method;
 *     the reference's Ddt is a stub, fdm.py:315-353, intended semantics tests/test_fdm.py:275-299). */
int pa_axpy(int dtype, long long n, double a, const void* x, const void* y, void* out, void* stream);

/* --- end-to-end entry with HOST buffers: copies x and rhs to the device, solves with CG,
 *     copies the solution back.  x_host, rhs_host: n0*n1*n2 elements of dtype (pinned
 *     memory recommended).  Device scratch is allocated and freed inside the call. */
int pa_cg_solve_host(const pa_grid* g, const pa_equation* eq, int nfaces,
                     const pa_face_bc* faces, int dtype, void* x_host, const void* rhs_host,
                     const pa_solver_cfg* cfg, pa_report* report);

/* --- multi-GPU (one process per GPU, slab decomposition along kernel axis 0).
 *     The reference has no distributed path; these are new.  NCCL is loaded at run time.
 *     pa_comm_unique_id: rank 0 creates the 128-byte NCCL id, the host layer broadcasts it.
 *     pa_cg_solve_dist : same contract as pa_cg_solve on the LOCAL slab described by `g`
 *     (n[0] includes the ghost planes, goff0/gn0/olo0/ohi0 place it in the global grid); the
 *     halo exchange of r and the all-reduces of the CG scalars run inside the call. */
int pa_comm_unique_id(void* out128);
int pa_comm_create(const void* id128, int rank, int nranks, void** comm_out);
int pa_comm_destroy(void* comm);
int pa_cg_solve_dist(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                     int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg, void* ws,
                     size_t ws_bytes, void* comm, int rank, int nranks, pa_report* report,
                     void* stream);

/* --- peer-memory mailboxes for the fused all-reduces (optional; without them the solvers use
 *     ncclAllReduce).  Every rank: pa_p2p_local_handle() -> 64-byte CUDA IPC handle of its mailbox;
 *     the host layer all-gathers the handles; pa_p2p_attach(all handles, rank, nranks) maps the peers
 *     (NVLink peer access).  From then on the slab-decomposed solvers sum their dot products over the
 *     ranks INSIDE the kernels that produce them (one thread per rank writes its partial sums into every
 *     peer's mailbox and adds the others' in rank order) and finalize the scalar stage there: CG
 *     {d.Ad}, {r.r, |dx|^2}; BiCGSTAB {r0.v}, {|s|^2, t.s, t.t, r0.t}, {|r|^2}; Jacobi {|dx|^2}.
 *     A peer that does not arrive within 60 s ends the solve with PA_ERR_NCCL (device-side watchdog)
 *     and disables the mailboxes. */
int pa_p2p_local_handle(void* out64);
int pa_p2p_attach(const void* handles, int rank, int nranks);
int pa_p2p_enabled(void);
/* --- halo exchange over peer memory (CG on slabs): the IPC allocation behind pa_p2p_local_handle also holds a
 *     landing zone of 2 slots x 2 sides x cap bytes (cap: PA_HALO_MIB MiB, default 8) into which the neighbours'
 *     phase-B kernels store their boundary planes of r directly (NVLink stores + one flag word per side), so the
 *     CG loop contains no NCCL call.  pa_p2p_halo_cap() = this rank's bytes per landing plane (0: none);
 *     every rank must use the same value: the host layer takes the minimum over the ranks and calls
 *     pa_p2p_set_halo_cap(min).  A plane larger than the cap keeps the ncclSend/ncclRecv exchange. */
long long pa_p2p_halo_cap(void);
int pa_p2p_set_halo_cap(long long bytes);
int pa_p2p_disable(void); /* every rank must agree: the host layer disables all if one rank failed to attach */

/* --- any of the three solvers on a slab (method = PA_METHOD_*); pa_cg_solve_dist is the CG case.
 *     Nonlinear advection (pa_op.adv_is_iterate) works on slabs too: the new iterate's ghost planes are exchanged
 *     after every update.
 *     BiCGSTAB: p and s get their ghost planes by one send/recv pair each before the operator
 *     application that reads them, and {r0.v}, {|s|^2}, {t.s, t.t, r0.t}, {|r|^2} are all-reduced
 *     (4 per iteration).  Jacobi: one exchange of the new iterate + one all-reduce {|dx|^2}. */
int pa_solve_dist(int method, const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, void* x, void* x_alt, const void* rhs, const pa_solver_cfg* cfg, void* ws,
                  size_t ws_bytes, void* comm, int rank, int nranks, pa_report* report, void* stream);

/* --- pa_euler_steps on a slab: every step ends with the exchange of the new field's boundary
 *     planes (no reduction). */
int pa_euler_steps_dist(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                        int dtype, void* phi, void* phi_alt, const void* rhs, double dt, int nsteps,
                        int* result_in_alt, void* comm, int rank, int nranks, void* stream);

/* --- ghost planes of a slab-decomposed field: first / last OWNED plane -> the neighbours' ghost planes
 *     (ncclSend / ncclRecv, a ring when `ring` != 0); returns when the exchange is complete.  Used by the explicit
 *     FDC operators on a SlabMesh (the reference's operators read one cell up and down every axis). */
int pa_halo_exchange(const pa_grid* g, int dtype, void* phi, int ring, void* comm, int rank, int nranks, void* stream);

/* --- instrumented CG pass (measurement only): `iters` iterations with every section bracketed
 *     by CUDA events on the launching stream.  out_ms[6] = average per iteration of
 *     {phase A (d update + d.Ad), phase B (x,r update), BC faces + shell norm, whole iteration,
 *      kernel launches} and [5] = 1 if the tiled kernels ran. */
/*     (`variant`: pa_solver_cfg.variant, + 0x100 for PA_FLAG_CONTRACT) */
int pa_cg_profile(const pa_grid* g, const pa_equation* eq, int nfaces, const pa_face_bc* faces,
                  int dtype, void* x, void* x_alt, const void* rhs, int iters, int variant, void* ws,
                  size_t ws_bytes, double* out_ms, void* stream);

enum { PA_METHOD_CG = 0, PA_METHOD_BICGSTAB = 1, PA_METHOD_JACOBI = 2 };

#ifdef __cplusplus
}
#endif
#endif /* PYAPES_B200_H */
