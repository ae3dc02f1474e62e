// Translation unit: shared-memory-resident kernels for 2-D grids (explicit Euler, whole-solve CG)
#include "kernels_resident.cuh"
namespace pa {
bool res_check_abort() { return res_abort_raised(); }
#define PA_INST(T)                                                                                                  \
  template bool launch_euler_resident<T>(cudaStream_t, const GridDev&, const pa_equation&, const EqDev<T>&, T*, T*, \
                                         const T*, T, int);                                                         \
  template bool launch_apply_direct2d<T, false>(cudaStream_t, const GridDev&, const EqDev<T>&, const T*, T*);           \
  template bool launch_apply_direct2d<T, true>(cudaStream_t, const GridDev&, const EqDev<T>&, const T*, T*);            \
  template bool launch_jacobi_resident<T>(cudaStream_t, const GridDev&, const pa_equation&, const EqDev<T>&, T*, T*,   \
                                          const T*, SolverState*, int);                                              \
  template bool launch_cg_resident<T>(cudaStream_t, const GridDev&, const EqDev<T>&, T*, T*, const T*, const T*, SolverState*, int, int);
PA_INST(double)
PA_INST(float)
#undef PA_INST
}  // namespace pa
