// Translation unit: register-tiled CG kernels (explicit instantiations; see api.cu extern templates)
#include "kernels_tiled.cuh"
namespace pa {
#define PA_INST(T)                                                                                          \
  template void launch_cg_phaseA<T>(cudaStream_t, const TilePlan&, const GridDev&, const EqDev<T>&, const T*, \
                                    const T*, T*, SolverState*, double*);                                   \
  template void launch_cg_phaseB<T>(cudaStream_t, const TilePlan&, const GridDev&, const EqDev<T>&, const T*, \
                                    T*, const T*, T*, SolverState*, double*);
PA_INST(double)
PA_INST(float)
#undef PA_INST
}  // namespace pa
