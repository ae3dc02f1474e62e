// Translation unit: TMA-staged CG kernels, double
#include "kernels_tma.cuh"
namespace pa {
#define PA_INST(T)                                                                                             \
  template void launch_cg_phaseA_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&, int, T*, \
                                        SolverState*, double*);                                                \
  template void launch_cg_phaseB_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&, int, T*, \
                                        T*, SolverState*, double*, int);                                       \
  template bool launch_cg_coop_tma<T>(cudaStream_t, const TmaPlan&, const GridDev&, const EqDev<T>&, T*, T*, T*, T*, \
                                      T*, SolverState*, double*);
PA_INST(double)
#undef PA_INST
}  // namespace pa
