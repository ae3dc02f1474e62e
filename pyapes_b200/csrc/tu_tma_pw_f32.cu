// Translation unit: TMA star engine, float
#include "kernels_tma_pw.cuh"
namespace pa {
#define PA_INST(MODE)                                                                                       \
  template bool launch_star_tma<float, MODE>(cudaStream_t, const GridDev&, const EqDev<float>&, const TilePlan&, \
                                           const float*, const float*, float*, float*, float, SolverState*, double*, int);
PA_INST(PW_RESID)
PA_INST(PW_JACOBI)
PA_INST(PW_EULER)
PA_INST(PW_APPLY_V)
PA_INST(PW_APPLY_T)
#undef PA_INST
template bool launch_bi_st_tma<float>(cudaStream_t, const GridDev&, const EqDev<float>&, const TilePlan&, const float*,
                                   const float*, const float*, float*, float*, SolverState*, double*, int);
}  // namespace pa
