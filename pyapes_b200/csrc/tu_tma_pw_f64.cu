// Translation unit: TMA star engine, double
#include "kernels_tma_pw.cuh"
namespace pa {
#define PA_INST(MODE)                                                                                       \
  template bool launch_star_tma<double, MODE>(cudaStream_t, const GridDev&, const EqDev<double>&, const TilePlan&, \
                                           const double*, const double*, double*, double*, double, SolverState*, double*, int);
PA_INST(PW_RESID)
PA_INST(PW_JACOBI)
PA_INST(PW_EULER)
PA_INST(PW_APPLY_V)
PA_INST(PW_APPLY_T)
#undef PA_INST
template bool launch_bi_st_tma<double>(cudaStream_t, const GridDev&, const EqDev<double>&, const TilePlan&, const double*,
                                   const double*, const double*, double*, double*, SolverState*, double*, int);
}  // namespace pa
