// Translation unit: TMA star engine, explicit operator application (PW_APPLY / PW_GRAD), double
#include "kernels_tma_pw.cuh"
namespace pa {
template bool launch_star_tma<double, PW_APPLY>(cudaStream_t, const GridDev&, const EqDev<double>&, const TilePlan&,
                                           const double*, const double*, double*, double*, double, SolverState*, double*, int);
template bool launch_star_grad<double>(cudaStream_t, const GridDev&, const EqDev<double>&, const TilePlan&, const double*,
                                  double*);
}  // namespace pa
