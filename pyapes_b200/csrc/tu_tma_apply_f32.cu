// Translation unit: TMA star engine, explicit operator application (PW_APPLY / PW_GRAD), float
#include "kernels_tma_pw.cuh"
namespace pa {
template bool launch_star_tma<float, PW_APPLY>(cudaStream_t, const GridDev&, const EqDev<float>&, const TilePlan&,
                                           const float*, const float*, float*, float*, float, SolverState*, double*, int);
template bool launch_star_grad<float>(cudaStream_t, const GridDev&, const EqDev<float>&, const TilePlan&, const float*,
                                  float*);
}  // namespace pa
