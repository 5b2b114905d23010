// sg2d_sep.cu -- separable fast path of the 2D filter (placeholder: planning disabled until the
// tiled kernel lands; every filter currently runs through sg2d_direct.cu).
#include "sg2d.h"

namespace sg2d {

void plan_separable(int nx, int ny, int, const double*, const float*, SepPlan* plan)
{
    plan->rank = 0;
    plan->nx = nx;
    plan->ny = ny;
    plan->max_err = 0.0f;
}

cudaError_t launch_separable(const Args2D&, const SepPlan&, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace sg2d
