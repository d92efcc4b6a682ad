// explicit instantiation: MsmRedLaunch<CurveBls> (see msm_common.cuh)
#include "msm_k_red.cuh"
namespace ab200 {
template struct MsmRedLaunch<CurveBls>;
}  // namespace ab200
