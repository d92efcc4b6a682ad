// explicit instantiation: MsmRedLaunch<CurveBn> (see msm_common.cuh)
#include "msm_k_red.cuh"
namespace ab200 {
template struct MsmRedLaunch<CurveBn>;
}  // namespace ab200
