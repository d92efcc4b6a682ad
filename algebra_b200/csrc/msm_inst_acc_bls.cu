// explicit instantiation: MsmAccLaunch<CurveBls> (see msm_common.cuh)
#include "msm_k_acc.cuh"
namespace ab200 {
template struct MsmAccLaunch<CurveBls>;
}  // namespace ab200
