// explicit instantiation: MsmAccLaunch<CurveBn> (see msm_common.cuh)
#include "msm_k_acc.cuh"
namespace ab200 {
template struct MsmAccLaunch<CurveBn>;
}  // namespace ab200
