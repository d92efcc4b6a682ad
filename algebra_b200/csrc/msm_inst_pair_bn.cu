// explicit instantiation: MsmPairLaunch<CurveBn> (see msm_common.cuh)
#include "msm_k_pair.cuh"
namespace ab200 {
template struct MsmPairLaunch<CurveBn>;
}  // namespace ab200
