// explicit instantiation: MsmPairLaunch<CurveBls> (see msm_common.cuh)
#include "msm_k_pair.cuh"
namespace ab200 {
template struct MsmPairLaunch<CurveBls>;
}  // namespace ab200
