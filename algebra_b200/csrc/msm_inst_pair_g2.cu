// explicit instantiation: MsmPairLaunch<CurveBlsG2> (see msm_common.cuh)
#include "msm_k_pair.cuh"
namespace ab200 {
template struct MsmPairLaunch<CurveBlsG2>;
}  // namespace ab200
