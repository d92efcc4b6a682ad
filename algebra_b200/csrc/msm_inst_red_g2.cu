// explicit instantiation: MsmRedLaunch<CurveBlsG2> (see msm_common.cuh)
#include "msm_k_red.cuh"
namespace ab200 {
template struct MsmRedLaunch<CurveBlsG2>;
}  // namespace ab200
