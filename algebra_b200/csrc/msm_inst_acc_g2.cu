// explicit instantiation: MsmAccLaunch<CurveBlsG2> (see msm_common.cuh)
#include "msm_k_acc.cuh"
namespace ab200 {
template struct MsmAccLaunch<CurveBlsG2>;
}  // namespace ab200
