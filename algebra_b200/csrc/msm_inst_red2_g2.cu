// explicit instantiation: MsmRedLaunch<CurveBlsG2>::combine / sum_or_affine (see msm_inst_red_g2.cu)
#include "msm_k_red.cuh"
namespace ab200 {
template int MsmRedLaunch<CurveBlsG2>::combine(const uint32_t *, int, int, uint32_t *, cudaStream_t);
template int MsmRedLaunch<CurveBlsG2>::sum_or_affine(bool, const uint32_t *, size_t, uint32_t *, cudaStream_t);
}  // namespace ab200
