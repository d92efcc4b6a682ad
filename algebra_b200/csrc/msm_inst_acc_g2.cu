// explicit instantiation: MsmAccLaunch<CurveBlsG2>::accumulate / occupancy (digits and merge: msm_inst_acc2_g2.cu)
#include "msm_k_acc.cuh"
namespace ab200 {
template int MsmAccLaunch<CurveBlsG2>::occupancy(bool);
template int MsmAccLaunch<CurveBlsG2>::accumulate(bool, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *,
                                                  uint32_t *, uint32_t *, uint32_t *, cudaStream_t);
}  // namespace ab200
