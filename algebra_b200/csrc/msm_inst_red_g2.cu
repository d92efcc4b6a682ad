// explicit instantiation: MsmRedLaunch<CurveBlsG2>::reduce (the other members: msm_inst_red2_g2.cu — split for build parallelism)
#include "msm_k_red.cuh"
namespace ab200 {
template int MsmRedLaunch<CurveBlsG2>::reduce(const uint32_t *, MsmGeom, int, uint32_t, uint32_t *, uint32_t *, uint32_t *, cudaStream_t);
}  // namespace ab200
