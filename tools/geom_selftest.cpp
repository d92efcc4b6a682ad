// Host-side exerciser of the MSM window / bucket-slice geometry (algebra_b200/csrc/msm_common.cuh: make_geom), the one piece of
// host logic every MSM kernel launch depends on.  Test tool (tests/test_msm_geometry.py): not part of the product library.
// Build: g++ -std=c++17 -fpermissive -w -x c++ -I/usr/local/cuda/include -shared -fPIC (device intrinsics stay uninstantiated).
#include "../algebra_b200/csrc/msm_common.cuh"

extern "C" void selftest_geom(int c, int scalar_bits, int slice, int slices, unsigned out[8]) {
    const ab200::MsmGeom g = ab200::make_geom(c, scalar_bits, slice, slices);
    out[0] = (unsigned)g.c; out[1] = (unsigned)g.W; out[2] = (unsigned)g.top_bits; out[3] = g.nb;
    out[4] = g.nb_top; out[5] = g.total_buckets; out[6] = g.off; out[7] = g.off_top;
}
