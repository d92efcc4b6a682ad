// msm_common.cuh — shared declarations of the MSM translation units: curve traits, window geometry, bucket I/O helpers and the
// launcher interfaces.  The kernels are templates over the curve; each (kernel group, curve) pair is instantiated in its own
// small .cu file (msm_inst_*.cu) so that the library builds in parallel, and msm.cu (host driver) only sees the launchers.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace ab200 {


// Curve traits: F = operations class of the coordinate field (Fp<P> for G1, Fp2<P> for G2), Fr = scalar-field parameter pack.
// (BlsFqRolled — Montgomery rows in a real loop, 4.6k instead of 7.6k SASS instructions — was measured SLOWER on B200:
// accumulate 338 -> 371 ms @2^26; the unrolled rows let ptxas interleave six carry chains.)
struct CurveBls {
    using F = Fp<BlsFq>;
    using Fr = BlsFr;
    static constexpr int SCALAR_BITS = 255;  // Fr::MODULUS_BIT_SIZE (variable_base/mod.rs:451)
    static constexpr bool AUTO_LEVELS = true;
    static constexpr int LEVEL_CAP = 4;      // automatic mode: at most this many batched-affine levels
    static constexpr int PAIR_MINB = 4;      // resident blocks per SM the pair-add kernels are compiled for (128 registers)
};
struct CurveBn {
    using F = Fp<BnFq>;
    using Fr = BnFr;
    static constexpr int SCALAR_BITS = 254;
    // 8-limb coordinates: the affine additions are 2.2x cheaper in multiplies but move almost as many bytes.  With the generation-1
    // kernel the levels measured slower (BN254 2^24: 49 -> 55 ms, round 1); with generation 2 they win: 49.4 ms without, 45.1 ms with
    // 3 levels, 45.9 ms with 4 (profiles/r02_levels_bn254_g2.log)
    static constexpr bool AUTO_LEVELS = true;
    static constexpr int LEVEL_CAP = 3;
    static constexpr int PAIR_MINB = 4;
};
// G2 of BLS12-381: coordinates in Fq2 (curves/bls12_381/src/curves/g2.rs:54), same scalar field
struct CurveBlsG2 {
    using F = Fp2<BlsFq>;
    using Fr = BlsFr;
    static constexpr int SCALAR_BITS = 255;
    static constexpr bool AUTO_LEVELS = true;   // G2 2^22: 100.4 ms without levels, 80.4 ms with 3, 74.4 ms with 4 (profiles/r02_levels_bn254_g2.log)
    static constexpr int LEVEL_CAP = 4;
    static constexpr int PAIR_MINB = 2;      // 24-word coordinates: 255 registers, two blocks per SM
};

struct MsmGeom {
    int c, W, top_bits;          // window bits, number of windows, bits in the top window
    uint32_t nb;                 // buckets of this slice per non-top window (whole window: 2^(c-1))
    uint32_t nb_top;             // buckets of this slice in the top window  (whole window: 2^top_bits)
    uint32_t total_buckets;      // (W-1)*nb + nb_top
    uint32_t off, off_top;       // first bucket of the slice inside a non-top / the top window (0 for the whole window)
};

// Bucket slice `slice` of `slices`: every window keeps the contiguous range [nbw*slice/slices, nbw*(slice+1)/slices) of its
// buckets (bucket j of a window has weight j+1, so a slice's sums carry `off` as an extra chunk offset in the reduction) and
// the slices' results add up to the complete MSM.  A window with fewer buckets than slices is given whole to slice 0.
inline MsmGeom make_geom(int c, int scalar_bits, int slice = 0, int slices = 1) {
    MsmGeom g;
    g.c = c;
    g.W = (scalar_bits + c - 1) / c;  // digits_count (:452)
    g.top_bits = scalar_bits - (g.W - 1) * c;
    const uint32_t full = 1u << (c - 1), full_top = 1u << g.top_bits;
    auto cut = [&](uint32_t nbw, int i) -> uint32_t {
        if (nbw < (uint32_t)slices) return i == 0 ? 0u : nbw;
        return (uint32_t)((uint64_t)nbw * (uint64_t)i / (uint64_t)slices);
    };
    g.off = cut(full, slice);
    g.nb = cut(full, slice + 1) - g.off;
    g.off_top = cut(full_top, slice);
    g.nb_top = cut(full_top, slice + 1) - g.off_top;
    g.total_buckets = (uint32_t)(g.W - 1) * g.nb + g.nb_top;
    return g;
}

static constexpr uint32_t kNoBucket = 0xffffffffu;
static constexpr uint32_t kFixupSmall = 16;

template <int L> __device__ __forceinline__ void load_xyzz(Xyzz<L> &b, const uint32_t *p) {
    load_limbs<L>(b.x, p);
    load_limbs<L>(b.y, p + L);
    load_limbs<L>(b.zz, p + 2 * L);
    load_limbs<L>(b.zzz, p + 3 * L);
}
template <int L> __device__ __forceinline__ void store_xyzz(uint32_t *p, const Xyzz<L> &b) {
    store_limbs<L>(p, b.x);
    store_limbs<L>(p + L, b.y);
    store_limbs<L>(p + 2 * L, b.zz);
    store_limbs<L>(p + 3 * L, b.zzz);
}

// ---- launchers (definitions: msm_k_pair.cuh / msm_k_acc.cuh / msm_k_red.cuh; instantiations: msm_inst_*.cu) ----
template <class C> struct MsmPairLaunch {
    // one batched-affine level over the buckets [0, nbg): variant 1 = thread-contiguous batches, 2 = warp-interleaved + cp.async
    static int run(int variant, bool first, const uint32_t *bases, const uint32_t *src, const uint32_t *offsets_in, const uint32_t *offsets_out,
                   const uint32_t *pairmap, uint32_t nbg, uint32_t batch, size_t out_cap, uint32_t *out, int shared_inv, int stagger, uint32_t base_stride, cudaStream_t st);
    // copy of the bases with one point per 128-byte line (base_stride = 32 words); only for points of <= 128 bytes
    static int pad_bases(const uint32_t *bases, size_t n, uint32_t *padded, cudaStream_t st);
};
template <class C> struct MsmAccLaunch {
    static int digits(int mode, const void *scalars, int kind, size_t nk, MsmGeom g, int w0, int w1, uint32_t *counts_or_cursor, uint32_t *sorted,
                      cudaStream_t st);
    static int occupancy(bool direct);   // resident 128-thread blocks per SM of the accumulate kernel
    // XYZZ accumulation + both fix-up kernels
    static int accumulate(bool direct, const uint32_t *pts, const uint32_t *sorted_idx, const uint32_t *offs, uint32_t nbk, uint32_t T,
                          uint32_t num_tasks, uint32_t *dst, uint32_t *head, uint32_t *tail, uint32_t *head_bucket, uint32_t *tail_bucket,
                          cudaStream_t st);
    static int merge(uint32_t *buckets, const uint32_t *extra, uint32_t nb, cudaStream_t st);
};
template <class C> struct MsmRedLaunch {
    // partials: W * chunks buckets of scratch; partials2: W * 32
    static int reduce(const uint32_t *buckets, MsmGeom g, int log_m, uint32_t chunks, uint32_t *partials, uint32_t *partials2, uint32_t *window_sums,
                      cudaStream_t st);
    static int combine(const uint32_t *window_sums, int W, int c, uint32_t *out, cudaStream_t st);
    static int sum_or_affine(bool to_affine, const uint32_t *d_in, size_t k, uint32_t *d_out, cudaStream_t st);
};

}  // namespace ab200
