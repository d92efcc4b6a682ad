// msm.cu — variable-base MSM over short-Weierstrass G1 (a = 0) for sm_100a: bucket method (Pippenger) with signed c-bit
// digits, counting-sort bucket assignment, batched-affine pre-reduction and task-balanced XYZZ accumulation.
//
// Replaces VariableBaseMSM::msm_unchecked for Projective<P> (ec/src/scalar_mul/variable_base/mod.rs:59-64
// -> msm_bigint_wnaf_parallel :437-503).  Same mathematics, GPU schedule:
//   reference (per window, serial over points)          here (all windows at once)
//   into_bigint + make_digits (:60-62, :754-794)    ->  msm_digits_kernel<HIST>: REDC + signed digits, histogram of (window, |digit|)
//   buckets[|d|-1] +=/-= base  (:467-475)           ->  exclusive scan; msm_digits_kernel<SCATTER> writes (index | sign) in bucket order;
//                                                       msm_pair_add_kernel x up to 4 levels: every bucket run is halved with AFFINE
//                                                       additions that share one inversion per batch (Montgomery's trick);
//                                                       msm_accumulate_kernel: one thread per task of T entries runs the reference's
//                                                       `Bucket += Affine` (bucket.rs:168-238) on what is left, flushing per bucket;
//                                                       msm_fixup_*: buckets that span tasks
//   running-sum  res += running_sum (:478-484)      ->  msm_bucket_reduce_kernel (running sum per chunk of 32 buckets + chunk offset times
//                                                       chunk total) and msm_sum_partials_kernel (tree over the chunks of a window)
//   window combine, c doublings per window (:489-502) -> msm_window_combine_kernel (Jacobian, one thread)
// EC addition is commutative/associative, so bucket order, atomics, pairing and chunking do not change the group element;
// results are compared with the reference after into_affine(), limb-exact.
// The digit recoding is the reference's make_digits (top window unsigned), so bucket counts per window are 2^(c-1), and
// 2^(lambda-(W-1)c) for the top window.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>
#include <vector>

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
    static constexpr int PAIR_MINB = 4;      // resident blocks per SM the pair-add kernels are compiled for (128 registers)
};
struct CurveBn {
    using F = Fp<BnFq>;
    using Fr = BnFr;
    static constexpr int SCALAR_BITS = 254;
    // 8-limb coordinates: the affine additions are 2.2x cheaper in multiplies but move almost as many bytes, and the levels
    // measured slower (BN254 2^24: 49 -> 55 ms), so the automatic mode keeps them for the 12-limb field only
    static constexpr bool AUTO_LEVELS = false;
    static constexpr int PAIR_MINB = 4;
};
// G2 of BLS12-381: coordinates in Fq2 (curves/bls12_381/src/curves/g2.rs:54), same scalar field
struct CurveBlsG2 {
    using F = Fp2<BlsFq>;
    using Fr = BlsFr;
    static constexpr int SCALAR_BITS = 255;
    static constexpr bool AUTO_LEVELS = false;
    static constexpr int PAIR_MINB = 2;      // 24-word coordinates: 255 registers, two blocks per SM
};

struct MsmGeom {
    int c, W, top_bits;          // window bits, number of windows, bits in the top window
    uint32_t nb;                 // buckets per non-top window = 2^(c-1)
    uint32_t nb_top;             // buckets in the top window   = 2^top_bits
    uint32_t total_buckets;      // (W-1)*nb + nb_top
};

static MsmGeom make_geom(int c, int scalar_bits) {
    MsmGeom g;
    g.c = c;
    g.W = (scalar_bits + c - 1) / c;  // digits_count (:452)
    g.top_bits = scalar_bits - (g.W - 1) * c;
    g.nb = 1u << (c - 1);
    g.nb_top = 1u << g.top_bits;
    g.total_buckets = (uint32_t)(g.W - 1) * g.nb + g.nb_top;
    return g;
}

// ------------------------------------------------------------------------------------------------
// digits: canonical scalar -> signed digits (make_digits, :754-794); MODE 0 = histogram, 1 = scatter
// ------------------------------------------------------------------------------------------------
template <class C, int MODE>
__global__ void __launch_bounds__(256) msm_digits_kernel(const void *__restrict__ scalars_v, int kind, size_t n, MsmGeom g, int w_lo, int w_hi,
                                                         uint32_t *__restrict__ counts_or_cursor, uint32_t *__restrict__ sorted) {
    using FR = Fp<typename C::Fr>;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[10];
#pragma unroll
    for (int j = 0; j < 10; j++) k[j] = 0;
    if (kind <= B200_SCALARS_BIGINT) {
        uint32_t s[8];
        load_limbs_nc<8>(s, (const uint32_t *)scalars_v + i * 8);
        if (kind == B200_SCALARS_FR_MONT) FR::from_mont(k, s);  // into_bigint (:60-62)
        else limbs_copy<8>(k, s);                               // msm_bigint: already canonical (:80-85)
    } else if (kind == B200_SCALARS_U8) k[0] = ((const uint8_t *)scalars_v)[i];
    else if (kind == B200_SCALARS_U16) k[0] = ((const uint16_t *)scalars_v)[i];
    else if (kind == B200_SCALARS_U32) k[0] = ((const uint32_t *)scalars_v)[i];
    else { uint2 v = ((const uint2 *)scalars_v)[i]; k[0] = v.x; k[1] = v.y; }
    // msm_signed's negative classes (variable_base/mod.rs:251-336): a scalar whose r - k fits 64 bits (NegU1..NegU64) is
    // handled as -(r - k), i.e. the point enters with the opposite sign and only ceil(64/c) windows are non-zero
    uint32_t flip = 0;
    if (kind <= B200_SCALARS_BIGINT && (k[2] | k[3] | k[4] | k[5] | k[6] | k[7])) {
        using R = typename C::Fr;
        uint32_t t[8];
        t[0] = ptx::sub_cc(R::MOD(0), k[0]);
#pragma unroll
        for (int j = 1; j < 8; j++) t[j] = ptx::subc_cc(R::MOD(j), k[j]);
        const uint32_t borrow = ptx::subc(0u, 0u);   // non-zero iff k > r (non-canonical msm_bigint input: left alone)
        if (!borrow && !(t[2] | t[3] | t[4] | t[5] | t[6] | t[7])) {
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = t[j];
            flip = 1;
        }
    }
    const int c = g.c;
    const uint32_t mask = (1u << c) - 1, half = 1u << (c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < w_hi; w++) {  // the carry chain needs every lower window; only [w_lo, w_hi) is emitted
        const int bit = w * c, wi = bit >> 5, sh = bit & 31;
        uint64_t two = ((uint64_t)k[wi + 1] << 32) | k[wi];
        uint32_t coef = ((uint32_t)(two >> sh) & mask) + carry;
        uint32_t mag, neg = 0;
        if (w == g.W - 1) {  // top digit stays unsigned (:789-791); bits above the declared scalar width are ignored
            mag = (((uint32_t)(two >> sh) & mask) & ((1u << g.top_bits) - 1)) + carry;
        } else {
            carry = (coef + half) >> c;
            if (carry) { mag = (1u << c) - coef; neg = 1; }  // digit = coef - 2^c in [-2^(c-1), 0)
            else mag = coef;
        }
        if (mag && w >= w_lo) {
            uint32_t gid = (uint32_t)w * g.nb + (mag - 1);
            if (MODE == 0) {
                atomicAdd(&counts_or_cursor[gid], 1u);
            } else {
                uint32_t pos = atomicAdd(&counts_or_cursor[gid], 1u);
                sorted[pos] = (uint32_t)i | ((neg ^ flip) << 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of u32 counts (total < 2^32): block totals -> serial scan of totals -> apply
// ------------------------------------------------------------------------------------------------
static constexpr int kScanThreads = 512, kScanPerThread = 8, kScanBlock = kScanThreads * kScanPerThread;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += y;
        }
        warp_sums[lane] = ws;  // inclusive
    }
    __syncthreads();
    uint32_t before = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    return before + x - v;  // exclusive
}

__global__ void __launch_bounds__(kScanThreads) scan_block_totals_kernel(const uint32_t *in, size_t n, uint32_t *block_totals) {
    size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++)
        if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_totals[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanThreads) scan_totals_kernel(uint32_t *block_totals, size_t nblocks) {
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (size_t base = 0; base < nblocks; base += kScanThreads) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_totals[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total);
        uint32_t r0 = running;
        if (i < nblocks) block_totals[i] = r0 + ex;
        __syncthreads();
        if (threadIdx.x == 0) running = r0 + total;
        __syncthreads();
    }
}
// out[i] = exclusive prefix; out[n] = grand total (out has n+1 entries)
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t *in, size_t n, const uint32_t *block_offsets, uint32_t *out) {
    size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    uint32_t v[kScanPerThread], s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPerThread; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

// ------------------------------------------------------------------------------------------------
// bucket accumulation — the hot loop (Bucket += Affine, bucket.rs:168-238), balanced by construction:
// the bucket-sorted entry array is cut into tasks of exactly T consecutive entries, one thread per task, whatever the
// bucket sizes are (a scalar distribution that piles everything into one bucket, or a short top window with 8 buckets
// holding n/8 points each, costs the same as the uniform case).  A thread walks its slice and flushes an accumulator at
// every bucket boundary: buckets that begin and end inside the slice are written straight to `buckets`; the piece of a
// bucket that began in an earlier task goes to head[t], the piece of a bucket that continues into the next task to tail[t].
// msm_fixup_* then add tail[t0] + head[t0+1..t1] for every bucket that spans tasks.  `buckets` is pre-zeroed (zz = zzz = 0
// is the XYZZ identity) so empty buckets need no writer.
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kNoBucket = 0xffffffffu;

template <int L> __device__ __forceinline__ void store_xyzz(uint32_t *p, const Xyzz<L> &b);
template <int L> __device__ __forceinline__ void load_xyzz(Xyzz<L> &b, const uint32_t *p);

// DIRECT = false: entry p is `sorted[p]` = (base index | sign<<31), gathered from `bases`;
// DIRECT = true : entry p is the affine point stored at bases[p] (output of the batched-affine pre-reduction), no sign.
template <class C, bool DIRECT>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                             const uint32_t *__restrict__ offsets, uint32_t total_buckets, uint32_t T,
                                                             uint32_t *__restrict__ buckets, uint32_t *__restrict__ head,
                                                             uint32_t *__restrict__ tail, uint32_t *__restrict__ head_bucket,
                                                             uint32_t *__restrict__ tail_bucket, uint32_t num_tasks) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tasks) return;
    // `offsets` may be a window group's slice of a larger offsets array: entry positions start at offsets[0]
    const uint32_t M = __ldg(offsets + total_buckets);
    const uint64_t lo64 = (uint64_t)__ldg(offsets) + (uint64_t)t * T;
    if (lo64 >= M) {
        head_bucket[t] = kNoBucket;
        tail_bucket[t] = kNoBucket;
        return;
    }
    const uint32_t lo = (uint32_t)lo64, hi = (uint32_t)min((uint64_t)M, lo64 + T);
    // b = last index with offsets[b] <= lo  (=> offsets[b] <= lo < offsets[b+1])
    uint32_t bl = 0, br = total_buckets;  // invariant: offsets[bl] <= lo, offsets[br] > lo (offsets[total] = M > lo)
    while (br - bl > 1) {
        uint32_t mid = bl + ((br - bl) >> 1);
        if (__ldg(offsets + mid) <= lo) bl = mid; else br = mid;
    }
    uint32_t b = bl, bucket_end = __ldg(offsets + b + 1);
    bool started_before = __ldg(offsets + b) < lo;
    uint32_t hb = kNoBucket;

    typename E::B acc;
    E::xyzz_set_zero(acc);
    uint32_t cx[L], cy[L], nx[L], ny[L];
    uint32_t e = DIRECT ? lo : __ldg(sorted + lo), e_next = 0;
    {
        const uint32_t *bp = bases + (size_t)(DIRECT ? e : (e & 0x7fffffffu)) * (2 * L);
        load_limbs_nc<L>(cx, bp);
        load_limbs_nc<L>(cy, bp + L);
    }
    for (uint32_t pos = lo; pos < hi; pos++) {
        const bool more = (pos + 1 < hi);
        if (more) {  // issue the next gather before the ~10 modmuls of this addition
            e_next = DIRECT ? pos + 1 : __ldg(sorted + pos + 1);
            const uint32_t *bp = bases + (size_t)(DIRECT ? e_next : (e_next & 0x7fffffffu)) * (2 * L);
            load_limbs_nc<L>(nx, bp);
            load_limbs_nc<L>(ny, bp + L);
        }
        if (pos == bucket_end) {  // bucket b is complete: flush, move to the (non-empty) bucket that owns `pos`
            if (started_before) { store_xyzz<L>(head + (size_t)t * (4 * L), acc); hb = b; }
            else store_xyzz<L>(buckets + (size_t)b * (4 * L), acc);
            E::xyzz_set_zero(acc);
            started_before = false;
            do { b++; bucket_end = __ldg(offsets + b + 1); } while (bucket_end <= pos);
        }
        E::madd(acc, cx, cy, !DIRECT && (e >> 31) != 0);
        if (more) {
            limbs_copy<L>(cx, nx);
            limbs_copy<L>(cy, ny);
            e = e_next;
        }
    }
    uint32_t tb = kNoBucket;
    if (bucket_end == hi) {  // the last bucket ends exactly with the slice
        if (started_before) { store_xyzz<L>(head + (size_t)t * (4 * L), acc); hb = b; }
        else store_xyzz<L>(buckets + (size_t)b * (4 * L), acc);
    } else if (started_before) {  // the whole slice is an inner piece of one bucket
        store_xyzz<L>(head + (size_t)t * (4 * L), acc);
        hb = b;
    } else {
        store_xyzz<L>(tail + (size_t)t * (4 * L), acc);
        tb = b;
    }
    head_bucket[t] = hb;
    tail_bucket[t] = tb;
}

// ------------------------------------------------------------------------------------------------
// Batched-affine pre-reduction (optional stage between the sort and the XYZZ accumulation).
// One level halves every bucket's run: output slot j of bucket b = in[2j] + in[2j+1] (or in[2j] alone when the run is odd),
// as AFFINE points.  Affine addition needs 1/(x2 - x1); a thread owns `batch` consecutive output slots and inverts all
// their denominators with ONE field inversion (Montgomery's trick: forward pass stores the running products in the
// output slots themselves, backward pass peels them off) — ~6 modmuls + inversion/batch per addition instead of the
// 10 of an XYZZ mixed addition.  Degenerate pairs keep the batch intact with a denominator of 1: an identity operand passes
// the other one through, equal points are doubled (denominator 2y), opposite points give the identity (0,0).
// Same group element as the reference's bucket sums (EC addition is associative); parity is checked after into_affine().
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_halve_counts_kernel(const uint32_t *__restrict__ offsets_in, uint32_t total_buckets,
                                                               uint32_t *__restrict__ counts_out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    uint32_t cnt = offsets_in[b + 1] - offsets_in[b];
    counts_out[b] = (cnt + 1) >> 1;
}

template <class F, bool FIRST>
__device__ __forceinline__ void pair_load_point(uint32_t *x, uint32_t *y, const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src,
                                                uint32_t k) {
    constexpr int L = F::L;
    if (FIRST) {
        const uint32_t e = __ldg(src + k);
        const uint32_t *bp = bases + (size_t)(e & 0x7fffffffu) * (2 * L);
        load_limbs_nc<L>(x, bp);
        load_limbs_nc<L>(y, bp + L);
        F::cneg(y, y, (e >> 31) != 0);   // -(0,0) stays (0,0)
    } else {
        const uint32_t *bp = src + (size_t)k * (2 * L);
        load_limbs_nc<L>(x, bp);
        load_limbs_nc<L>(y, bp + L);
    }
}
// x coordinate only (the forward pass needs y only for the rare degenerate pairs)
template <class F, bool FIRST>
__device__ __forceinline__ void pair_load_x(uint32_t *x, const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src, uint32_t k) {
    constexpr int L = F::L;
    const uint32_t *bp = FIRST ? bases + (size_t)(__ldg(src + k) & 0x7fffffffu) * (2 * L) : src + (size_t)k * (2 * L);
    load_limbs_nc<L>(x, bp);
}

enum { PAIR_PASS1 = 0, PAIR_PASS2 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };
// classify (P1, P2) and produce the denominator of the slope (ONE for the degenerate kinds)
template <class F> __device__ __forceinline__ int pair_classify(uint32_t *den, const uint32_t *x1, const uint32_t *y1, const uint32_t *x2,
                                                                const uint32_t *y2, bool has2) {
    constexpr int L = F::L;
    const bool z1 = limbs_is_zero<L>(x1) && limbs_is_zero<L>(y1);
    const bool z2 = !has2 || (limbs_is_zero<L>(x2) && limbs_is_zero<L>(y2));
    F::set_one(den);
    if (z2) return PAIR_PASS1;            // also covers "both identity" (P1 = (0,0) passes through)
    if (z1) return PAIR_PASS2;
    if (limbs_eq<L>(x1, x2)) {
        if (limbs_eq<L>(y1, y2) && !limbs_is_zero<L>(y1)) { F::dbl(den, y1); return PAIR_DBL; }
        return PAIR_INF;
    }
    F::sub(den, x2, x1);
    return PAIR_ADD;
}

// Latency hiding in this kernel is left to occupancy (128 registers -> 16 warps per SM).  Measured alternatives @2^26,
// accumulation phase with 4 levels: plain loads 285 ms; next-slot operands held in registers (198 regs, 8 warps/SM) 329 ms;
// prefetch.global.L2 of the next slot's operands (fetches whole 128-byte lines for 96-byte points) 346 ms.
template <class C, bool FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_pair_add_kernel(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src,
                                                           const uint32_t *__restrict__ offsets_in, const uint32_t *__restrict__ offsets_out,
                                                           uint32_t total_buckets, uint32_t batch, uint32_t *__restrict__ out,
                                                           uint32_t num_threads) {
    using F = typename C::F;
    constexpr int L = F::L;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_threads) return;
    const uint32_t M = __ldg(offsets_out + total_buckets);
    const uint64_t lo64 = (uint64_t)t * batch;
    if (lo64 >= M) return;
    const uint32_t lo = (uint32_t)lo64, hi = (uint32_t)min((uint64_t)M, lo64 + batch);
    uint32_t bl = 0, br = total_buckets;   // last b with offsets_out[b] <= lo
    while (br - bl > 1) {
        uint32_t mid = bl + ((br - bl) >> 1);
        if (__ldg(offsets_out + mid) <= lo) bl = mid; else br = mid;
    }
    uint32_t x1[L], y1[L], x2[L], y2[L], den[L], run[L];
    F::set_one(run);
    // walk state = the bucket owning the NEXT slot (one ahead of the slot being computed)
    uint32_t b = bl, out_end = __ldg(offsets_out + b + 1), out_beg = __ldg(offsets_out + b), in_beg = __ldg(offsets_in + b),
             in_end = __ldg(offsets_in + b + 1);
    // ---- forward: running product of the denominators, parked in the x-half of each output slot (x coordinates only)
    uint32_t k = in_beg + 2 * (lo - out_beg);
    bool has2 = k + 1 < in_end;
    for (uint32_t p = lo; p < hi; p++) {
        uint32_t kn = 0;
        bool has2n = false;
        if (p + 1 < hi) {
            while (p + 1 >= out_end) {
                b++;
                out_beg = out_end;
                out_end = __ldg(offsets_out + b + 1);
                in_beg = __ldg(offsets_in + b);
                in_end = __ldg(offsets_in + b + 1);
            }
            kn = in_beg + 2 * (p + 1 - out_beg);
            has2n = kn + 1 < in_end;
        }
        if (has2) {
            pair_load_x<F, FIRST>(x1, bases, src, k);
            pair_load_x<F, FIRST>(x2, bases, src, k + 1);
            if (limbs_is_zero<L>(x1) || limbs_is_zero<L>(x2) || limbs_eq<L>(x1, x2)) {   // rare: identity operand / equal x
                pair_load_point<F, FIRST>(x1, y1, bases, src, k);
                pair_load_point<F, FIRST>(x2, y2, bases, src, k + 1);
                const int kind = pair_classify<F>(den, x1, y1, x2, y2, true);
                if (kind >= PAIR_ADD) F::mul(run, run, den);
            } else {
                F::sub(den, x2, x1);
                F::mul(run, run, den);
            }
        }
        store_limbs<L>(out + (size_t)p * (2 * L), run);
        k = kn;
        has2 = has2n;
    }
    uint32_t inv[L];
    F::inv(inv, run);
    // ---- backward: peel the inverses off and write the sums; the walk state now sits on the bucket of slot hi-1
    k = in_beg + 2 * (hi - 1 - out_beg);
    has2 = k + 1 < in_end;
    for (uint32_t p = hi; p-- > lo;) {
        uint32_t kn = 0;
        bool has2n = false;
        if (p > lo) {
            while (p - 1 < out_beg) {
                b--;
                out_end = out_beg;
                out_beg = __ldg(offsets_out + b);
                in_beg = __ldg(offsets_in + b);
                in_end = __ldg(offsets_in + b + 1);
            }
            kn = in_beg + 2 * (p - 1 - out_beg);
            has2n = kn + 1 < in_end;
        }
        pair_load_point<F, FIRST>(x1, y1, bases, src, k);
        if (has2) pair_load_point<F, FIRST>(x2, y2, bases, src, k + 1);
        const int kind = pair_classify<F>(den, x1, y1, x2, y2, has2);
        uint32_t *o = out + (size_t)p * (2 * L);
        if (kind >= PAIR_ADD) {
            uint32_t dinv[L], lam[L], t3[L];
            if (p > lo) { load_limbs<L>(t3, out + (size_t)(p - 1) * (2 * L)); F::mul(dinv, inv, t3); }   // inv * prefix_{p-1} = 1/den
            else limbs_copy<L>(dinv, inv);
            F::mul(inv, inv, den);
            if (kind == PAIR_ADD) {
                F::sub(lam, y2, y1);
            } else {                      // doubling: slope = 3 x^2 / (2 y)
                F::sqr(lam, x1);
                F::dbl(t3, lam);
                F::add(lam, lam, t3);
                limbs_copy<L>(x2, x1);
            }
            F::mul(lam, lam, dinv);
            F::sqr(t3, lam);
            F::sub(t3, t3, x1);
            F::sub(t3, t3, x2);           // x3
            F::sub(x2, x1, t3);
            F::mul(x2, lam, x2);
            F::sub(x2, x2, y1);           // y3 = lam (x1 - x3) - y1
            store_limbs<L>(o, t3);
            store_limbs<L>(o + L, x2);
        } else if (kind == PAIR_PASS1) {
            store_limbs<L>(o, x1);
            store_limbs<L>(o + L, y1);
        } else if (kind == PAIR_PASS2) {
            store_limbs<L>(o, x2);
            store_limbs<L>(o + L, y2);
        } else {
            F::set_zero(x1);
            store_limbs<L>(o, x1);
            store_limbs<L>(o + L, x1);
        }
        k = kn;
        has2 = has2n;
    }
}

// bucket b = tail[t0] + head[t0+1] + ... + head[t1], t1 = task holding the bucket's last entry.
// small spans: one thread per task boundary; long spans (heavy buckets): one block per bucket.
static constexpr uint32_t kFixupSmall = 16;
template <class C>
__global__ void __launch_bounds__(128) msm_fixup_small_kernel(const uint32_t *__restrict__ offsets, uint32_t T, const uint32_t *__restrict__ head,
                                                              const uint32_t *__restrict__ tail, const uint32_t *__restrict__ tail_bucket,
                                                              uint32_t num_tasks, uint32_t *__restrict__ buckets) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tasks) return;
    const uint32_t b = tail_bucket[t];
    if (b == kNoBucket) return;
    const uint32_t t1 = (__ldg(offsets + b + 1) - 1 - __ldg(offsets)) / T;
    if (t1 - t > kFixupSmall) return;
    typename E::B acc, x;
    load_xyzz<L>(acc, tail + (size_t)t * (4 * L));
    for (uint32_t k = t + 1; k <= t1; k++) {
        load_xyzz<L>(x, head + (size_t)k * (4 * L));
        E::xyzz_add(acc, x);
    }
    store_xyzz<L>(buckets + (size_t)b * (4 * L), acc);
}
template <class C>
__global__ void __launch_bounds__(128) msm_fixup_big_kernel(const uint32_t *__restrict__ offsets, uint32_t T, const uint32_t *__restrict__ head,
                                                            const uint32_t *__restrict__ tail, const uint32_t *__restrict__ tail_bucket,
                                                            uint32_t num_tasks, uint32_t *__restrict__ buckets) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    extern __shared__ uint32_t sm[];
    const uint32_t t = blockIdx.x;
    if (t >= num_tasks) return;
    const uint32_t b = tail_bucket[t];
    if (b == kNoBucket) return;
    const uint32_t t1 = (__ldg(offsets + b + 1) - 1 - __ldg(offsets)) / T;
    if (t1 - t <= kFixupSmall) return;
    typename E::B acc, x;
    E::xyzz_set_zero(acc);
    if (threadIdx.x == 0) load_xyzz<L>(acc, tail + (size_t)t * (4 * L));
    for (uint32_t k = t + 1 + threadIdx.x; k <= t1; k += blockDim.x) {
        load_xyzz<L>(x, head + (size_t)k * (4 * L));
        E::xyzz_add(acc, x);
    }
    store_xyzz<L>(sm + threadIdx.x * (4 * L), acc);
    __syncthreads();
    for (uint32_t s2 = blockDim.x >> 1; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) {
            load_xyzz<L>(acc, sm + threadIdx.x * (4 * L));
            load_xyzz<L>(x, sm + (threadIdx.x + s2) * (4 * L));
            E::xyzz_add(acc, x);
            store_xyzz<L>(sm + threadIdx.x * (4 * L), acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        load_xyzz<L>(acc, sm);
        store_xyzz<L>(buckets + (size_t)b * (4 * L), acc);
    }
}

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

// ------------------------------------------------------------------------------------------------
// bucket reduction.  Window w needs S_w = sum_j (j+1) * B_w[j]  (:478-484).  Thread t of a window takes buckets
// [t*m, (t+1)*m): running sum gives  sum_l (l+1)*B[t*m+l]  and the chunk total R_t; adding (t*m) * R_t (double-and-add)
// makes its contribution complete.  partial index = window * chunks_stride + t.
// ------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) msm_bucket_reduce_kernel(const uint32_t *__restrict__ buckets, MsmGeom g, int log_m,
                                                                uint32_t chunks_per_window, uint32_t *__restrict__ partials) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w = tid / chunks_per_window, t = tid % chunks_per_window;
    if (w >= (uint32_t)g.W) return;
    const uint32_t nbw = (w == (uint32_t)g.W - 1) ? g.nb_top : g.nb;
    const uint32_t m = 1u << log_m;
    typename E::B run, sum;
    E::xyzz_set_zero(run);
    E::xyzz_set_zero(sum);
    const uint32_t lo = t * m;
    if (lo < nbw) {
        const uint32_t hi = min(lo + m, nbw);
        const uint32_t *base = buckets + ((size_t)w * g.nb) * (4 * L);
        for (uint32_t j = hi; j-- > lo;) {
            typename E::B b;
            load_xyzz<L>(b, base + (size_t)j * (4 * L));
            E::xyzz_add(run, b);
            E::xyzz_add(sum, run);
        }
        // sum += lo * run
        if (lo != 0 && !E::xyzz_is_zero(run)) {
            typename E::B acc;
            E::xyzz_set_zero(acc);
            for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
                if (!E::xyzz_is_zero(acc)) E::xyzz_dbl(acc);
                if ((lo >> bit) & 1) E::xyzz_add(acc, run);
            }
            E::xyzz_add(sum, acc);
        }
    }
    store_xyzz<L>(partials + ((size_t)w * chunks_per_window + t) * (4 * L), sum);
}

// one block per window: strided sums then a shared-memory tree; result -> window_sums[w]
template <class C>
__global__ void __launch_bounds__(128) msm_sum_partials_kernel(const uint32_t *__restrict__ partials, uint32_t chunks_per_window,
                                                               uint32_t *__restrict__ window_sums) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    extern __shared__ uint32_t sm[];
    const uint32_t w = blockIdx.x;
    typename E::B acc;
    E::xyzz_set_zero(acc);
    for (uint32_t t = threadIdx.x; t < chunks_per_window; t += blockDim.x) {
        typename E::B b;
        load_xyzz<L>(b, partials + ((size_t)w * chunks_per_window + t) * (4 * L));
        E::xyzz_add(acc, b);
    }
    store_xyzz<L>(sm + threadIdx.x * (4 * L), acc);
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            typename E::B a, b;
            load_xyzz<L>(a, sm + threadIdx.x * (4 * L));
            load_xyzz<L>(b, sm + (threadIdx.x + s) * (4 * L));
            E::xyzz_add(a, b);
            store_xyzz<L>(sm + threadIdx.x * (4 * L), a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        typename E::B a;
        load_xyzz<L>(a, sm);
        store_xyzz<L>(window_sums + (size_t)w * (4 * L), a);
    }
}

// total = sum_w 2^(c*w) * S_w by Horner (:489-502); Jacobian result (x, y, z) -> out (3L words)
template <class C> __global__ void msm_window_combine_kernel(const uint32_t *__restrict__ window_sums, int W, int c, uint32_t *__restrict__ out) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    typename E::J total;
    E::jac_set_zero(total);
    for (int w = W - 1; w >= 0; w--) {
        typename E::B b;
        typename E::J j;
        load_xyzz<L>(b, window_sums + (size_t)w * (4 * L));
        E::xyzz_to_jac(j, b);
        E::jac_add(total, j);  // Projective += &Bucket (bucket.rs:345-359)
        if (w > 0)
            for (int d = 0; d < c; d++) E::jac_dbl(total);
    }
    store_limbs<L>(out, total.x);
    store_limbs<L>(out + L, total.y);
    store_limbs<L>(out + 2 * L, total.z);
}

// sum of k Jacobian points (multi-GPU gather reduce), one thread
template <class C> __global__ void jac_sum_kernel(const uint32_t *__restrict__ pts, size_t k, uint32_t *__restrict__ out) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    typename E::J total;
    E::jac_set_zero(total);
    for (size_t i = 0; i < k; i++) {
        typename E::J j;
        load_limbs<L>(j.x, pts + i * 3 * L);
        load_limbs<L>(j.y, pts + i * 3 * L + L);
        load_limbs<L>(j.z, pts + i * 3 * L + 2 * L);
        E::jac_add(total, j);
    }
    if (E::jac_is_zero(total)) E::jac_set_zero(total);
    store_limbs<L>(out, total.x);
    store_limbs<L>(out + L, total.y);
    store_limbs<L>(out + 2 * L, total.z);
}
template <class C> __global__ void jac_to_affine_kernel(const uint32_t *__restrict__ pts, size_t k, uint32_t *__restrict__ out) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    typename E::J j;
    load_limbs<L>(j.x, pts + i * 3 * L);
    load_limbs<L>(j.y, pts + i * 3 * L + L);
    load_limbs<L>(j.z, pts + i * 3 * L + 2 * L);
    uint32_t ax[L], ay[L];
    E::jac_to_affine(ax, ay, j);
    store_limbs<L>(out + i * 2 * L, ax);
    store_limbs<L>(out + i * 2 * L + L, ay);
}

// ------------------------------------------------------------------------------------------------
// Pair-add, second generation: warp-interleaved slots + asynchronous operand staging.
//   * a WARP owns 32*batch consecutive output slots and lane l takes slots l, l+32, l+64, ... — every Montgomery-trick chain is
//     still private to one thread, but the 32 lanes of a load/store touch 32 consecutive slots (coalesced streaming for the
//     levels >= 2 and for the parked prefix products);
//   * the operands of the NEXT slot are fetched with cp.async (LDGSTS) into a per-thread shared-memory strip while the current
//     slot's multiplications run, so the random 96-byte gathers of level 1 no longer stall the integer pipe and cost no registers;
//   * slot -> input-pair mapping comes from `pairmap` (msm_pairmap_kernel), not from a per-thread bucket walk.
// Same arithmetic, same degenerate-pair handling and same output layout as msm_pair_add_kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// per-thread strip: element e, 16-byte chunk j of thread t lives at uint4 index (e*(L/4) + j)*128 + t (conflict-free)
template <int L> __device__ __forceinline__ void strip_fetch(uint32_t sbase, int e, const uint32_t *g) {
#pragma unroll
    for (int j = 0; j < L / 4; j++) cp_async16(sbase + ((uint32_t)((e * (L / 4) + j) * 128 + threadIdx.x) << 4), g + 4 * j);
}
template <int L> __device__ __forceinline__ void strip_read(uint32_t *r, const uint4 *sm, int e) {
#pragma unroll
    for (int j = 0; j < L / 4; j++) {
        const uint4 v = sm[(e * (L / 4) + j) * 128 + threadIdx.x];
        r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
    }
}

// pairmap[p] = (index of the first input entry of output slot p) | (the slot has a second entry) << 31.
// One block per 1024 consecutive slots: two threads locate the first and last bucket of the block's slot range, the bucket
// offsets in between are staged in shared memory and every slot finds its bucket by a short binary search there (any bucket
// size distribution costs the same; a run of > 2048 buckets inside one block — mostly empty ones — falls back to global loads).
__global__ void __launch_bounds__(256) msm_pairmap_kernel(const uint32_t *__restrict__ offsets_in, const uint32_t *__restrict__ offsets_out,
                                                          uint32_t nb, uint32_t *__restrict__ pairmap) {
    constexpr uint32_t SPAN = 2048;
    __shared__ uint32_t s_off[SPAN + 1];
    __shared__ uint32_t s_b[2];
    const uint32_t M = __ldg(offsets_out + nb);
    const uint64_t P0l = (uint64_t)blockIdx.x * 1024;
    if (P0l >= M) return;
    const uint32_t P0 = (uint32_t)P0l, P1 = (uint32_t)min((uint64_t)M, P0l + 1024) - 1;
    if (threadIdx.x < 2) {
        const uint32_t p = threadIdx.x ? P1 : P0;
        uint32_t lo = 0, hi = nb;   // offsets_out[lo] <= p < offsets_out[hi]
        while (hi - lo > 1) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (__ldg(offsets_out + mid) <= p) lo = mid; else hi = mid;
        }
        s_b[threadIdx.x] = lo;
    }
    __syncthreads();
    const uint32_t b_first = s_b[0], span = s_b[1] - s_b[0] + 1;
    const bool staged = span <= SPAN;
    if (staged)
        for (uint32_t i = threadIdx.x; i <= span; i += 256) s_off[i] = __ldg(offsets_out + b_first + i);
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < 1024; q += 256) {
        const uint32_t p = P0 + q;
        if (p > P1) break;
        uint32_t lo = 0, hi = span;   // off[lo] <= p < off[hi]
        while (hi - lo > 1) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            const uint32_t v = staged ? s_off[mid] : __ldg(offsets_out + b_first + mid);
            if (v <= p) lo = mid; else hi = mid;
        }
        const uint32_t b = b_first + lo, out_beg = staged ? s_off[lo] : __ldg(offsets_out + b);
        const uint32_t in_beg = __ldg(offsets_in + b), in_end = __ldg(offsets_in + b + 1);
        const uint32_t k = in_beg + 2 * (p - out_beg);
        pairmap[p] = k | ((k + 1 < in_end) ? 0x80000000u : 0u);
    }
}

template <class C, bool FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_pair_add2_kernel(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src,
                                                                const uint32_t *__restrict__ pairmap, const uint32_t *__restrict__ offsets_out,
                                                                uint32_t nb, uint32_t batch, uint32_t *__restrict__ out) {
    using F = typename C::F;
    constexpr int L = F::L;
    extern __shared__ uint4 strip[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(strip);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t M = __ldg(offsets_out + nb);
    const uint64_t w_lo = (uint64_t)(blockIdx.x * 4 + (threadIdx.x >> 5)) * 32 * batch;
    if (w_lo + lane >= M) return;
    const uint32_t w_hi = (uint32_t)min((uint64_t)M, w_lo + (uint64_t)32 * batch);
    const uint32_t p0 = (uint32_t)w_lo + lane;
    const uint32_t cnt = (w_hi - p0 + 31) >> 5;   // slots p0 + 32*i, i < cnt
    // address of the point behind input entry `k` (FIRST: through the sorted index, e = index | sign << 31)
    auto point = [&](uint32_t k, uint32_t e) -> const uint32_t * {
        return FIRST ? bases + (size_t)(e & 0x7fffffffu) * (2 * L) : src + (size_t)k * (2 * L);
    };
    uint32_t x1[L], y1[L], x2[L], y2[L], den[L], run[L];
    // software pipeline over the slots of this lane: (m, e1, e2) describe a slot (pairmap word and, for FIRST, the two sorted
    // entries); *_c = slot being computed, *_n = next slot (operands in flight), m_nn = pairmap word two slots ahead
    uint32_t m_c, m_n = 0, m_nn = 0, e1_c = 0, e2_c = 0, e1_n = 0, e2_n = 0;

    // ---------------- forward: running product of the denominators (x coordinates only), parked in the x-half of the slots
    m_c = __ldg(pairmap + p0);
    if (FIRST) { e1_c = __ldg(src + (m_c & 0x7fffffffu)); if (m_c >> 31) e2_c = __ldg(src + (m_c & 0x7fffffffu) + 1); }
    // (forward uses strip elements {0,2} for even slots and {1,3} for odd ones: a true double buffer)
    if (m_c >> 31) { strip_fetch<L>(sbase, 0, point(m_c & 0x7fffffffu, e1_c)); strip_fetch<L>(sbase, 2, point((m_c & 0x7fffffffu) + 1, e2_c)); }
    cp_async_commit();
    if (cnt > 1) {
        m_n = __ldg(pairmap + p0 + 32);
        if (FIRST) { e1_n = __ldg(src + (m_n & 0x7fffffffu)); if (m_n >> 31) e2_n = __ldg(src + (m_n & 0x7fffffffu) + 1); }
    }
    if (cnt > 2) m_nn = __ldg(pairmap + p0 + 64);
    F::set_one(run);
    for (uint32_t i = 0; i < cnt; i++) {
        cp_async_wait_all();
        const bool has2 = (m_c >> 31) != 0;
        const int eb = (int)(i & 1);
        if (has2) { strip_read<L>(x1, strip, eb); strip_read<L>(x2, strip, 2 + eb); }
        if (i + 1 < cnt && (m_n >> 31)) {
            strip_fetch<L>(sbase, 1 - eb, point(m_n & 0x7fffffffu, e1_n));
            strip_fetch<L>(sbase, 3 - eb, point((m_n & 0x7fffffffu) + 1, e2_n));
        }
        cp_async_commit();
        uint32_t e1_nn = 0, e2_nn = 0, m_n3 = 0;
        if (FIRST && i + 2 < cnt) { e1_nn = __ldg(src + (m_nn & 0x7fffffffu)); if (m_nn >> 31) e2_nn = __ldg(src + (m_nn & 0x7fffffffu) + 1); }
        if (i + 3 < cnt) m_n3 = __ldg(pairmap + p0 + 32 * (i + 3));
        if (has2) {
            if (limbs_is_zero<L>(x1) || limbs_is_zero<L>(x2) || limbs_eq<L>(x1, x2)) {   // rare: identity operand / equal x
                const uint32_t k = m_c & 0x7fffffffu;
                pair_load_point<F, FIRST>(x1, y1, bases, src, k);
                pair_load_point<F, FIRST>(x2, y2, bases, src, k + 1);
                const int kind = pair_classify<F>(den, x1, y1, x2, y2, true);
                if (kind >= PAIR_ADD) F::mul(run, run, den);
            } else {
                F::sub(den, x2, x1);
                F::mul(run, run, den);
            }
        }
        store_limbs<L>(out + (size_t)(p0 + 32 * i) * (2 * L), run);
        m_c = m_n; e1_c = e1_n; e2_c = e2_n;
        m_n = m_nn; e1_n = e1_nn; e2_n = e2_nn;
        m_nn = m_n3;
    }
    uint32_t inv[L];
    F::inv(inv, run);

    // ---------------- backward: peel the inverses off and write the sums; strip elements 0..3 = x1, y1, x2, y2, 4 = prefix
    auto fetch_bwd = [&](uint32_t m, uint32_t e1, uint32_t e2, uint32_t i) {   // operands of slot i and the prefix parked in slot i-1
        const uint32_t k = m & 0x7fffffffu;
        const uint32_t *a = point(k, e1);
        strip_fetch<L>(sbase, 0, a);
        strip_fetch<L>(sbase, 1, a + L);
        if (m >> 31) {
            const uint32_t *b = point(k + 1, e2);
            strip_fetch<L>(sbase, 2, b);
            strip_fetch<L>(sbase, 3, b + L);
        }
        if (i > 0) strip_fetch<L>(sbase, 4, out + (size_t)(p0 + 32 * (i - 1)) * (2 * L));
    };
    m_c = __ldg(pairmap + p0 + 32 * (cnt - 1));
    e1_c = e2_c = e1_n = e2_n = 0;
    m_n = m_nn = 0;
    if (FIRST) { e1_c = __ldg(src + (m_c & 0x7fffffffu)); if (m_c >> 31) e2_c = __ldg(src + (m_c & 0x7fffffffu) + 1); }
    fetch_bwd(m_c, e1_c, e2_c, cnt - 1);
    cp_async_commit();
    if (cnt > 1) {
        m_n = __ldg(pairmap + p0 + 32 * (cnt - 2));
        if (FIRST) { e1_n = __ldg(src + (m_n & 0x7fffffffu)); if (m_n >> 31) e2_n = __ldg(src + (m_n & 0x7fffffffu) + 1); }
    }
    if (cnt > 2) m_nn = __ldg(pairmap + p0 + 32 * (cnt - 3));
    for (uint32_t i = cnt; i-- > 0;) {
        cp_async_wait_all();
        const bool has2 = (m_c >> 31) != 0;
        uint32_t dinv[L];
        strip_read<L>(x1, strip, 0);
        strip_read<L>(y1, strip, 1);
        if (has2) { strip_read<L>(x2, strip, 2); strip_read<L>(y2, strip, 3); }
        if (i > 0) { strip_read<L>(dinv, strip, 4); F::mul(dinv, inv, dinv); }   // inv * prefix_{i-1} = 1/den_i
        else limbs_copy<L>(dinv, inv);
        if (FIRST) {
            F::cneg(y1, y1, (e1_c >> 31) != 0);
            if (has2) F::cneg(y2, y2, (e2_c >> 31) != 0);
        }
        if (i > 0) fetch_bwd(m_n, e1_n, e2_n, i - 1);
        cp_async_commit();
        uint32_t e1_nn = 0, e2_nn = 0, m_n3 = 0;
        if (FIRST && i >= 2) { e1_nn = __ldg(src + (m_nn & 0x7fffffffu)); if (m_nn >> 31) e2_nn = __ldg(src + (m_nn & 0x7fffffffu) + 1); }
        if (i >= 3) m_n3 = __ldg(pairmap + p0 + 32 * (i - 3));
        const int kind = pair_classify<F>(den, x1, y1, x2, y2, has2);
        uint32_t *o = out + (size_t)(p0 + 32 * i) * (2 * L);
        if (kind >= PAIR_ADD) {
            uint32_t lam[L], t3[L];
            F::mul(inv, inv, den);
            if (kind == PAIR_ADD) {
                F::sub(lam, y2, y1);
            } else {                      // doubling: slope = 3 x^2 / (2 y)
                F::sqr(lam, x1);
                F::dbl(t3, lam);
                F::add(lam, lam, t3);
                limbs_copy<L>(x2, x1);
            }
            F::mul(lam, lam, dinv);
            F::sqr(t3, lam);
            F::sub(t3, t3, x1);
            F::sub(t3, t3, x2);           // x3
            F::sub(x2, x1, t3);
            F::mul(x2, lam, x2);
            F::sub(x2, x2, y1);           // y3 = lam (x1 - x3) - y1
            store_limbs<L>(o, t3);
            store_limbs<L>(o + L, x2);
        } else if (kind == PAIR_PASS1) {
            store_limbs<L>(o, x1);
            store_limbs<L>(o + L, y1);
        } else if (kind == PAIR_PASS2) {
            store_limbs<L>(o, x2);
            store_limbs<L>(o + L, y2);
        } else {
            F::set_zero(x1);
            store_limbs<L>(o, x1);
            store_limbs<L>(o + L, x1);
        }
        m_c = m_n; e1_c = e1_n; e2_c = e2_n;
        m_n = m_nn; e1_n = e1_nn; e2_n = e2_nn;
        m_nn = m_n3;
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
// Options of a call.  Defaults come from the per-thread setters (b200_set_msm_window / _affine_levels) and, for the tuning
// knobs without a public setter, from environment variables read ONCE per process (never on the call path).
struct EnvKnobs {
    int pair_variant = 2;                 // 1 = thread-contiguous batches, 2 = warp-interleaved + cp.async staging
    double level_budget_bytes = 24e9;     // scratch allowed for the affine level arrays (window groups are sized to fit)
    int force_chunks = 0;                 // test hook: split device-resident inputs into this many chunks
    int l2_fetch_granularity = 0;         // cudaLimitMaxL2FetchGranularity during the MSM (0 = leave alone)
    EnvKnobs() {
        if (const char *e = getenv("B200_MSM_PAIR_VARIANT")) pair_variant = atoi(e) == 1 ? 1 : 2;
        if (const char *e = getenv("B200_MSM_LEVEL_BUDGET_GB")) { double v = atof(e); if (v > 0.01) level_budget_bytes = v * 1e9; }
        if (const char *e = getenv("B200_MSM_FORCE_CHUNKS")) force_chunks = atoi(e);
        if (const char *e = getenv("B200_L2_FETCH_GRANULARITY")) l2_fetch_granularity = atoi(e);
    }
};
static const EnvKnobs &env_knobs() {
    static const EnvKnobs k;
    return k;
}

static thread_local int t_window_override = 0;
static thread_local int t_affine_levels = -1;   // batched-affine pre-reduction levels; -1 = automatic
struct MsmTimings {
    float ms[7] = {0, 0, 0, 0, 0, 0, 0};
    int c = 0, W = 0;
    unsigned long long bucket_adds = 0;
};
static thread_local MsmTimings t_last;

int msm_set_affine_levels(int levels) {
    if (levels < -1 || levels > 6) { set_last_error("affine levels must be in [-1, 6]"); return B200_EINVAL; }
    t_affine_levels = levels;
    return 0;
}

int msm_get_window() { return t_window_override; }
int msm_get_affine_levels() { return t_affine_levels; }
int msm_set_window(int c) {
    if (c < 0 || c > 24) { set_last_error("window size must be in [1,24] (0 = automatic)"); return B200_EINVAL; }
    t_window_override = c;
    return 0;
}

// Window choice: a time model fitted to the B200 sweeps committed in profiles/ (n = 2^26 and 2^23, BLS12-381):
//   accumulation 0.33 ns per (point, window) with the batched-affine levels, reduction 2.6 ns per bucket,
//   scatter + histogram 0.02 ns per entry, + contention when the top window has fewer than ~2^10 buckets.
int msm_auto_window(size_t n, int scalar_bits) {
    if (n < 32) return 3;  // same floor as the reference (:445-449)
    double best = 1e300;
    int best_c = 3;
    for (int c = 4; c <= 23; c++) {
        MsmGeom g = make_geom(c, scalar_bits);
        if ((double)g.total_buckets * 192.0 > 24e9) continue;
        const double entries = (double)n * g.W;
        double t = 0.33 * entries + 2.6 * (double)g.total_buckets + 0.02 * entries;
        if (g.top_bits < 10) t += 0.15 * (double)n;   // hot top-window buckets serialise the atomics
        t += 2000.0 * g.W;                             // per-window fixed costs (reduction tree, combine doublings)
        if (t < best) { best = t; best_c = c; }
    }
    return best_c;
}

// buckets[b] += extra[b]  (chunked paths: every chunk after the first accumulates into `extra`)
template <class C>
__global__ void __launch_bounds__(128) msm_merge_kernel(uint32_t *__restrict__ buckets, const uint32_t *__restrict__ extra, uint32_t total_buckets) {
    using F = typename C::F;
    using E = Ec<F>;
    constexpr int L = F::L;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    typename E::B x, y;
    load_xyzz<L>(y, extra + (size_t)b * (4 * L));
    if (E::xyzz_is_zero(y)) return;
    load_xyzz<L>(x, buckets + (size_t)b * (4 * L));
    E::xyzz_add(x, y);
    store_xyzz<L>(buckets + (size_t)b * (4 * L), x);
}

static int scalar_kind_bits(int kind, int field_bits) {
    switch (kind) {
        case B200_SCALARS_U8: return 8;
        case B200_SCALARS_U16: return 16;
        case B200_SCALARS_U32: return 32;
        case B200_SCALARS_U64: return 64;
        default: return field_bits;  // Fr::MODULUS_BIT_SIZE (variable_base/mod.rs:451)
    }
}
size_t scalar_kind_bytes(int kind) {
    switch (kind) {
        case B200_SCALARS_U8: return 1;
        case B200_SCALARS_U16: return 2;
        case B200_SCALARS_U32: return 4;
        case B200_SCALARS_U64: return 8;
        default: return 32;
    }
}

// stream-ordered scratch and events owned by a session: everything is released on every exit path (cudaFreeAsync on the
// session's stream / cudaEventDestroy), including the error returns of AB_CUDA / AB_LAUNCHED.
struct StreamArena {
    cudaStream_t st = 0;
    std::vector<void *> ptrs;
    template <class T> int alloc(T **p, size_t bytes) {
        void *q = nullptr;
        AB_CUDA(cudaMallocAsync(&q, bytes ? bytes : 16, st));
        ptrs.push_back(q);
        *p = (T *)q;
        return 0;
    }
    void release(void *p) {
        if (!p) return;
        for (auto &q : ptrs)
            if (q == p) { cudaFreeAsync(q, st); q = nullptr; return; }
    }
    ~StreamArena() {
        for (void *q : ptrs)
            if (q) cudaFreeAsync(q, st);
    }
};
struct EventSet {
    std::vector<cudaEvent_t> ev;
    int make(cudaEvent_t *e) {
        AB_CUDA(cudaEventCreate(e));
        ev.push_back(*e);
        return 0;
    }
    ~EventSet() {
        for (auto e : ev) cudaEventDestroy(e);
    }
};

static int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

struct MsmSessionBase {
    virtual ~MsmSessionBase() {}
    // n_total: expected total number of pairs (window choice); max_chunk: largest chunk handed to add_chunk
    virtual int begin(size_t n_total, size_t max_chunk, int kind, cudaStream_t st) = 0;
    // digit-sort + accumulate one chunk into the session's buckets; `ready` (may be null) is waited on first.  Asynchronous.
    virtual int add_chunk(const void *d_bases, const void *d_scalars, size_t nk, cudaEvent_t ready) = 0;
    // bucket reduction + window combine -> d_out (3 coordinates, Jacobian).  Asynchronous on the session's stream.
    virtual int finish(void *d_out) = 0;
    // after the stream has been synchronised: fold the per-phase events into the thread's "last timings"
    virtual int collect_timings() = 0;
    virtual int coord_words() const = 0;   // 32-bit words per coordinate (12 / 8 / 24)
};

template <class C> struct MsmSession final : MsmSessionBase {
    using F = typename C::F;
    static constexpr int L = F::L;
    MsmGeom g;
    int kind = 0, levels_opt = -1, chunks_done = 0;
    size_t scalar_bytes = 32, n_seen = 0, max_chunk = 0;
    cudaStream_t st = 0;
    StreamArena arena;
    EventSet events;
    std::vector<cudaEvent_t> chunk_ev;   // 5 per chunk
    cudaEvent_t e_begin = nullptr, e_acc_done = nullptr, e_red = nullptr, e_end = nullptr;
    uint32_t *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *sorted = nullptr, *block_totals = nullptr;
    uint32_t *buckets = nullptr, *extra = nullptr;
    size_t scan_blocks = 0;
    int restore_l2_gran = -1;

    int coord_words() const override { return L; }

    ~MsmSession() override {
        if (restore_l2_gran >= 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)restore_l2_gran);
    }

    int begin(size_t n_total, size_t max_chunk_, int kind_, cudaStream_t st_) override {
        st = st_;
        arena.st = st_;
        kind = kind_;
        max_chunk = max_chunk_;
        levels_opt = t_affine_levels;
        if (max_chunk >= ((size_t)1 << 31)) { set_last_error("chunk must be < 2^31 pairs"); return B200_ETOOLARGE; }
        const int scalar_bits = scalar_kind_bits(kind, C::SCALAR_BITS);
        scalar_bytes = scalar_kind_bytes(kind);
        int c = t_window_override ? t_window_override : msm_auto_window(std::max(n_total, max_chunk), scalar_bits);
        if (c > scalar_bits) c = scalar_bits;
        g = make_geom(c, scalar_bits);
        const size_t nb_total = g.total_buckets;
        const size_t max_entries = max_chunk * (size_t)g.W;
        if (max_entries >= ((size_t)1 << 31)) { set_last_error("chunk pairs * windows must be < 2^31"); return B200_ETOOLARGE; }
        if (env_knobs().l2_fetch_granularity > 0) {
            size_t cur = 0;
            if (cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity) == cudaSuccess) {
                restore_l2_gran = (int)cur;
                cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)env_knobs().l2_fetch_granularity);
            }
        }
        scan_blocks = (nb_total + kScanBlock - 1) / kScanBlock;
        if (int rc = arena.alloc(&counts, nb_total * 4)) return rc;
        if (int rc = arena.alloc(&offsets, (nb_total + 1) * 4)) return rc;
        if (int rc = arena.alloc(&cursor, nb_total * 4)) return rc;
        if (int rc = arena.alloc(&sorted, std::max<size_t>(max_entries, 1) * 4)) return rc;
        if (int rc = arena.alloc(&block_totals, std::max<size_t>(scan_blocks, 1) * 4)) return rc;
        if (int rc = arena.alloc(&buckets, nb_total * 4 * L * 4)) return rc;
        if (int rc = events.make(&e_begin)) return rc;
        if (int rc = events.make(&e_acc_done)) return rc;
        if (int rc = events.make(&e_red)) return rc;
        if (int rc = events.make(&e_end)) return rc;
        AB_CUDA(cudaEventRecord(e_begin, st));
        return 0;
    }

    int scan(const uint32_t *in, size_t n, uint32_t *out) {   // out[0..n] = exclusive prefix sums of in[0..n)
        const size_t blocks = (n + kScanBlock - 1) / kScanBlock;
        scan_block_totals_kernel<<<(unsigned)blocks, kScanThreads, 0, st>>>(in, n, block_totals);
        AB_LAUNCHED();
        scan_totals_kernel<<<1, kScanThreads, 0, st>>>(block_totals, blocks);
        AB_LAUNCHED();
        scan_apply_kernel<<<(unsigned)blocks, kScanThreads, 0, st>>>(in, n, block_totals, out);
        AB_LAUNCHED();
        return 0;
    }

    // XYZZ accumulation of `entries` bucket-ordered entries (bases gathered through `sorted_idx`, or direct affine points)
    // over the buckets [0, nbk) described by `offs`, written to `dst` (nbk buckets).
    int accumulate(const uint32_t *pts, const uint32_t *sorted_idx, const uint32_t *offs, size_t nbk, size_t entries, uint32_t *dst) {
        if (!entries) return 0;
        // balanced tasks: T entries per thread such that the grid is a whole number of waves (no tail), T <= 512
        int blocks_per_sm = 3;
        if (sorted_idx) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, msm_accumulate_kernel<C, false>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, msm_accumulate_kernel<C, true>, 128, 0);
        const double wave = (double)sm_count() * std::max(blocks_per_sm, 1) * 128.0;
        const double per_thread = (double)entries / wave;
        uint32_t T;
        if (per_thread <= 8.0) T = 8;
        else {
            const double waves = std::ceil(per_thread / 512.0);
            T = (uint32_t)std::ceil((double)entries / (waves * wave));
            T = std::min(512u, std::max(8u, T));
        }
        const uint32_t num_tasks = (uint32_t)((entries + T - 1) / T);
        uint32_t *head = nullptr, *tail = nullptr, *head_bucket = nullptr, *tail_bucket = nullptr;
        if (int rc = arena.alloc(&head, (size_t)num_tasks * 4 * L * 4)) return rc;
        if (int rc = arena.alloc(&tail, (size_t)num_tasks * 4 * L * 4)) return rc;
        if (int rc = arena.alloc(&head_bucket, (size_t)num_tasks * 4)) return rc;
        if (int rc = arena.alloc(&tail_bucket, (size_t)num_tasks * 4)) return rc;
        const unsigned grid = (num_tasks + 127) / 128;
        if (sorted_idx)
            msm_accumulate_kernel<C, false><<<grid, 128, 0, st>>>(pts, sorted_idx, offs, (uint32_t)nbk, T, dst, head, tail, head_bucket, tail_bucket, num_tasks);
        else
            msm_accumulate_kernel<C, true><<<grid, 128, 0, st>>>(pts, nullptr, offs, (uint32_t)nbk, T, dst, head, tail, head_bucket, tail_bucket, num_tasks);
        AB_LAUNCHED();
        msm_fixup_small_kernel<C><<<grid, 128, 0, st>>>(offs, T, head, tail, tail_bucket, num_tasks, dst);
        AB_LAUNCHED();
        msm_fixup_big_kernel<C><<<num_tasks, 128, 128 * 4 * L * 4, st>>>(offs, T, head, tail, tail_bucket, num_tasks, dst);
        AB_LAUNCHED();
        arena.release(head);
        arena.release(tail);
        arena.release(head_bucket);
        arena.release(tail_bucket);
        return 0;
    }

    // batched-affine levels + XYZZ accumulation of the buckets [b0, b1) (a group of whole windows) of the current chunk
    int reduce_group(const uint32_t *bas, size_t b0, size_t b1, size_t group_entries, int levels, uint32_t *target) {
        const size_t nbg = b1 - b0;
        const bool forced = levels_opt >= 0;
        const int variant = env_knobs().pair_variant;
        size_t cur_entries = group_entries;   // upper bound on the entries of the current level
        const uint32_t *cur_src = sorted, *cur_offsets = offsets + b0;
        uint32_t *lvl_pts[2] = {nullptr, nullptr}, *lvl_off[2] = {nullptr, nullptr};
        int levels_done = 0;
        const double lanes_per_wave = (double)sm_count() * C::PAIR_MINB * 128;   // resident blocks of 128 threads
        for (int lv = 0; lv < levels; lv++) {
            // sum_b ceil(cnt_b / 2) <= min((entries + non-empty buckets) / 2, entries)
            const size_t out_cap = std::min((cur_entries + nbg) / 2 + 1, cur_entries);
            // one inversion (~570 modmuls) per `batch` additions.  The grid is sized to a whole number of waves of resident
            // threads so that no SM idles through a partial last wave; in automatic mode a level that cannot give every
            // resident thread >= 256 slots is left to the XYZZ kernel instead (measured: thin levels are slower).
            const double per_lane = (double)out_cap / lanes_per_wave;
            if (!forced && per_lane < 256.0) break;
            const double waves = std::max(1.0, std::ceil(per_lane / 1024.0));
            uint32_t batch = (uint32_t)std::ceil((double)out_cap / (waves * lanes_per_wave));
            batch = std::min(1024u, std::max(forced ? 8u : 256u, batch));
            uint32_t *pts = nullptr, *off2 = nullptr;
            if (int rc = arena.alloc(&pts, out_cap * 2 * L * 4)) return rc;
            if (int rc = arena.alloc(&off2, (nbg + 1) * 4)) return rc;
            msm_halve_counts_kernel<<<(unsigned)((nbg + 255) / 256), 256, 0, st>>>(cur_offsets, (uint32_t)nbg, counts);
            AB_LAUNCHED();
            if (int rc = scan(counts, nbg, off2)) return rc;
            if (variant == 2) {
                uint32_t *pairmap = nullptr;
                if (int rc = arena.alloc(&pairmap, out_cap * 4)) return rc;
                msm_pairmap_kernel<<<(unsigned)((out_cap + 1023) / 1024), 256, 0, st>>>(cur_offsets, off2, (uint32_t)nbg, pairmap);
                AB_LAUNCHED();
                const size_t warps = (out_cap + (size_t)32 * batch - 1) / ((size_t)32 * batch);
                const unsigned pg = (unsigned)((warps + 3) / 4);
                const size_t smem = (size_t)5 * (L / 4) * 128 * 16;
                if (lv == 0) {
                    AB_CUDA(cudaFuncSetAttribute(msm_pair_add2_kernel<C, true, C::PAIR_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    msm_pair_add2_kernel<C, true, C::PAIR_MINB><<<pg, 128, smem, st>>>(bas, cur_src, pairmap, off2, (uint32_t)nbg, batch, pts);
                } else {
                    AB_CUDA(cudaFuncSetAttribute(msm_pair_add2_kernel<C, false, C::PAIR_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    msm_pair_add2_kernel<C, false, C::PAIR_MINB><<<pg, 128, smem, st>>>(bas, cur_src, pairmap, off2, (uint32_t)nbg, batch, pts);
                }
                AB_LAUNCHED();
                arena.release(pairmap);
            } else {
                const uint32_t nthreads = (uint32_t)((out_cap + batch - 1) / batch);
                const unsigned pg = (nthreads + 127) / 128;
                if (lv == 0) msm_pair_add_kernel<C, true, C::PAIR_MINB><<<pg, 128, 0, st>>>(bas, cur_src, cur_offsets, off2, (uint32_t)nbg, batch, pts, nthreads);
                else msm_pair_add_kernel<C, false, C::PAIR_MINB><<<pg, 128, 0, st>>>(bas, cur_src, cur_offsets, off2, (uint32_t)nbg, batch, pts, nthreads);
                AB_LAUNCHED();
            }
            // the level before the previous one is no longer read
            arena.release(lvl_pts[lv & 1]);
            arena.release(lvl_off[lv & 1]);
            lvl_pts[lv & 1] = pts;
            lvl_off[lv & 1] = off2;
            cur_src = pts;
            cur_offsets = off2;
            cur_entries = out_cap;
            levels_done++;
        }
        int rc;
        if (levels_done) rc = accumulate(cur_src, nullptr, cur_offsets, nbg, cur_entries, target + b0 * 4 * L);
        else rc = accumulate(bas, sorted, offsets + b0, nbg, group_entries, target + b0 * 4 * L);   // no level ran: gather through the sorted indices
        for (int q = 0; q < 2; q++) { arena.release(lvl_pts[q]); arena.release(lvl_off[q]); }
        return rc;
    }

    int add_chunk(const void *d_bases_v, const void *d_scalars, size_t nk, cudaEvent_t ready) override {
        if (nk > max_chunk) { set_last_error("chunk larger than announced at begin"); return B200_EINVAL; }
        const uint32_t *bas = (const uint32_t *)d_bases_v;
        const size_t nb_total = g.total_buckets;
        cudaEvent_t e[5];
        for (int i = 0; i < 5; i++) {
            if (int rc = events.make(&e[i])) return rc;
            chunk_ev.push_back(e[i]);
        }
        if (ready) AB_CUDA(cudaStreamWaitEvent(st, ready, 0));
        uint32_t *target = buckets;
        if (chunks_done > 0) {
            if (!extra)
                if (int rc = arena.alloc(&extra, nb_total * 4 * L * 4)) return rc;
            target = extra;
        }
        AB_CUDA(cudaEventRecord(e[0], st));
        AB_CUDA(cudaMemsetAsync(counts, 0, nb_total * 4, st));
        const unsigned dblocks = (unsigned)((nk + 255) / 256);
        if (nk) {
            msm_digits_kernel<C, 0><<<dblocks, 256, 0, st>>>(d_scalars, kind, nk, g, 0, g.W, counts, nullptr);
            AB_LAUNCHED();
        }
        AB_CUDA(cudaEventRecord(e[1], st));
        if (int rc = scan(counts, nb_total, offsets)) return rc;
        AB_CUDA(cudaMemcpyAsync(cursor, offsets, nb_total * 4, cudaMemcpyDeviceToDevice, st));
        AB_CUDA(cudaEventRecord(e[2], st));
        // scatter in groups of windows so that the active write fronts (one 32-byte sector per bucket of the group) stay
        // L2-resident: random 4-byte stores into a multi-GB array otherwise cost a DRAM sector each (measured 2.5x slower)
        if (nk) {
            const size_t front_bytes = (size_t)g.nb * 32;
            const int group = (int)std::max<size_t>(1, ((size_t)48 << 20) / front_bytes);
            for (int w0 = 0; w0 < g.W; w0 += group) {
                msm_digits_kernel<C, 1><<<dblocks, 256, 0, st>>>(d_scalars, kind, nk, g, w0, std::min(g.W, w0 + group), cursor, sorted);
                AB_LAUNCHED();
            }
        }
        AB_CUDA(cudaEventRecord(e[3], st));
        AB_CUDA(cudaMemsetAsync(target, 0, nb_total * 4 * L * 4, st));
        if (nk) {
            // batched-affine levels: automatic = keep halving while buckets still hold >= ~8 entries, at most 4 levels
            // (measured at 2^26: 339 / 324 / 310 / 303 / 300 ms of accumulation for 0..4 levels)
            int levels = levels_opt;
            if (levels < 0) {
                levels = 0;
                if (C::AUTO_LEVELS)
                    for (double l = (double)nk / (double)g.nb; l >= 16.0 && levels < 4; l *= 0.5) levels++;
                // the first level must be able to give every resident thread a batch of >= 256 additions
                if ((double)nk * g.W * 0.5 / ((double)sm_count() * 512.0) < 256.0) levels = 0;
            }
            if (levels == 0) {
                if (int rc = accumulate(bas, sorted, offsets, nb_total, nk * (size_t)g.W, target)) return rc;
            } else {
                // window groups: the level arrays of one group (level 1: entries/2 affine points, level 2: half of that, two
                // levels alive at a time) must fit the scratch budget; groups are whole windows, balanced in size
                const double per_window = ((double)nk * 0.5 + (double)g.nb * 0.5) * 2 * L * 4 * 1.5 + (double)nk * 0.5 * 4;
                int gw = (int)std::floor(env_knobs().level_budget_bytes / per_window);
                // ... and a group must give every resident thread a batch of >= 256 additions at level 1
                const int gw_min = (int)std::ceil(256.0 * (double)sm_count() * 512.0 / ((double)nk * 0.5));
                gw = std::max(std::max(1, gw_min), std::min(g.W, gw));
                gw = std::min(g.W, gw);
                const int ngroups = (g.W + gw - 1) / gw;
                gw = (g.W + ngroups - 1) / ngroups;
                for (int w0 = 0; w0 < g.W; w0 += gw) {
                    const int w1 = std::min(g.W, w0 + gw);
                    const size_t b0 = (size_t)w0 * g.nb, b1 = (w1 == g.W) ? nb_total : (size_t)w1 * g.nb;
                    if (int rc = reduce_group(bas, b0, b1, nk * (size_t)(w1 - w0), levels, target)) return rc;
                }
            }
        }
        if (chunks_done > 0) {
            msm_merge_kernel<C><<<(unsigned)((nb_total + 127) / 128), 128, 0, st>>>(buckets, extra, (uint32_t)nb_total);
            AB_LAUNCHED();
        }
        AB_CUDA(cudaEventRecord(e[4], st));
        chunks_done++;
        n_seen += nk;
        return 0;
    }

    int finish(void *d_out_v) override {
        uint32_t *d_out = (uint32_t *)d_out_v;
        const size_t nb_total = g.total_buckets;
        if (chunks_done == 0) AB_CUDA(cudaMemsetAsync(buckets, 0, nb_total * 4 * L * 4, st));
        AB_CUDA(cudaEventRecord(e_acc_done, st));
        // reduction geometry: chunk of m = 2^log_m buckets per thread
        int log_m = 5;
        while (log_m > 0 && (g.nb >> log_m) < 64) log_m--;
        const uint32_t max_nb = std::max(g.nb, g.nb_top);
        const uint32_t chunks = (max_nb + (1u << log_m) - 1) >> log_m;
        uint32_t *partials = nullptr, *window_sums = nullptr;
        if (int rc = arena.alloc(&partials, (size_t)g.W * chunks * 4 * L * 4)) return rc;
        if (int rc = arena.alloc(&window_sums, (size_t)g.W * 4 * L * 4)) return rc;
        const unsigned rthreads = (unsigned)g.W * chunks;
        msm_bucket_reduce_kernel<C><<<(rthreads + 127) / 128, 128, 0, st>>>(buckets, g, log_m, chunks, partials);
        AB_LAUNCHED();
        msm_sum_partials_kernel<C><<<g.W, 128, 128 * 4 * L * 4, st>>>(partials, chunks, window_sums);
        AB_LAUNCHED();
        AB_CUDA(cudaEventRecord(e_red, st));
        msm_window_combine_kernel<C><<<1, 32, 0, st>>>(window_sums, g.W, g.c, d_out);
        AB_LAUNCHED();
        AB_CUDA(cudaEventRecord(e_end, st));
        return 0;
    }

    int collect_timings() override {
        for (int i = 0; i < 7; i++) t_last.ms[i] = 0.f;
        for (size_t k = 0; k + 4 < chunk_ev.size(); k += 5)
            for (int i = 0; i < 4; i++) {
                float ms = 0.f;
                AB_CUDA(cudaEventElapsedTime(&ms, chunk_ev[k + i], chunk_ev[k + i + 1]));
                t_last.ms[i] += ms;  // digits+hist, scan, scatter, accumulate(+fixups, merge)
            }
        AB_CUDA(cudaEventElapsedTime(&t_last.ms[4], e_acc_done, e_red));
        AB_CUDA(cudaEventElapsedTime(&t_last.ms[5], e_red, e_end));
        AB_CUDA(cudaEventElapsedTime(&t_last.ms[6], e_begin, e_end));  // includes waiting for transfers on the host paths
        t_last.c = g.c;
        t_last.W = g.W;
        t_last.bucket_adds = (unsigned long long)n_seen * g.W;  // upper bound: zero digits are skipped
        return 0;
    }
};

MsmSessionBase *msm_session_create(int curve) {
    switch (curve) {
        case B200_CURVE_BLS12_381: return new MsmSession<CurveBls>();
        case B200_CURVE_BN254: return new MsmSession<CurveBn>();
        case B200_CURVE_BLS12_381_G2: return new MsmSession<CurveBlsG2>();
    }
    return nullptr;
}
int msm_coord_words(int curve) {
    switch (curve) {
        case B200_CURVE_BLS12_381: return 12;
        case B200_CURVE_BN254: return 8;
        case B200_CURVE_BLS12_381_G2: return 24;
    }
    return 0;
}
static uint32_t curve_one_limb(int curve, int i) {   // Montgomery ONE of the coordinate field, 32-bit word i
    switch (curve) {
        case B200_CURVE_BLS12_381: return BlsFq::ONE(i);
        case B200_CURVE_BN254: return BnFq::ONE(i);
        default: return i < 12 ? BlsFq::ONE(i) : 0u;   // Fq2 one = (1, 0)
    }
}
// Projective::zero() = (1,1,0) (group.rs:142-158) as 3 coordinates of L words, written as u64 limbs
void msm_write_zero(int curve, uint64_t *out_xyz) {
    const int L = msm_coord_words(curve);
    for (int i = 0; i < 3 * L / 2; i++) out_xyz[i] = 0;
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < L; i++) out_xyz[(k * L + i) / 2] |= (uint64_t)curve_one_limb(curve, i) << (32 * ((k * L + i) & 1));
}

// One MSM over K chunks of device-resident (or in-flight: `ready` events) inputs.  Each chunk is digit-sorted and accumulated
// as soon as its event fires, so the host paths overlap the PCIe transfer of chunk k+1 with the arithmetic of chunk k; the
// bucket reduction and window combine run once at the end.
int msm_dispatch(int curve, int kind, const void *d_bases, const void *d_scalars, size_t n, uint64_t *out_xyz_host, cudaStream_t st, int K,
                 const size_t *chunk_off, const cudaEvent_t *ready) {
    if (kind < B200_SCALARS_FR_MONT || kind > B200_SCALARS_U64) { set_last_error("unknown scalar kind"); return B200_EINVAL; }
    if (!out_xyz_host || (n && (!d_bases || !d_scalars))) { set_last_error("null pointer"); return B200_EINVAL; }
    const int L = msm_coord_words(curve);
    if (!L) { set_last_error("unknown curve id"); return B200_EINVAL; }
    if (n == 0) { msm_write_zero(curve, out_xyz_host); return 0; }
    if (n >= ((size_t)1 << 31)) { set_last_error("n must be < 2^31"); return B200_ETOOLARGE; }
    // device-resident inputs too large for one 31-bit entry index space (n * windows >= 2^31): split into equal chunks
    std::vector<size_t> auto_off;
    const int force = env_knobs().force_chunks;   // test hook: exercise the chunked device path at small n
    if (K <= 1) {
        int kk = 1;
        if (n > ((size_t)1 << 27)) kk = (int)((n + ((size_t)1 << 27) - 1) >> 27);
        else if (force > 1 && n >= 64) kk = force;
        for (int k = 0; k <= kk; k++) auto_off.push_back(n * (size_t)k / kk);
        K = kk;
        chunk_off = auto_off.data();
        ready = nullptr;
    }
    size_t max_chunk = 0;
    for (int k = 0; k < K; k++) max_chunk = std::max(max_chunk, chunk_off[k + 1] - chunk_off[k]);
    std::unique_ptr<MsmSessionBase> s(msm_session_create(curve));
    const size_t sb = scalar_kind_bytes(kind);
    int rc = s->begin(n, max_chunk, kind, st);
    for (int k = 0; k < K && !rc; k++)
        rc = s->add_chunk((const char *)d_bases + chunk_off[k] * (size_t)(2 * L * 4), (const char *)d_scalars + chunk_off[k] * sb,
                          chunk_off[k + 1] - chunk_off[k], ready ? ready[k] : nullptr);
    uint32_t *d_out = nullptr;
    if (!rc) {
        cudaError_t e = cudaMallocAsync(&d_out, 3 * L * 4, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    }
    if (!rc) rc = s->finish(d_out);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out_xyz_host, d_out, 3 * L * 4, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H", __FILE__, __LINE__);
    }
    if (d_out) cudaFreeAsync(d_out, st);
    cudaError_t e = cudaStreamSynchronize(st);   // also on the error paths: nothing of this call is in flight when it returns
    if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    if (!rc) rc = s->collect_timings();
    return rc;   // ~MsmSession frees the scratch and the events
}

int msm_last_timings(float *ms7, int *c, int *windows, unsigned long long *bucket_adds) {
    if (ms7) for (int i = 0; i < 7; i++) ms7[i] = t_last.ms[i];
    if (c) *c = t_last.c;
    if (windows) *windows = t_last.W;
    if (bucket_adds) *bucket_adds = t_last.bucket_adds;
    return 0;
}

template <class C> static void launch_sum(bool to_affine, const uint32_t *d_in, size_t k, uint32_t *d_out, cudaStream_t st) {
    if (to_affine) jac_to_affine_kernel<C><<<(unsigned)((k + 31) / 32), 32, 0, st>>>(d_in, k, d_out);
    else jac_sum_kernel<C><<<1, 32, 0, st>>>(d_in, k, d_out);
}
int g1_sum_dispatch(int curve, const uint64_t *pts_host, size_t k, uint64_t *out_host, bool to_affine) {
    const int L = msm_coord_words(curve);
    if (!L) { set_last_error("unknown curve id"); return B200_EINVAL; }
    if (!pts_host || !out_host) { set_last_error("null pointer"); return B200_EINVAL; }
    const size_t in_words = k * 3 * L, out_words = to_affine ? k * 2 * L : 3 * L;
    cudaStream_t st = 0;
    StreamArena arena;
    arena.st = st;
    uint32_t *d_in = nullptr, *d_out = nullptr;
    if (int rc = arena.alloc(&d_in, std::max<size_t>(in_words, 1) * 4)) return rc;
    if (int rc = arena.alloc(&d_out, std::max<size_t>(out_words, 1) * 4)) return rc;
    AB_CUDA(cudaMemcpyAsync(d_in, pts_host, in_words * 4, cudaMemcpyHostToDevice, st));
    if (curve == B200_CURVE_BLS12_381) launch_sum<CurveBls>(to_affine, d_in, k, d_out, st);
    else if (curve == B200_CURVE_BN254) launch_sum<CurveBn>(to_affine, d_in, k, d_out, st);
    else launch_sum<CurveBlsG2>(to_affine, d_in, k, d_out, st);
    AB_LAUNCHED();
    AB_CUDA(cudaMemcpyAsync(out_host, d_out, out_words * 4, cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace ab200
