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
#include <vector>

#include "common.cuh"
#include "ec.cuh"

namespace ab200 {

struct CurveBls {
    using Fq = BlsFq;
    // Hot-loop field type.  BlsFqRolled (row pairs in a real loop, 4.6k instead of 7.6k SASS instructions) was measured
    // SLOWER on B200 (accumulate 338 -> 371 ms @2^26): the unrolled rows let ptxas interleave six carry chains.
    using FqAcc = BlsFq;
    using Fr = BlsFr;
    static constexpr int SCALAR_BITS = 255;  // Fr::MODULUS_BIT_SIZE (variable_base/mod.rs:451)
};
struct CurveBn {
    using Fq = BnFq;
    using FqAcc = BnFq;
    using Fr = BnFr;
    static constexpr int SCALAR_BITS = 254;
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
                sorted[pos] = (uint32_t)i | (neg << 31);
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

template <class P> __device__ __forceinline__ void store_xyzz(uint32_t *p, const Xyzz<P> &b);
template <class P> __device__ __forceinline__ void load_xyzz(Xyzz<P> &b, const uint32_t *p);

// DIRECT = false: entry p is `sorted[p]` = (base index | sign<<31), gathered from `bases`;
// DIRECT = true : entry p is the affine point stored at bases[p] (output of the batched-affine pre-reduction), no sign.
template <class C, bool DIRECT>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint32_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                             const uint32_t *__restrict__ offsets, uint32_t total_buckets, uint32_t T,
                                                             uint32_t *__restrict__ buckets, uint32_t *__restrict__ head,
                                                             uint32_t *__restrict__ tail, uint32_t *__restrict__ head_bucket,
                                                             uint32_t *__restrict__ tail_bucket, uint32_t num_tasks) {
    using P = typename C::FqAcc;
    using E = Ec<P>;
    constexpr int L = P::L;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tasks) return;
    const uint32_t M = __ldg(offsets + total_buckets);
    const uint64_t lo64 = (uint64_t)t * T;
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
            if (started_before) { store_xyzz<P>(head + (size_t)t * (4 * L), acc); hb = b; }
            else store_xyzz<P>(buckets + (size_t)b * (4 * L), acc);
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
        if (started_before) { store_xyzz<P>(head + (size_t)t * (4 * L), acc); hb = b; }
        else store_xyzz<P>(buckets + (size_t)b * (4 * L), acc);
    } else if (started_before) {  // the whole slice is an inner piece of one bucket
        store_xyzz<P>(head + (size_t)t * (4 * L), acc);
        hb = b;
    } else {
        store_xyzz<P>(tail + (size_t)t * (4 * L), acc);
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

template <class P, bool FIRST>
__device__ __forceinline__ void pair_load_point(uint32_t *x, uint32_t *y, const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src,
                                                uint32_t k) {
    constexpr int L = P::L;
    if (FIRST) {
        const uint32_t e = __ldg(src + k);
        const uint32_t *bp = bases + (size_t)(e & 0x7fffffffu) * (2 * L);
        load_limbs_nc<L>(x, bp);
        load_limbs_nc<L>(y, bp + L);
        Fp<P>::cneg(y, y, (e >> 31) != 0);   // -(0,0) stays (0,0)
    } else {
        const uint32_t *bp = src + (size_t)k * (2 * L);
        load_limbs_nc<L>(x, bp);
        load_limbs_nc<L>(y, bp + L);
    }
}
// x coordinate only (the forward pass needs y only for the rare degenerate pairs)
template <class P, bool FIRST>
__device__ __forceinline__ void pair_load_x(uint32_t *x, const uint32_t *__restrict__ bases, const uint32_t *__restrict__ src, uint32_t k) {
    constexpr int L = P::L;
    const uint32_t *bp = FIRST ? bases + (size_t)(__ldg(src + k) & 0x7fffffffu) * (2 * L) : src + (size_t)k * (2 * L);
    load_limbs_nc<L>(x, bp);
}

enum { PAIR_PASS1 = 0, PAIR_PASS2 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };
// classify (P1, P2) and produce the denominator of the slope (ONE for the degenerate kinds)
template <class P> __device__ __forceinline__ int pair_classify(uint32_t *den, const uint32_t *x1, const uint32_t *y1, const uint32_t *x2,
                                                                const uint32_t *y2, bool has2) {
    using F = Fp<P>;
    constexpr int L = P::L;
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
    using P = typename C::Fq;
    using F = Fp<P>;
    constexpr int L = P::L;
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
            pair_load_x<P, FIRST>(x1, bases, src, k);
            pair_load_x<P, FIRST>(x2, bases, src, k + 1);
            if (limbs_is_zero<L>(x1) || limbs_is_zero<L>(x2) || limbs_eq<L>(x1, x2)) {   // rare: identity operand / equal x
                pair_load_point<P, FIRST>(x1, y1, bases, src, k);
                pair_load_point<P, FIRST>(x2, y2, bases, src, k + 1);
                const int kind = pair_classify<P>(den, x1, y1, x2, y2, true);
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
        pair_load_point<P, FIRST>(x1, y1, bases, src, k);
        if (has2) pair_load_point<P, FIRST>(x2, y2, bases, src, k + 1);
        const int kind = pair_classify<P>(den, x1, y1, x2, y2, has2);
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
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tasks) return;
    const uint32_t b = tail_bucket[t];
    if (b == kNoBucket) return;
    const uint32_t t1 = (__ldg(offsets + b + 1) - 1) / T;
    if (t1 - t > kFixupSmall) return;
    typename E::B acc, x;
    load_xyzz<P>(acc, tail + (size_t)t * (4 * L));
    for (uint32_t k = t + 1; k <= t1; k++) {
        load_xyzz<P>(x, head + (size_t)k * (4 * L));
        E::xyzz_add(acc, x);
    }
    store_xyzz<P>(buckets + (size_t)b * (4 * L), acc);
}
template <class C>
__global__ void __launch_bounds__(128) msm_fixup_big_kernel(const uint32_t *__restrict__ offsets, uint32_t T, const uint32_t *__restrict__ head,
                                                            const uint32_t *__restrict__ tail, const uint32_t *__restrict__ tail_bucket,
                                                            uint32_t num_tasks, uint32_t *__restrict__ buckets) {
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
    extern __shared__ uint32_t sm[];
    const uint32_t t = blockIdx.x;
    if (t >= num_tasks) return;
    const uint32_t b = tail_bucket[t];
    if (b == kNoBucket) return;
    const uint32_t t1 = (__ldg(offsets + b + 1) - 1) / T;
    if (t1 - t <= kFixupSmall) return;
    typename E::B acc, x;
    E::xyzz_set_zero(acc);
    if (threadIdx.x == 0) load_xyzz<P>(acc, tail + (size_t)t * (4 * L));
    for (uint32_t k = t + 1 + threadIdx.x; k <= t1; k += blockDim.x) {
        load_xyzz<P>(x, head + (size_t)k * (4 * L));
        E::xyzz_add(acc, x);
    }
    store_xyzz<P>(sm + threadIdx.x * (4 * L), acc);
    __syncthreads();
    for (uint32_t s2 = blockDim.x >> 1; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) {
            load_xyzz<P>(acc, sm + threadIdx.x * (4 * L));
            load_xyzz<P>(x, sm + (threadIdx.x + s2) * (4 * L));
            E::xyzz_add(acc, x);
            store_xyzz<P>(sm + threadIdx.x * (4 * L), acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        load_xyzz<P>(acc, sm);
        store_xyzz<P>(buckets + (size_t)b * (4 * L), acc);
    }
}

template <class P> __device__ __forceinline__ void load_xyzz(Xyzz<P> &b, const uint32_t *p) {
    constexpr int L = P::L;
    load_limbs<L>(b.x, p);
    load_limbs<L>(b.y, p + L);
    load_limbs<L>(b.zz, p + 2 * L);
    load_limbs<L>(b.zzz, p + 3 * L);
}
template <class P> __device__ __forceinline__ void store_xyzz(uint32_t *p, const Xyzz<P> &b) {
    constexpr int L = P::L;
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
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
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
            load_xyzz<P>(b, base + (size_t)j * (4 * L));
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
    store_xyzz<P>(partials + ((size_t)w * chunks_per_window + t) * (4 * L), sum);
}

// one block per window: strided sums then a shared-memory tree; result -> window_sums[w]
template <class C>
__global__ void __launch_bounds__(128) msm_sum_partials_kernel(const uint32_t *__restrict__ partials, uint32_t chunks_per_window,
                                                               uint32_t *__restrict__ window_sums) {
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
    extern __shared__ uint32_t sm[];
    const uint32_t w = blockIdx.x;
    typename E::B acc;
    E::xyzz_set_zero(acc);
    for (uint32_t t = threadIdx.x; t < chunks_per_window; t += blockDim.x) {
        typename E::B b;
        load_xyzz<P>(b, partials + ((size_t)w * chunks_per_window + t) * (4 * L));
        E::xyzz_add(acc, b);
    }
    store_xyzz<P>(sm + threadIdx.x * (4 * L), acc);
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            typename E::B a, b;
            load_xyzz<P>(a, sm + threadIdx.x * (4 * L));
            load_xyzz<P>(b, sm + (threadIdx.x + s) * (4 * L));
            E::xyzz_add(a, b);
            store_xyzz<P>(sm + threadIdx.x * (4 * L), a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        typename E::B a;
        load_xyzz<P>(a, sm);
        store_xyzz<P>(window_sums + (size_t)w * (4 * L), a);
    }
}

// total = sum_w 2^(c*w) * S_w by Horner (:489-502); Jacobian result (x, y, z) -> out (3L words)
template <class C> __global__ void msm_window_combine_kernel(const uint32_t *__restrict__ window_sums, int W, int c, uint32_t *__restrict__ out) {
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    typename E::J total;
    E::jac_set_zero(total);
    for (int w = W - 1; w >= 0; w--) {
        typename E::B b;
        typename E::J j;
        load_xyzz<P>(b, window_sums + (size_t)w * (4 * L));
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
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
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
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
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
// host driver
// ------------------------------------------------------------------------------------------------
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

int msm_set_window(int c) {
    if (c < 0 || c > 24) { set_last_error("window size must be in [1,24] (0 = automatic)"); return B200_EINVAL; }
    t_window_override = c;
    return 0;
}

// Window choice: a time model fitted to the B200 sweep committed in profiles/r01_window_sweep.jsonl (n = 2^26, BLS12-381):
//   accumulation 0.39 ns per (point, window) [x L^2 scaling for the 8-limb curve], reduction 2.6 ns per bucket,
//   scatter + histogram 0.02 ns per entry, + contention when the top window has fewer than ~2^10 buckets.
int msm_auto_window(size_t n, int scalar_bits) {
    if (n < 32) return 3;  // same floor as the reference (:445-449)
    double best = 1e300;
    int best_c = 3;
    for (int c = 4; c <= 23; c++) {
        MsmGeom g = make_geom(c, scalar_bits);
        if ((double)g.total_buckets * 192.0 > 24e9) continue;
        const double entries = (double)n * g.W;
        double t = 0.33 * entries + 2.6 * (double)g.total_buckets + 0.02 * entries;   // 0.33: with the batched-affine levels (0.39 without)
        if (g.top_bits < 10) t += 0.15 * (double)n;   // hot top-window buckets serialise the atomics
        // tiny inputs: keep enough accumulation tasks (>= 8 entries each) to fill the machine
        t += 2000.0 * g.W;                             // per-window fixed costs (reduction tree, combine doublings)
        if (t < best) { best = t; best_c = c; }
    }
    return best_c;
}

// buckets[b] += extra[b]  (chunked host path: every chunk after the first accumulates into `extra`)
template <class C>
__global__ void __launch_bounds__(128) msm_merge_kernel(uint32_t *__restrict__ buckets, const uint32_t *__restrict__ extra, uint32_t total_buckets) {
    using P = typename C::Fq;
    using E = Ec<P>;
    constexpr int L = P::L;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    typename E::B x, y;
    load_xyzz<P>(y, extra + (size_t)b * (4 * L));
    if (E::xyzz_is_zero(y)) return;
    load_xyzz<P>(x, buckets + (size_t)b * (4 * L));
    E::xyzz_add(x, y);
    store_xyzz<P>(buckets + (size_t)b * (4 * L), x);
}

// One MSM = K input chunks (K = 1 for device-resident inputs).  Each chunk is digit-sorted and accumulated on its own as
// soon as its `ready` event fires, so the host path overlaps the PCIe transfer of chunk k+1 with the arithmetic of chunk k;
// chunk 0 accumulates straight into `buckets`, later chunks into a second array that is merged bucket-wise.  The bucket
// reduction and window combine run once at the end.
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

struct MsmChunks {
    int K = 1;
    const size_t *offset = nullptr;      // K+1 element offsets into bases / scalars
    const cudaEvent_t *ready = nullptr;  // K events (or nullptr: data already resident)
};

template <class C> static int msm_run(const uint32_t *d_bases, const void *d_scalars, int kind, size_t n, uint32_t *d_out, cudaStream_t st,
                                      const MsmChunks &ch) {
    constexpr int L = C::Fq::L;
    if (n >= ((size_t)1 << 31)) { set_last_error("n must be < 2^31"); return B200_ETOOLARGE; }
    const int scalar_bits = scalar_kind_bits(kind, C::SCALAR_BITS);
    const size_t scalar_bytes = scalar_kind_bytes(kind);
    int c = t_window_override ? t_window_override : msm_auto_window(n, scalar_bits);
    if (c > scalar_bits) c = scalar_bits;
    const MsmGeom g = make_geom(c, scalar_bits);
    const size_t nb_total = g.total_buckets;
    const size_t one_chunk[2] = {0, n};
    const int K = ch.K > 1 ? ch.K : 1;
    const size_t *coff = K > 1 ? ch.offset : one_chunk;
    size_t n_max = 0;
    for (int k = 0; k < K; k++) n_max = std::max(n_max, coff[k + 1] - coff[k]);
    const size_t max_entries = n_max * (size_t)g.W;
    if (max_entries >= ((size_t)1 << 32)) { set_last_error("n * windows must be < 2^32"); return B200_ETOOLARGE; }
    // batched-affine levels: automatic = keep halving while buckets still hold >= ~8 entries, at most 4 levels (measured
    // at 2^26: 339 / 324 / 310 / 303 / 300 ms of accumulation for 0..4 levels), and only if the level-1 array fits comfortably
    int levels = t_affine_levels;
    if (levels < 0) {
        const double load = (double)n_max / (double)g.nb;
        levels = 0;
        for (double l = load; l >= 16.0 && levels < 4; l *= 0.5) levels++;
        if ((double)max_entries * 0.5 * 2 * L * 4 > 64e9) levels = 0;
        // 8-limb curves: the affine additions are 2.2x cheaper in multiplies but move almost as many bytes, and the levels
        // measured slower (BN254 2^24: 49 -> 55 ms), so the automatic mode keeps them for the 12-limb field only
        if (L < 12) levels = 0;
    }

    std::vector<cudaEvent_t> ev((size_t)K * 5 + 3);
    for (auto &e : ev) AB_CUDA(cudaEventCreate(&e));
    uint32_t *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *sorted = nullptr, *block_totals = nullptr;
    uint32_t *buckets = nullptr, *extra = nullptr, *partials = nullptr, *window_sums = nullptr;
    const size_t scan_blocks = (nb_total + kScanBlock - 1) / kScanBlock;
    AB_CUDA(cudaMallocAsync(&counts, nb_total * 4, st));
    AB_CUDA(cudaMallocAsync(&offsets, (nb_total + 1) * 4, st));
    AB_CUDA(cudaMallocAsync(&cursor, nb_total * 4, st));
    AB_CUDA(cudaMallocAsync(&sorted, std::max<size_t>(max_entries, 1) * 4, st));
    AB_CUDA(cudaMallocAsync(&block_totals, scan_blocks * 4, st));
    AB_CUDA(cudaMallocAsync(&buckets, nb_total * 4 * L * 4, st));
    if (K > 1) AB_CUDA(cudaMallocAsync(&extra, nb_total * 4 * L * 4, st));
    // accumulation tasks: T consecutive sorted entries per thread (>= ~128k tasks when the input allows it)
    uint32_t T = 512;
    while (T > 8 && max_entries / T < (1u << 17)) T >>= 1;
    const uint32_t max_tasks = (uint32_t)((max_entries + T - 1) / T);
    uint32_t *head = nullptr, *tail = nullptr, *head_bucket = nullptr, *tail_bucket = nullptr;
    AB_CUDA(cudaMallocAsync(&head, (size_t)std::max(max_tasks, 1u) * 4 * L * 4, st));
    AB_CUDA(cudaMallocAsync(&tail, (size_t)std::max(max_tasks, 1u) * 4 * L * 4, st));
    AB_CUDA(cudaMallocAsync(&head_bucket, (size_t)std::max(max_tasks, 1u) * 4, st));
    AB_CUDA(cudaMallocAsync(&tail_bucket, (size_t)std::max(max_tasks, 1u) * 4, st));
    // reduction geometry: chunk of m = 2^log_m buckets per thread
    int log_m = 5;
    while (log_m > 0 && (g.nb >> log_m) < 64) log_m--;
    const uint32_t max_nb = std::max(g.nb, g.nb_top);
    const uint32_t chunks = (max_nb + (1u << log_m) - 1) >> log_m;
    AB_CUDA(cudaMallocAsync(&partials, (size_t)g.W * chunks * 4 * L * 4, st));
    AB_CUDA(cudaMallocAsync(&window_sums, (size_t)g.W * 4 * L * 4, st));

    cudaEvent_t *e_begin = &ev[(size_t)K * 5], *e_acc_done = &ev[(size_t)K * 5 + 1], *e_end = &ev[(size_t)K * 5 + 2];
    AB_CUDA(cudaEventRecord(*e_begin, st));
    for (int k = 0; k < K; k++) {
        const size_t nk = coff[k + 1] - coff[k];
        cudaEvent_t *e = &ev[(size_t)k * 5];
        if (ch.ready) AB_CUDA(cudaStreamWaitEvent(st, ch.ready[k], 0));
        const void *scal = (const char *)d_scalars + coff[k] * scalar_bytes;
        const uint32_t *bas = d_bases + coff[k] * (2 * L);
        uint32_t *target = k == 0 ? buckets : extra;
        AB_CUDA(cudaEventRecord(e[0], st));
        AB_CUDA(cudaMemsetAsync(counts, 0, nb_total * 4, st));
        const unsigned dblocks = (unsigned)((nk + 255) / 256);
        if (nk) {
            msm_digits_kernel<C, 0><<<dblocks, 256, 0, st>>>(scal, kind, nk, g, 0, g.W, counts, nullptr);
            AB_LAUNCHED();
        }
        AB_CUDA(cudaEventRecord(e[1], st));
        scan_block_totals_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, nb_total, block_totals);
        AB_LAUNCHED();
        scan_totals_kernel<<<1, kScanThreads, 0, st>>>(block_totals, scan_blocks);
        AB_LAUNCHED();
        scan_apply_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, nb_total, block_totals, offsets);
        AB_LAUNCHED();
        AB_CUDA(cudaMemcpyAsync(cursor, offsets, nb_total * 4, cudaMemcpyDeviceToDevice, st));
        AB_CUDA(cudaEventRecord(e[2], st));
        // scatter in groups of windows so that the active write fronts (one 32-byte sector per bucket of the group) stay
        // L2-resident: random 4-byte stores into a multi-GB array otherwise cost a DRAM sector each (measured 2.5x slower)
        if (nk) {
            const size_t front_bytes = (size_t)g.nb * 32;
            int group = (int)std::max<size_t>(1, ((size_t)48 << 20) / front_bytes);
            for (int w0 = 0; w0 < g.W; w0 += group) {
                msm_digits_kernel<C, 1><<<dblocks, 256, 0, st>>>(scal, kind, nk, g, w0, std::min(g.W, w0 + group), cursor, sorted);
                AB_LAUNCHED();
            }
        }
        AB_CUDA(cudaEventRecord(e[3], st));
        AB_CUDA(cudaMemsetAsync(target, 0, nb_total * 4 * L * 4, st));
        // optional batched-affine pre-reduction: each level halves the entries (as affine points) before the XYZZ accumulation
        size_t cur_entries = nk * (size_t)g.W;   // upper bound on the entries of the current level
        const uint32_t *cur_src = sorted, *cur_offsets = offsets;
        uint32_t *lvl_pts[2] = {nullptr, nullptr}, *lvl_off[2] = {nullptr, nullptr};
        int levels_done = 0;
        for (int lv = 0; lv < levels && nk; lv++) {
            // sum_b ceil(cnt_b / 2) <= min((entries + non-empty buckets) / 2, entries)
            const size_t out_cap = std::min((cur_entries + nb_total) / 2 + 1, cur_entries);
            // one inversion (~570 modmuls) per `batch` additions: never below 256 in automatic mode, where a level that cannot
            // fill the machine with 256-slot threads is left to the XYZZ kernel instead
            const bool forced = t_affine_levels >= 0;
            uint32_t batch = 1024;
            while (batch > (forced ? 32u : 256u) && out_cap / batch < (1u << 16)) batch >>= 1;
            if (!forced && out_cap / batch < (1u << 16)) break;   // measured: 2^20 inputs got slower (12.0 -> 13.8 ms) with thin levels
            uint32_t *pts = nullptr, *off2 = nullptr;
            AB_CUDA(cudaMallocAsync(&pts, out_cap * 2 * L * 4, st));
            AB_CUDA(cudaMallocAsync(&off2, (nb_total + 1) * 4, st));
            msm_halve_counts_kernel<<<(unsigned)((nb_total + 255) / 256), 256, 0, st>>>(cur_offsets, (uint32_t)nb_total, counts);
            AB_LAUNCHED();
            scan_block_totals_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, nb_total, block_totals);
            AB_LAUNCHED();
            scan_totals_kernel<<<1, kScanThreads, 0, st>>>(block_totals, scan_blocks);
            AB_LAUNCHED();
            scan_apply_kernel<<<(unsigned)scan_blocks, kScanThreads, 0, st>>>(counts, nb_total, block_totals, off2);
            AB_LAUNCHED();
            const uint32_t nthreads = (uint32_t)((out_cap + batch - 1) / batch);
            // 128 threads x 4 resident blocks (128 registers, no spills) measured best: forcing 5 / 6 blocks per SM (96 / 80
            // registers with spills) gave 324 / 345 ms of accumulation instead of 286 ms @2^26
            const unsigned pg = (nthreads + 127) / 128;
            if (lv == 0) msm_pair_add_kernel<C, true, 4><<<pg, 128, 0, st>>>(bas, cur_src, cur_offsets, off2, (uint32_t)nb_total, batch, pts, nthreads);
            else msm_pair_add_kernel<C, false, 4><<<pg, 128, 0, st>>>(bas, cur_src, cur_offsets, off2, (uint32_t)nb_total, batch, pts, nthreads);
            AB_LAUNCHED();
            // the level before the previous one is no longer read
            if (lvl_pts[lv & 1]) { AB_CUDA(cudaFreeAsync(lvl_pts[lv & 1], st)); AB_CUDA(cudaFreeAsync(lvl_off[lv & 1], st)); }
            lvl_pts[lv & 1] = pts;
            lvl_off[lv & 1] = off2;
            cur_src = pts;
            cur_offsets = off2;
            cur_entries = out_cap;
            levels_done++;
        }
        const uint32_t num_tasks = (uint32_t)((cur_entries + T - 1) / T);
        if (num_tasks > max_tasks) { set_last_error("internal: task count exceeds scratch"); return B200_EINVAL; }
        if (num_tasks) {
            if (levels_done)
                msm_accumulate_kernel<C, true><<<(num_tasks + 127) / 128, 128, 0, st>>>(cur_src, nullptr, cur_offsets, (uint32_t)nb_total, T, target, head,
                                                                                      tail, head_bucket, tail_bucket, num_tasks);
            else
                msm_accumulate_kernel<C, false><<<(num_tasks + 127) / 128, 128, 0, st>>>(bas, sorted, offsets, (uint32_t)nb_total, T, target, head, tail,
                                                                                       head_bucket, tail_bucket, num_tasks);
            AB_LAUNCHED();
            msm_fixup_small_kernel<C><<<(num_tasks + 127) / 128, 128, 0, st>>>(cur_offsets, T, head, tail, tail_bucket, num_tasks, target);
            AB_LAUNCHED();
            msm_fixup_big_kernel<C><<<num_tasks, 128, 128 * 4 * L * 4, st>>>(cur_offsets, T, head, tail, tail_bucket, num_tasks, target);
            AB_LAUNCHED();
        }
        for (int q = 0; q < 2; q++)
            if (lvl_pts[q]) { AB_CUDA(cudaFreeAsync(lvl_pts[q], st)); AB_CUDA(cudaFreeAsync(lvl_off[q], st)); }
        if (k > 0) {
            msm_merge_kernel<C><<<(unsigned)((nb_total + 127) / 128), 128, 0, st>>>(buckets, extra, (uint32_t)nb_total);
            AB_LAUNCHED();
        }
        AB_CUDA(cudaEventRecord(e[4], st));
    }
    AB_CUDA(cudaEventRecord(*e_acc_done, st));
    const unsigned rthreads = (unsigned)g.W * chunks;
    msm_bucket_reduce_kernel<C><<<(rthreads + 127) / 128, 128, 0, st>>>(buckets, g, log_m, chunks, partials);
    AB_LAUNCHED();
    msm_sum_partials_kernel<C><<<g.W, 128, 128 * 4 * L * 4, st>>>(partials, chunks, window_sums);
    AB_LAUNCHED();
    cudaEvent_t e_red;
    AB_CUDA(cudaEventCreate(&e_red));
    AB_CUDA(cudaEventRecord(e_red, st));
    msm_window_combine_kernel<C><<<1, 32, 0, st>>>(window_sums, g.W, g.c, d_out);
    AB_LAUNCHED();
    AB_CUDA(cudaEventRecord(*e_end, st));

    for (uint32_t *p : {counts, offsets, cursor, sorted, block_totals, buckets, extra, partials, window_sums, head, tail, head_bucket, tail_bucket})
        if (p) AB_CUDA(cudaFreeAsync(p, st));
    AB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 7; i++) t_last.ms[i] = 0.f;
    for (int k = 0; k < K; k++)
        for (int i = 0; i < 4; i++) {
            float ms = 0.f;
            AB_CUDA(cudaEventElapsedTime(&ms, ev[(size_t)k * 5 + i], ev[(size_t)k * 5 + i + 1]));
            t_last.ms[i] += ms;  // digits+hist, scan, scatter, accumulate(+fixups, merge)
        }
    AB_CUDA(cudaEventElapsedTime(&t_last.ms[4], *e_acc_done, e_red));
    AB_CUDA(cudaEventElapsedTime(&t_last.ms[5], e_red, *e_end));
    AB_CUDA(cudaEventElapsedTime(&t_last.ms[6], *e_begin, *e_end));  // includes waiting for transfers on the host path
    for (auto &e : ev) cudaEventDestroy(e);
    cudaEventDestroy(e_red);
    t_last.c = g.c;
    t_last.W = g.W;
    t_last.bucket_adds = (unsigned long long)n * g.W;  // upper bound: zero digits are skipped
    return 0;
}

int msm_dispatch(int curve, int kind, const void *d_bases, const void *d_scalars, size_t n, uint64_t *out_xyz_host, cudaStream_t st, int K,
                 const size_t *chunk_off, const cudaEvent_t *ready) {
    if (kind < B200_SCALARS_FR_MONT || kind > B200_SCALARS_U64) { set_last_error("unknown scalar kind"); return B200_EINVAL; }
    if (!out_xyz_host || (n && (!d_bases || !d_scalars))) { set_last_error("null pointer"); return B200_EINVAL; }
    if (curve != B200_CURVE_BLS12_381 && curve != B200_CURVE_BN254) { set_last_error("unknown curve id"); return B200_EINVAL; }
    const int L = curve == B200_CURVE_BLS12_381 ? 12 : 8;
    if (n == 0) {  // Projective::zero() = (1,1,0) (group.rs:142-158)
        for (int i = 0; i < 3 * L / 2; i++) out_xyz_host[i] = 0;
        for (int i = 0; i < L; i++) {
            uint32_t one = curve == B200_CURVE_BLS12_381 ? BlsFq::ONE(i) : BnFq::ONE(i);
            for (int k = 0; k < 2; k++) out_xyz_host[(k * L + i) / 2] |= (uint64_t)one << (32 * ((k * L + i) & 1));
        }
        return 0;
    }
    MsmChunks ch;
    ch.K = K;
    ch.offset = chunk_off;
    ch.ready = ready;
    // device-resident inputs too large for one 32-bit entry index space (n * windows >= 2^32): split into equal chunks
    std::vector<size_t> auto_off;
    const char *force = getenv("B200_MSM_FORCE_CHUNKS");   // test hook: exercise the chunked device path at small n
    if (K <= 1 && (n > ((size_t)1 << 27) || (force && atoi(force) > 1 && n >= 64))) {
        const int kk = n > ((size_t)1 << 27) ? (int)((n + ((size_t)1 << 27) - 1) >> 27) : atoi(force);
        for (int k = 0; k <= kk; k++) auto_off.push_back(n * (size_t)k / kk);
        ch.K = kk;
        ch.offset = auto_off.data();
        ch.ready = nullptr;
    }
    uint32_t *d_out = nullptr;
    AB_CUDA(cudaMallocAsync(&d_out, 3 * L * 4, st));
    int rc = curve == B200_CURVE_BLS12_381 ? msm_run<CurveBls>((const uint32_t *)d_bases, d_scalars, kind, n, d_out, st, ch)
                                           : msm_run<CurveBn>((const uint32_t *)d_bases, d_scalars, kind, n, d_out, st, ch);
    if (rc) return rc;
    AB_CUDA(cudaMemcpyAsync(out_xyz_host, d_out, 3 * L * 4, cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaFreeAsync(d_out, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int msm_last_timings(float *ms7, int *c, int *windows, unsigned long long *bucket_adds) {
    if (ms7) for (int i = 0; i < 7; i++) ms7[i] = t_last.ms[i];
    if (c) *c = t_last.c;
    if (windows) *windows = t_last.W;
    if (bucket_adds) *bucket_adds = t_last.bucket_adds;
    return 0;
}

int g1_sum_dispatch(int curve, const uint64_t *pts_host, size_t k, uint64_t *out_host, bool to_affine) {
    if (curve != B200_CURVE_BLS12_381 && curve != B200_CURVE_BN254) { set_last_error("unknown curve id"); return B200_EINVAL; }
    if (!pts_host || !out_host) { set_last_error("null pointer"); return B200_EINVAL; }
    const int L = curve == B200_CURVE_BLS12_381 ? 12 : 8;
    const size_t in_words = k * 3 * L, out_words = to_affine ? k * 2 * L : 3 * L;
    uint32_t *d_in = nullptr, *d_out = nullptr;
    cudaStream_t st = 0;
    AB_CUDA(cudaMallocAsync(&d_in, std::max<size_t>(in_words, 1) * 4, st));
    AB_CUDA(cudaMallocAsync(&d_out, std::max<size_t>(out_words, 1) * 4, st));
    AB_CUDA(cudaMemcpyAsync(d_in, pts_host, in_words * 4, cudaMemcpyHostToDevice, st));
    if (to_affine) {
        if (curve == B200_CURVE_BLS12_381) jac_to_affine_kernel<CurveBls><<<(unsigned)((k + 31) / 32), 32, 0, st>>>(d_in, k, d_out);
        else jac_to_affine_kernel<CurveBn><<<(unsigned)((k + 31) / 32), 32, 0, st>>>(d_in, k, d_out);
    } else {
        if (curve == B200_CURVE_BLS12_381) jac_sum_kernel<CurveBls><<<1, 32, 0, st>>>(d_in, k, d_out);
        else jac_sum_kernel<CurveBn><<<1, 32, 0, st>>>(d_in, k, d_out);
    }
    AB_LAUNCHED();
    AB_CUDA(cudaMemcpyAsync(out_host, d_out, out_words * 4, cudaMemcpyDeviceToHost, st));
    AB_CUDA(cudaFreeAsync(d_in, st));
    AB_CUDA(cudaFreeAsync(d_out, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace ab200
