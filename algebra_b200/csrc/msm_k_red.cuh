// msm_k_red.cuh — bucket reduction, window combine, Jacobian sum / normalisation kernels and their launchers.
#pragma once
#include "msm_common.cuh"

namespace ab200 {

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

template <class C>
int MsmRedLaunch<C>::reduce(const uint32_t *buckets, MsmGeom g, int log_m, uint32_t chunks, uint32_t *partials, uint32_t *window_sums, cudaStream_t st) {
    constexpr int L = C::F::L;
    const unsigned rthreads = (unsigned)g.W * chunks;
    msm_bucket_reduce_kernel<C><<<(rthreads + 127) / 128, 128, 0, st>>>(buckets, g, log_m, chunks, partials);
    AB_LAUNCHED();
    AB_CUDA(cudaFuncSetAttribute(msm_sum_partials_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * L * 4));
    msm_sum_partials_kernel<C><<<g.W, 128, 128 * 4 * L * 4, st>>>(partials, chunks, window_sums);
    AB_LAUNCHED();
    return 0;
}
template <class C> int MsmRedLaunch<C>::combine(const uint32_t *window_sums, int W, int c, uint32_t *out, cudaStream_t st) {
    msm_window_combine_kernel<C><<<1, 32, 0, st>>>(window_sums, W, c, out);
    AB_LAUNCHED();
    return 0;
}
template <class C> int MsmRedLaunch<C>::sum_or_affine(bool to_affine, const uint32_t *d_in, size_t k, uint32_t *d_out, cudaStream_t st) {
    if (to_affine) jac_to_affine_kernel<C><<<(unsigned)((k + 31) / 32), 32, 0, st>>>(d_in, k, d_out);
    else jac_sum_kernel<C><<<1, 32, 0, st>>>(d_in, k, d_out);
    AB_LAUNCHED();
    return 0;
}

}  // namespace ab200
