// gen.cu — synthetic inputs generated on the device (benchmarks, large-size property tests).
//   scalars: uniform Fr by mask-and-reject, the rule of Fp::rand (ff/src/fields/models/fp/mod.rs:521-548)
//   bases:   P_i = b_i * G (b_i a 64-bit splitmix64 stream), fixed-base comb over an 8 x 256 table of multiples of G,
//            normalised to affine with a per-thread batch inversion (the role of Projective::normalize_batch,
//            group.rs:302-319) — so that the exact MSM answer is (sum s_i * b_i mod r) * G.
#include <map>
#include <mutex>

#include "common.cuh"
#include "ec.cuh"

namespace ab200 {

__host__ __device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <class P> __global__ void gen_scalars_kernel(uint64_t seed, size_t n, uint32_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v[8];
    for (uint64_t attempt = 0;; attempt++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t w = splitmix64_at(seed ^ (attempt * 0xD1B54A32D192ED03ull), i * 4 + k);
            v[2 * k] = (uint32_t)w;
            v[2 * k + 1] = (uint32_t)(w >> 32);
        }
        v[7] &= (P::BITS % 32) ? ((1u << (P::BITS % 32)) - 1) : 0xffffffffu;
        // accept iff v < p
        uint32_t t = ptx::sub_cc(v[0], P::MOD(0));
#pragma unroll
        for (int k = 1; k < 8; k++) t = ptx::subc_cc(v[k], P::MOD(k));
        (void)t;
        if (ptx::subc(0u, 0u)) break;
    }
    store_limbs<8>(out + i * 8, v);
}

// table[w*256 + d] = (d * 256^w) * G, affine Montgomery (d = 0 -> (0,0))
template <class P> __global__ void gen_table_kernel(uint32_t *table) {
    using E = Ec<P>;
    using F = Fp<P>;
    constexpr int L = P::L;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 8 * 256) return;
    int w = t >> 8, d = t & 255;
    typename E::J base, acc;
#pragma unroll
    for (int i = 0; i < L; i++) { base.x[i] = P::GEN_X(i); base.y[i] = P::GEN_Y(i); base.z[i] = P::ONE(i); }
    for (int k = 0; k < 8 * w; k++) E::jac_dbl(base);
    E::jac_set_zero(acc);
    for (int bit = 7; bit >= 0; bit--) {
        E::jac_dbl(acc);
        if ((d >> bit) & 1) E::jac_add(acc, base);
    }
    uint32_t ax[L], ay[L];
    E::jac_to_affine(ax, ay, acc);
    store_limbs<L>(table + (size_t)t * 2 * L, ax);
    store_limbs<L>(table + (size_t)t * 2 * L + L, ay);
    (void)sizeof(F);
}

static constexpr int kGenBatch = 8;
template <class P> __global__ void __launch_bounds__(64) gen_bases_kernel(uint64_t seed, size_t n, const uint32_t *__restrict__ table,
                                                                         uint32_t *__restrict__ bases, uint64_t *__restrict__ bvals) {
    using E = Ec<P>;
    using F = Fp<P>;
    constexpr int L = P::L;
    size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kGenBatch;
    if (i0 >= n) return;
    typename E::B pts[kGenBatch];
    uint32_t prefix[kGenBatch][L];
    int cnt = 0;
    for (; cnt < kGenBatch && i0 + cnt < n; cnt++) {
        uint64_t b = splitmix64_at(seed, i0 + cnt) | 1ull;
        if (bvals) bvals[i0 + cnt] = b;
        typename E::B acc;
        E::xyzz_set_zero(acc);
        for (int w = 0; w < 8; w++) {
            uint32_t d = (uint32_t)(b >> (8 * w)) & 255u;
            uint32_t px[L], py[L];
            const uint32_t *tp = table + (size_t)(w * 256 + d) * 2 * L;
            load_limbs_nc<L>(px, tp);
            load_limbs_nc<L>(py, tp + L);
            E::madd(acc, px, py, false);
        }
        pts[cnt] = acc;
        if (cnt == 0) limbs_copy<L>(prefix[0], acc.zzz);
        else F::mul(prefix[cnt], prefix[cnt - 1], acc.zzz);
    }
    // b*G with b odd and < 2^64 << r is never the identity, so every zzz is invertible
    uint32_t inv[L];
    F::inv(inv, prefix[cnt - 1]);
    for (int k = cnt - 1; k >= 0; k--) {
        uint32_t zi[L], t[L], ax[L], ay[L];
        if (k > 0) F::mul(zi, inv, prefix[k - 1]);
        else limbs_copy<L>(zi, inv);            // zi = 1/zzz_k
        F::mul(inv, inv, pts[k].zzz);
        F::mul(t, pts[k].zz, zi);               // zz/zzz = 1/z ; (1/z)^2 = 1/zz
        F::sqr(t, t);
        F::mul(ax, pts[k].x, t);
        F::mul(ay, pts[k].y, zi);
        store_limbs<L>(bases + (i0 + k) * 2 * L, ax);
        store_limbs<L>(bases + (i0 + k) * 2 * L + L, ay);
    }
}

static std::mutex g_table_mutex;
static std::map<std::pair<int, int>, uint32_t *> g_tables;

template <class P> static int gen_bases_run(int curve, uint64_t seed, size_t n, void *d_bases, void *d_b, cudaStream_t st) {
    constexpr int L = P::L;
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    uint32_t *table = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_table_mutex);
        auto key = std::make_pair(dev, curve);
        auto it = g_tables.find(key);
        if (it == g_tables.end()) {
            AB_CUDA(cudaMalloc(&table, (size_t)8 * 256 * 2 * L * 4));
            gen_table_kernel<P><<<(8 * 256 + 63) / 64, 64, 0, st>>>(table);
            AB_LAUNCHED();
            AB_CUDA(cudaStreamSynchronize(st));
            g_tables[key] = table;
        } else table = it->second;
    }
    size_t threads = (n + kGenBatch - 1) / kGenBatch;
    gen_bases_kernel<P><<<(unsigned)((threads + 63) / 64), 64, 0, st>>>(seed, n, table, (uint32_t *)d_bases, (uint64_t *)d_b);
    AB_LAUNCHED();
    return 0;
}

int gen_bases_dispatch(int curve, uint64_t seed, size_t n, void *d_bases, void *d_b, cudaStream_t st) {
    if (!d_bases && n) { set_last_error("null pointer"); return B200_EINVAL; }
    if (n == 0) return 0;
    if (curve == B200_CURVE_BLS12_381) return gen_bases_run<BlsFq>(curve, seed, n, d_bases, d_b, st);
    if (curve == B200_CURVE_BN254) return gen_bases_run<BnFq>(curve, seed, n, d_bases, d_b, st);
    set_last_error("unknown curve id");
    return B200_EINVAL;
}
int gen_scalars_dispatch(int field, uint64_t seed, size_t n, void *d_scalars, cudaStream_t st) {
    if (!d_scalars && n) { set_last_error("null pointer"); return B200_EINVAL; }
    if (n == 0) return 0;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (field == B200_FIELD_BLS12_381_FR) gen_scalars_kernel<BlsFr><<<blocks, 256, 0, st>>>(seed, n, (uint32_t *)d_scalars);
    else if (field == B200_FIELD_BN254_FR) gen_scalars_kernel<BnFr><<<blocks, 256, 0, st>>>(seed, n, (uint32_t *)d_scalars);
    else { set_last_error("unknown scalar field id"); return B200_EINVAL; }
    AB_LAUNCHED();
    return 0;
}

}  // namespace ab200
