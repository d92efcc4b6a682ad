// capi.cu — the extern "C" boundary declared in include/algebra_b200.h.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace ab200 {
int ntt_dispatch(int field, void *d_data, uint32_t log_n, int inverse, const uint64_t *coset, cudaStream_t st);
int ntt_clear_cache();
int msm_dispatch(int curve, int kind, const void *d_bases, const void *d_scalars, size_t n, uint64_t *out_xyz_host, cudaStream_t st, int K,
                 const size_t *chunk_off, const cudaEvent_t *ready);
size_t scalar_kind_bytes(int kind);
int msm_set_window(int c);
int msm_set_affine_levels(int levels);
int msm_auto_window(size_t n, int scalar_bits);
int msm_last_timings(float *ms7, int *c, int *windows, unsigned long long *bucket_adds);
int g1_sum_dispatch(int curve, const uint64_t *pts_host, size_t k, uint64_t *out_host, bool to_affine);
int fp_op_dispatch(int field, int op, const void *a, const void *b, void *out, size_t n, int reps, cudaStream_t st);
int ec_op_dispatch(int curve, int op, const void *a, const void *b, void *out, size_t n, cudaStream_t st);
int gen_bases_dispatch(int curve, uint64_t seed, size_t n, void *d_bases, void *d_b, cudaStream_t st);
int gen_scalars_dispatch(int field, uint64_t seed, size_t n, void *d_scalars, cudaStream_t st);
int batch_mul_dispatch(int curve, const uint64_t *base_xy, const void *d_scalars, size_t n, void *d_out, cudaStream_t st);
int normalize_batch_dispatch(int curve, const void *d_xyz, size_t n, void *d_out, cudaStream_t st);
}  // namespace ab200
using namespace ab200;

extern "C" {

int b200_msm_sw_g1_scalars_dev(int curve, int scalar_kind, const void *d_bases, const void *d_scalars, size_t n, uint64_t *out_xyz, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    return msm_dispatch(curve, scalar_kind, d_bases, d_scalars, n, out_xyz, (cudaStream_t)stream, 1, nullptr, nullptr);
}
int b200_msm_sw_g1_dev(int curve, const void *d_bases, const void *d_scalars, size_t n, uint64_t *out_xyz, void *stream) {
    return b200_msm_sw_g1_scalars_dev(curve, B200_SCALARS_FR_MONT, d_bases, d_scalars, n, out_xyz, stream);
}
int b200_msm_sw_g1(int curve, const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out_xyz) {
    return b200_msm_sw_g1_scalars(curve, B200_SCALARS_FR_MONT, bases, scalars, n, out_xyz);
}

int b200_msm_sw_g1_scalars(int curve, int scalar_kind, const uint64_t *bases, const void *scalars, size_t n, uint64_t *out_xyz) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    if (scalar_kind < B200_SCALARS_FR_MONT || scalar_kind > B200_SCALARS_U64) { set_last_error("unknown scalar kind"); return B200_EINVAL; }
    if (curve != B200_CURVE_BLS12_381 && curve != B200_CURVE_BN254) { set_last_error("unknown curve id"); return B200_EINVAL; }
    if (!out_xyz || (n && (!bases || !scalars))) { set_last_error("null pointer"); return B200_EINVAL; }
    if (n == 0) return msm_dispatch(curve, scalar_kind, nullptr, nullptr, 0, out_xyz, 0, 1, nullptr, nullptr);
    const size_t sb = scalar_kind_bytes(scalar_kind);
    const size_t N = curve == B200_CURVE_BLS12_381 ? 6 : 4;
    // Pipeline over K input chunks on two streams: the copy stream moves (scalars_k, bases_k) for k = 0..K-1 back to back,
    // the compute stream sorts and accumulates chunk k as soon as it has landed (msm.cu: msm_run).  With pinned host
    // memory the PCIe time of chunks 1..K-1 hides behind the arithmetic of the chunks before them.
    static thread_local cudaStream_t streams[64][2] = {};
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_last_error("device index out of range"); return B200_EINVAL; }
    if (!streams[dev][0]) {
        AB_CUDA(cudaStreamCreateWithFlags(&streams[dev][0], cudaStreamNonBlocking));
        AB_CUDA(cudaStreamCreateWithFlags(&streams[dev][1], cudaStreamNonBlocking));
    }
    cudaStream_t st = streams[dev][0], copy_st = streams[dev][1];
    // chunk boundaries in 16ths of n.  Growing chunks: only the first, small transfer is exposed and every later one is
    // shorter than the arithmetic of the chunk before it; few chunks, because small chunks run the accumulation less
    // efficiently (short bucket runs) and each extra chunk costs a bucket-merge pass.  B200_MSM_CHUNKS="2,8,16" overrides (tuning knob).
    // measured @2^26 (e2e ms): {2,7,16} 404, {2,6,16} 405, {2,8,16} 418, {2,4,8,16} 419, {4,16} 432, {3,16} 447, unchunked 511
    int bounds[9] = {0, 2, 7, 16, 0, 0, 0, 0, 0};
    int K = n >= ((size_t)1 << 22) ? 3 : 1;
    if (const char *e = getenv("B200_MSM_CHUNKS")) {
        int k = 0, v = 0;
        const char *q = e;
        while (*q && k < 8) {
            v = atoi(q);
            if (v <= bounds[k] || v > 16) break;
            bounds[++k] = v;
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        if (k >= 1 && bounds[k] == 16 && n >= 1024) K = k;
    }
    size_t off[9];
    cudaEvent_t ready[8], alloc_done;
    for (int k = 0; k <= K; k++) off[k] = K == 1 ? (size_t)k * n : n / 16 * (size_t)bounds[k];
    off[K] = n;
    void *d_bases = nullptr, *d_scalars = nullptr;
    AB_CUDA(cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming));
    AB_CUDA(cudaMallocAsync(&d_bases, n * 2 * N * 8, st));
    AB_CUDA(cudaMallocAsync(&d_scalars, n * sb, st));
    AB_CUDA(cudaEventRecord(alloc_done, st));
    AB_CUDA(cudaStreamWaitEvent(copy_st, alloc_done, 0));
    for (int k = 0; k < K; k++) {
        const size_t lo = off[k], cnt = off[k + 1] - off[k];
        AB_CUDA(cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
        AB_CUDA(cudaMemcpyAsync((char *)d_scalars + lo * sb, (const char *)scalars + lo * sb, cnt * sb, cudaMemcpyHostToDevice, copy_st));
        AB_CUDA(cudaMemcpyAsync((char *)d_bases + lo * 2 * N * 8, (const char *)bases + lo * 2 * N * 8, cnt * 2 * N * 8, cudaMemcpyHostToDevice, copy_st));
        AB_CUDA(cudaEventRecord(ready[k], copy_st));
    }
    int rc = msm_dispatch(curve, scalar_kind, d_bases, d_scalars, n, out_xyz, st, K, off, ready);
    cudaStreamSynchronize(copy_st);
    cudaFreeAsync(d_bases, st);
    cudaFreeAsync(d_scalars, st);
    cudaStreamSynchronize(st);
    cudaEventDestroy(alloc_done);
    for (int k = 0; k < K; k++) cudaEventDestroy(ready[k]);
    return rc;
}

int b200_set_msm_window(int c) { return msm_set_window(c); }
int b200_set_msm_affine_levels(int levels) { return msm_set_affine_levels(levels); }
int b200_msm_window_for(int curve, size_t n) {
    if (curve != B200_CURVE_BLS12_381 && curve != B200_CURVE_BN254) return B200_EINVAL;
    return msm_auto_window(n, curve == B200_CURVE_BLS12_381 ? 255 : 254);
}
int b200_g1_sum(int curve, const uint64_t *points_xyz, size_t k, uint64_t *out_xyz) { if (int irc = ensure_device_init()) return irc; return g1_sum_dispatch(curve, points_xyz, k, out_xyz, false); }
int b200_g1_into_affine(int curve, const uint64_t *xyz, uint64_t *out_xy) { if (int irc = ensure_device_init()) return irc; return g1_sum_dispatch(curve, xyz, 1, out_xy, true); }

int b200_ntt_fr_dev(int field, void *d_data, uint32_t log_n, int inverse, const uint64_t *coset_offset, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = ntt_dispatch(field, d_data, log_n, inverse, coset_offset, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int b200_ntt_fr(int field, uint64_t *data, uint32_t log_n, int inverse, const uint64_t *coset_offset) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    if (!data) { set_last_error("null data pointer"); return B200_EINVAL; }
    if (field != B200_FIELD_BLS12_381_FR && field != B200_FIELD_BN254_FR) { set_last_error("unknown scalar field id"); return B200_EINVAL; }
    if (log_n > (uint32_t)(field == B200_FIELD_BLS12_381_FR ? 32 : 28)) {
        set_last_error("log_n exceeds TWO_ADICITY (Radix2EvaluationDomain::new would return None)");
        return B200_ETOOLARGE;
    }
    const size_t bytes = ((size_t)32) << log_n;
    cudaStream_t st = 0;
    void *d = nullptr;
    AB_CUDA(cudaMallocAsync(&d, bytes, st));
    AB_CUDA(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, st));
    int rc = ntt_dispatch(field, d, log_n, inverse, coset_offset, st);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H", __FILE__, __LINE__);
    }
    cudaFreeAsync(d, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    return rc;
}
int b200_clear_cache(void) {
    int rc = ntt_clear_cache();
    // also hand the stream-ordered scratch retained by the default memory pool back to the driver (a 2^26 MSM keeps tens
    // of GB of level buffers cached for the next call otherwise)
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
    }
    return rc;
}

size_t b200_poly_mul_size(int field, size_t la, size_t lb) {
    if ((field != B200_FIELD_BLS12_381_FR && field != B200_FIELD_BN254_FR) || la == 0 || lb == 0) return 0;
    const size_t need = la + lb - 1;
    size_t n = 1;
    uint32_t lg = 0;
    while (n < need) { n <<= 1; lg++; }
    return lg > (uint32_t)(field == B200_FIELD_BLS12_381_FR ? 32 : 28) ? 0 : n;
}

static int poly_mul_on_device(int field, const void *d_a, size_t la, const void *d_b, size_t lb, void *d_out, cudaStream_t st) {
    const size_t n = b200_poly_mul_size(field, la, lb);
    if (!n) { set_last_error("empty operand or product degree past TWO_ADICITY"); return la && lb ? B200_ETOOLARGE : B200_EINVAL; }
    uint32_t lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    void *tmp = nullptr;
    AB_CUDA(cudaMallocAsync(&tmp, n * 32, st));
    // evaluate_over_domain: zero-pad to n, forward transform
    AB_CUDA(cudaMemsetAsync(d_out, 0, n * 32, st));
    AB_CUDA(cudaMemsetAsync(tmp, 0, n * 32, st));
    AB_CUDA(cudaMemcpyAsync(d_out, d_a, la * 32, cudaMemcpyDeviceToDevice, st));
    AB_CUDA(cudaMemcpyAsync(tmp, d_b, lb * 32, cudaMemcpyDeviceToDevice, st));
    int rc = ntt_dispatch(field, d_out, lg, 0, nullptr, st);
    if (!rc) rc = ntt_dispatch(field, tmp, lg, 0, nullptr, st);
    // self_evals *= &other_evals ; interpolate
    if (!rc) rc = fp_op_dispatch(field == B200_FIELD_BLS12_381_FR ? 1 : 3, 0, d_out, tmp, d_out, n, 1, st);
    if (!rc) rc = ntt_dispatch(field, d_out, lg, 1, nullptr, st);
    cudaFreeAsync(tmp, st);
    return rc;
}

int b200_poly_mul_fr_dev(int field, const void *d_a, size_t la, const void *d_b, size_t lb, void *d_out, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    if (!d_a || !d_b || !d_out) { set_last_error("null pointer"); return B200_EINVAL; }
    int rc = poly_mul_on_device(field, d_a, la, d_b, lb, d_out, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int b200_poly_mul_fr(int field, const uint64_t *a, size_t la, const uint64_t *b, size_t lb, uint64_t *out) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    if (!a || !b || !out) { set_last_error("null pointer"); return B200_EINVAL; }
    const size_t n = b200_poly_mul_size(field, la, lb);
    if (!n) { set_last_error("empty operand or product degree past TWO_ADICITY"); return la && lb ? B200_ETOOLARGE : B200_EINVAL; }
    cudaStream_t st = 0;
    void *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    AB_CUDA(cudaMallocAsync(&d_a, la * 32, st));
    AB_CUDA(cudaMallocAsync(&d_b, lb * 32, st));
    AB_CUDA(cudaMallocAsync(&d_o, n * 32, st));
    AB_CUDA(cudaMemcpyAsync(d_a, a, la * 32, cudaMemcpyHostToDevice, st));
    AB_CUDA(cudaMemcpyAsync(d_b, b, lb * 32, cudaMemcpyHostToDevice, st));
    int rc = poly_mul_on_device(field, d_a, la, d_b, lb, d_o, st);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, d_o, n * 32, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync D2H", __FILE__, __LINE__);
    }
    cudaFreeAsync(d_a, st); cudaFreeAsync(d_b, st); cudaFreeAsync(d_o, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    return rc;
}

int b200_gen_bases_dev(int curve, uint64_t seed, size_t n, void *d_bases, void *d_b, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = gen_bases_dispatch(curve, seed, n, d_bases, d_b, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_gen_scalars_dev(int field, uint64_t seed, size_t n, void *d_scalars, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = gen_scalars_dispatch(field, seed, n, d_scalars, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_g1_batch_mul_dev(int curve, const uint64_t *base_xy, const void *d_scalars, size_t n, void *d_out_xy, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = batch_mul_dispatch(curve, base_xy, d_scalars, n, d_out_xy, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_g1_normalize_batch_dev(int curve, const void *d_xyz, size_t n, void *d_out_xy, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = normalize_batch_dispatch(curve, d_xyz, n, d_out_xy, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_fp_op_dev(int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n, int reps, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = fp_op_dispatch(field, op, d_a, d_b, d_out, n, reps, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_ec_op_dev(int curve, int op, const void *d_a, const void *d_b, void *d_out, size_t n, void *stream) {
    { int irc = ensure_device_init(); if (irc) return irc; }
    int rc = ec_op_dispatch(curve, op, d_a, d_b, d_out, n, (cudaStream_t)stream);
    if (rc) return rc;
    AB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int b200_msm_last_timings(float *ms7, int *c, int *windows, unsigned long long *bucket_adds) { return msm_last_timings(ms7, c, windows, bucket_adds); }

}  // extern "C"
