// CPU test of the C++ host mirror (include/algebra_b200.hpp): domain construction / coset / element logic and the MSM
// length-mismatch contract.  Prints values that tests/test_cpp_mirror.py compares with the oracle.  No GPU needed for the
// printed part; with `gpu` as argv[1] it also runs one tiny MSM + NTT through the mirror.
#include <cstdio>
#include <cstring>
#include "../../include/algebra_b200.hpp"
using namespace ab200;
static void print(const char *name, const Fr &v) {
    printf("%s %016lx%016lx%016lx%016lx\n", name, (unsigned long)v[3], (unsigned long)v[2], (unsigned long)v[1], (unsigned long)v[0]);
}
int main(int argc, char **argv) {
    auto d = Radix2EvaluationDomain<Bls12_381G1>::make(1000).value();
    printf("size %lu log %u\n", (unsigned long)d.size, d.log_size_of_group);
    print("group_gen", d.group_gen); print("group_gen_inv", d.group_gen_inv); print("size_inv", d.size_inv);
    Fr off = d.size_as_field_element;  // 1024 as a field element: a non-trivial offset
    auto c = d.get_coset(off).value();
    print("offset", c.offset); print("offset_inv", c.offset_inv); print("offset_pow_size", c.offset_pow_size); print("element5", c.element(5));
    printf("too_big %d\n", (int)Radix2EvaluationDomain<Bls12_381G1>::make(((size_t)1 << 32) + 1).has_value());
    printf("bn_too_big %d\n", (int)Radix2EvaluationDomain<Bn254G1>::make(((size_t)1 << 28) + 1).has_value());
    std::vector<Affine<Bls12_381G1>> bases(3); std::vector<Fr> scalars(2);
    auto r = VariableBaseMSM<Bls12_381G1>::msm(bases, scalars);
    printf("mismatch_err %d min_len %zu\n", (int)std::holds_alternative<size_t>(r), std::get<size_t>(r));
    if (argc > 1 && !strcmp(argv[1], "gpu")) {
        scalars.resize(3);
        auto p = std::get<0>(VariableBaseMSM<Bls12_381G1>::msm(bases, scalars));   // identity bases -> identity
        printf("gpu_msm_identity %d\n", (int)(p.z == Fq<Bls12_381G1>{}));
        // every MSM entry point through the header only, on all devices of the box: bases {G x 5000}, scalars i -> (sum i) * G
        {
            using C = Bls12_381G1;
            Affine<C> G{};
            for (int i = 0; i < 6; i++) {
                G.x[i] = (uint64_t)BlsFq::GEN_X(2 * i) | ((uint64_t)BlsFq::GEN_X(2 * i + 1) << 32);
                G.y[i] = (uint64_t)BlsFq::GEN_Y(2 * i) | ((uint64_t)BlsFq::GEN_Y(2 * i + 1) << 32);
            }
            auto fr_of = [](uint64_t k) {
                uint32_t l[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
                Fp<BlsFr>::to_mont(l, l);
                Fr r;
                for (int i = 0; i < 4; i++) r[i] = (uint64_t)l[2 * i] | ((uint64_t)l[2 * i + 1] << 32);
                return r;
            };
            const size_t n = 5000;
            std::vector<Affine<C>> B(n, G);
            std::vector<Fr> S(n);
            uint64_t total = 0;
            for (size_t i = 0; i < n; i++) { S[i] = fr_of(i + 1); total += i + 1; }
            auto aff = [](const Projective<C> &p) { return VariableBaseMSM<C>::into_affine(p); };
            auto same = [](const Affine<C> &a, const Affine<C> &b) { return a.x == b.x && a.y == b.y; };
            const Affine<C> want = aff(VariableBaseMSM<C>::msm_unchecked({G}, {fr_of(total)}));
            const int devices = b200_device_count();
            printf("gpu_devices %d\n", devices);
            printf("gpu_msm_single %d\n", (int)same(aff(VariableBaseMSM<C>::msm_unchecked(B, S)), want));
            printf("gpu_msm_multi %d\n", (int)same(aff(VariableBaseMSM<C>::msm_unchecked_multi(B, S, devices)), want));
            printf("gpu_msm_chunks %d\n", (int)same(aff(VariableBaseMSM<C>::msm_chunks(B, S, 1200)), want));
            {   // three bucket slices over the same inputs, summed with b200_g1_sum
                Projective<C> parts[3], sum{};
                for (int i = 0; i < 3; i++) parts[i] = VariableBaseMSM<C>::msm_bucket_slice(B, S, i, 3);
                const int rc = b200_g1_sum(C::ID, reinterpret_cast<const uint64_t *>(parts), 3, reinterpret_cast<uint64_t *>(&sum));
                printf("gpu_msm_slices %d\n", (int)(rc == 0 && same(aff(sum), want)));
            }
            ResidentBases<C> srs(B, devices);
            printf("gpu_msm_resident %d\n", (int)same(aff(std::get<0>(srs.msm(S))), want));
            S.pop_back();
            auto r2 = srs.msm(S);
            printf("gpu_resident_mismatch %d\n", (int)(std::holds_alternative<size_t>(r2) && std::get<size_t>(r2) == n - 1));
        }
        std::vector<Fr> v(8, d.size_inv);
        auto dom8 = Radix2EvaluationDomain<Bls12_381G1>::make(8).value();
        auto w = dom8.ifft(dom8.fft(v));
        printf("gpu_ntt_roundtrip %d\n", (int)(w == v));
    }
    return 0;
}
