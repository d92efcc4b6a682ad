"""Window and bucket-slice geometry of the MSM (algebra_b200/csrc/msm_common.cuh: make_geom), compiled for the host: the window
count follows the reference's digits_count (ec/src/scalar_mul/variable_base/mod.rs:451-453), the signed-digit bucket count is
2^(c-1) with an unsigned top window, and the S bucket slices of b200_set_msm_bucket_slice partition every window."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def geom(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("geom") / "libgeom_selftest.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fpermissive", "-w", "-x", "c++", "-I" + cuda_inc, "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tools", "geom_selftest.cpp")])
    lib = ctypes.CDLL(so)

    def call(c, bits, s=0, S=1):
        out = (ctypes.c_uint * 8)()
        lib.selftest_geom(c, bits, s, S, out)
        return dict(zip(("c", "W", "top_bits", "nb", "nb_top", "total", "off", "off_top"), out))
    return call


@pytest.mark.parametrize("bits", [255, 254, 64, 32, 16, 8, 1])
def test_whole_window_geometry(geom, bits):
    for c in range(1, min(bits, 24) + 1):
        g = geom(c, bits)
        W = (bits + c - 1) // c                          # digits_count
        assert g["W"] == W and g["top_bits"] == bits - (W - 1) * c and 1 <= g["top_bits"] <= c
        assert g["nb"] == 1 << (c - 1) and g["nb_top"] == 1 << g["top_bits"]      # signed digits below, unsigned top digit
        assert g["total"] == (W - 1) * g["nb"] + g["nb_top"] and g["off"] == 0 and g["off_top"] == 0


@pytest.mark.parametrize("S", [2, 3, 5, 8, 64])
def test_bucket_slices_partition_every_window(geom, S):
    for bits in (255, 254, 64, 8):
        for c in range(1, min(bits, 24) + 1):
            whole = geom(c, bits)
            parts = [geom(c, bits, s, S) for s in range(S)]
            for key_n, key_off, full in (("nb", "off", whole["nb"]), ("nb_top", "off_top", whole["nb_top"])):
                pos = 0
                for p in parts:                           # contiguous, in order, nothing lost or doubled
                    assert p[key_off] == pos or p[key_n] == 0
                    pos += p[key_n]
                assert pos == full
                sizes = [p[key_n] for p in parts]
                if full >= S:
                    assert max(sizes) - min(sizes) <= 1   # balanced
                else:
                    assert sizes[0] == full and not any(sizes[1:])   # a window with fewer buckets than slices goes whole to slice 0
            for p in parts:
                assert p["W"] == whole["W"] and p["top_bits"] == whole["top_bits"]
                assert p["total"] == (p["W"] - 1) * p["nb"] + p["nb_top"]
