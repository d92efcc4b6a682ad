"""The C++ host mirror (include/algebra_b200.hpp) compiled with g++ and linked against libalgebra_b200.so: its domain
parameters must equal the oracle's, `new` must refuse sizes past TWO_ADICITY and `msm` must return Err(min_len)."""
import os
import subprocess

import pytest

from oracle import pyoracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_test")
    libdir = os.path.join(ROOT, "algebra_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", out,
                           "-L" + libdir, "-lalgebra_b200", "-Wl,-rpath," + libdir])
    return out


def parse(text):
    return dict(line.split(" ", 1) for line in text.strip().splitlines())


def test_cpp_mirror_host_logic(binary):
    kv = parse(subprocess.check_output([binary], text=True))
    fr = O.BLS12_381_FR
    dom = O.Radix2Domain(fr, 1000)
    assert kv["size"] == "1024 log 10"
    mont = lambda v: "%064x" % fr.to_mont(v)
    assert kv["group_gen"] == mont(dom.group_gen) and kv["group_gen_inv"] == mont(dom.group_gen_inv) and kv["size_inv"] == mont(dom.size_inv)
    co = O.Radix2Domain(fr, 1000, 1024)
    assert kv["offset"] == mont(1024) and kv["offset_inv"] == mont(co.offset_inv) and kv["offset_pow_size"] == mont(co.offset_pow_size)
    assert kv["element5"] == mont(co.element(5))
    assert kv["too_big"] == "0" and kv["bn_too_big"] == "0"
    assert kv["mismatch_err"] == "1 min_len 2"


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(binary):
    kv = parse(subprocess.check_output([binary, "gpu"], text=True))
    assert kv["gpu_msm_identity"] == "1" and kv["gpu_ntt_roundtrip"] == "1"
    # single-device, all-devices (b200_msm_sw_g1_multi), streaming and resident-bases entry points agree with (sum i) * G
    assert int(kv["gpu_devices"]) >= 1
    for k in ("gpu_msm_single", "gpu_msm_multi", "gpu_msm_chunks", "gpu_msm_slices", "gpu_msm_resident", "gpu_resident_mismatch"):
        assert kv[k] == "1", k
