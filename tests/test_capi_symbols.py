"""CPU checks of the boundary: the shared library loads and exports every symbol include/algebra_b200.h declares, the
ctypes signature table covers exactly that set, and compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "algebra_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from algebra_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    assert sorted(_lib.SIGNATURES) == syms
    lib = _lib.lib()
    for s in syms:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.b200_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import algebra_b200 as ab
    b = np.zeros((4, 12), dtype=np.uint64)
    s = np.zeros((4, 4), dtype=np.uint64)
    with pytest.raises(ab._lib.B200Error):
        ab.msm(0, b, s)
    with pytest.raises(ab._lib.B200Error):
        ab.Radix2EvaluationDomain.new(0, 8).fft(np.zeros((8, 4), dtype=np.uint64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "algebra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f


def test_host_mirror_domain_logic():
    """Radix2EvaluationDomain mirror: constructor / getters against the oracle's domain parameters (no GPU needed)."""
    import algebra_b200 as ab
    from oracle import coracle as C
    from oracle import pyoracle as O
    for fid, ofid, fr in ((0, 1, O.BLS12_381_FR), (1, 3, O.BN254_FR)):
        for log_n in (0, 1, 5, 16, 24, fr.two_adicity):
            d = ab.Radix2EvaluationDomain.new(fid, 1 << log_n)
            g, gi, ni = C.domain_params(ofid, log_n)
            assert (d.group_gen() == g).all() and (d.group_gen_inv() == gi).all() and (d.size_inv() == ni).all()
            assert d.size == 1 << log_n and d.log_size_of_group == log_n
        assert ab.Radix2EvaluationDomain.new(fid, (1 << fr.two_adicity) + 1) is None
        d = ab.Radix2EvaluationDomain.new(fid, 100)
        assert d.size == 128
        co = d.get_coset(fr.generator)
        od = O.Radix2Domain(fr, 128, fr.generator)
        assert fr.decode(co.coset_offset_pow_size())[0] == od.offset_pow_size
        assert fr.decode(co.coset_offset_inv())[0] == od.offset_inv
        assert [fr.decode(e)[0] for e in list(co.elements())[:5]] == [od.element(i) for i in range(5)]
        assert fr.decode(co.element(77))[0] == od.element(77)
