#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the hot path into small fixtures.

Run in the BUILD container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Sources (all under /root/reference; parsed as DATA, never imported or copied as code):
  * curves/bls12_381/src/curves/tests/g1_uncompressed_valid_test_vectors.dat
      [0*G, 1*G, ..., 999*G] in zkcrypto big-endian encoding
      (format: curves/bls12_381/src/curves/util.rs:61-93,139-171;
       checked by curves/bls12_381/src/curves/tests/mod.rs:70-111)
  * test-curves/src/bls12_381/fq.rs:20-106   INV, R, R2, GENERATOR(mont) for Fq
  * test-curves/src/bls12_381/fr.rs:12-28    INV, MODULUS limbs for Fr
  * curves/bls12_381/src/fields/tests.rs:21-32      -1 in Montgomery form
  * curves/bls12_381/src/fields/tests.rs:1231-1345  Fq2 square / mul KATs
      (each Fq2 product pins 3-4 Fq products: (a0b0 - a1b1) + (a0b1 + a1b0)u)

Outputs (committed):
  tests/golden/bls12_381_g1_multiples.npy   uint64 (1000, 12): canonical LE limbs x|y, (0,0)=inf
  tests/golden/bls12_381_g2_multiples.npy   uint64 (1000, 24): x.c0|x.c1|y.c0|y.c1 of i*G2 from
      curves/bls12_381/src/curves/tests/g2_uncompressed_valid_test_vectors.dat
  tests/golden/kats.json                    constants + Fq2 KATs as hex strings
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def limbs_to_int(limbs):
    return sum(v << (64 * i) for i, v in enumerate(limbs))


def parse_bigint_blocks(text):
    """all `[0x.., 0x.., ...]` limb lists in order of appearance"""
    out = []
    for m in re.finditer(r"\[\s*((?:0x[0-9a-fA-F_]+\s*,?\s*)+)\]", text):
        limbs = [int(x.replace("_", ""), 16) for x in re.findall(r"0x[0-9a-fA-F_]+", m.group(1))]
        out.append(limbs)
    return out


def g1_table():
    raw = open(f"{REF}/curves/bls12_381/src/curves/tests/g1_uncompressed_valid_test_vectors.dat", "rb").read()
    assert len(raw) == 96000
    tab = np.zeros((1000, 12), dtype=np.uint64)
    for i in range(1000):
        rec = bytearray(raw[96 * i:96 * (i + 1)])
        flags = rec[0] >> 5
        rec[0] &= 0x1F
        compressed, infinity = bool(flags & 4), bool(flags & 2)
        assert not compressed
        x = int.from_bytes(rec[:48], "big")
        y = int.from_bytes(rec[48:], "big")
        if infinity:
            assert x == 0 and y == 0 and i == 0
        for j in range(6):
            tab[i, j] = (x >> (64 * j)) & (2**64 - 1)
            tab[i, 6 + j] = (y >> (64 * j)) & (2**64 - 1)
    return tab


def g2_table():
    """[0*G2 .. 999*G2], zkcrypto order x.c1 | x.c0 | y.c1 | y.c0 big-endian (curves/bls12_381/src/curves/util.rs:211-256) ->
    uint64 (1000, 24): canonical LE limbs of x.c0 | x.c1 | y.c0 | y.c1 (arkworks' QuadExtField order), all-zero = infinity"""
    raw = open(f"{REF}/curves/bls12_381/src/curves/tests/g2_uncompressed_valid_test_vectors.dat", "rb").read()
    assert len(raw) == 192000
    tab = np.zeros((1000, 24), dtype=np.uint64)
    for i in range(1000):
        rec = bytearray(raw[192 * i:192 * (i + 1)])
        flags = rec[0] >> 5
        rec[0] &= 0x1F
        assert not (flags & 4)
        xc1, xc0, yc1, yc0 = (int.from_bytes(rec[48 * k:48 * (k + 1)], "big") for k in range(4))
        if flags & 2:
            assert xc1 == xc0 == yc1 == yc0 == 0 and i == 0
        for k, v in enumerate((xc0, xc1, yc0, yc1)):
            for j in range(6):
                tab[i, 6 * k + j] = (v >> (64 * j)) & (2**64 - 1)
    return tab


def fq_constants():
    src = open(f"{REF}/test-curves/src/bls12_381/fq.rs").read()
    inv = int(re.search(r"FqConfig::INV,\s*(0x[0-9a-f_]+)", src).group(1).replace("_", ""), 16)
    body = src[src.index("fn test_constants"):]
    blocks = parse_bigint_blocks(body)
    # order in the file: R, R2, TRACE, MODULUS_MINUS_ONE_DIV_TWO, TRACE_MINUS_ONE_DIV_TWO, GENERATOR
    names = ["R", "R2", "TRACE", "MODULUS_MINUS_ONE_DIV_TWO", "TRACE_MINUS_ONE_DIV_TWO", "GENERATOR_MONT"]
    d = {"INV": hex(inv)}
    for n, b in zip(names, blocks):
        assert len(b) == 6
        d[n] = hex(limbs_to_int(b))
    return d


def fr_constants():
    src = open(f"{REF}/test-curves/src/bls12_381/fr.rs").read()
    inv = int(re.search(r"FrConfig::INV,\s*(0x[0-9a-f_]+)", src).group(1).replace("_", ""), 16)
    mod = parse_bigint_blocks(src[src.index("fn test_modulus"):])[0]
    return {"INV": hex(inv), "MODULUS": hex(limbs_to_int(mod))}


def fq_field_kats():
    src = open(f"{REF}/curves/bls12_381/src/fields/tests.rs").read()
    out = {}
    neg = src[src.index("fn test_negative_one"):src.index("fn test_frob_coeffs")]
    out["neg_one_mont"] = hex(limbs_to_int(parse_bigint_blocks(neg)[0]))

    sq = src[src.index("fn test_fq2_squaring"):src.index("fn test_fq2_mul")]
    b = [limbs_to_int(x) for x in parse_bigint_blocks(sq) if len(x) == 6]
    # a0, a1, r0, r1
    out["fq2_square"] = {"a": [hex(b[0]), hex(b[1])], "r": [hex(b[2]), hex(b[3])]}

    mu = src[src.index("fn test_fq2_mul"):src.index("fn test_fq2_inverse")]
    b = [limbs_to_int(x) for x in parse_bigint_blocks(mu) if len(x) == 6]
    out["fq2_mul"] = {"a": [hex(b[0]), hex(b[1])], "b": [hex(b[2]), hex(b[3])], "r": [hex(b[4]), hex(b[5])]}

    iv = src[src.index("fn test_fq2_inverse"):src.index("fn test_fq2_addition")]
    b = [limbs_to_int(x) for x in parse_bigint_blocks(iv) if len(x) == 6]
    out["fq2_inverse"] = {"a": [hex(b[0]), hex(b[1])], "r": [hex(b[2]), hex(b[3])]}
    return out


def main():
    np.save(os.path.join(HERE, "bls12_381_g1_multiples.npy"), g1_table())
    np.save(os.path.join(HERE, "bls12_381_g2_multiples.npy"), g2_table())
    kats = {
        "source": "arkworks-rs/algebra v0.6.0 (af564e48) — see make_golden.py docstring for file:line",
        "bls12_381_fq": fq_constants(),
        "bls12_381_fr": fr_constants(),
        "bls12_381_fq_field": fq_field_kats(),
    }
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
