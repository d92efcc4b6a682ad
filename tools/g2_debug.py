#!/usr/bin/env python3
"""G2 MSM debugging matrix (GPU box): tiny inputs x window sizes, pass/fail against the oracle — isolates which stage of the
pipeline (accumulate / bucket reduce / partial sums / window combine) first produces a wrong group element."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algebra_b200 as ab
from algebra_b200 import variable_base as M
from oracle import pyoracle as O

G2, fr = O.BLS12_381_G2, O.BLS12_381_FR
P = [G2.mul(G2.G, k) for k in (5, 9, 11, 13)]
cases = [
    ("1 pt, s=1", [P[0]], [1]), ("1 pt, s=2", [P[0]], [2]), ("1 pt, s=3", [P[0]], [3]), ("1 pt, s=5", [P[0]], [5]),
    ("1 pt, s=8", [P[0]], [8]), ("1 pt, s=9", [P[0]], [9]), ("1 pt, s=64", [P[0]], [64]), ("1 pt, s=2^40+1", [P[0]], [(1 << 40) + 1]),
    ("1 pt, s=r-1", [P[0]], [fr.p - 1]), ("1 pt, s=r-3", [P[0]], [fr.p - 3]),
    ("2 pts, (1,1)", P[:2], [1, 1]), ("2 pts, (1,2)", P[:2], [1, 2]), ("2 pts, (2,1)", P[:2], [2, 1]), ("2 pts, (1,3)", P[:2], [1, 3]),
    ("2 pts, (3,2)", P[:2], [3, 2]), ("2 pts, (3,r-2)", P[:2], [3, fr.p - 2]), ("2 pts, (1,r-1)", P[:2], [1, fr.p - 1]),
    ("2 pts same bucket (3,3)", P[:2], [3, 3]), ("3 pts (1,2,3)", P[:3], [1, 2, 3]), ("4 pts (1,2,3,4)", P, [1, 2, 3, 4]),
    ("4 pts random", P, [0x1234567890abcdef123, 0xfedcba9876543210, 7, fr.p - 12345]),
]
for curve, name in ((0, "G1"), (2, "G2")):
    cv = O.CURVES[curve]
    print("=====", name)
    for c in (0, 1, 2, 3, 4, 7):
        M.set_window(c)
        row = []
        for label, pts, sc in cases:
            if curve == 0:
                pts = [cv.mul(cv.G, k) for k in (5, 9, 11, 13)][:len(pts)]
                want = cv.encode_affine([cv.jac_to_affine(None)])[0] * 0
                acc = None
                for Pt, s in zip(pts, sc):
                    acc = cv.add(acc, cv.mul(Pt, s))
                want = cv.encode_affine([acc])[0]
            else:
                want = cv.encode_affine([cv.naive_msm(pts, sc)])[0]
            got = ab.into_affine(curve, ab.msm(curve, cv.encode_affine(pts), fr.encode(sc)))
            row.append("." if (got == want).all() else "X")
        print("c=%d  %s" % (c, "".join(row)))
    M.set_window(0)
print("legend:", "; ".join("%d=%s" % (i, l) for i, (l, _, _) in enumerate(cases)))
