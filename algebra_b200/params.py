"""Host-side constants of the hot path (moduli, limb counts, generators) and Montgomery-limb marshalling.

Layout = arkworks' in-memory layout: `Fp` is N little-endian u64 limbs in Montgomery form, R = 2^(64N)
(ff/src/fields/models/fp/mod.rs:107-115); numpy arrays of dtype uint64 with a trailing axis of N (field
elements), 2N (affine points: x then y, (0,0) = identity) or 3N (Jacobian x, y, z)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MASK64 = (1 << 64) - 1


@dataclass(frozen=True)
class PrimeField:
    name: str
    fid: int                 # id used by b200_fp_op_dev
    modulus: int
    limbs: int               # N (u64)
    generator: int           # multiplicative generator (#[generator], e.g. curves/bls12_381/src/fields/fr.rs:5)

    @property
    def bits(self) -> int:
        return self.modulus.bit_length()

    @property
    def R(self) -> int:
        return (1 << (64 * self.limbs)) % self.modulus

    @property
    def two_adicity(self) -> int:
        s, t = 0, self.modulus - 1
        while t % 2 == 0:
            t //= 2
            s += 1
        return s

    @property
    def two_adic_root_of_unity(self) -> int:
        """GENERATOR^((p-1)/2^s)  (ff-macros/src/montgomery/mod.rs:44-55)"""
        return pow(self.generator, (self.modulus - 1) >> self.two_adicity, self.modulus)

    # ---- marshalling -----------------------------------------------------------------------
    def to_limbs(self, value: int) -> np.ndarray:
        """canonical int -> N Montgomery limbs"""
        m = (value % self.modulus) * self.R % self.modulus
        return np.array([(m >> (64 * i)) & MASK64 for i in range(self.limbs)], dtype=np.uint64)

    def from_limbs(self, limbs) -> int:
        m = sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(limbs, dtype=np.uint64).reshape(-1)))
        return m * pow(self.R, -1, self.modulus) % self.modulus

    def encode(self, values) -> np.ndarray:
        out = np.empty((len(values), self.limbs), dtype=np.uint64)
        for i, v in enumerate(values):
            out[i] = self.to_limbs(v)
        return out

    def decode(self, arr) -> list[int]:
        arr = np.asarray(arr, dtype=np.uint64).reshape(-1, self.limbs)
        return [self.from_limbs(r) for r in arr]


BLS12_381_FQ = PrimeField("bls12_381_fq", 0, 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB, 6, 2)
BLS12_381_FR = PrimeField("bls12_381_fr", 1, 52435875175126190479447740508185965837690552500527637822603658699938581184513, 4, 7)
BN254_FQ = PrimeField("bn254_fq", 2, 21888242871839275222246405745257275088696311157297823662689037894645226208583, 4, 3)
BN254_FR = PrimeField("bn254_fr", 3, 21888242871839275222246405745257275088548364400416034343698204186575808495617, 4, 5)


@dataclass(frozen=True)
class G1Curve:
    name: str
    cid: int                 # B200_CURVE_*
    ntt_field_id: int        # B200_FIELD_* of the scalar field
    fq: PrimeField
    fr: PrimeField
    coeff_b: object          # int (G1) or (c0, c1) (G2)
    generator: tuple         # affine (x, y), canonical ints (G2: each coordinate is (c0, c1))
    ext_degree: int = 1      # coordinates live in Fq (1) or Fq2 (2)

    @property
    def N(self) -> int:
        """u64 limbs per coordinate"""
        return self.fq.limbs * self.ext_degree


BLS12_381_G1 = G1Curve(
    "bls12_381_g1", 0, 0, BLS12_381_FQ, BLS12_381_FR, 4,
    (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
     1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569))
BN254_G1 = G1Curve("bn254_g1", 1, 1, BN254_FQ, BN254_FR, 3, (1, 2))
BLS12_381_G2 = G1Curve(
    "bls12_381_g2", 2, 0, BLS12_381_FQ, BLS12_381_FR, (4, 4),
    ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
      3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
     (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
      927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)), 2)
CURVES = {0: BLS12_381_G1, 1: BN254_G1, 2: BLS12_381_G2}
SCALAR_FIELDS = {0: BLS12_381_FR, 1: BN254_FR}   # keyed by B200_FIELD_*
