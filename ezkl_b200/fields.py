"""Host-side BN254 scalar-field constants and wire-format conversion (python ints <-> 4 x u64 Montgomery limbs).

Host logic of the product (the analogue of what halo2's EvaluationDomain::new does with field ops on the CPU); the bulk
arithmetic all happens on the device.  Constants: SURVEY.md Appendix A.
"""
import numpy as np

FR_MODULUS = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
FQ_MODULUS = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
FR_S = 28
FR_ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c   # 7^((r-1)/2^28)
FR_ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23           # Fr::ZETA
_R = 1 << 256
_MASK = (1 << 64) - 1


def fr_to_limbs(x: int) -> np.ndarray:
    m = (x % FR_MODULUS) * _R % FR_MODULUS
    return np.array([(m >> (64 * i)) & _MASK for i in range(4)], dtype=np.uint64)


def fr_from_limbs(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(4)
    m = sum(int(a[i]) << (64 * i) for i in range(4))
    return m * pow(_R, -1, FR_MODULUS) % FR_MODULUS


def fr_inv(x: int) -> int:
    return pow(x, -1, FR_MODULUS)


def fq_to_limbs(x: int) -> np.ndarray:
    m = (x % FQ_MODULUS) * _R % FQ_MODULUS
    return np.array([(m >> (64 * i)) & _MASK for i in range(4)], dtype=np.uint64)
