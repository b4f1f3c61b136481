"""mat_mul_b200 -- B200-native (sm_100a) implementation of the data-parallel hot
path of kurtosis/mat_mul's TensorGame environment: batched transition,
synthetic-demonstration generation and change-of-basis augmentation, behind
the reference's Python API (utils / datasets / act) and a C ABI
(include/tensorgame.h).  No CPU fallback."""
from ._lib import FLAG_NULL, FLAG_RANGE, FLAG_TERMINAL, TensorGameError, build  # noqa: F401

__version__ = "0.1.0"
