"""secp256k1-zkp_b200 -- B200-native batch secp256k1 engine behind the libsecp256k1-zkp C API.

The product is the C-ABI shared library ``libsecp256k1_b200.so`` (hand-written sm_100a CUDA, see
``csrc/``).  This package is only the thin host-side mirror used by the tests and the benchmark:
``host.engine.Engine`` wraps the C ABI declared in ``include/secp256k1_b200.h`` through ctypes.
There is no CPU compute path: if the library or a B200 is missing, constructing an Engine raises.

(The directory name contains a hyphen, as the task specifies; import it as ``secp256k1_zkp_b200``,
a two-line alias package at the repository root.)
"""
from .host.engine import (Engine, MultiEngine, B200Error, load_library, library_path,  # noqa: F401
                          PK_OPAQUE64, PK_XONLY32, PK_COMPRESSED33, SIG_OPAQUE64, SIG_COMPACT64,
                          FE_MUL, FE_SQR, FE_ADD, FE_SUB, FE_HALF, FE_INV, FE_SQRT, FE_NEG,
                          FE_INV_FERMAT, FE_IS_SQUARE, FE_IS_SQUARE_SQRT, FE_MUL_PORTABLE, FE_SQR_PORTABLE)
