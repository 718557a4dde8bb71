"""argtypes/restype declarations for the non-GEMM entry points of libduett_b200.so (include/duett_b200.h)."""
from __future__ import annotations

import ctypes as C


def declare(L: C.CDLL) -> None:
    pass
