"""ctypes binding of the C ABI declared in include/neusky_b200.h.

There is no fallback: if the shared library is missing or a symbol is absent the import of an
op raises.  ``neusky_b200.build.build()`` (or ``python -m neusky_b200.build``) creates it.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libneusky_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "neusky_b200.h")

_lib = None


def declared_symbols():
    """Every ``nsk_*`` function declared in the public header."""
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(nsk_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"neusky_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m neusky_b200.build`. "
                "There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        lib.nsk_last_error.restype = ctypes.c_char_p
        for name in ("nsk_reni_weights_floats", "nsk_ddf_simt_weights_floats", "nsk_ddf_tc_weights_bytes", "nsk_sdf_simt_weights_floats", "nsk_sdf_tc_weights_bytes", "nsk_ddf_tc2_weights_bytes", "nsk_reni_bwd_weights_floats", "nsk_reni_fused_weights_bytes", "nsk_reni_bwd_workspace_floats", "nsk_proposal_mlp_floats"):
            if hasattr(lib, name):
                getattr(lib, name).restype = ctypes.c_int64
        _lib = lib
    return _lib


# CUDA kernels launched per successful ABI call (1 unless listed); bench.py reports the running total
# as ``gpu_launches``.
KERNELS_PER_CALL = {"nsk_reni_decode_fwd": 2, "nsk_reni_decode_rows_fwd": 2, "nsk_reni_decode_bwd": 3}
launches = 0


def check(status: int, what: str) -> None:
    global launches
    if status != 0:
        raise RuntimeError(f"{what} failed: {load().nsk_last_error().decode()}")
    launches += KERNELS_PER_CALL.get(what, 1)
