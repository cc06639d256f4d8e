"""ctypes binding of libcats_b200.so (the C-ABI declared in include/cats_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly with the build
command, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcats_b200.so"
if os.environ.get("CATS_B200_LIB"):  # development aid: load an experimental build of the same ABI
    LIB_PATH = Path(os.environ["CATS_B200_LIB"]).resolve()

N_PARAMS = 34
N_SCAL = 32
N_STATS = 16
MAX_Z = 8

FLAG_BURNIN, FLAG_LOG, FLAG_AVG_PRICE, FLAG_REWARD_RMS, FLAG_STRICT = 1, 2, 4, 8, 16
STATUS_DEPVAL_DIV0, STATUS_PROFIT_NAN, STATUS_Z_MISMATCH, STATUS_TAPE_RANGE = 1, 2, 4, 8
OK, EINVAL, ECUDA, ESMEM = 0, -1, -2, -3

STAT_NAMES = ["Y_real", "wb", "Un", "bankruptcy_rate", "avg_price", "Y_nominal_tot", "gdp_deflator",
              "inflationRate", "consumption", "totK", "dividendsB", "profitsB", "GB", "deficitGDP",
              "bonds", "price_k"]

# every symbol include/cats_b200.h and include/cats_b200_debug.h declare (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "cats_abi_version", "cats_last_error", "cats_blob_bytes", "cats_field_info", "cats_field_name",
    "cats_field_stride",
    "cats_tape_bytes", "cats_tape_offsets", "cats_init", "cats_reset_masked", "cats_step", "cats_q_step", "cats_host_scratch_bytes",
    "cats_step_host", "cats_host_block_offsets", "cats_workset_bytes", "cats_workset_field_info", "cats_launch_info",
    "cats_launch_count", "cats_debug_phase_cycles", "cats_debug_set_group", "cats_debug_set_static",
    "cats_debug_set_warps", "cats_debug_set_pipe", "cats_debug_set_host_zero_copy", "cats_launch_shape",
]


ABI_VERSION = 2


class StepArgs(C.Structure):
    """struct cats_step_args (include/cats_b200.h)."""
    _fields_ = [
        ("d_blobs", C.c_void_p), ("B", C.c_int64), ("W", C.c_int32), ("F", C.c_int32),
        ("N", C.c_int32), ("d_params", C.c_void_p), ("params_stride", C.c_int64),
        ("seed", C.c_uint64), ("economy_base", C.c_uint64), ("n_periods", C.c_int32),
        ("flags", C.c_uint32), ("n_agents", C.c_int32), ("d_actions", C.c_void_p),
        ("d_action_mask", C.c_void_p), ("bankruptcy_reward", C.c_double), ("d_obs", C.c_void_p),
        ("d_reward", C.c_void_p), ("d_terminated", C.c_void_p), ("d_status", C.c_void_p),
        ("d_stats", C.c_void_p), ("d_tape", C.c_void_p), ("tape_z", C.c_int32 * 3),
        ("max_z", C.c_int32), ("stop_phase", C.c_int32), ("d_debug", C.c_void_p), ("stream", C.c_void_p),
        ("d_active", C.c_void_p), ("z_all", C.c_int32 * 3),
    ]


class QArgs(C.Structure):
    """struct cats_q_args (include/cats_b200.h)."""
    _fields_ = [
        ("d_Q", C.c_void_p), ("B", C.c_int64), ("A", C.c_int32), ("n", C.c_int32 * 4),
        ("obs_lo", C.c_double * 2), ("obs_hi", C.c_double * 2), ("act_lo", C.c_double * 2),
        ("act_hi", C.c_double * 2), ("alpha", C.c_double), ("gamma", C.c_double), ("epsilon", C.c_double),
        ("seed", C.c_uint64), ("economy_base", C.c_uint64), ("step", C.c_uint32), ("do_train", C.c_int32),
        ("do_act", C.c_int32), ("d_obs", C.c_void_p), ("d_reward", C.c_void_p), ("d_terminated", C.c_void_p),
        ("d_last", C.c_void_p), ("d_actions", C.c_void_p), ("d_active", C.c_void_p), ("stream", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """Load the in-tree library; raise ImportError (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            f"(`bash {_PKG / 'csrc' / 'build.sh'}` or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "r_mabm_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    L.cats_abi_version.restype = C.c_int
    L.cats_last_error.restype = C.c_char_p
    L.cats_blob_bytes.restype = C.c_size_t
    L.cats_blob_bytes.argtypes = [C.c_int] * 3
    L.cats_workset_bytes.restype = C.c_size_t
    L.cats_workset_bytes.argtypes = [C.c_int] * 3
    L.cats_field_info.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.cats_workset_field_info.argtypes = L.cats_field_info.argtypes
    L.cats_field_stride.argtypes = [C.c_char_p]
    L.cats_field_stride.restype = C.c_int
    L.cats_field_name.restype = C.c_char_p
    L.cats_field_name.argtypes = [C.c_int]
    L.cats_tape_bytes.restype = C.c_size_t
    L.cats_tape_bytes.argtypes = [C.c_int] * 6
    L.cats_tape_offsets.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_uint64)]
    L.cats_init.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                            C.c_void_p]
    L.cats_reset_masked.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p]
    L.cats_step.argtypes = [C.POINTER(StepArgs)]
    L.cats_q_step.argtypes = [C.POINTER(QArgs)]
    L.cats_host_scratch_bytes.restype = C.c_size_t
    L.cats_host_scratch_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.cats_step_host.argtypes = [C.POINTER(StepArgs), C.c_void_p]
    L.cats_host_block_offsets.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.cats_launch_info.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int32)] * 4
    L.cats_launch_shape.argtypes = [C.c_int] * 3 + [C.c_int64] + [C.POINTER(C.c_int32)] * 3
    L.cats_launch_count.restype = C.c_uint64
    L.cats_debug_phase_cycles.argtypes = [C.c_void_p]
    L.cats_debug_phase_cycles.restype = None
    L.cats_debug_set_group.argtypes = [C.c_int]
    L.cats_debug_set_group.restype = None
    L.cats_debug_set_static.argtypes = [C.c_int]
    L.cats_debug_set_static.restype = None
    L.cats_debug_set_warps.argtypes = [C.c_int]
    L.cats_debug_set_warps.restype = None
    L.cats_debug_set_pipe.argtypes = [C.c_int]
    L.cats_debug_set_pipe.restype = None
    L.cats_debug_set_host_zero_copy.argtypes = [C.c_int]
    L.cats_debug_set_host_zero_copy.restype = None
    if L.cats_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.cats_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    """Map C-ABI error codes onto the reference's exception conventions (ValueError for bad
    arguments, /root/reference/rmabm/environments/cats.py:141-154,350-353)."""
    if rc == OK:
        return
    msg = load().cats_last_error().decode()
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ESMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"cats_b200 error {rc}: {msg}")


def field_info(W: int, F: int, N: int, name: str, workset: bool = False):
    off, cnt, dt = C.c_uint64(0), C.c_int32(0), C.c_int32(0)
    fn = load().cats_workset_field_info if workset else load().cats_field_info
    check(fn(W, F, N, name.encode(), C.byref(off), C.byref(cnt), C.byref(dt)))
    return int(off.value), int(cnt.value), int(dt.value)


def field_stride(name: str) -> int:
    """Element stride of a state field (2 for the interleaved household records, else 1)."""
    st = int(load().cats_field_stride(name.encode()))
    if st < 1:
        raise ValueError(f"unknown field '{name}'")
    return st


def field_names() -> list[str]:
    out, i = [], 0
    L = load()
    while True:
        n = L.cats_field_name(i)
        if n is None:
            return out
        out.append(n.decode())
        i += 1
