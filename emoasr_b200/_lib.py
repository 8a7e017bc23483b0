"""ctypes binding of libemoasr_b200.so (include/emoasr_b200.h).

There is NO CPU fallback: if the library is missing or cannot be loaded every op raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EMOASR_B200_LIB: an instrumented copy of the library (python -m emoasr_b200.build --prof; tools/ only)
LIB_PATH = os.environ.get("EMOASR_B200_LIB") or os.path.join(_HERE, "lib", "libemoasr_b200.so")

OP_RNNT_JOINT_FWD, OP_RNNT_JOINT_BWD, OP_CTC, OP_CTC_HEAD, OP_RNNT_JOINT_FULL = 0, 1, 2, 3, 4
PREC_FP32, PREC_BF16 = 0, 1
ABI_VERSION = 10

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_SZ = _c.c_size_t

_SIGNATURES = {
    "emo_abi_version": (_I, []),
    "emo_last_error_string": (_c.c_char_p, []),
    "emo_workspace_bytes": (_SZ, [_I] * 7),
    "emo_launch_count": (_I, [_I] * 7),
    "emo_rnnt_joint_supported": (_I, [_I] * 6),
    "emo_rnnt_lattice_fwd_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "emo_rnnt_dense_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "emo_rnnt_dense_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "emo_rnnt_joint_fwd": (_I, [_P] * 7 + [_I] * 7 + [_P, _P, _P, _SZ, _P]),
    "emo_rnnt_joint_bwd": (_I, [_P] * 12 + [_I] * 7 + [_P, _P, _P, _P, _P, _SZ, _P]),
    "emo_rnnt_joint_full_supported": (_I, [_I] * 7),
    "emo_rnnt_joint_full_workspace_bytes": (_SZ, [_I] * 8),
    "emo_rnnt_joint_full_fwd": (_I, [_P] * 11 + [_I] * 8 + [_P, _P, _P, _SZ, _P]),
    "emo_rnnt_joint_full_bwd": (_I, [_P] * 10 + [_I] * 8 + [_P] * 9 + [_SZ, _P]),
    "emo_rnnt_align": (_I, [_P] * 4 + [_I] * 3 + [_P, _P]),
    "emo_ctc_fwd": (_I, [_P] * 4 + [_I] * 6 + [_P, _P, _P, _P, _P]),
    "emo_ctc_bwd": (_I, [_P] * 8 + [_I] * 6 + [_P, _I, _P, _P]),
    "emo_ctc_align_workspace_bytes": (_SZ, [_I] * 3),
    "emo_ctc_align": (_I, [_P] * 4 + [_I] * 5 + [_P, _P, _SZ, _P]),
    "emo_ctc_head_supported": (_I, [_I] * 5),
    "emo_ctc_head_workspace_bytes": (_SZ, [_I] * 6),
    "emo_ctc_head_fwd": (_I, [_P] * 6 + [_I] * 7 + [_P] * 6 + [_SZ, _P]),
    "emo_ctc_head_bwd": (_I, [_P] * 11 + [_I] * 6 + [_P] * 4 + [_SZ, _P]),
    "emo_rnnt_step_workspace_bytes": (_SZ, [_I]),
    "emo_rnnt_joint_step": (_I, [_P] * 5 + [_I] * 3 + [_P] * 3 + [_SZ, _P]),
    "emo_rnnt_greedy_step": (_I, [_P] * 5 + [_I] * 6 + [_P] * 5 + [_I] + [_P] * 4 + [_SZ, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


class EmoLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmoLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m emoasr_b200.build` "
            "(there is no CPU or eager fallback for this path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.emo_abi_version() != ABI_VERSION:
        raise EmoLibraryError(f"ABI mismatch: library {lib.emo_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().emo_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def workspace_bytes(op, precision, B, T, U1, J, V):
    return int(load().emo_workspace_bytes(op, precision, B, T, U1, J, V))


def launch_count(op, precision, B, T, U1, J, V):
    return int(load().emo_launch_count(op, precision, B, T, U1, J, V))
