"""ctypes loader for oracle/lattice.c (ORACLE - test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_lattice.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_rnnt_lattice.restype = ctypes.c_double
        _lib.oracle_rnnt_lattice.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        _lib.oracle_ctc_lattice.restype = ctypes.c_double
        _lib.oracle_ctc_lattice.argtypes = [dp, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_int, dp]
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def rnnt_lattice(lp2, T, U):
    """lp2 (T_max, U1_max, 2) -> cost, gamma2."""
    lp2 = np.ascontiguousarray(lp2, dtype=np.float64)
    g = np.empty_like(lp2)
    cost = lib().oracle_rnnt_lattice(_dp(lp2), lp2.shape[0], lp2.shape[1], int(T), int(U), _dp(g))
    return cost, g


def ctc_lattice(lp, y, blank=0):
    """lp (T,V) log-probs, y (U,) -> nll, occ (T,V)."""
    lp = np.ascontiguousarray(lp, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.int64)
    occ = np.empty_like(lp)
    nll = lib().oracle_ctc_lattice(_dp(lp), lp.shape[0], lp.shape[1],
                                   y.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), len(y), blank, _dp(occ))
    return nll, occ
