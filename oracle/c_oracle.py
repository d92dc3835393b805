"""ctypes loader for oracle/libqaoa_oracle.so (C/OpenMP restatement) -- TEST / BASELINE ONLY."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libqaoa_oracle.so")
_lib = None


def build():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        lib = ctypes.CDLL(_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.qref_num_threads.restype = ctypes.c_int
        lib.qref_set_num_threads.argtypes = [ctypes.c_int]
        lib.qref_maxcut_diag.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, dp, dp]
        for name in ("qref_energy_c128", "qref_energy_c64"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_void_p]
        for name in ("qref_energy_c128_nb", "qref_energy_c64_nb"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_void_p]
        lib.qref_qubo_diag.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp]
        _lib = lib
    return _lib


def num_threads():
    return load().qref_num_threads()


def use_all_cores():
    """Launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU baseline uses every core it may."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().qref_set_num_threads(int(n))
    return num_threads()


def maxcut_diag(n, edges):
    lib = load()
    u = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
    v = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
    w = np.ascontiguousarray([e[2] for e in edges], dtype=np.float64)
    out = np.empty(1 << n, dtype=np.float64)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    lib.qref_maxcut_diag(n, len(edges), u.ctypes.data_as(ip), v.ctypes.data_as(ip), w.ctypes.data_as(dp),
                         out.ctypes.data_as(dp))
    return out


def qubo_diag(Q, c=None, b=0.0):
    lib = load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    n = Q.shape[0]
    cc = np.ascontiguousarray(np.zeros(n) if c is None else c, dtype=np.float64)
    out = np.empty(1 << n, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.qref_qubo_diag(n, Q.ctypes.data_as(dp), cc.ctypes.data_as(dp), float(b), out.ctypes.data_as(dp))
    return out


def energy(n, diag, angles, dtype="complex128", n_beta=1):
    """n_beta = 1: X mixer; n_beta = n: XMultiAngle (flat reference layout, qaoa/qaoa.py:937-990)."""
    lib = load()
    diag = np.ascontiguousarray(diag, dtype=np.float64)
    angles = np.ascontiguousarray(angles, dtype=np.float64)
    out = np.empty(4, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    fn = lib.qref_energy_c128_nb if dtype == "complex128" else lib.qref_energy_c64_nb
    rc = fn(n, diag.ctypes.data_as(dp), angles.size // (1 + n_beta), n_beta, angles.ctypes.data_as(dp),
            out.ctypes.data_as(dp), None)
    if rc:
        raise MemoryError("C oracle could not allocate the statevector")
    return out
