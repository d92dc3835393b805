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
        lib.qref_qubo_diag_fast.restype = ctypes.c_int
        lib.qref_qubo_diag_fast.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp]
        lib.qref_maxcut_diag_u8.restype = ctypes.c_int
        lib.qref_maxcut_diag_u8.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ip, ctypes.POINTER(ctypes.c_uint8)]
        lib.qref_run.restype = ctypes.c_int
        lib.qref_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
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


def qubo_diag(Q, c=None, b=0.0, fast=False):
    """fast=True: O(1) per basis state (split tables), a few ulp from the term-by-term sum; for n >= 26."""
    lib = load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    n = Q.shape[0]
    cc = np.ascontiguousarray(np.zeros(n) if c is None else c, dtype=np.float64)
    out = np.empty(1 << n, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    if fast:
        if lib.qref_qubo_diag_fast(n, Q.ctypes.data_as(dp), cc.ctypes.data_as(dp), float(b), out.ctypes.data_as(dp)):
            raise MemoryError("C oracle: qubo_diag_fast could not allocate its tables")
    else:
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


def maxcut_diag_u8(n, edges):
    """Integer-weighted MaxCut cost, one byte per basis state (cost = byte): 2^n bytes instead of 2^(n+3)."""
    lib = load()
    u = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
    v = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
    w = np.ascontiguousarray([int(round(e[2])) for e in edges], dtype=np.int32)
    assert all(float(e[2]) == int(round(e[2])) for e in edges), "u8 diagonal: integer weights only"
    out = np.empty(1 << n, dtype=np.uint8)
    ip = ctypes.POINTER(ctypes.c_int32)
    rc = lib.qref_maxcut_diag_u8(n, len(edges), u.ctypes.data_as(ip), v.ctypes.data_as(ip), w.ctypes.data_as(ip),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    if rc:
        raise ValueError("u8 diagonal: total weight above 255 or negative weights")
    return out


MIXERS = {"x": 0, "xmulti": 1, "xy": 2, "grover": 3}
INITS = {"plus": 0, "dicke": 1}


def run(n, diag, angles, dtype="complex128", mixer="x", init=("plus", 0), sub=("plus", 0), pairs=None, trotter=1,
        blocked=True, c0=0.0, delta=1.0):
    """General entry of the C oracle: (E[cost], E[cost^2], min, max, norm).

    diag: float64 array of 2^n costs, or uint8 array (cost = c0 + delta * byte).  mixer: "x" | "xmulti" | "xy"
    (ordered `pairs`) | "grover" (|s> = sub).  init / sub: ("plus", 0) or ("dicke", k).  angles: the reference's flat
    layout (qaoa/qaoa.py:937-990).  blocked: cache-blocked X-type layers (False: one sweep per gate)."""
    lib = load()
    diag = np.ascontiguousarray(diag)
    assert diag.size == 1 << n
    if diag.dtype == np.uint8:
        kind = 1
    else:
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        kind = 0
    angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
    n_beta = n if mixer == "xmulti" else 1
    assert angles.size % (1 + n_beta) == 0
    pr = None
    npairs = 0
    if pairs is not None and len(pairs):
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1)
        npairs = pr.size // 2
    out = np.empty(5, dtype=np.float64)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    rc = lib.qref_run(n, 0 if dtype == "complex64" else 1, kind, diag.ctypes.data_as(ctypes.c_void_p), float(c0),
                      float(delta), INITS[init[0]], int(init[1]), MIXERS[mixer], INITS[sub[0]], int(sub[1]),
                      pr.ctypes.data_as(ip) if pr is not None else None, npairs, int(trotter), int(bool(blocked)),
                      angles.size // (1 + n_beta), angles.ctypes.data_as(dp), out.ctypes.data_as(dp))
    if rc == 1:
        raise MemoryError("C oracle could not allocate the statevector")
    if rc:
        raise ValueError("C oracle: bad argument")
    return out


def host_mem_available_gib():
    """MemAvailable of the host in GiB (0 when /proc/meminfo is unreadable): full-size oracle runs check it first."""
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0
