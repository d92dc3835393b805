"""ctypes binding of libqaoa_b200.so (C ABI declared in include/qaoa_b200.h).

This is the only place where Python touches the CUDA engine.  There is NO CPU
fallback: if the shared library is missing or no B200 is visible, constructing
an :class:`Engine` raises.  The engine replaces the reference's
``transpile -> backend.run -> get_counts -> problem.cost`` evaluation path
(reference qaoa/qaoa.py:516-519, 592-614, 653-661, 691-740, 912-922).
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libqaoa_b200.so")

DTYPE_C64 = 0
DTYPE_C128 = 1
MIXER_X, MIXER_XMULTI, MIXER_XY, MIXER_GROVER, MIXER_NODE_GROVER = 0, 1, 2, 3, 4
INIT_PLUS, INIT_DICKE, INIT_STATEVECTOR, INIT_PLUS_PHASES = 0, 1, 2, 3

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_uint8_p = ctypes.POINTER(ctypes.c_uint8)
_c_uint64_p = ctypes.POINTER(ctypes.c_uint64)


class EngineError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def load_library():
    """Loads libqaoa_b200.so and declares every prototype of include/qaoa_b200.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise EngineError(
            "libqaoa_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C qaoa_b200/csrc` (there is no CPU fallback)" % _LIB_PATH
        )
    lib = ctypes.CDLL(_LIB_PATH)
    H = ctypes.c_void_p
    sig = {
        "qaoa_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(H)]),
        "qaoa_destroy": (ctypes.c_int, [H]),
        "qaoa_last_error": (ctypes.c_char_p, [H]),
        "qaoa_global_error": (ctypes.c_char_p, []),
        "qaoa_set_option": (ctypes.c_int, [H, ctypes.c_char_p, ctypes.c_double]),
        "qaoa_get_counter": (ctypes.c_double, [H, ctypes.c_char_p]),
        "qaoa_set_cost_maxkcut": (
            ctypes.c_int,
            [H, ctypes.c_int, ctypes.c_int, _c_int32_p, _c_int32_p, _c_double_p, ctypes.c_int, _c_uint8_p,
             ctypes.c_int, ctypes.c_int],
        ),
        "qaoa_set_cost_maxkcut_terms": (
            ctypes.c_int,
            [H, ctypes.c_int, ctypes.c_int, _c_int32_p, _c_int32_p, _c_double_p, _c_int32_p, ctypes.c_int, ctypes.c_int,
             _c_uint8_p, ctypes.c_int, ctypes.c_int],
        ),
        "qaoa_set_cost_qubo": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_double]),
        "qaoa_set_cost_diagonal": (ctypes.c_int, [H, _c_double_p]),
        "qaoa_read_cost_diagonal": (ctypes.c_int, [H, ctypes.c_size_t, ctypes.c_size_t, _c_double_p]),
        "qaoa_cost_minmax": (ctypes.c_int, [H, _c_double_p, _c_double_p]),
        "qaoa_set_mixer": (
            ctypes.c_int,
            [H, ctypes.c_int, _c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_double_p],
        ),
        "qaoa_set_initial": (ctypes.c_int, [H, ctypes.c_int, ctypes.c_int, _c_double_p]),
        "qaoa_energy": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, _c_double_p]),
        "qaoa_energy_batch": (
            ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, _c_double_p]),
        "qaoa_energy_batch_resident": (
            ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int]),
        "qaoa_fetch_results": (ctypes.c_int, [H, ctypes.c_int, _c_double_p]),
        "qaoa_read_statevector": (ctypes.c_int, [H, _c_double_p]),
        "qaoa_continue_from_kept_state": (ctypes.c_int, [H, ctypes.c_uint64]),
        "qaoa_debug_tile_program": (
            ctypes.c_int, [H, ctypes.c_int, _c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, _c_double_p]),
        "qaoa_debug_plan": (
            ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                           ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]),
        "qaoa_sample": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_uint64_p]),
        "qaoa_extreme_solutions": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  _c_uint64_p, ctypes.POINTER(ctypes.c_int), _c_double_p]),
        "qaoa_shard_chunk_probs": (ctypes.c_int, [H, _c_double_p, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)]),
        "qaoa_shard_locate": (ctypes.c_int, [H, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), _c_double_p, _c_uint64_p]),
        "qaoa_shard_cvar_digit": (ctypes.c_int, [H, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, _c_double_p,
                                                 ctypes.POINTER(ctypes.c_int)]),
        "qaoa_get_stream": (ctypes.c_int, [H, ctypes.POINTER(ctypes.c_ulonglong)]),
        "qaoa_get_stream2": (ctypes.c_int, [H, ctypes.POINTER(ctypes.c_ulonglong)]),
        "qaoa_get_timeline": (ctypes.c_int, [H, _c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
        "qaoa_shard_chunk_info": (ctypes.c_int, [H, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
        "qaoa_shard_run_pass_chunk": (ctypes.c_int, [H, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
        "qaoa_shard_finish_device": (ctypes.c_int, [H, ctypes.c_void_p]),
        "qaoa_shard_grover_step": (ctypes.c_int, [H, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int, _c_double_p, _c_double_p]),
        "qaoa_cvar": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double, _c_double_p]),
        "qaoa_cvar_stats": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double, _c_double_p]),
        "qaoa_cost_distribution": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, _c_double_p,
                                                  _c_double_p, _c_double_p]),
        # sharded registers (one engine per GPU)
        "qaoa_create_sharded": (
            ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(H)]),
        "qaoa_shard_info": (ctypes.c_int, [H, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_int)]),
        "qaoa_shard_local_ptr": (ctypes.c_int, [H, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
        "qaoa_shard_set_peers": (ctypes.c_int, [H, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
        "qaoa_shard_export": (ctypes.c_int, [H, ctypes.c_int, ctypes.c_char_p]),
        "qaoa_shard_import": (ctypes.c_int, [H, ctypes.c_int, ctypes.c_char_p]),
        "qaoa_shard_close_peers": (ctypes.c_int, [H]),
        "qaoa_shard_begin": (ctypes.c_int, [H, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
        "qaoa_shard_pass_is_cross": (ctypes.c_int, [H, ctypes.c_int]),
        "qaoa_shard_run_pass": (ctypes.c_int, [H, ctypes.c_int]),
        "qaoa_shard_sync": (ctypes.c_int, [H]),
        "qaoa_shard_finish": (ctypes.c_int, [H, _c_double_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(_c_double_p)


def debug_plan(n_qubits, dtype="complex64", rank=0, world=1, symmetric=True, mixer=MIXER_X, p=1, schedule=0, low_bits=0,
               cross_spare=True):
    """The tile plan the engine would build for this configuration (dict); needs no GPU -- the planner is host code."""
    import json

    lib = load_library()
    dt = {"complex64": DTYPE_C64, "complex128": DTYPE_C128}.get(dtype, dtype)
    options = (schedule & 3) | ((low_bits & 7) << 2) | (0 if cross_spare else 32)
    buf = ctypes.create_string_buffer(1 << 20)
    rc = lib.qaoa_debug_plan(int(n_qubits), int(dt), int(rank), int(world), int(bool(symmetric)), int(mixer), int(p),
                             int(options), buf, len(buf))
    if rc != 0:
        raise EngineError(lib.qaoa_global_error().decode())
    return json.loads(buf.value.decode())


class Engine:
    """One device-resident QAOA register of ``n_qubits`` qubits on ``cuda:device``."""

    def __init__(self, n_qubits, dtype="complex64", device=0, rank=0, world=1, shard_symmetric=False):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        self.n_qubits = int(n_qubits)
        self.dtype = {"complex64": DTYPE_C64, "complex128": DTYPE_C128, DTYPE_C64: DTYPE_C64,
                      DTYPE_C128: DTYPE_C128}[dtype]
        self.n_beta = 1
        self.n_gamma = 1
        self.n_init = 0
        self.rank, self.world = int(rank), int(world)
        rc = self._lib.qaoa_create_sharded(self.n_qubits, self.dtype, int(device), self.rank, self.world,
                                           ctypes.byref(self._h))
        if rc != 0:
            raise EngineError(self._lib.qaoa_global_error().decode())
        self.local_qubits = self.n_qubits - (self.world.bit_length() - 1)
        self.shard_symmetric = bool(shard_symmetric)
        if self.shard_symmetric:
            # sharded symmetric register (MaxCut + X / multi-angle X + |+>, n even, complex64): each rank holds
            # 2^(n-1-log2 world) stored amplitudes in twisted coordinates (csrc/tile_kernel.cuh)
            self.set_option("shard_symmetric", 1)
            self.local_qubits -= 1

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise EngineError(self._lib.qaoa_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.qaoa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self._lib.qaoa_set_option(self._h, key.encode(), float(value)))

    def counter(self, key):
        return self._lib.qaoa_get_counter(self._h, key.encode())

    # -- cost -------------------------------------------------------------
    def set_cost_maxkcut(self, n_nodes, edges, bits_per_node, lut, fixed_node=False, fixed_colour=0):
        self.n_gamma = 1
        u = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
        v = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
        w = np.ascontiguousarray([e[2] for e in edges], dtype=np.float64)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        assert lut.size == 1 << bits_per_node
        self._check(self._lib.qaoa_set_cost_maxkcut(
            self._h, int(n_nodes), len(edges), u.ctypes.data_as(_c_int32_p), v.ctypes.data_as(_c_int32_p),
            _dptr(w), int(bits_per_node), lut.ctypes.data_as(_c_uint8_p), int(bool(fixed_node)), int(fixed_colour)))

    def set_cost_maxkcut_terms(self, n_nodes, edges, term_of_edge, n_terms, bits_per_node, lut, fixed_node=False,
                               fixed_colour=0):
        """Same cost as set_cost_maxkcut, one gamma per term (MaxCutFree / MaxCutOrbit); n <= 14 / 13 only."""
        u = np.ascontiguousarray([e[0] for e in edges], dtype=np.int32)
        v = np.ascontiguousarray([e[1] for e in edges], dtype=np.int32)
        w = np.ascontiguousarray([e[2] for e in edges], dtype=np.float64)
        t = np.ascontiguousarray(term_of_edge, dtype=np.int32)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        assert t.size == len(edges) and lut.size == 1 << bits_per_node
        self._check(self._lib.qaoa_set_cost_maxkcut_terms(
            self._h, int(n_nodes), len(edges), u.ctypes.data_as(_c_int32_p), v.ctypes.data_as(_c_int32_p), _dptr(w),
            t.ctypes.data_as(_c_int32_p), int(n_terms), int(bits_per_node), lut.ctypes.data_as(_c_uint8_p),
            int(bool(fixed_node)), int(fixed_colour)))
        self.n_gamma = int(n_terms)

    def set_cost_qubo(self, Q, c=None, b=0.0):
        self.n_gamma = 1
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        n = Q.shape[0]
        c = np.zeros(n) if c is None else np.ascontiguousarray(c, dtype=np.float64)
        self._check(self._lib.qaoa_set_cost_qubo(self._h, n, _dptr(Q), _dptr(c), float(b)))

    def set_cost_diagonal(self, diag):
        self.n_gamma = 1
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        assert diag.size == 1 << self.n_qubits
        self._check(self._lib.qaoa_set_cost_diagonal(self._h, _dptr(diag)))

    def read_cost_diagonal(self, offset=0, count=None):
        """cost(x) for x in [offset, offset+count); a sharded engine indexes its own slice from 0."""
        count = (1 << self.local_qubits) - offset if count is None else count
        out = np.empty(count, dtype=np.float64)
        self._check(self._lib.qaoa_read_cost_diagonal(self._h, offset, count, _dptr(out)))
        return out

    def cost_minmax(self):
        lo, hi = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.qaoa_cost_minmax(self._h, ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value

    # -- mixer / initial state -------------------------------------------
    def set_mixer(self, kind, pairs=None, trotter=1, sub_kind=INIT_PLUS, sub_k=0, sub_vec=None):
        if pairs is not None and len(pairs):
            pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1)
            pp, npairs = pr.ctypes.data_as(_c_int32_p), pr.size // 2
        else:
            pr, pp, npairs = None, None, 0
        sv = None
        if sub_vec is not None:
            sv = np.ascontiguousarray(np.asarray(sub_vec, dtype=np.complex128)).view(np.float64)
        self._check(self._lib.qaoa_set_mixer(self._h, int(kind), pp, npairs, int(trotter), int(sub_kind),
                                             int(sub_k), _dptr(sv) if sv is not None else None))
        self.n_beta = self.n_qubits if kind == MIXER_XMULTI else 1

    def set_initial(self, kind, k=0, vec=None):
        sv = None
        if vec is not None:
            sv = np.ascontiguousarray(np.asarray(vec, dtype=np.complex128)).view(np.float64)
            assert sv.size == 2 << self.n_qubits
        self._check(self._lib.qaoa_set_initial(self._h, int(kind), int(k), _dptr(sv) if sv is not None else None))
        self.n_init = int(k) if kind == INIT_PLUS_PHASES else 0

    # -- evaluation ------------------------------------------------------
    def energy(self, angles):
        """Returns (E[cost], E[cost^2], min, max, norm) for one flat angle vector."""
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        per = self.n_gamma + self.n_beta
        body = angles.size - self.n_init
        if body <= 0 or body % per:
            raise ValueError("angle vector length %d is not %d + a multiple of %d" % (angles.size, self.n_init, per))
        out = np.empty(5, dtype=np.float64)
        self._check(self._lib.qaoa_energy(self._h, body // per, _dptr(angles), angles.size, _dptr(out)))
        return out

    def energy_batch(self, angles):
        """angles: (B, n_angles) -> (B, 5)."""
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        assert angles.ndim == 2
        B, na = angles.shape
        per = self.n_gamma + self.n_beta
        body = na - self.n_init
        if body <= 0 or body % per:
            raise ValueError("angle vector length %d is not %d + a multiple of %d" % (na, self.n_init, per))
        out = np.empty((B, 5), dtype=np.float64)
        self._check(self._lib.qaoa_energy_batch(self._h, body // per, _dptr(angles), na, B, _dptr(out)))
        return out

    def energy_batch_resident(self, angles):
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        B, na = angles.shape
        per = self.n_gamma + self.n_beta
        self._check(self._lib.qaoa_energy_batch_resident(self._h, (na - self.n_init) // per, _dptr(angles), na, B))

    def fetch_results(self, B):
        out = np.empty((B, 5), dtype=np.float64)
        self._check(self._lib.qaoa_fetch_results(self._h, B, _dptr(out)))
        return out

    def read_statevector(self):
        out = np.empty(2 << self.n_qubits, dtype=np.float64)
        self._check(self._lib.qaoa_read_statevector(self._h, _dptr(out)))
        return out.view(np.complex128)

    def final_probabilities(self, angles):
        """|psi(angles)|^2 for every basis state (host array of 2^n doubles): one evaluation with the final state kept in
        HBM, read back once.  For statistics of the final state that have no device reduction yet (the exact CVaR of
        real-valued costs); the evolution itself runs on the device."""
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        self.set_option("keep_state", 1)
        try:
            self.energy(angles)
            psi = self.read_statevector()
        finally:
            self.set_option("keep_state", 0)
        pr = psi.real ** 2 + psi.imag ** 2
        return pr / pr.sum()

    def continue_from_kept_state(self, flip_mask=0):
        """X gates between layers (the reference's flip=True, qaoa/qaoa.py:505-509): the state kept by the last evaluation
        (option keep_state=1), with X applied on the qubits of ``flip_mask`` (bit q = qubit q), becomes the initial
        state, device to device; the next evaluation continues the circuit with its remaining layers."""
        self._check(self._lib.qaoa_continue_from_kept_state(self._h, int(flip_mask)))
        self.n_init = 0

    # -- sampling / cost distribution of the final state -------------------------------------------------
    def _angles(self, angles):
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        per = self.n_gamma + self.n_beta
        body = angles.size - self.n_init
        if body <= 0 or body % per:
            raise ValueError("angle vector length %d is not %d + a multiple of %d" % (angles.size, self.n_init, per))
        return angles, body // per

    def sample(self, angles, shots, seed=None):
        """``shots`` basis indices drawn from |psi(angles)|^2 (replaces backend.run(shots) + get_counts,
        reference qaoa/qaoa.py:1157-1172).  Bit i of an index is qubit i."""
        angles, p = self._angles(angles)
        u = np.ascontiguousarray(np.random.default_rng(seed).random(int(shots)), dtype=np.float64)
        out = np.empty(int(shots), dtype=np.uint64)
        self._check(self._lib.qaoa_sample(self._h, p, _dptr(angles), angles.size, int(shots), _dptr(u),
                                          out.ctypes.data_as(_c_uint64_p)))
        return out

    def extreme_solutions(self, angles, want_max=True, cap=4096):
        """(indices, cost): the basis states of the support of |psi(angles)> with the largest (smallest) cost."""
        angles, p = self._angles(angles)
        out = np.empty(int(cap), dtype=np.uint64)
        cnt, cost = ctypes.c_int(), ctypes.c_double()
        self._check(self._lib.qaoa_extreme_solutions(self._h, p, _dptr(angles), angles.size, int(bool(want_max)), int(cap),
                                                     out.ctypes.data_as(_c_uint64_p), ctypes.byref(cnt), ctypes.byref(cost)))
        return out[: min(cnt.value, int(cap))].copy(), cost.value

    def cost_distribution(self, angles):
        """(cost values, probabilities) of the final state; integer-valued costs with range <= 255 only."""
        angles, p = self._angles(angles)
        probs = np.empty(256, dtype=np.float64)
        c0, delta = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.qaoa_cost_distribution(self._h, p, _dptr(angles), angles.size, _dptr(probs),
                                                     ctypes.byref(c0), ctypes.byref(delta)))
        keep = probs > 0
        return (c0.value + delta.value * np.arange(256))[keep], probs[keep]

    def cvar(self, angles, alpha):
        """exact CVaR_alpha (mean cost over the best alpha of the probability mass) of the final state, any cost type"""
        angles, p = self._angles(angles)
        out = np.empty(2, dtype=np.float64)
        self._check(self._lib.qaoa_cvar(self._h, p, _dptr(angles), angles.size, float(alpha), _dptr(out)))
        return float(out[0])

    def cvar_stats(self, angles, alpha):
        """(E[cost], E[cost^2], min, max, norm, CVaR_alpha) of ONE evaluation (the kept-state run behind the CVaR)"""
        angles, p = self._angles(angles)
        out = np.empty(7, dtype=np.float64)
        self._check(self._lib.qaoa_cvar_stats(self._h, p, _dptr(angles), angles.size, float(alpha), _dptr(out)))
        return np.array([out[2], out[3], out[4], out[5], out[6], out[0]])

    # -- sharded registers: one engine per GPU (include/qaoa_b200.h, "Sharded registers") -------------
    def shard_local_ptr(self, which):
        out = ctypes.c_void_p()
        self._check(self._lib.qaoa_shard_local_ptr(self._h, int(which), ctypes.byref(out)))
        return out.value

    def shard_set_peers(self, which, ptrs):
        arr = (ctypes.c_void_p * self.world)(*[ctypes.c_void_p(p) for p in ptrs])
        self._check(self._lib.qaoa_shard_set_peers(self._h, int(which), arr))

    def shard_export(self, which):
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.qaoa_shard_export(self._h, int(which), buf))
        return buf.raw

    def shard_import(self, which, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        self._check(self._lib.qaoa_shard_import(self._h, int(which), blob))

    def shard_close_peers(self):
        self._check(self._lib.qaoa_shard_close_peers(self._h))

    def shard_begin(self, angles):
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        per = self.n_gamma + self.n_beta
        if angles.size == 0 or angles.size % per:
            raise ValueError("angle vector length %d is not a multiple of %d" % (angles.size, per))
        n_passes = ctypes.c_int()
        self._check(self._lib.qaoa_shard_begin(self._h, angles.size // per, _dptr(angles), angles.size,
                                               ctypes.byref(n_passes)))
        return n_passes.value

    def shard_pass_is_cross(self, i):
        rc = self._lib.qaoa_shard_pass_is_cross(self._h, int(i))
        if rc < 0:
            raise EngineError("no evaluation in flight or pass index out of range")
        return bool(rc)

    def shard_run_pass(self, i):
        self._check(self._lib.qaoa_shard_run_pass(self._h, int(i)))

    def shard_chunk_probs(self):
        """probability mass of every 2^16-amplitude chunk of the kept local shard"""
        n = ctypes.c_ulonglong()
        self._check(self._lib.qaoa_shard_chunk_probs(self._h, None, 0, ctypes.byref(n)))
        out = np.empty(int(n.value), dtype=np.float64)
        self._check(self._lib.qaoa_shard_chunk_probs(self._h, _dptr(out), out.size, ctypes.byref(n)))
        return out

    def shard_locate(self, chunk_of, resid):
        chunk_of = np.ascontiguousarray(chunk_of, dtype=np.uint32)
        resid = np.ascontiguousarray(resid, dtype=np.float64)
        out = np.empty(chunk_of.size, dtype=np.uint64)
        if chunk_of.size:
            self._check(self._lib.qaoa_shard_locate(self._h, int(chunk_of.size), chunk_of.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                                    _dptr(resid), out.ctypes.data_as(_c_uint64_p)))
        return out

    def shard_cvar_key_bits(self):
        kb = ctypes.c_int()
        self._check(self._lib.qaoa_shard_cvar_digit(self._h, 0, 0, 0, None, ctypes.byref(kb)))
        return int(kb.value)

    def shard_cvar_digit(self, prefix, shift, check):
        out = np.empty(512, dtype=np.float64)
        self._check(self._lib.qaoa_shard_cvar_digit(self._h, int(prefix), int(shift), int(bool(check)), _dptr(out), None))
        return out

    def stream_handle(self):
        """the engine's cudaStream_t (integer) -- for torch.cuda.ExternalStream"""
        out = ctypes.c_ulonglong()
        self._check(self._lib.qaoa_get_stream(self._h, ctypes.byref(out)))
        return int(out.value)

    def timeline(self):
        """launches of the last evaluation timed with option time_passes: rows {program id, chunk, stream, start ms, end ms}"""
        n = ctypes.c_int()
        self._check(self._lib.qaoa_get_timeline(self._h, None, 0, ctypes.byref(n)))
        out = np.zeros((max(1, n.value), 5), dtype=np.float64)
        self._check(self._lib.qaoa_get_timeline(self._h, _dptr(out), int(n.value), ctypes.byref(n)))
        return out[: n.value]

    def stream2_handle(self):
        out = ctypes.c_ulonglong()
        self._check(self._lib.qaoa_get_stream2(self._h, ctypes.byref(out)))
        return int(out.value)

    def shard_chunk_info(self, n_passes):
        """(number of chunks, per-pass flag: may be launched chunk by chunk) of the evaluation in flight"""
        nch = ctypes.c_int()
        closed = (ctypes.c_int * max(1, n_passes))()
        self._check(self._lib.qaoa_shard_chunk_info(self._h, ctypes.byref(nch), closed, n_passes))
        return int(nch.value), [bool(closed[i]) for i in range(n_passes)]

    def shard_run_pass_chunk(self, i, chunk=-1, stream=0, ctas=0):
        self._check(self._lib.qaoa_shard_run_pass_chunk(self._h, int(i), int(chunk), int(stream), int(ctas)))

    def shard_finish_device(self, device_ptr):
        """raw partial sums {sum p, sum p c, sum p c^2, min, max} written to DEVICE memory at device_ptr (5 doubles)"""
        self._check(self._lib.qaoa_shard_finish_device(self._h, ctypes.c_void_p(int(device_ptr))))

    def shard_grover_step(self, step, angles, dot_global=None):
        """one step of the sharded Grover circuit: (local partial dot) for step < p, raw local statistics for step = p"""
        angles, p = self._angles(angles)
        out = np.zeros(5, dtype=np.float64)
        dg = None if dot_global is None else np.ascontiguousarray(dot_global, dtype=np.float64)
        self._check(self._lib.qaoa_shard_grover_step(self._h, int(step), p, _dptr(angles), angles.size,
                                                     _dptr(dg) if dg is not None else None, _dptr(out)))
        return out[:2].copy() if step < p else out

    def shard_sync(self):
        self._check(self._lib.qaoa_shard_sync(self._h))

    def shard_finish(self):
        out = np.empty(5, dtype=np.float64)
        self._check(self._lib.qaoa_shard_finish(self._h, _dptr(out)))
        return out

    def debug_tile_program(self, hb, ops, reps=5, tau=0.1):
        """Profiling hook: mean ms per launch of an arbitrary (type, frame, mask) step program."""
        arr = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1)
        ms = ctypes.c_double()
        self._check(self._lib.qaoa_debug_tile_program(self._h, int(hb), arr.ctypes.data_as(_c_int32_p), arr.size // 3,
                                                      int(reps), float(tau), ctypes.byref(ms)))
        return ms.value
