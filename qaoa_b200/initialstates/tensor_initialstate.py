"""Tensor: `num` copies of a sub-state side by side, copy i on qubits [i*m, (i+1)*m)
(reference qaoa/initialstates/tensor_initialstate.py:9-57).  Lowered as the Kronecker product of the copies'
amplitudes (copy 0 = the least significant index bits)."""

import numpy as np

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng


def amplitudes_of(state, n_qubits=None):
    """statevector of a supported initial state on its own qubits (Plus, Dicke, StateVector, LessThanK, Dicke1_2, Tensor)"""
    if hasattr(state, "state_vector"):
        return np.asarray(state.state_vector(), dtype=np.complex128)
    n = n_qubits if n_qubits is not None else state.N_qubits
    kind, k, vec = state.state_descriptor()
    if kind == _eng.INIT_PLUS:
        return np.full(1 << n, 2.0 ** (-0.5 * n), dtype=np.complex128)
    if kind == _eng.INIT_DICKE:
        v = np.array([1.0 if bin(x).count("1") == k else 0.0 for x in range(1 << n)], dtype=np.complex128)
        return v / np.linalg.norm(v)
    if kind == _eng.INIT_STATEVECTOR:
        return np.asarray(vec, dtype=np.complex128)
    raise NotImplementedError("%s has no amplitude description" % type(state).__name__)


class Tensor(InitialState):
    def __init__(self, subcircuit, num: int, label=None) -> None:
        super().__init__(label=label)
        self.num = num
        self.subcircuit = subcircuit
        self.N_qubits = self.num * self.subcircuit.N_qubits

    def state_vector(self):
        one = amplitudes_of(self.subcircuit)
        out = np.ones(1, dtype=np.complex128)
        for _ in range(self.num):
            out = np.kron(one, out)     # the new copy takes the more significant bits
        return out

    def state_descriptor(self):
        if self.N_qubits != self.num * self.subcircuit.N_qubits:
            raise ValueError("Tensor of %d x %d qubits cannot live on %d qubits" % (
                self.num, self.subcircuit.N_qubits, self.N_qubits))
        return _eng.INIT_STATEVECTOR, 0, self.state_vector()
