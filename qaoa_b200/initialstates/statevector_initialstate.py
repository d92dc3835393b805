"""User-supplied initial statevector, little-endian index
(reference qaoa/initialstates/statevector_initialstate.py:19-33)."""

import numpy as np

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng


class StateVector(InitialState):
    def __init__(self, statevector) -> None:
        super().__init__()
        self.statevector = statevector

    def state_descriptor(self):
        vec = np.asarray(self.statevector, dtype=np.complex128).reshape(-1)
        if vec.size != 1 << self.N_qubits:
            raise ValueError("statevector has %d amplitudes, expected 2^%d" % (vec.size, self.N_qubits))
        norm = np.linalg.norm(vec)
        if not np.isclose(norm, 1.0, atol=1e-8):
            # qiskit's initialize() refuses unnormalised vectors as well
            raise ValueError("Sum of amplitudes-squared is not 1, but %r." % float(norm * norm))
        return _eng.INIT_STATEVECTOR, 0, vec
