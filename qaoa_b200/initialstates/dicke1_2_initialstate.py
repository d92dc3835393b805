"""Dicke1_2: the equal superposition of the weight-1 and weight-2 strings of three qubits, the "max balanced" colour
encoding for k = 6 (reference qaoa/initialstates/dicke1_2_initialstate.py:10-41).  Amplitudes (phases included) are those
of the reference's fixed circuit, evaluated once on the host."""

import numpy as np

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng
from qaoa_b200.utils.gatesim import SmallState


class Dicke1_2(InitialState):
    def __init__(self) -> None:
        self.k = 6
        self.N_qubits = 3

    def state_vector(self):
        s = SmallState(3)      # dicke1_2_initialstate.py:28-41
        s.x(0)
        s.evolve_yyy(np.pi / 4)
        s.xx_plus_yy(np.arcsin(2 * np.sqrt(2) / 3), np.pi / 2, 0, 1)
        s.xx_plus_yy(-np.pi / 2, np.pi / 2, 0, 2)
        s.x(1)
        s.cz(1, 2)
        s.x(1)
        if self.N_qubits == 3:
            return s.v
        full = np.zeros(1 << self.N_qubits, dtype=np.complex128)   # qubits beyond the first three stay |0>
        full[:8] = s.v
        return full

    def state_descriptor(self):
        return _eng.INIT_STATEVECTOR, 0, self.state_vector()
