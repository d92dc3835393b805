"""LessThanK: k feasible computational-basis states of ceil(log2 k) qubits in equal superposition
(reference qaoa/initialstates/lessthank_initialstate.py:8-152; construction of arXiv:2411.08594).
The reference gives one fixed circuit per k in {3, 5, 6, 7} (Hadamards for powers of two); the amplitudes those
circuits prepare are evaluated once on the host (qaoa_b200/utils/gatesim.py) and uploaded as a statevector."""

import numpy as np

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng
from qaoa_b200.utils.gatesim import SmallState


class LessThanK(InitialState):
    def __init__(self, k: int) -> None:
        if not LessThanK.is_power_of_two_or_between_2_and_8(k):
            raise ValueError("k must be a power of two or between 2 and 8")
        self.k = k
        self.N_qubits = int(np.ceil(np.log2(self.k)))

    def is_power_of_two_or_between_2_and_8(k):
        return bool(2 <= k <= 8 or (k > 0 and (k & (k - 1)) == 0))

    def state_vector(self):
        """amplitudes of the reference's circuit for this k (lessthank_initialstate.py:69-152).  The circuit touches
        the first ceil(log2 k) qubits only: on a larger register (QAOA sets N_qubits to the problem's) every other
        qubit stays |0>, exactly as the reference's circuit leaves it."""
        kb = int(np.ceil(np.log2(self.k)))
        s = SmallState(kb)
        third = np.arccos(1 / np.sqrt(3)) * 2
        if self.k == 3:        # :79-93
            s.ry(third, 1)
            s.xx_plus_yy(np.pi / 2, -np.pi / 2, 0, 1)
        elif self.k == 5:      # :95-107
            s.ry(np.arcsin(1 / np.sqrt(5)) * 2, 0)
            s.h(1, ctrl=0, ctrl_state=0)
            s.h(2, ctrl=0, ctrl_state=0)
        elif self.k == 6:      # :109-127
            s.ry(np.pi / 2, 2)
            s.ry(third, 1)
            s.xx_plus_yy(np.pi / 2, -np.pi / 2, 0, 1)
        elif self.k == 7:      # :129-152
            s.ry(np.arcsin(1 / np.sqrt(7)) * 2, 0)
            s.cx(0, 1)
            s.ry(np.pi / 2, 2, ctrl=0, ctrl_state=0)
            s.ry(third, 1, ctrl=0, ctrl_state=0)
            s.xx_plus_yy(np.pi / 2, -np.pi / 2, 0, 1)
        else:                  # powers of two: Hadamards on the whole register (:69-77)
            return np.full(1 << self.N_qubits, 2.0 ** (-0.5 * self.N_qubits), dtype=np.complex128)
        if self.N_qubits < kb:
            raise ValueError("LessThanK(%d) needs %d qubits" % (self.k, kb))
        full = np.zeros(1 << self.N_qubits, dtype=np.complex128)
        full[: 1 << kb] = s.v
        return full

    def state_descriptor(self):
        return _eng.INIT_STATEVECTOR, 0, self.state_vector()
