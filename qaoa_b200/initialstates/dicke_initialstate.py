"""Dicke state |D^n_k>: the uniform, real, positive superposition of all weight-k strings — what
the Baertschi-Eidenbenz circuit of the reference prepares
(reference qaoa/initialstates/dicke_initialstate.py:23-56)."""

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng


class Dicke(InitialState):
    def __init__(self, k: int, N: int | None = None, label: str | None = None) -> None:
        super().__init__(label=label)
        self.k = k
        self.N_qubits = N if N is not None else k

    def state_descriptor(self):
        return _eng.INIT_DICKE, int(self.k), None
