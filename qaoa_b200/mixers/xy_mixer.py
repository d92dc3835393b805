"""XY mixer: XXPlusYY(beta/2) applied sequentially along a ring, a chain or explicit pairs
(reference qaoa/mixers/xy_mixer.py:42-79).  The gates do not commute: order is semantics."""

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng


class XY(Mixer):
    def __init__(self, topology=None, case="ring") -> None:
        self.topology = topology
        self.case = case
        self.circuit = None

    def create_circuit(self):
        if not self.topology:
            self.topology = XY.generate_pairs(self.N_qubits, case=self.case)
        self.circuit = None

    @staticmethod
    def generate_pairs(n, case="ring"):
        assert case in ["ring", "chain"]
        if n < 2:
            return []
        pairs = [[q, q + 1] for q in range(n - 1)]
        if case == "ring":
            pairs.append([n - 1, 0])
        return pairs

    def lower_mixer(self, engine, trotter=1):
        self.create_circuit()
        engine.set_mixer(_eng.MIXER_XY, pairs=[list(p) for p in self.topology], trotter=trotter)
