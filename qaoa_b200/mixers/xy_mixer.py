"""XY mixer: XXPlusYY(beta/2) applied sequentially along a ring, a chain or explicit pairs
(reference qaoa/mixers/xy_mixer.py:42-79).  The gates do not commute: order is semantics.

Convention ([3p], qiskit is not importable here): XXPlusYYGate(theta) acts on span{|01>, |10>} as
[[cos(theta/2), -i sin(theta/2)], [-i sin(theta/2), cos(theta/2)]], so one mixer gate rotates by beta/4 with hopping term
-i sin(beta/4).  ``angle_sign=-1`` (an extension, not in the reference's signature) runs every gate with -beta, i.e. with
the opposite hopping sign: the reference's PortOpt notebook is reproduced by that convention only (DESIGN.md section 4,
"Open parity question"), which a run against qiskit has to settle."""

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng


class XY(Mixer):
    def __init__(self, topology=None, case="ring", angle_sign=1) -> None:
        self.topology = topology
        self.case = case
        self.circuit = None
        if angle_sign not in (1, -1):
            raise ValueError("angle_sign must be +1 or -1")
        self.angle_sign = angle_sign

    def engine_betas(self, betas):
        """this layer's beta -> the angle the device's XY gates take (sign convention, see the module docstring)"""
        sign = getattr(self, "angle_sign", 1)
        return [sign * b for b in betas]

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
