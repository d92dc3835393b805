"""Grover mixer U_S^dag X^n C^{n-1}P(-beta) X^n U_S = 1 + (e^{-i beta} - 1)|s><s| with
|s> = U_S|0> given by an InitialState (reference qaoa/mixers/grover_mixer.py:24-82).
Only |s> matters, so the sub-circuit is lowered as a state, not as gates."""

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng


class Grover(Mixer):
    def __init__(self, subcircuit, label=None) -> None:
        super().__init__(label=label)
        self.subcircuit = subcircuit
        if hasattr(subcircuit, "N_qubits"):
            self.N_qubits = subcircuit.N_qubits

    def create_circuit(self):
        if not hasattr(self.subcircuit, "N_qubits") or self.subcircuit.N_qubits != self.N_qubits:
            self.subcircuit.setNumQubits(self.N_qubits)
        self.circuit = None

    def lower_mixer(self, engine, trotter=1):
        self.create_circuit()
        kind, k, vec = self.subcircuit.state_descriptor()
        engine.set_mixer(_eng.MIXER_GROVER, trotter=trotter, sub_kind=kind, sub_k=k, sub_vec=vec)
