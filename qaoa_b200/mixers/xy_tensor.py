"""XYTensor: an XY mixer on every k-qubit node register of a one-hot Max-k-Cut encoding, all registers sharing beta
(reference qaoa/mixers/xy_tensor.py:10-60: Tensor(XY(topology) on k qubits, num_V)).  Lowered as the engine's XY mixer
with the node topology replicated per register (gates of different registers act on disjoint qubits and commute; inside
a register the reference's gate order is kept)."""

from .base_mixer import Mixer
from .xy_mixer import XY
from qaoa_b200 import engine as _eng


class XYTensor(Mixer):
    def __init__(self, k_cuts: int, topology=None) -> None:
        self.k_cuts = k_cuts
        self.topology = topology

    def create_circuit(self) -> None:
        num_V = self.N_qubits / self.k_cuts
        if not float(num_V).is_integer():
            raise ValueError("Total qubits=" + str(self.N_qubits) + " is not a multiple of " + str(self.k_cuts))
        self.num_V = int(num_V)
        self.circuit = None

    def node_pairs(self):
        return [list(p) for p in (self.topology or XY.generate_pairs(self.k_cuts, case="ring"))]

    def lower_mixer(self, engine, trotter=1):
        self.create_circuit()
        k = self.k_cuts
        pairs = [[v * k + a, v * k + b] for v in range(self.num_V) for (a, b) in self.node_pairs()]
        engine.set_mixer(_eng.MIXER_XY, pairs=pairs, trotter=trotter)
