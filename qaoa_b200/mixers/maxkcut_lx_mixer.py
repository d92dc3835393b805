"""MaxKCutLX: the logical-X mixer of the binary Max-k-Cut encodings (reference qaoa/mixers/maxkcut_lx_mixer.py:12-155).

Constructor, argument checks and error types are the reference's (`:29-61`), so code that builds the mixer keeps working
up to the point where the reference would execute it.  The mixer itself is REJECTED at lowering: the reference realises
it as `PauliEvolutionGate(op, time=beta)` per node register (`:150-155`) with `op` a sum of NON-commuting Pauli terms
(`:70-139`), i.e. as whatever product formula qiskit's default synthesis emits for that operator -- the unitary is defined
by a third-party synthesis pass (term order, grouping, version), not by the reference's own code, and cannot be restated
without qiskit.  `MaxKCutGrover` (both forms) and `XYTensor` cover the constrained k-cut mixers on the device.
"""
import numpy as np

from .base_mixer import Mixer


class MaxKCutLX(Mixer):
    not_lowerable = True

    def __init__(self, k_cuts: int, color_encoding: str, topology: str = "standard"):
        if (k_cuts < 2) or (k_cuts > 8):
            raise ValueError("k_cuts must be 2 or more, and is not implemented for k_cuts > 8")
        self.k_cuts = k_cuts
        self.k_bits = int(np.ceil(np.log2(k_cuts)))
        self.color_encoding = color_encoding
        self.topology = topology
        if self.is_power_of_two():
            raise ValueError("k_cuts is a power of two. Use e.g. X-mixer instead.")
        if not color_encoding:
            raise ValueError("please specify a color encoding")
        if k_cuts == 3 and (topology not in ["standard", "ring"]):
            raise ValueError('topology must be in ["standard", "ring"]')
        known = {3: ["LessThanK"], 5: ["LessThanK"], 6: ["LessThanK", "Dicke1_2", "max_balanced"], 7: ["LessThanK"]}
        if color_encoding not in known[k_cuts]:
            raise ValueError("invalid or missing color_encoding")

    def is_power_of_two(self) -> bool:
        return self.k_cuts > 0 and (self.k_cuts & (self.k_cuts - 1)) == 0

    def lower_mixer(self, engine, trotter=1):
        raise NotImplementedError(
            "MaxKCutLX is a qiskit PauliEvolutionGate of non-commuting terms: its unitary is fixed by qiskit's synthesis "
            "pass, not by the reference, and is rejected by the B200 engine (never routed to a CPU simulator); "
            "use MaxKCutGrover or XYTensor"
        )
