"""X mixer with one angle per qubit; parameter index == qubit index
(reference qaoa/mixers/x_multiangle_mixer.py:24-50, binding order qaoa/qaoa.py:494-498)."""

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng


class XMultiAngle(Mixer):
    def get_num_parameters(self):
        return self.N_qubits

    def lower_mixer(self, engine, trotter=1):
        engine.set_mixer(_eng.MIXER_XMULTI, trotter=trotter)
