"""Transverse-field mixer: RX(-2 beta) = exp(+i beta X) on every qubit
(reference qaoa/mixers/x_mixer.py:28-35)."""

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng


class X(Mixer):
    def __init__(self) -> None:
        self.circuit = None

    def lower_mixer(self, engine, trotter=1):
        engine.set_mixer(_eng.MIXER_X, trotter=trotter)
