"""Mixer plugin base (public surface of reference qaoa/mixers/base_mixer.py:44-141).

A supported mixer lowers itself onto the CUDA engine with ``lower_mixer(engine, trotter)``
instead of producing a qiskit circuit.
"""

from abc import ABC


class MixerBase(ABC):
    def __init__(self, label=None) -> None:
        self.circuit = None
        self.label = label if label is not None else self.__class__.__name__

    def setNumQubits(self, n):
        self.N_qubits = n


class Mixer(MixerBase):
    def __init_subclass__(cls, **kwargs):
        # subclasses are not required to call super().__init__(): give them a label anyway
        super().__init_subclass__(**kwargs)
        original = cls.__dict__.get("__init__")
        if original is not None:
            def init_with_label(self, *args, __orig=original, **kw):
                __orig(self, *args, **kw)
                if not hasattr(self, "label"):
                    self.label = type(self).__name__
                if not hasattr(self, "circuit"):
                    self.circuit = None
            init_with_label.__doc__ = original.__doc__
            cls.__init__ = init_with_label

    def create_circuit(self):
        """No gates are ever executed; kept for call-compatibility (qaoa/qaoa.py:442)."""
        self.circuit = None

    def get_num_parameters(self):
        return 1

    def engine_betas(self, betas):
        """this layer's mixer parameters -> the angles the device mixer takes (identity unless parameters are shared)"""
        return list(betas)

    def lower_mixer(self, engine, trotter=1):
        raise NotImplementedError(
            "%s cannot be lowered onto the B200 engine (only X, XMultiAngle, XY and Grover are supported; "
            "gate-level mixers are rejected, never routed to a CPU simulator)" % type(self).__name__
        )

    def __str__(self):
        return getattr(self, "label", type(self).__name__)
