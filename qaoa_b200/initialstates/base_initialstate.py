"""Initial-state plugin base (public surface of reference
qaoa/initialstates/base_initialstate.py:44-135).  Supported states describe themselves with
``state_descriptor()`` -> (kind, k, vector) and are prepared directly in HBM."""

from abc import ABC


class BaseInitialState(ABC):
    def __init__(self, label=None) -> None:
        self.circuit = None
        self.label = label if label is not None else self.__class__.__name__

    def setNumQubits(self, n):
        self.N_qubits = n


class InitialState(BaseInitialState):
    def __init_subclass__(cls, **kwargs):
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
        self.circuit = None

    def get_num_parameters(self):
        return 0

    def state_descriptor(self):
        raise NotImplementedError(
            "%s cannot be lowered onto the B200 engine (supported: Plus, Dicke, StateVector)" % type(self).__name__
        )

    def lower_initial(self, engine):
        kind, k, vec = self.state_descriptor()
        engine.set_initial(kind, k=k, vec=vec)

    def __str__(self):
        return getattr(self, "label", type(self).__name__)
