"""|+>^n followed by one RZ(phi_q) per qubit (public surface of reference
qaoa/initialstates/plus_parameterized_initialstate.py:25-64): ``num_phases`` parameters (default N_qubits) that
lead the flat angle vector (``n_init``, qaoa/qaoa.py:937-990).  Registers beyond one CTA run layer by layer on the tile passes (X-type mixers)."""

from .base_initialstate import InitialState
from qaoa_b200 import engine as _eng


class PlusParameterized(InitialState):
    def __init__(self, num_phases=None) -> None:
        super().__init__()
        self.num_phases = num_phases

    def get_num_parameters(self):
        return self.num_phases if self.num_phases is not None else self.N_qubits

    def lower_initial(self, engine):
        if self.get_num_parameters() > self.N_qubits:
            raise ValueError("num_phases must not exceed N_qubits")
        engine.set_initial(_eng.INIT_PLUS_PHASES, self.get_num_parameters())
