"""MaxKCutFeasible: every node register in a superposition of feasible colour strings only
(reference qaoa/initialstates/maxkcut_feasible_initialstate.py:11-103): a Tensor of LessThanK(k) / Dicke1_2 (binary
encoding) or Dicke(1, k) (one-hot encoding) node states."""


from .base_initialstate import InitialState
from .dicke_initialstate import Dicke
from .dicke1_2_initialstate import Dicke1_2
from .lessthank_initialstate import LessThanK
from .tensor_initialstate import Tensor
from qaoa_b200 import engine as _eng


class MaxKCutFeasible(InitialState):
    def __init__(self, k_cuts: int, problem_encoding: str, color_encoding: str = "LessThanK") -> None:
        self.k_cuts = k_cuts
        self.problem_encoding = problem_encoding
        self.color_encoding = color_encoding
        if problem_encoding not in ["onehot", "binary"]:
            raise ValueError('case must be in ["onehot", "binary"]')
        if problem_encoding == "binary":
            if k_cuts == 6 and (color_encoding not in ["Dicke1_2", "LessThanK"]):
                raise ValueError('color_encoding must be in ["LessThanK", "Dicke1_2"]')
        # strings (character i = qubit i of the node register) outside the feasible set, :53-68
        if self.k_cuts == 3:
            self.infeasible = ["11"]
        elif self.k_cuts == 5:
            self.infeasible = ["100", "111", "101"] if self.color_encoding == "max_balanced" else ["101", "110", "111"]
        elif self.k_cuts == 6:
            self.infeasible = ["000", "111"] if self.color_encoding in ["Dicke1_2", "max_balanced"] else ["110", "111"]
        elif self.k_cuts == 7:
            self.infeasible = ["111"]

    def node_state(self):
        if self.problem_encoding == "binary":
            if self.k_cuts == 6 and self.color_encoding == "Dicke1_2":
                return Dicke1_2()
            return LessThanK(self.k_cuts)
        return Dicke(1, self.k_cuts)

    def create_circuit(self) -> None:
        one = self.node_state()
        num_V = self.N_qubits / one.N_qubits
        if not float(num_V).is_integer():
            raise ValueError("Total qubits=" + str(self.N_qubits) + " is not a multiple of " + str(one.N_qubits))
        self.k_bits = one.N_qubits
        self.num_V = int(num_V)
        self.tensor = Tensor(one, self.num_V)
        self.circuit = None

    def state_vector(self):
        self.create_circuit()
        return self.tensor.state_vector()

    def state_descriptor(self):
        return _eng.INIT_STATEVECTOR, 0, self.state_vector()
