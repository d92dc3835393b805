"""MaxKCutGrover: Grover mixer over the feasible colour strings of a Max-k-Cut encoding
(reference qaoa/mixers/maxkcut_grover_mixer.py:12-136).

tensorized=False: ONE Grover mixer 1 + (e^{-i beta} - 1)|S><S| whose state is the tensor product of the node states,
|S> = |s_node>^(x)num_V -- lowered as the engine's Grover mixer over an uploaded statevector.
tensorized=True: the tensor product of num_V node-level Grover mixers sharing one beta -- lowered as the engine's
node Grover mixer (a dense reflection on every b-qubit node register).
Binary encodings only (the one-hot problem class is outside this engine's scope)."""

import numpy as np

from .base_mixer import Mixer
from qaoa_b200 import engine as _eng
from qaoa_b200.initialstates.dicke_initialstate import Dicke
from qaoa_b200.initialstates.plus_initialstate import Plus
from qaoa_b200.initialstates.dicke1_2_initialstate import Dicke1_2
from qaoa_b200.initialstates.lessthank_initialstate import LessThanK
from qaoa_b200.initialstates.tensor_initialstate import Tensor, amplitudes_of


class MaxKCutGrover(Mixer):
    def __init__(self, k_cuts: int, problem_encoding: str, color_encoding: str, tensorized: bool) -> None:
        if (k_cuts < 2) or (k_cuts > 8):
            raise ValueError("k_cuts must be 2 or more, and is not implemented for k_cuts > 8")
        if problem_encoding not in ["onehot", "binary"]:
            raise ValueError('problem_encoding must be in ["onehot", "binary"]')
        self.k_cuts = k_cuts
        self.problem_encoding = problem_encoding
        self.color_encoding = color_encoding
        self.tensorized = tensorized
        if (self.problem_encoding == "binary") and self.is_power_of_two():
            print("k_cuts is a power of two. You might want to use the X-mixer instead.")
        # for k=6, max_balanced == Dicke1_2
        if k_cuts == 6 and (color_encoding not in ["max_balanced", "Dicke1_2", "LessThanK"]):
            raise ValueError('color_encoding must be in ["LessThanK", "Dicke1_2", max_balanced]')

    def is_power_of_two(self) -> bool:
        return bool(self.k_cuts > 0 and (self.k_cuts & (self.k_cuts - 1)) == 0)

    def set_numV(self, k):
        num_V = self.N_qubits / k
        if not float(num_V).is_integer():
            raise ValueError("Total qubits=" + str(self.N_qubits) + " is not a multiple of " + str(k))
        self.num_V = int(num_V)

    def node_state(self):
        """the one-node state the reference builds its Grover mixer from (maxkcut_grover_mixer.py:100-118)"""
        if self.problem_encoding == "binary":
            self.k_bits = int(np.ceil(np.log2(self.k_cuts)))
            self.set_numV(self.k_bits)
            if self.is_power_of_two():
                one = Plus()
                one.N_qubits = self.k_bits
            elif self.k_cuts == 6 and self.color_encoding in ["max_balanced", "Dicke1_2"]:
                one = Dicke1_2()
            else:
                one = LessThanK(self.k_cuts)
        else:
            self.set_numV(self.k_cuts)
            one = Dicke(1, self.k_cuts)
        return one

    def create_circuit(self) -> None:
        self.node_state()
        self.circuit = None

    def lower_mixer(self, engine, trotter=1):
        one = self.node_state()
        if self.tensorized:
            if one.N_qubits > 3:
                raise NotImplementedError("tensorised Grover mixers over node registers of more than 3 qubits")
            engine.set_mixer(_eng.MIXER_NODE_GROVER, trotter=trotter, sub_k=one.N_qubits, sub_vec=amplitudes_of(one))
        else:
            vec = Tensor(one, self.num_V).state_vector()
            engine.set_mixer(_eng.MIXER_GROVER, trotter=trotter, sub_kind=_eng.INIT_STATEVECTOR, sub_vec=vec)
