"""Max-k-Cut with k a power of two, binary node encoding.

Cost semantics of the reference class (reference qaoa/problems/maxkcut_binary_powertwo.py:38-142).
``method`` only selected between two gate decompositions of the same diagonal there; it is
validated and stored but has no numerical effect here.
"""

import numpy as np

from .graph_problem import GraphProblem


class MaxKCutBinaryPowerOfTwo(GraphProblem):
    def __init__(self, G, k_cuts: int, method: str = "Diffusion", fix_one_node: bool = False) -> None:
        MaxKCutBinaryPowerOfTwo.validate_parameters(k_cuts, method)
        self.k_cuts = k_cuts
        self.method = method
        super().__init__(G, int(np.ceil(np.log2(k_cuts))), fix_one_node)
        self.construct_colors()

    @staticmethod
    def is_power_of_two(k) -> bool:
        return k > 0 and (k & (k - 1)) == 0

    @staticmethod
    def validate_parameters(k, method) -> None:
        if not MaxKCutBinaryPowerOfTwo.is_power_of_two(k):
            raise ValueError("k_cuts must be a power of two")
        if k < 2 or k > 8:
            raise ValueError("k_cuts must be 2 or more, and is not implemented for k_cuts > 8")
        valid_methods = ["PauliBasis", "Diffusion"]
        if method not in valid_methods:
            raise ValueError("method must be in " + str(valid_methods))

    def construct_colors(self):
        b = self.N_qubits_per_node
        # colour i+1 <-> the b-character binary string of i (most significant character first)
        self.colors = {"color%d" % (i + 1): [format(i, "0%db" % b)] for i in range(self.k_cuts)}
        self.bitstring_to_color = {s: name for name, group in self.colors.items() for s in group}
