"""Max-k-Cut for k in {3, 5, 6, 7} with surplus binary labels merged into colour classes.

Only the COST side of the reference class is provided (colour classes of reference
qaoa/problems/maxkcut_binary_fullH.py:119-192); its hard-coded gate tables are irrelevant for a
diagonal kernel.  ``fix_one_node`` follows GraphProblem semantics.
"""

import numpy as np

from .graph_problem import GraphProblem

_CLASSES = {
    (3, "LessThanK"): [["00"], ["01"], ["10", "11"]],
    (5, "LessThanK"): [["000"], ["001"], ["010"], ["011"], ["100", "101", "110", "111"]],
    (5, "max_balanced"): [["000", "001"], ["010"], ["011"], ["100", "101"], ["110", "111"]],
    (6, "LessThanK"): [["000"], ["001"], ["010"], ["011"], ["100"], ["101", "110", "111"]],
    (6, "max_balanced"): [["000", "001"], ["010"], ["011"], ["100", "101"], ["110"], ["111"]],
    (7, "LessThanK"): [["000"], ["001"], ["010"], ["011"], ["100"], ["101"], ["110", "111"]],
}


class MaxKCutBinaryFullH(GraphProblem):
    def __init__(self, G, k_cuts: int, color_encoding: str, method: str = "Diffusion",
                 fix_one_node: bool = False) -> None:
        MaxKCutBinaryFullH.validate_parameters(k_cuts, method, fix_one_node)
        self.k_cuts = k_cuts
        self.method = method
        self.color_encoding = color_encoding
        super().__init__(G, int(np.ceil(np.log2(k_cuts))), fix_one_node)
        self.construct_colors()

    @staticmethod
    def validate_parameters(k, method, fix_one_node) -> None:
        """the reference's checks and messages (maxkcut_binary_fullH.py:90-117)"""
        if k not in (3, 5, 6, 7):
            raise ValueError("k_cuts must be in " + str([3, 5, 6, 7]))
        if method not in ("PauliBasis", "PowerOfTwo", "Diffusion"):
            raise ValueError("method must be in " + str(["PauliBasis", "PowerOfTwo", "Diffusion"]))
        if method == "PowerOfTwo" and fix_one_node:
            raise ValueError('For the PowerOfTwo method it "fix_one_node" is not implemented. '
                             'Use "PauliBasis" or "Diffusion" instead.')

    def construct_colors(self):
        groups = _CLASSES.get((self.k_cuts, self.color_encoding))
        if groups is None:
            raise ValueError("invalid or unspecified color_encoding")
        self.colors = {"color%d" % (i + 1): list(g) for i, g in enumerate(groups)}
        self.bitstring_to_color = {s: name for name, group in self.colors.items() for s in group}
