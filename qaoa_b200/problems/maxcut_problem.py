"""MaxCut = Max-2-Cut (reference qaoa/problems/maxcut_problem.py:23-39)."""

from .maxkcut_binary_powertwo import MaxKCutBinaryPowerOfTwo


class MaxCut(MaxKCutBinaryPowerOfTwo):
    def __init__(self, G, method: str = "Diffusion", fix_one_node: bool = False) -> None:
        super().__init__(G, k_cuts=2, method=method, fix_one_node=fix_one_node)
