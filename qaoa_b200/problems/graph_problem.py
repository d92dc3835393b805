"""Graph-colouring problems (Max-k-Cut family) as cost diagonals.

Mirrors the cost side of the reference's GraphProblem (reference
qaoa/problems/graph_problem.py:31-57 ctor, :120-150 labels, :152-176 cost).  The circuit side
(:79-116 and the edge gadgets) is replaced by the engine's on-device diagonal generator.
"""

import numpy as np

from .base_problem import Problem
from qaoa_b200.utils.graphutils import GraphHandler


class GraphProblem(Problem):
    def __init__(self, G, N_qubits_per_node=1, fix_one_node: bool = False) -> None:
        super().__init__()
        self.fix_one_node = fix_one_node
        self.graph_handler = GraphHandler(G)
        self.num_V = self.graph_handler.G.number_of_nodes()
        self.N_qubits_per_node = N_qubits_per_node
        self.N_qubits = (self.num_V - self.fix_one_node) * self.N_qubits_per_node

    # -- labels and colours (subclasses provide self.colors / self.bitstring_to_color) -------
    def same_color(self, str1: str, str2: str) -> bool:
        return self.bitstring_to_color.get(str1) == self.bitstring_to_color.get(str2)

    def slice_string(self, string: str) -> list:
        b = self.N_qubits_per_node
        labels = [string[i:i + b] for i in range(0, (self.num_V - self.fix_one_node) * b, b)]
        if self.fix_one_node:
            labels.append(self.colors["color1"][0])
        return labels

    def cost(self, string: str) -> float | int:
        if len(string) != self.N_qubits:
            raise ValueError(
                f"Expected a string of length {self.N_qubits}, " f"but received length {len(string)}."
            )
        lab = self.slice_string(string)
        G = self.graph_handler.G
        total = 0
        for u, v in G.edges():
            if not self.same_color(lab[u], lab[v]):
                total += G[u][v].get("weight", 1)
        return total

    # -- lowering -----------------------------------------------------------------------------
    def colour_lut(self):
        """uint8 LUT indexed by the b-bit integer whose bit j is qubit v*b+j.

        A node label is the string slice whose FIRST character is the node's lowest qubit, so the
        integer is converted to that string before the reference's colour dictionary is consulted.
        """
        b = self.N_qubits_per_node
        ids = {name: i for i, name in enumerate(self.colors.keys())}
        lut = np.zeros(1 << b, dtype=np.uint8)
        for value in range(1 << b):
            label = "".join("1" if (value >> j) & 1 else "0" for j in range(b))
            lut[value] = ids[self.bitstring_to_color[label]]
        fixed = ids[self.bitstring_to_color[self.colors["color1"][0]]]
        return lut, fixed

    def lower_cost(self, engine):
        lut, fixed = self.colour_lut()
        engine.set_cost_maxkcut(self.num_V, self.graph_handler.edge_list(), self.N_qubits_per_node, lut,
                                fixed_node=self.fix_one_node, fixed_colour=fixed)
