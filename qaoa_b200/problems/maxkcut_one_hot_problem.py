"""Max-k-Cut in the one-hot encoding: node v owns the k qubits [v k, (v+1) k), exactly one of which is 1 in a
feasible string (reference qaoa/problems/maxkcut_one_hot_problem.py:9-116).

cost(string): the colour of a node is read off the LAST '1' of its segment (`binstringToLabels`, :44-68: the first '1'
of the reversed segment; an empty segment is an error), cost = sum of w_uv over edges whose end colours differ (:70-89).
The graph is used with its own integer node ids and every edge needs a "weight" (:85-87).

Phase separator.  The reference's circuit (:91-116) applies, per edge and per colour c, CX - RZ(w gamma) - CX on the
qubit pair (k i + c, k j + c), i.e. exp(-i (w gamma / 2) sum_c Z Z).  On one-hot strings sum_c Z Z = k for equal colours
and k - 4 for different ones, so up to a global phase the circuit is exp(+i * 2 gamma * cost): TWICE the angle of
exp(+i gamma cost).  The engine therefore runs its diagonal phase separator with 2 gamma (`engine_gammas`), so that
identical angles give the reference's results.  Strings outside the one-hot subspace are never populated by the one-hot
initial states and mixers (Dicke(1, k) per node, XY / Grover over it); their diagonal entries are set to 0."""

import numpy as np

from .base_problem import Problem


class MaxKCutOneHot(Problem):
    def __init__(self, G, k_cuts: int) -> None:
        super().__init__()
        if (k_cuts < 2) or (k_cuts > 8):
            raise ValueError("k_cuts must be 2 or more, and is not implemented for k_cuts > 8")
        self.G = G
        self.num_V = self.G.number_of_nodes()
        self.k_cuts = k_cuts
        self.N_qubits = self.num_V * self.k_cuts

    def binstringToLabels(self, string: str) -> str:
        k = self.k_cuts
        labels = ""
        for v in range(self.num_V):
            segment = string[v * k: (v + 1) * k]
            last = segment.rfind("1")
            if last == -1:
                raise ValueError(f"Segment {segment} from {string} is not a valid encoding")
            labels += str(len(segment) - 1 - last)
        return labels

    def cost(self, string: str):
        labels = self.binstringToLabels(string)
        C = 0
        for (i, j) in self.G.edges():
            li = min(self.k_cuts - 1, int(labels[int(i)]))
            lj = min(self.k_cuts - 1, int(labels[int(j)]))
            if li != lj:
                C += self.G[i][j]["weight"]
        return C

    def isFeasible(self, string):
        k = self.k_cuts
        return all(string[v * k: (v + 1) * k].count("1") == 1 for v in range(self.num_V))

    def engine_gammas(self, gammas):
        return [2.0 * g for g in gammas]

    def cost_diagonal(self):
        """cost for every basis index (bit q of the index = character q of the string), 0 where a segment is empty"""
        n, k = self.N_qubits, self.k_cuts
        if n > 26:
            raise NotImplementedError("one-hot Max-k-Cut beyond 26 qubits (host-generated diagonal)")
        x = np.arange(1 << n, dtype=np.int64)
        labels, valid = [], np.ones(1 << n, dtype=bool)
        for v in range(self.num_V):
            seg = (x >> (v * k)) & ((1 << k) - 1)
            top = np.zeros(1 << n, dtype=np.int64)          # position of the highest set bit = last '1' of the segment
            for c in range(k):
                top = np.where((seg >> c) & 1, c, top)
            valid &= seg != 0
            labels.append(np.minimum(k - 1, k - 1 - top))
        diag = np.zeros(1 << n, dtype=np.float64)
        for (i, j) in self.G.edges():
            diag += np.where(labels[int(i)] != labels[int(j)], float(self.G[i][j]["weight"]), 0.0)
        diag[~valid] = 0.0
        return diag

    def lower_cost(self, engine):
        engine.set_cost_diagonal(self.cost_diagonal())
