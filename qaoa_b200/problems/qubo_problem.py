"""Quadratic unconstrained binary optimisation: cost(x) = -(x^T Q x + c^T x + b).

Argument checks and cost follow the reference (reference qaoa/problems/qubo_problem.py:37-108);
the rz/rzz circuit (:111-150) is replaced by the engine's QUBO diagonal generator.
"""

import numpy as np

from .base_problem import Problem


class QUBO(Problem):
    def __init__(self, Q=None, c=None, b=None) -> None:
        super().__init__()
        assert type(Q) is np.ndarray, "Q needs to be a numpy ndarray, but is " + str(type(Q))
        assert Q.ndim == 2, "Q needs to be a 2-dimensional numpy ndarray, but has dim " + str(Q.ndim)
        assert Q.shape[0] == Q.shape[1], "Q needs to be a square matrix, but is " + str(Q.shape)
        n = Q.shape[0]
        self.N_qubits = n
        self.QUBO_Q = Q
        if c is None:
            c = np.zeros(n)
        assert type(c) is np.ndarray, "c needs to be a numpy ndarray, but is " + str(type(c))
        assert c.ndim == 1, "c needs to be a 1-dimensional numpy ndarray, but has dim " + str(c.ndim)
        assert c.shape[0] == n, "c is of size " + str(c.shape[0]) + " but should be compatible size to Q, meaning " + str(n)
        self.QUBO_c = c
        if b is None:
            b = 0.0
        assert np.isscalar(b), "b is expected to be scalar, but is " + str(b)
        self.QUBO_b = b

    def cost(self, string):
        return self.qubo_cost(string)

    def qubo_cost(self, string):
        x = np.fromiter((int(ch) for ch in string), dtype=np.int64, count=len(string))
        return -(x @ self.QUBO_Q @ x + self.QUBO_c @ x + self.QUBO_b)

    def lower_cost(self, engine):
        engine.set_cost_qubo(np.asarray(self.QUBO_Q, dtype=np.float64), np.asarray(self.QUBO_c, dtype=np.float64),
                             float(self.QUBO_b))
