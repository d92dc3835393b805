"""Portfolio optimisation as a QUBO (reference qaoa/problems/portfolio_problem.py:28-108)."""

import itertools
import math

import numpy as np

from .qubo_problem import QUBO


class PortfolioOptimization(QUBO):
    def __init__(self, risk, budget, cov_matrix, exp_return, penalty=0) -> None:
        self.risk = risk
        self.budget = budget
        self.cov_matrix = cov_matrix
        self.exp_return = exp_return
        self.penalty = penalty
        self.N_qubits = len(self.exp_return)
        # min x^T Q x + c^T x + b  (mapping of portfolio_problem.py:48-54)
        ones = np.ones_like(self.cov_matrix)
        Q = self.risk * self.cov_matrix + self.penalty * ones
        c = -self.exp_return - 2 * self.penalty * self.budget * np.ones_like(self.exp_return)
        b = self.penalty * self.budget * self.budget
        super().__init__(Q=Q, c=c, b=b)

    def _bits(self, string):
        x = np.fromiter((int(ch) for ch in string), dtype=np.int64, count=len(string))
        assert len(x) == len(self.exp_return), (
            "bitstring  " + string + " of wrong size. Expected " + str(len(self.exp_return)) + " but got " + str(len(x))
        )
        return x

    def cost(self, string):
        x = self._bits(string)
        value = self.risk * (x @ self.cov_matrix @ x) - self.exp_return @ x
        value += self.penalty * (x.sum() - self.budget) ** 2
        return -value

    def isFeasible(self, string):
        return math.isclose(self._bits(string).sum() - self.budget, 0, abs_tol=1e-7)

    def brute_force_solve(self):
        best, best_s = -np.inf, None
        n = self.N_qubits
        for ones in itertools.combinations(range(n), self.budget):
            s = "".join("1" if i in ones else "0" for i in range(n))
            v = self.cost(s)
            if v > best:
                best, best_s = v, s
        return best_s
