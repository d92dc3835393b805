"""Exact cover with optional subset weights, as a penalised QUBO
(reference qaoa/problems/exactcover_problem.py:31-143)."""

import itertools
import math

import numpy as np

from .qubo_problem import QUBO


class ExactCover(QUBO):
    def __init__(self, columns, weights=None, penalty_factor=None, allow_infeasible=False,
                 hamming_weight=None, scale_problem=False) -> None:
        self.columns = columns
        self.weights = weights
        self.penalty_factor = penalty_factor
        self.allow_infeasible = allow_infeasible
        rows, ncols = columns.shape
        if weights is None:
            self.weights = np.zeros(ncols)
        if hamming_weight is not None:
            assert type(hamming_weight) == int, "hamming_weight must be int, but is " + str(hamming_weight)
            assert 0 < hamming_weight < ncols, (
                "hamming_weight must between 0 and " + str(ncols) + ", but is " + str(hamming_weight))
        self.hamming_weight = hamming_weight
        if penalty_factor is None:
            if np.all(self.weights == 0):
                self.penalty_factor = 1
            elif self.hamming_weight:
                ordered = np.sort(self.weights)
                self.penalty_factor = np.sum(ordered[-hamming_weight:]) - np.sum(ordered[:hamming_weight])
            else:
                self.penalty_factor = np.sum(self.weights[self.weights > 0])
        if scale_problem:
            self._scale_problem()
        # C(x) = x^T Q x + c^T x + b   (exactcover_problem.py:78-80)
        c = self.weights - 2 * self.penalty_factor * (np.ones(rows) @ self.columns)
        Q = self.penalty_factor * (self.columns.T @ self.columns)
        b = self.penalty_factor * rows
        super().__init__(Q, c, b)
        self.N_qubits = ncols

    def _violation(self, x):
        return np.sum((1 - (self.columns @ x)) ** 2)

    def cost(self, string):
        x = np.fromiter((int(ch) for ch in string), dtype=np.int64, count=len(string))
        return -(self.weights @ x + self.penalty_factor * self._violation(x))

    def isFeasible(self, string):
        x = np.fromiter((int(ch) for ch in string), dtype=np.int64, count=len(string))
        return math.isclose(self._violation(x), 0, abs_tol=1e-7) or self.allow_infeasible

    def _scale_problem(self):
        assert self.hamming_weight is not None, (
            "scaling currently only supported for exact problems with a given hamming_weight")
        hm = self.hamming_weight
        rows = self.columns.shape[0]
        w = np.array(self.weights, dtype=float)
        omega = w - np.sum(np.sort(w)[:hm]) / hm            # lower bound of the objective -> 0
        so = np.sort(omega)
        omega_penalty = np.sum(so[-hm:]) - np.sum(so[:hm])
        scaling = 1.0 / (omega_penalty * (1 + rows * (hm - 1) ** 2))
        self.weights = scaling * omega
        self.penalty_factor = scaling * omega_penalty

    def brute_force_solve(self, return_num_feasible=False):
        n = self.N_qubits
        if self.hamming_weight is not None:
            strings = ("".join("1" if i in ones else "0" for i in range(n))
                       for ones in itertools.combinations(range(n), self.hamming_weight))
        else:
            strings = ("".join(bits) for bits in itertools.product("01", repeat=n))
        best, best_s, feasible = -np.inf, None, 0
        for s in strings:
            v = self.cost(s)
            if v > best:
                best, best_s = v, s
            feasible += self.isFeasible(s)
        return (best_s, feasible) if return_num_feasible else best_s
