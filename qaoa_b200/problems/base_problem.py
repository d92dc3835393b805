"""Problem plugin base classes.

Same public surface as the reference (reference qaoa/problems/base_problem.py:29-149):
``cost(string)``, ``isFeasible(string)``, ``get_num_parameters()``, ``computeMinMaxCosts()``,
attributes ``N_qubits`` / ``N_ancilla_qubits`` / ``circuit``.  Instead of a qiskit circuit, a
supported problem *lowers* itself onto the CUDA engine through ``lower_cost(engine)``; the phase
separator is then the diagonal ``exp(+i gamma cost(x))`` (sign pinned by the reference's own
validator, qaoa/utils/validation.py:43).
"""

from abc import ABC, abstractmethod
import itertools


class BaseProblem(ABC):
    def __init__(self) -> None:
        self.circuit = None
        self.N_ancilla_qubits = 0


class Problem(BaseProblem):
    @abstractmethod
    def cost(self, string):
        """Value to MAXIMISE for the bit-string ``string`` (string[i] = qubit i)."""

    def create_circuit(self):
        """The engine never executes gates: there is no circuit to build.

        Kept so code written against the reference (which calls ``create_circuit()`` before
        reading ``get_num_parameters()``, qaoa/qaoa.py:441-448) still runs.
        """
        self.circuit = None

    def lower_cost(self, engine):
        """Installs this problem's cost diagonal on ``engine`` (qaoa_b200.engine.Engine).

        The default uses the escape hatch: it evaluates ``cost`` for every bit-string on the host
        (Python loop, as the reference's computeMinMaxCosts does) and uploads the diagonal.  It is
        meant for user-defined problems of modest size; built-in problems override it with an
        on-device generator.
        """
        import numpy as np

        n = self.N_qubits
        if n > 22:
            raise NotImplementedError(
                "%s has no device lowering and 2^%d host cost() calls are not practical" % (type(self).__name__, n)
            )
        diag = np.empty(1 << n)
        for x in range(1 << n):
            diag[x] = self.cost("".join("1" if (x >> i) & 1 else "0" for i in range(n)))
        engine.set_cost_diagonal(diag)

    def validate_circuit(self, t=1, flip=True, atol=1e-8, rtol=1e-8, engine=None):
        """(ok, report): the device's phase separator exp(+i t c_dev(x)) against exp(+i t cost(string)) for every string, up
        to a global phase (reference base_problem.py:136-149 -> utils/validation.py:16-79); <= 16 qubits."""
        from qaoa_b200.utils import validation

        return validation.check_phase_separator_exact_problem(self, t=t, flip=flip, atol=atol, rtol=rtol, engine=engine)

    def engine_gammas(self, gammas):
        """this layer's cost parameters -> the gammas the device phase separator exp(+i gamma cost) takes (identity
        unless the reference's phase-separator CIRCUIT realises a different multiple of its cost function)"""
        return list(gammas)

    def isFeasible(self, string):
        return True

    def get_num_parameters(self):
        return 1

    _MINMAX_BRUTE_FORCE_QUBITS = 16

    def computeMinMaxCosts(self, engine=None):
        """(min, max) of -cost over the feasible strings (reference base_problem.py:121-134, a Python loop over all 2^n
        strings).  Unconstrained problems beyond 16 qubits -- or whenever an ``engine`` is given -- take the extremes of
        the cost diagonal from a device reduction instead (qaoa_cost_minmax); constrained problems keep the loop."""
        unconstrained = type(self).isFeasible is Problem.isFeasible
        if unconstrained and (engine is not None or self.N_qubits > self._MINMAX_BRUTE_FORCE_QUBITS):
            if engine is None:
                from qaoa_b200.engine import Engine

                engine = Engine(self.N_qubits, dtype="complex128")
            self.lower_cost(engine)
            cmin, cmax = engine.cost_minmax()
            return -cmax, -cmin
        lo, hi = float("inf"), float("-inf")
        for bits in itertools.product("01", repeat=self.N_qubits):
            s = "".join(bits)
            if self.isFeasible(s):
                value = -self.cost(s)
                lo, hi = min(lo, value), max(hi, value)
        return lo, hi
