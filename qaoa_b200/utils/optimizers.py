"""Optimizer shims with the protocol the reference expects from qiskit-algorithms:
``opt = cls(**options); res = opt.minimize(fun, x0=...)`` with ``res.x`` / ``res.fun``
(reference qaoa/qaoa.py:858-883).  qiskit-algorithms is not a dependency of this package;
COBYLA and NELDER_MEAD delegate to scipy exactly like qiskit-algorithms' SciPyOptimizer does
([3p] defaults: maxiter=1000, rhobeg=1.0, tol=None for COBYLA).  Any user class honouring the
protocol is accepted as well.
"""

import numpy as np
from scipy.optimize import minimize as _scipy_minimize


class OptimizerResult:
    def __init__(self):
        self.x = None
        self.fun = None
        self.nfev = None
        self.nit = None

    def __repr__(self):
        return "OptimizerResult(fun=%r, nfev=%r)" % (self.fun, self.nfev)


class _SciPy:
    method = None

    def __init__(self, tol=None, **options):
        self._tol = tol
        self._options = options

    def minimize(self, fun, x0, jac=None, bounds=None):
        raw = _scipy_minimize(fun, np.asarray(x0, dtype=float), method=self.method, jac=jac, bounds=bounds,
                              tol=self._tol, options=self._options)
        res = OptimizerResult()
        res.x, res.fun = raw.x, raw.fun
        res.nfev, res.nit = raw.get("nfev"), raw.get("nit")
        return res

    def __str__(self):
        return self.__class__.__name__


class COBYLA(_SciPy):
    method = "COBYLA"

    def __init__(self, maxiter=1000, disp=False, rhobeg=1.0, tol=None, **kwargs):
        super().__init__(tol=tol, maxiter=maxiter, disp=disp, rhobeg=rhobeg, **kwargs)


class NELDER_MEAD(_SciPy):
    method = "Nelder-Mead"

    def __init__(self, maxiter=None, maxfev=1000, disp=False, xatol=1e-4, tol=None, adaptive=False, **kwargs):
        opts = dict(maxfev=maxfev, disp=disp, xatol=xatol, adaptive=adaptive, **kwargs)
        if maxiter is not None:
            opts["maxiter"] = maxiter
        super().__init__(tol=tol, **opts)


class L_BFGS_B(_SciPy):
    method = "L-BFGS-B"

    def __init__(self, maxfun=15000, maxiter=15000, ftol=2.220446049250313e-15, eps=1e-8, **kwargs):
        super().__init__(tol=None, maxfun=maxfun, maxiter=maxiter, ftol=ftol, eps=eps, **kwargs)


class SPSA:
    """Plain simultaneous-perturbation stochastic approximation (Spall 1998 gains).

    Deterministic for a given ``seed``.  This is NOT a re-implementation of qiskit-algorithms'
    calibrated SPSA; it honours the same protocol so scripts that pass ``[SPSA, {"maxiter": 10}]``
    (reference unittests/test_qaoa_core.py:78) keep working.
    """

    def __init__(self, maxiter=100, a=0.2, c=0.1, alpha=0.602, gamma=0.101, A=None, seed=0, **_ignored):
        self.maxiter, self.a, self.c, self.alpha, self.gamma = maxiter, a, c, alpha, gamma
        self.A = 0.1 * maxiter if A is None else A
        self.seed = seed

    def minimize(self, fun, x0, jac=None, bounds=None):
        rng = np.random.default_rng(self.seed)
        x = np.asarray(x0, dtype=float).copy()
        nfev = 0
        for k in range(self.maxiter):
            ak = self.a / (k + 1 + self.A) ** self.alpha
            ck = self.c / (k + 1) ** self.gamma
            delta = rng.choice([-1.0, 1.0], size=x.size)
            plus, minus = fun(x + ck * delta), fun(x - ck * delta)
            nfev += 2
            x = x - ak * (plus - minus) / (2.0 * ck) * delta
        res = OptimizerResult()
        res.x, res.fun, res.nfev, res.nit = x, fun(x), nfev + 1, self.maxiter
        return res

    def __str__(self):
        return "SPSA"
