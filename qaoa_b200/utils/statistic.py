"""Streaming sample statistics and their exact (shots -> infinity) counterpart.

``Statistic`` keeps the public behaviour of the reference's class of the same name
(reference qaoa/utils/statistic.py:44-142: West/Welford weighted mean and variance,
extremes with their bit-strings, CVaR over the sorted sample list).  ``ExactStatistic``
carries the numbers the CUDA engine returns for one evaluation and answers the same
getters, so ``OptResult.add_iteration`` (reference qaoa/qaoa.py:77-92) works unchanged.
"""

import numpy as np


class Statistic:
    def __init__(self, cvar=1):
        self.cvar = cvar
        self.reset()

    def reset(self):
        self.W = 0
        self.E = 0
        self.S = 0
        self.maxval = float("-inf")
        self.minval = float("inf")
        self.maxSols = []
        self.minSols = []
        self.all_values = np.array([])

    def add_sample(self, value, weight, string):
        # extremes first (ties extend the solution lists)
        if value > self.maxval:
            self.maxval, self.maxSols = value, [string]
        elif value == self.maxval:
            self.maxSols.append(string)
        if value < self.minval:
            self.minval, self.minSols = value, [string]
        elif value == self.minval:
            self.minSols.append(string)
        # weighted incremental mean / sum of squares (West 1979)
        self.W += weight
        previous = self.E
        self.E = previous + (weight / self.W) * (value - previous)
        self.S += weight * (value - previous) * (value - self.E)
        if self.cvar < 1:
            at = np.searchsorted(self.all_values, value)
            self.all_values = np.insert(self.all_values, at, np.full(int(weight), value))

    def get_E(self):
        return self.E

    def get_Variance(self):
        return self.S / (self.W - 1)

    def get_max(self):
        return self.maxval

    def get_min(self):
        return self.minval

    def get_max_sols(self):
        return self.maxSols

    def get_min_sols(self):
        return self.minSols

    def get_CVaR(self):
        if self.cvar < 1:
            keep = int(np.round(self.cvar * len(self.all_values)))
            return np.sum(self.all_values[-keep:]) / keep
        return self.get_E()


class ExactStatistic:
    """Result of one exact evaluation: E[cost], E[cost^2], min/max cost over the support.

    Differences from the sampled ``Statistic`` (documented in DESIGN.md): the variance is the
    population variance ``E[c^2] - E[c]^2`` (the reference divides by W-1, statistic.py:103, which
    is undefined for exact probabilities) and the extremes are taken over the whole support of
    |psi|^2 instead of the sampled strings.
    """

    def __init__(self, cvar=1):
        self.cvar = cvar
        self.reset()

    def reset(self):
        self.E = 0.0
        self.E2 = 0.0
        self.minval = float("inf")
        self.maxval = float("-inf")
        self.cvar_value = None
        self.maxSols = []
        self.minSols = []
        self._sampled = None

    def set_exact(self, e, e2, cmin, cmax, cvar_value=None):
        self.E, self.E2, self.minval, self.maxval = float(e), float(e2), float(cmin), float(cmax)
        self.cvar_value = cvar_value

    def add_sample(self, value, weight, string):
        """Sampled use of the same object (classical post-processing of histograms, utils/post.py): after reset() the
        first add_sample switches to the reference's streaming statistics until the next reset()."""
        if self._sampled is None:
            self._sampled = Statistic(cvar=self.cvar)
        sm = self._sampled
        sm.add_sample(value, weight, string)
        self.E, self.minval, self.maxval = sm.E, sm.minval, sm.maxval
        self.maxSols, self.minSols = sm.maxSols, sm.minSols

    def get_E(self):
        return self.E

    def get_Variance(self):
        if self._sampled is not None:
            return self._sampled.get_Variance()
        return max(self.E2 - self.E * self.E, 0.0)

    def get_max(self):
        return self.maxval

    def get_min(self):
        return self.minval

    def get_max_sols(self):
        return self.maxSols

    def get_min_sols(self):
        return self.minSols

    def get_CVaR(self):
        if self._sampled is not None:
            return self._sampled.get_CVaR()
        if self.cvar < 1:
            if self.cvar_value is None:
                raise NotImplementedError("exact CVaR was not computed for this evaluation")
            return self.cvar_value
        return self.E
