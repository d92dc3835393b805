"""Emulates the reference's NOISY pipeline for the XY-ring CVaR(0.1) run of examples/PortfolioOptimization/PortOpt.ipynb on
the CPU oracle (test tooling, not collected by pytest): 1024 shots per evaluation, Statistic's CVaR (mean of the best 102
samples), 10 x 10 landscape over [0, 2 pi)^2, COBYLA from the arg-min -- once with the XY convention used in this repository
(hopping term -i sin(beta/4), [3p] XXPlusYYGate) and once with the opposite hopping sign.  qiskit-aer printed
cost(depth 1) = -0.28425 (tests/golden/notebook_runs.json).

    python tests/emulate_portopt_xy.py [seeds=8]

Round-1 result (8 seeds each): convention used here -0.161 ... -0.179 (never below -0.18); opposite sign -0.2817 ... -0.2875.
See DESIGN.md section 4, "Open parity question (XY mixer)".
"""
import json
import os
import sys

import numpy as np
from scipy.optimize import minimize

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import qaoa_oracle as orc  # noqa: E402
import test_host_api as T  # noqa: E402


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    with open(os.path.join(HERE, "golden", "notebook_runs.json")) as f:
        run = json.load(f)["runs"]["PortOpt"]
    po, n = T._portopt(run), run["n_assets"]
    diag = np.array([po.cost("".join("1" if (x >> i) & 1 else "0" for i in range(n))) for x in range(1 << n)])
    spec = orc.Spec(n, diag, mixer=("xy", orc.xy_pairs(n, "ring")), init=orc.dicke_state(n, run["budget"]))

    def sampled_cvar(a, rng, alpha=0.1, shots=1024):
        p = np.abs(orc.evolve(spec, np.asarray(a, dtype=float))) ** 2
        c = np.sort(diag[rng.choice(p.size, size=shots, p=p / p.sum())])
        return -c[-int(round(alpha * shots)):].mean()

    grid = np.linspace(0, 2 * np.pi, 10, endpoint=False)
    for sign, label in ((+1, "convention used here (-i sin)"), (-1, "opposite hopping sign")):
        finals = []
        for seed in range(seeds):
            rng = np.random.default_rng(seed)
            land = np.array([[sampled_cvar([g, sign * b], rng) for g in grid] for b in grid])
            ib, ig = np.unravel_index(np.argmin(land), land.shape)
            res = minimize(lambda a: sampled_cvar([a[0], sign * a[1]], rng), np.array([grid[ig], grid[ib]]), method="COBYLA",
                           options={"maxiter": 1000, "rhobeg": 1.0})
            finals.append(round(float(res.fun), 4))
        print("%-32s cost(depth 1) over %d seeds: %s   (qiskit-aer printed %.5f)" % (label, seeds, finals, run["cost"][1][0]))


if __name__ == "__main__":
    main()
