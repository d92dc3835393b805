"""BASELINE config 4: QUBO / Portfolio n=30, Grover mixer + Dicke initial state, and multi-angle X mixer, p=3.
Synthetic inputs as in SURVEY.md section 8d.  Prints one JSON line per case (ms per evaluation, algorithmic GB/s)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from qaoa_b200 import QAOA, initialstates, mixers, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = 3


def timed(q, n_angles, label, dtype):
    eng = q._get_engine()
    ang = np.random.default_rng(1234).uniform(0, 2 * np.pi, size=n_angles)
    if "QUBO" in label:
        per = n_angles // p
        ang[0::per] *= 0.02           # gammas: costs are O(100)
    out = eng.energy(ang)
    eng.counter("reset")
    t0 = time.perf_counter()
    for _ in range(reps):
        out = eng.energy(ang)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    gb = eng.counter("bytes_alg") / reps / 1e9
    print(json.dumps({"case": label, "n": n, "p": p, "dtype": dtype, "ms_per_eval": ms, "algorithmic_gb": gb,
                      "gbs": gb / ms * 1e3, "launches_per_eval": eng.counter("launches") / reps,
                      "E": float(out[0]), "norm": float(out[4])}), flush=True)
    eng.close()


rng = np.random.RandomState(0)
A = rng.randn(n, 101)
cov, mu = np.cov(A), A.mean(axis=1)
k = n // 3
for dtype in ("complex64", "complex128"):
    q = QAOA(problem=problems.PortfolioOptimization(risk=0.5, budget=k, cov_matrix=cov, exp_return=mu, penalty=0),
             mixer=mixers.Grover(initialstates.Dicke(k)), initialstate=initialstates.Dicke(k), dtype=dtype)
    timed(q, 2 * p, "Portfolio + Grover(Dicke(%d)) + Dicke(%d)" % (k, k), dtype)
rs = np.random.RandomState(42)
Q, c = rs.rand(n, n), 3 * (rs.rand(n) - 0.5)
q = QAOA(problem=problems.QUBO(Q=Q, c=c, b=0.0), mixer=mixers.XMultiAngle(), initialstate=initialstates.Plus(), dtype="complex64")
timed(q, p * (1 + n), "QUBO + XMultiAngle + Plus", "complex64")
q = QAOA(problem=problems.QUBO(Q=Q, c=c, b=0.0), mixer=mixers.X(), initialstate=initialstates.Plus(), dtype="complex64")
timed(q, 2 * p, "QUBO + X + Plus", "complex64")
q = QAOA(problem=problems.QUBO(Q=Q, c=c, b=0.0), mixer=mixers.XY(case="ring"), initialstate=initialstates.Dicke(k), dtype="complex64")
timed(q, 2 * p, "QUBO + XY(ring) + Dicke(%d)" % k, "complex64")
