"""BASELINE config 3 through the reference-facing API: MaxCut 3-regular n=32 (seed 42), X mixer, |+>, complex64,
`QAOA.optimize(depth=5)` on one B200 -- landscape seed, COBYLA per depth, INTERP between depths.

  python scripts/config3_optimize.py [n=32] [depth=5] [maxiter=25] [grid=12] [out=gpurun_out/config3_optimize.jsonl] [dtype=complex64]

(n=8 depth=3 maxiter=1000 grid=20 dtype=complex128 is BASELINE config 1, the README quick example.)

One JSON line per stage is appended (and flushed) as soon as the stage ends, so a run cut short by a time limit keeps
what it measured: the landscape (points, seconds), then per depth the number of energy evaluations COBYLA made, the wall
time, ms per evaluation as the optimiser sees it (host angles in, host statistics out, Python bookkeeping included) and
the best Exp.  `maxiter` bounds COBYLA (the reference default is 1000); the per-evaluation cost does not depend on it.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import networkx as nx  # noqa: E402


def main():
    arg = sys.argv[1:]
    n = int(arg[0]) if len(arg) > 0 else 32
    depth = int(arg[1]) if len(arg) > 1 else 5
    maxiter = int(arg[2]) if len(arg) > 2 else 25
    grid = int(arg[3]) if len(arg) > 3 else 12
    path = arg[4] if len(arg) > 4 else os.path.join(ROOT, "gpurun_out", "config3_optimize.jsonl")
    dtype = arg[5] if len(arg) > 5 else "complex64"
    os.makedirs(os.path.dirname(path), exist_ok=True)
    out = open(path, "a")

    def emit(rec):
        out.write(json.dumps(rec) + "\n")
        out.flush()
        os.fsync(out.fileno())
        print(json.dumps(rec), flush=True)

    t_start = time.perf_counter()
    from qaoa_b200 import QAOA, initialstates, mixers, problems
    from qaoa_b200.qaoa import B200Backend
    from qaoa_b200.utils import COBYLA

    G = nx.random_regular_graph(3, n, seed=42)
    q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(),
             backend=B200Backend(device=0, dtype=dtype), optimizer=[COBYLA, {"maxiter": maxiter}])
    t0 = time.perf_counter()
    q.createParameterizedCircuit(1)          # lowers problem / mixer / state: diagonal generated on the device
    q._evaluate(np.array([[0.3, 0.2]]))      # first evaluation builds the plan and the diagonal streams
    emit({"stage": "setup", "n": n, "dtype": dtype, "import_s": t0 - t_start, "lower_and_first_eval_s": time.perf_counter() - t0})

    spec = {"gamma": [0, 2 * np.pi, grid], "beta": [0, 2 * np.pi, grid]}
    t0 = time.perf_counter()
    q.sample_cost_landscape(spec)
    dt = time.perf_counter() - t0
    emit({"stage": "landscape", "n": n, "points": grid * grid, "seconds": dt, "ms_per_point": 1e3 * dt / (grid * grid),
          "min_exp": float(q.Exp_sampled_p1.min())})

    for d in range(1, depth + 1):
        t0 = time.perf_counter()
        q.optimize(depth=d, angles=spec)
        dt = time.perf_counter() - t0
        r = q.optimization_results[d]
        emit({"stage": "optimize", "n": n, "depth": d, "evaluations": r.num_fval(), "seconds": dt,
              "ms_per_evaluation": 1e3 * dt / max(1, r.num_fval()), "best_exp": float(q.get_Exp(d)),
              "angles": [float(a) for a in q.get_angles(d)], "cobyla_maxiter": maxiter})
    eng = q._get_engine()
    emit({"stage": "done", "n": n, "total_s": time.perf_counter() - t_start,
          "symmetric_register": bool(eng.counter("symmetric_plans") >= 1),
          "exp_per_depth": [float(e) for e in q.get_Exp()]})


if __name__ == "__main__":
    main()
