"""Short driver for ncu: one evaluation of each BASELINE config 4 case (bench.py config4_leg's engines) at size n."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from qaoa_b200 import engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "grover,xmulti,xy"
rng = np.random.default_rng(1234)
ang2 = rng.uniform(0, 2 * np.pi, size=2 * p)
Qp, cp, bp, k = bench.portfolio_instance(n)
Q, c = bench.qubo_instance(n)
if "grover" in which:
    E = eng.Engine(n, "complex64")
    E.set_option("feasible_subspace", 0)
    E.set_cost_qubo(Qp, cp, bp)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
    print("grover full", E.energy(ang2))
    E.close()
if "xmulti" in which:
    E = eng.Engine(n, "complex64")
    E.set_cost_qubo(Q, c, 0.0)
    E.set_mixer(eng.MIXER_XMULTI)
    am = rng.uniform(0, 2 * np.pi, size=p * (1 + n))
    am[0::1 + n] *= 0.02
    print("xmulti", E.energy(am))
    E.close()
if "xy" in which:
    E = eng.Engine(n, "complex64")
    E.set_cost_qubo(Q, c, 0.0)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_XY, pairs=[[i, (i + 1) % n] for i in range(n)])
    a = ang2.copy()
    a[0::2] *= 0.02
    print("xy", E.energy(a))
    E.close()
