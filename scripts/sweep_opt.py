"""Sweep an engine option and time energy evaluations: python sweep_opt.py n p key v1 v2 ..."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from qaoa_b200 import engine as eng
import parity_helpers as H
n, p, key = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
vals = [float(v) for v in sys.argv[4:]]
edges, colors, table = H.regular_maxcut(n)
lut, _ = H.lut_from_table(table, 1)
E = eng.Engine(n, "complex64")
E.set_cost_maxkcut(n, edges, 1, lut)
ang = np.random.default_rng(1234).uniform(0, 2 * np.pi, size=2 * p)
E.energy(ang)
for v in vals:
    E.set_option(key, v)
    E.energy(ang)
    ts = []
    for _ in range(4):
        a = time.time(); out = E.energy(ang); ts.append(time.time() - a)
    print("n=%d p=%d %s=%g : %.2f ms  E=%.6f" % (n, p, key, v, 1e3 * min(ts), out[0]), flush=True)
