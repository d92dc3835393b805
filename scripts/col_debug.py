"""debug: column kernels at one size / depth (run in a fresh process: a kernel fault poisons the context)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_helpers as H
from qaoa_b200 import engine as eng
n, sym, p = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
edges, colors, table = H.regular_maxcut(n + (n & 1))
edges = [e for e in edges if e[0] < n and e[1] < n]
lut, _ = H.lut_from_table(table, 1)
E = eng.Engine(n, "complex64")
E.set_option("symmetric", sym)
E.set_cost_maxkcut(n, edges, 1, lut)
ang = np.random.default_rng(n).uniform(0, 2 * np.pi, size=2 * p)
old = E.energy(ang)
d = eng.debug_plan(n, "complex64", symmetric=bool(sym), p=p)
print(n, sym, p, [(ps["prog"], d["shapes"][ps["shape"]]["pos"][5]) for ps in d["passes"]], flush=True)
E.set_option("column_kernels", 1)
if len(sys.argv) > 4:
    E.set_option("column_sync", int(sys.argv[4]))
try:
    new = E.energy(ang)
    new2 = E.energy(ang)
    print("  repeat", new2[0] - new[0])
    print("  OK", old[0], new[0], abs(old[0] - new[0]) / abs(old[0]), flush=True)
except Exception as exc:
    print("  FAIL", str(exc)[:200], flush=True)
