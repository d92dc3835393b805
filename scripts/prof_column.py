"""Short driver for ncu: one n-qubit evaluation with the column kernels on (BOUND2 / FINAL2 on column_prog_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from qaoa_b200 import engine as eng
import parity_helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
edges, colors, table = H.regular_maxcut(n)
lut, _ = H.lut_from_table(table, 1)
E = eng.Engine(n, "complex64")
E.set_option("column_kernels", 1)
E.set_cost_maxkcut(n, edges, 1, lut)
ang = np.random.default_rng(1234).uniform(0, 2 * np.pi, size=2 * p)
for _ in range(2):
    out = E.energy(ang)
print(out)
