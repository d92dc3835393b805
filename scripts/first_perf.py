"""Quick timing of the tile path (not the bench): ms per energy evaluation vs n, p=5."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, networkx as nx
from qaoa_b200 import engine as eng
import parity_helpers as H

for n in [int(a) for a in sys.argv[1:]] or [24, 28, 30, 32]:
    t0 = time.time()
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E = eng.Engine(n, "complex64")
    E.set_cost_maxkcut(n, edges, 1, lut)
    t1 = time.time()
    rng = np.random.default_rng(1234)
    p = 5
    ang = rng.uniform(0, 2 * np.pi, size=2 * p)
    out = E.energy(ang)   # builds plan + streams
    t2 = time.time()
    ts = []
    for _ in range(3):
        a = time.time(); out = E.energy(ang); ts.append(time.time() - a)
    npass = E.counter("n_passes")
    gb = (2 * npass - 2) * 8 * 2.0**n / 1e9
    print("n=%d setup %.2fs first %.2fs | eval %.2f ms (best of 3) passes=%d -> %.0f GB/s actual | E=%.6f norm=%.6f"
          % (n, t1 - t0, t2 - t1, 1e3 * min(ts), npass, gb / min(ts), out[0], out[4]), flush=True)
    E.close()
