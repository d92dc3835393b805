"""Landscape throughput (BASELINE config 2: n=24, p=1, 100x100) against the number of points per launch chunk."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from qaoa_b200 import engine as eng  # noqa: E402

n, g = 24, 100
E = eng.Engine(n, "complex64")
E.set_cost_maxkcut(n, bench.workload(n), 1, np.array([0, 1], dtype=np.uint8))
gam, bet = np.meshgrid(np.linspace(0, 2 * np.pi, g, endpoint=False), np.linspace(0, 2 * np.pi, g, endpoint=False))
rows = np.stack([gam.ravel(), bet.ravel()], axis=1)
E.energy_batch(rows[:64])
for bp in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 8, 16, 32, 64]:
    E.set_option("batch_points", bp)
    E.energy_batch(rows[:256])
    t0 = time.perf_counter()
    out = E.energy_batch(rows)
    dt = time.perf_counter() - t0
    print(json.dumps({"batch_points": bp, "points_per_s": len(rows) / dt, "min": float(-out[:, 0].max())}), flush=True)
