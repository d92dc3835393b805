"""Landscape throughput (BASELINE config 2): MaxCut 3-regular n=24, X mixer, p=1, 100x100 grid."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, networkx as nx
from qaoa_b200 import QAOA, problems, mixers, initialstates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
g = int(sys.argv[2]) if len(sys.argv) > 2 else 100
G = nx.random_regular_graph(3, n, seed=42)
q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(), dtype="complex64")
grid = {"gamma": [0, 2 * np.pi, g], "beta": [0, 2 * np.pi, g]}
q.sample_cost_landscape({"gamma": [0, 2 * np.pi, 4], "beta": [0, 2 * np.pi, 4]})
for bp in [int(a) for a in sys.argv[3:]] or [0]:
    q._get_engine().set_option("batch_points", bp)
    t0 = time.time(); q.sample_cost_landscape(grid); dt = time.time() - t0
    print("n=%d %dx%d grid batch_points=%d: %.3f s -> %.0f points/s (min Exp %.6f)" % (n, g, g, bp, dt, g * g / dt, q.Exp_sampled_p1.min()), flush=True)
eng = q._get_engine()
eng.set_option("time_passes", 1)
eng.counter("reset")
t0 = time.time(); q.sample_cost_landscape(grid); dt = time.time() - t0
names = ["interpreter", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2", "Final2", "BoundX", "FinalX", "Bound2L4", "Final2L4"]
for i, nm in enumerate(names):
    c = eng.counter("pass_count_%d" % i)
    if c:
        ms = eng.counter("pass_ms_%d" % i)
        print("  %-10s launches %4d  total %8.2f ms  per point %6.2f us  %6.0f GB/s" % (nm, c, ms, 1e3 * ms / (g * g), eng.counter("pass_bytes_%d" % i) / ms / 1e6))
print("  wall %.3f s; launches %d" % (dt, eng.counter("launches")))
for sched in (1,):
    eng.set_option("schedule", sched)
    t0 = time.time(); q.sample_cost_landscape(grid); dt = time.time() - t0
    print("schedule=%d: %.3f s -> %.0f points/s" % (sched, dt, g * g / dt))
