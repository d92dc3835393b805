"""Half-register (symmetric) vs full-register timing: MaxCut 3-regular, X mixer, Plus, complex64."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, networkx as nx
from qaoa_b200 import engine as eng
names = ["interpreter", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2", "Final2", "BoundX", "FinalX", "Bound2L4", "Final2L4"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
G = nx.random_regular_graph(3, n, seed=42)
edges = [(u, v, 1.0) for u, v in G.edges()]
ang = np.random.default_rng(1234).uniform(0, 2 * np.pi, size=2 * p)
modes = (1,) if len(sys.argv) > 4 and sys.argv[4] == "symonly" else (1, 0)
for sym in modes:
    E = eng.Engine(n, "complex64")
    E.set_option("symmetric", sym)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    E.set_option("time_passes", 1)
    for _ in range(3):
        out = E.energy(ang)
    E.counter("reset")
    ms = []
    for _ in range(reps):
        E.energy_batch_resident(ang[None, :]); out = E.fetch_results(1)[0]
        ms.append(E.counter("last_device_ms"))
    print("n=%d p=%d symmetric=%d: %.3f ms / evaluation (best %.3f)  E=%.9f norm=%.7f  plans_sym=%d  GB alg/eval %.1f" % (
        n, p, sym, np.mean(ms), np.min(ms), out[0], out[4], E.counter("symmetric_plans"), E.counter("bytes_alg") / reps / 1e9), flush=True)
    for i, nm in enumerate(names):
        c = E.counter("pass_count_%d" % i)
        if c:
            t = E.counter("pass_ms_%d" % i)
            print("   %-10s %3d launches/eval  %8.3f ms each  %6.0f GB/s" % (nm, c / reps, t / c, E.counter("pass_bytes_%d" % i) / t / 1e6), flush=True)
    E.close()
