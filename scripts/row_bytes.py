"""DRAM efficiency of the tile access pattern vs contiguous row length (c low bits => 2^c*8 B rows)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from qaoa_b200 import engine as eng
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
E = eng.Engine(n, "complex64")
GB = 2 * 8 * 2.0**n / 1e9
for pf in (0,):
    E.set_option("prefetch_dist", pf)
    for c in (5, 4, 3, 2):
        for hb in (c, 12, n - (14 - c)):
            ms = E.debug_tile_program(c * 100 + hb, [(1, 0, 0), (2, 0, 0)], reps=5)
            print("rows of %3d B, other bits from %2d: %7.3f ms %6.0f GB/s" % (8 << c, hb, ms, GB / ms * 1e3), flush=True)
