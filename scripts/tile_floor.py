"""Where does a tile pass spend its time?  Times step programs of increasing content."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from qaoa_b200 import engine as eng
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
E = eng.Engine(n, "complex64")
LOAD, STORE, INIT, PHASE, ROUND, WT, BT, RED = 1, 2, 3, 4, 5, 6, 7, 8
GB = 2 * 8 * 2.0**n / 1e9
progs = {
    "load+store A->A": [(LOAD, 0, 0), (STORE, 0, 0)],
    "load WT BT store": [(LOAD, 0, 0), (WT, 0, 0), (BT, 0, 0), (STORE, 1, 0)],
    "interior (10 q)": [(LOAD, 0, 0), (ROUND, 0, 0x1e), (WT, 0, 0), (ROUND, 0, 0x18), (BT, 0, 0), (ROUND, 0, 0x1e), (STORE, 1, 0)],
    "boundary-like (24 q, no phase)": [(LOAD, 0, 0), (ROUND, 0, 0x1e), (WT, 0, 0), (ROUND, 0, 0x18), (BT, 0, 0), (ROUND, 0, 0x1e),
                                        (ROUND, 0, 0x1f), (WT, 0, 0), (ROUND, 0, 0x1f), (BT, 0, 0), (ROUND, 0, 0x1e), (STORE, 0, 0)],
    "init+store (write only)": [(INIT, 0, 0), (STORE, 0, 0)],
}
for pf in (0, 148):
    E.set_option("prefetch_dist", pf)
    for hb in (4, n - 10):
        for name, ops in progs.items():
            ms = E.debug_tile_program(hb, ops, reps=5)
            gb = GB / 2 if name.startswith("init") else GB
            print("prefetch=%3d hb=%2d %-32s %8.3f ms  %7.0f GB/s" % (pf, hb, name, ms, gb / ms * 1e3 / 1e0 / 1e0 * 1e-0), flush=True)
