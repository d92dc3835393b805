import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import parity_helpers as H
from oracle import c_oracle
from qaoa_b200.sharded import LocalShardGroup
for (n, world, dtype, sym) in [(24, 2, "complex128", False), (24, 2, "complex64", False), (26, 2, "complex128", False)]:
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    grp = LocalShardGroup(n, world, dtype=dtype, symmetric=sym)
    grp.set_cost_maxkcut(n, edges, 1, lut)
    d8 = c_oracle.maxcut_diag_u8(n, edges)
    rng = np.random.default_rng(77)
    for p in (1, 2, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        grp.chunk_major = False
        whole = grp.energy(ang)
        grp.chunk_major = True
        got = grp.energy(ang)
        ref = c_oracle.run(n, d8, ang, "complex128", blocked=True)
        print(n, world, dtype, p, grp.last_chunks, "whole", whole[0], "chunk", got[0], "ref", ref[0], "norm", got[4], flush=True)
    grp.close()
