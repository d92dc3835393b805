"""Localises errors of the sharded symmetric register: multi-angle X with a single non-zero beta."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import parity_helpers as H
from oracle import c_oracle
from qaoa_b200 import engine as eng
from qaoa_b200.sharded import LocalShardGroup

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
edges, colors, table = H.regular_maxcut(n)
lut, _ = H.lut_from_table(table, 1)
diag = c_oracle.maxcut_diag(n, edges)
grp = LocalShardGroup(n, world, dtype="complex64", symmetric=True)
grp.set_cost_maxkcut(n, edges, 1, lut)
grp.set_mixer(eng.MIXER_XMULTI)
for p in (1, 2):
    for layer in range(p):
        for q in list(range(n)) + [-1]:
            ang = np.zeros(p * (1 + n))
            ang[0::(1 + n)] = [0.7, 1.3][:p]
            if q >= 0:
                ang[layer * (1 + n) + 1 + q] = 0.9
            else:
                ang[layer * (1 + n) + 1: layer * (1 + n) + 1 + n] = 0.9
            ref = c_oracle.energy(n, diag, ang, "complex128", n_beta=n)
            got = grp.energy(ang)
            bad = abs(got[0] - ref[0]) > 1e-4 * abs(ref[0]) or abs(got[4] - 1) > 1e-4
            print("p=%d layer=%d q=%2d  E=%.6f ref=%.6f norm=%.6f %s" % (p, layer, q, got[0], ref[0], got[4], "BAD" if bad else ""))
