"""GPU: the EXPERIMENTAL column kernels (csrc/column_kernel.cuh, engine option column_kernels, off by default) against the
tile kernels and the C oracle.  Kept in a file of its own that sorts after every other GPU test: the product path does not
depend on these kernels, and a failure here must not cut the run of the suite short under `pytest -x`."""
import numpy as np
import pytest

import parity_helpers as H

pytestmark = pytest.mark.gpu


def _engine(n, dtype):
    from qaoa_b200 import engine as eng
    return eng, eng.Engine(n, dtype=dtype)


@pytest.mark.parametrize("n,symmetric", [(26, 0), (28, 1), (28, 0), (30, 1)])
def test_column_kernels_match_the_tile_kernels(n, symmetric):
    """engine option column_kernels (experimental, off by default): BOUND2 / FINAL2 on the warp-decoupled TMA / mbarrier
    kernels of csrc/column_kernel.cuh -- same plans, coefficients, tables and diagonal streams as the tile kernels;
    energies against them (repeatedly: the kernels are asynchronous-proxy code) and against the C oracle"""
    from oracle import c_oracle

    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_option("symmetric", symmetric)
    E.set_cost_maxkcut(n, edges, 1, lut)
    d8 = c_oracle.maxcut_diag_u8(n, edges)
    rng = np.random.default_rng(n)
    for p in (1, 2, 3, 5):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        E.set_option("column_kernels", 0)
        old = E.energy(ang)
        E.set_option("column_kernels", 1)
        ref = c_oracle.run(n, d8, ang, "complex128", blocked=True)
        for _ in range(4):
            new = E.energy(ang)
            assert H.rel_err(new[0], ref[0]) < 1e-4 and H.rel_err(new[1], ref[1]) < 2e-4, (p, new, ref)
            assert new[2] == ref[2] and new[3] == ref[3] and abs(new[4] - 1) < 1e-4
            assert H.rel_err(new[0], old[0]) < 1e-9 and H.rel_err(new[1], old[1]) < 1e-9, (p, new, old)
    E.close()
