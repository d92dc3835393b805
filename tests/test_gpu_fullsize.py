"""GPU parity at the FULL BASELINE sizes against the C oracle (oracle/c/qaoa_ref.c, pinned to the numpy oracle in
tests/test_oracle.py): config 3 (MaxCut n=32, X, Plus, p=5, complex64) and config 4 (n=30, p=3: Portfolio + Grover(Dicke)
+ Dicke, QUBO + multi-angle X + Plus, QUBO + XY ring + Dicke).  Tolerance: 1e-4 relative in complex64 (north_star).

The oracle needs host memory (n=32: 32 GiB state + 4 GiB byte diagonal; n=30: 8 GiB state + 8 GiB f64 diagonal) and a
minute or two of host CPU per case; a case is skipped -- loudly -- when the box does not have the memory.
"""
import numpy as np
import pytest

import parity_helpers as H

pytestmark = pytest.mark.gpu


def _need_host_gib(gib):
    from oracle import c_oracle
    have = c_oracle.host_mem_available_gib()
    if have < gib:
        pytest.skip("host has %.0f GiB available, the oracle run needs %.0f GiB" % (have, gib))


def _oracle():
    from oracle import c_oracle
    c_oracle.build()
    c_oracle.use_all_cores()
    return c_oracle


def _close(got, ref, tol=1e-4):
    assert H.rel_err(got[0], ref[0]) < tol, (got, ref)
    assert H.rel_err(got[1], ref[1]) < 2 * tol, (got, ref)
    assert abs(got[4] - 1.0) < 1e-4, got


def test_config3_n32_p5_complex64_against_the_c_oracle():
    """BASELINE config 3 at full size: symmetric half register AND full register vs the oracle (one oracle run)."""
    _need_host_gib(44)
    from qaoa_b200 import engine as eng

    co = _oracle()
    n, p = 32, 5
    edges, colors, table = H.regular_maxcut(n)
    ang = np.random.default_rng(32).uniform(0, 2 * np.pi, size=2 * p)
    d8 = co.maxcut_diag_u8(n, edges)
    ref = co.run(n, d8, ang, "complex64", blocked=True)
    got = {}
    for sym in (1, 0):
        E = eng.Engine(n, "complex64")
        E.set_option("symmetric", sym)
        E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
        # bit-exact diagonal on a slice at the very top of the register
        off = (1 << n) - 4096
        assert np.array_equal(E.read_cost_diagonal(off, 4096), d8[off:].astype(np.float64))
        got[sym] = E.energy(ang)
        assert (E.counter("symmetric_plans") >= 1) == bool(sym)
        E.close()
        _close(got[sym], ref)
        assert got[sym][2] == ref[2] and got[sym][3] == ref[3]
    assert H.rel_err(got[1][0], got[0][0]) < 1e-6


def _config4_qubo(n):
    rs = np.random.RandomState(42)
    return rs.rand(n, n), 3 * (rs.rand(n) - 0.5)


def _gamma_scaled(ang, per, scale):
    ang = ang.copy()
    ang[0::per] *= scale      # costs are O(100): keep gamma * cost inside a few turns
    return ang


def test_config4_n30_qubo_multiangle_x_against_the_c_oracle():
    _need_host_gib(24)
    from qaoa_b200 import engine as eng

    co = _oracle()
    n, p = 30, 3
    Q, c = _config4_qubo(n)
    ang = _gamma_scaled(np.random.default_rng(30).uniform(0, 2 * np.pi, size=p * (1 + n)), 1 + n, 0.02)
    diag = co.qubo_diag(Q, c, 0.0, fast=True)
    ref = co.run(n, diag, ang, "complex64", mixer="xmulti", blocked=True)
    E = eng.Engine(n, "complex64")
    E.set_cost_qubo(Q, c, 0.0)
    off = (1 << n) - 1024
    dev = E.read_cost_diagonal(off, 1024)
    assert np.allclose(dev, diag[off:], rtol=0, atol=3e-9 * np.abs(diag).max())   # 32-bit fixed point diagonal
    E.set_mixer(eng.MIXER_XMULTI)
    got = E.energy(ang)
    E.close()
    _close(got, ref)


def _portfolio(n):
    rng = np.random.RandomState(0)
    A = rng.randn(n, 101)
    cov, mu = np.cov(A), A.mean(axis=1)
    k = n // 3
    ones = np.ones_like(cov)
    penalty, risk = 0.0, 0.5
    Q = risk * cov + penalty * ones
    c = -mu - 2 * penalty * k * np.ones_like(mu)
    return Q, c, penalty * k * k, k


@pytest.mark.parametrize("feasible_subspace", [1, 0])
def test_config4_n30_portfolio_grover_dicke_against_the_c_oracle(feasible_subspace):
    """Dicke(k) + Grover(Dicke(k)): the compact C(n,k) register and the full 2^n register vs the oracle's full register"""
    _need_host_gib(24)
    from qaoa_b200 import engine as eng

    co = _oracle()
    n, p = 30, 3
    Q, c, b, k = _portfolio(n)
    ang = _gamma_scaled(np.random.default_rng(31).uniform(0, 2 * np.pi, size=2 * p), 2, 3.0)
    diag = co.qubo_diag(Q, c, b, fast=True)
    ref = co.run(n, diag, ang, "complex64", mixer="grover", init=("dicke", k), sub=("dicke", k))
    E = eng.Engine(n, "complex64")
    E.set_option("feasible_subspace", feasible_subspace)
    E.set_cost_qubo(Q, c, b)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
    got = E.energy(ang)
    E.close()
    _close(got, ref)
    # extremes over the support = extremes over the weight-k strings
    assert abs(got[2] - ref[2]) <= 1e-6 * abs(ref[2]) and abs(got[3] - ref[3]) <= 1e-6 * abs(ref[3])


def test_config4_n30_qubo_xy_ring_dicke_against_the_c_oracle():
    _need_host_gib(24)
    from qaoa_b200 import engine as eng

    co = _oracle()
    n, p = 30, 3
    Q, c = _config4_qubo(n)
    k = n // 3
    ang = _gamma_scaled(np.random.default_rng(33).uniform(0, 2 * np.pi, size=2 * p), 2, 0.02)
    pairs = [[i, i + 1] for i in range(n - 1)] + [[n - 1, 0]]
    diag = co.qubo_diag(Q, c, 0.0, fast=True)
    ref = co.run(n, diag, ang, "complex64", mixer="xy", init=("dicke", k), pairs=pairs)
    E = eng.Engine(n, "complex64")
    E.set_cost_qubo(Q, c, 0.0)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_XY, pairs=pairs)
    got = E.energy(ang)
    E.close()
    _close(got, ref)
