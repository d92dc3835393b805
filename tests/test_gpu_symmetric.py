"""GPU parity of the symmetric (Z2) register: MaxCut-like costs with cost(x) = cost(~x), X / multi-angle X mixer,
|+> and even n keep psi(x) = psi(~x), so the tile path stores the half register with the top qubit = 0 and pairs
every element with its mirror image for the top qubit's rotation (qaoa_b200/csrc/tile_kernel.cuh, "symmetric
registers").  Checked against the oracle AND against the same engine with the option switched off.

Tolerances as everywhere: 1e-9 relative in complex128, 1e-4 in complex64 (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
import parity_helpers as H

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-4, "complex128": 1e-9}


def _engine(n, dtype):
    from qaoa_b200 import engine as eng
    return eng, eng.Engine(n, dtype=dtype)


def _check(out, ref, dtype, what=""):
    e, e2, mn, mx = ref
    tol = TOL[dtype]
    assert H.rel_err(out[0], e) < tol, (what, out, ref)
    assert H.rel_err(out[1], e2) < 2 * tol, (what, out, ref)
    assert abs(out[2] - mn) <= 1e-6 * max(1, abs(mn)) and abs(out[3] - mx) <= 1e-6 * max(1, abs(mx)), (what, out, ref)
    assert abs(out[4] - 1.0) < (1e-4 if dtype == "complex64" else 1e-11), (what, out)


def _both(E, ang):
    """evaluates with the symmetric register on and off; returns (on, off) and checks which plan kind ran"""
    E.set_option("symmetric", 1)
    on = E.energy(ang)
    assert E.counter("symmetric_plans") >= 1, "the symmetric plan was not selected"
    E.set_option("symmetric", 0)
    off = E.energy(ang)
    assert E.counter("symmetric_plans") == 0
    E.set_option("symmetric", 1)
    return on, off


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [16, 18, 20, 22])
def test_symmetric_register_x_mixer(n, dtype):
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table))
    rng = np.random.default_rng(7 + n)
    for p in (1, 2, 3, 5):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ref = orc.energy(spec, ang)
        on, off = _both(E, ang)
        _check(on, ref, dtype, "symmetric n=%d p=%d" % (n, p))
        _check(off, ref, dtype, "full n=%d p=%d" % (n, p))
        assert H.rel_err(on[0], off[0]) < TOL[dtype] * 0.1
    # the interpreter kernel runs the same plan (plain loads, per-slot butterfly forms)
    ang = rng.uniform(-1.2, 1.2, size=6)
    ref = orc.energy(spec, ang)
    E.set_option("force_generic", 1)
    _check(E.energy(ang), ref, dtype, "interpreter n=%d" % n)
    E.set_option("force_generic", 0)
    _check(E.energy(ang), ref, dtype, "compiled n=%d" % n)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_symmetric_register_cot_form_and_special_angles(dtype):
    """beta = pi/2 exactly and close to it (cot-form butterflies, interpreter), beta = 0, gamma = 0"""
    n = 16
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table))
    for ang in ([0.7, np.pi / 2], [0.7, np.pi / 2 - 1e-9, 0.3, 0.4], [0.4, 0.0], [0.0, 0.3], [1.1, 1.5707, 0.2, -1.5709, 0.9, 0.3]):
        ang = np.array(ang)
        ref = orc.energy(spec, ang)
        on, off = _both(E, ang)
        _check(on, ref, dtype, "symmetric %s" % ang)
        _check(off, ref, dtype, "full %s" % ang)


def test_symmetric_register_beta_period_keeps_grids_on_the_compiled_kernels():
    """X mixer on a symmetric register: energies have period pi/2 in beta, the engine reduces beta to [-pi/4, pi/4].
    The reference's default landscape grids (endpoint=False, 4 | num) contain beta = pi/2 and 3 pi/2 exactly; those rows
    must neither need the cot form (interpreter kernel) nor differ from the oracle."""
    n = 18
    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table))
    betas = np.linspace(0, 2 * np.pi, 8, endpoint=False)
    rows = np.array([[0.8, b] for b in betas])
    E.set_option("time_passes", 1)
    E.counter("reset")
    got = E.energy_batch(rows)
    assert E.counter("symmetric_plans") >= 1
    assert E.counter("pass_count_0") == 0, "some points fell back to the interpreter kernel"
    for r, g in zip(rows, got):
        ref = orc.energy(spec, r)
        _check(g, ref, "complex64", "beta = %.4f" % r[1])
    # period pi/2: rows 0, 2, 4, 6 (beta = 0, pi/2, pi, 3 pi/2) agree, and so do rows 1, 3, 5, 7
    assert np.allclose(got[0::2, 0], got[0, 0], rtol=1e-5) and np.allclose(got[1::2, 0], got[1, 0], rtol=1e-5)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [16, 20])
def test_symmetric_register_multiangle_and_weights(n, dtype):
    """per-qubit betas (the top qubit's own beta drives the mirror pairing) on a weighted MaxCut whose real-valued
    costs use the U32FX (complex64) / F64 (complex128) diagonal"""
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    rng = np.random.default_rng(3 + n)
    edges = [(u, v, float(rng.uniform(0.2, 1.7))) for (u, v, w) in edges]
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    E.set_mixer(eng.MIXER_XMULTI)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table), mixer=("xmulti",))
    for p in (1, 3):
        ang = rng.uniform(-1.3, 1.3, size=p * (1 + n))
        ref = orc.energy(spec, ang)
        on, off = _both(E, ang)
        _check(on, ref, dtype, "symmetric xmulti n=%d p=%d" % (n, p))
        _check(off, ref, dtype, "full xmulti n=%d p=%d" % (n, p))


def test_symmetric_register_is_not_used_without_the_symmetry():
    """QUBO costs (no bit-flip symmetry), odd n, Dicke states: the full register is used, results as before"""
    eng, E = _engine(16, "complex128")
    Q, c, b = H.random_qubo(16)
    E.set_cost_qubo(Q, c, b)
    x = np.arange(1 << 16)
    bits = ((x[:, None] >> np.arange(16)) & 1).astype(float)
    diag = -(np.einsum("xi,ij,xj->x", bits, Q, bits) + bits @ c + b)
    ang = np.array([0.01, 0.4, 0.02, 0.3])
    out = E.energy(ang)
    assert E.counter("symmetric_plans") == 0
    _check(out, orc.energy(orc.Spec(16, diag), ang), "complex128", "qubo")

    eng, E = _engine(17, "complex128")
    edges, colors, table = H.regular_maxcut(17)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(17, edges, 1, lut)
    ang = np.array([0.3, 0.4, 0.2, 0.3])
    out = E.energy(ang)
    assert E.counter("symmetric_plans") == 0
    _check(out, orc.energy(orc.Spec(17, H.maxcut_diag(17, edges, colors, table)), ang), "complex128", "odd n")


@pytest.mark.parametrize("n,dtype,low_bits,ps", [(24, "complex64", 5, (1, 2, 3)), (24, "complex64", 0, (1, 2, 3)),
                                                 (26, "complex64", 0, (2,)), (24, "complex128", 0, (1, 3))])
def test_symmetric_register_large_schedules(n, dtype, low_bits, ps):
    """the schedules the benchmark sizes use: n=24 picks 4 low bits (FIRST + FINAL2L4 at p=1: the landscape
    configuration; classic INTERIOR / BOUNDARY passes beyond), low_bits=5 and n=26 give three tile shapes = the
    balanced schedule of n=32 (FIRST, MID on the mirrored low tile; BOUND2, FINAL2 on ordinary high tiles);
    complex128 runs the balanced schedule at n=24.  Oracle: the C twin (pinned to numpy in tests/test_oracle.py)."""
    from oracle import c_oracle
    eng, E = _engine(n, dtype)
    if low_bits:
        E.set_option("low_bits", low_bits)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = c_oracle.maxcut_diag(n, edges)
    rng = np.random.default_rng(11 + n)
    for p in ps:
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)
        ref = c_oracle.energy(n, diag, ang, "complex128")
        E.counter("reset")
        on, off = _both(E, ang)
        assert E.counter("prog_launches") >= 4
        _check(on, ref, dtype, "symmetric n=%d p=%d" % (n, p))
        _check(off, ref, dtype, "full n=%d p=%d" % (n, p))
        E.set_option("force_generic", 1)
        _check(E.energy(ang), ref, dtype, "symmetric, interpreter n=%d p=%d" % (n, p))
        E.set_option("force_generic", 0)
    assert E.counter("state_bytes") >= (1 << n) * (8 if dtype == "complex64" else 16)   # the full register ran as well


def test_symmetric_register_batched_points_and_kept_state():
    """several points per launch (landscape batches) on the half register, then a call that needs the whole final
    state (sampling switches to the full register and back)"""
    n, p = 18, 2
    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table))
    rng = np.random.default_rng(5)
    rows = rng.uniform(-1.2, 1.2, size=(7, 2 * p))
    E.set_option("batch_points", 3)
    got = E.energy_batch(rows)
    assert E.counter("symmetric_plans") >= 1
    half_bytes = E.counter("state_bytes")
    assert half_bytes == 3 * (1 << (n - 1)) * 8
    for i in range(7):
        _check(got[i], orc.energy(spec, rows[i]), "complex64", "point %d" % i)
    idx = E.sample(rows[0], 2000, seed=1)
    assert idx.max() < (1 << n) and E.counter("state_bytes") >= (1 << n) * 8
    pr = np.abs(orc.evolve(spec, rows[0])) ** 2
    assert pr[idx].mean() > 1.2 / (1 << n)   # samples concentrate where |psi|^2 is large
    again = E.energy_batch(rows)
    assert np.allclose(again[:, :2], got[:, :2], rtol=1e-6)


def test_symmetric_register_through_the_reference_api():
    """QAOA(problem=MaxCut, mixer=X, initialstate=Plus) picks the symmetric register on its own: landscape and
    optimiser results equal those of the full register"""
    import networkx as nx
    from qaoa_b200 import QAOA, initialstates, mixers, problems

    G = nx.random_regular_graph(3, 16, seed=42)
    res = []
    for sym in (1, 0):
        q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(), dtype="complex128")
        q._get_engine().set_option("symmetric", sym)
        q.sample_cost_landscape({"gamma": [0, np.pi, 6], "beta": [0, np.pi, 5]})
        q.optimize(depth=2)
        assert (q._get_engine().counter("symmetric_plans") >= 1) == bool(sym)
        res.append((q.Exp_sampled_p1.copy(), q.get_Exp(1), q.get_Exp(2)))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-9, atol=1e-11)
    assert abs(res[0][1] - res[1][1]) < 1e-6 * abs(res[1][1])
    assert abs(res[0][2] - res[1][2]) < 1e-6 * abs(res[1][2])
