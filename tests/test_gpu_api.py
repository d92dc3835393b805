"""GPU: the reference-facing Python API (QAOA class) on the real engine vs the same driver on the
oracle stand-in: landscapes, optimizer trajectories, Grover/Dicke and multi-angle pipelines."""
import networkx as nx
import numpy as np
import pytest

from qaoa_b200 import QAOA, initialstates, mixers, problems
from qaoa_b200.utils import COBYLA
from test_host_api import make

pytestmark = pytest.mark.gpu


def both(problem_factory, mixer_factory, init_factory, dtype, **kw):
    dev = QAOA(problem=problem_factory(), mixer=mixer_factory(), initialstate=init_factory(), dtype=dtype, **kw)
    ref, _ = make(problem_factory(), mixer=mixer_factory(), init=init_factory(), **kw)
    return dev, ref


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-9), ("complex64", 1e-4)])
def test_readme_example_landscape_and_trajectory(dtype, tol):
    """BASELINE config 1: README quick example (MaxCut 3-regular n=8 seed 42, X, Plus): landscape then
    optimize to depth 3; the optimizer trajectory must follow the oracle's to the stated tolerance."""
    G = nx.random_regular_graph(3, 8, seed=42)
    opt = [COBYLA, {"maxiter": 80, "tol": 1e-5}]
    dev, ref = both(lambda: problems.MaxCut(G), mixers.X, initialstates.Plus, dtype, optimizer=opt)
    dev.sample_cost_landscape()
    ref.sample_cost_landscape()
    assert dev.Exp_sampled_p1.shape == (20, 20)
    assert np.allclose(dev.Exp_sampled_p1, ref.Exp_sampled_p1, rtol=tol, atol=tol)
    assert np.allclose(dev.Var_sampled_p1, ref.Var_sampled_p1, rtol=10 * tol, atol=10 * tol)
    assert np.array_equal(dev.MaxCost_sampled_p1, ref.MaxCost_sampled_p1)
    dev.optimize(depth=3)
    ref.optimize(depth=3)
    # Trajectory parity: every iterate the device-driven optimiser visited must carry the oracle's
    # energy at THOSE angles.  (Comparing the two runs iterate by iterate is ill-posed: the seed is
    # a symmetric grid point, COBYLA's first comparisons are exact ties and 1e-16 noise picks a branch.)
    fake = ref._engine
    for d in (1, 2, 3):
        a = dev.optimization_results[d]
        assert a.num_fval() >= 3
        oracle_vals = -fake.energy_batch(np.asarray(a.angles))[:, 0]
        assert np.allclose(a.Exp, oracle_vals, rtol=tol, atol=tol), d
        assert dev.get_Exp(d) <= dev.Exp_sampled_p1.min() + 10 * tol
    assert dev.get_Exp(3) < dev.get_Exp(1) - 0.3
    assert dev.get_Exp(3) == pytest.approx(ref.get_Exp(3), abs=0.15)   # same basin quality
    ok, rep = dev.validate_circuit()
    assert ok, rep


def test_portfolio_grover_dicke_and_multiangle_qubo():
    """BASELINE config 4 at test size: Portfolio + Grover(Dicke) + Dicke, and QUBO + XMultiAngle + Plus."""
    rng = np.random.RandomState(0)
    A = rng.randn(10, 60)
    cov, mu = np.cov(A), A.mean(axis=1)
    mk = lambda: problems.PortfolioOptimization(risk=0.5, budget=4, cov_matrix=cov, exp_return=mu, penalty=0)
    dev, ref = both(mk, lambda: mixers.Grover(initialstates.Dicke(4)), lambda: initialstates.Dicke(4), "complex128",
                    optimizer=[COBYLA, {"maxiter": 30}])
    grid = {"gamma": [0, 2 * np.pi, 6], "beta": [0, 2 * np.pi, 6]}
    dev.optimize(depth=2, angles=grid)
    ref.optimize(depth=2, angles=grid)
    assert np.allclose(dev.Exp_sampled_p1, ref.Exp_sampled_p1, rtol=1e-9, atol=1e-12)
    for d in (1, 2):
        a = dev.optimization_results[d]
        assert np.allclose(a.Exp, -ref._engine.energy_batch(np.asarray(a.angles))[:, 0], rtol=1e-9, atol=1e-12)

    rs = np.random.RandomState(42)
    Q, c = rs.rand(9, 9), 3 * (rs.rand(9) - 0.5)
    dev, ref = both(lambda: problems.QUBO(Q=Q, c=c, b=0.0), mixers.XMultiAngle, initialstates.Plus, "complex128",
                    optimizer=[COBYLA, {"maxiter": 25}])
    grid = {"gamma": [0, 1.0, 5], "beta": [0, np.pi, 5]}
    dev.optimize(depth=2, angles=grid)
    ref.optimize(depth=2, angles=grid)
    assert len(dev.get_angles(2)) == 20
    for d in (1, 2):
        a = dev.optimization_results[d]
        assert np.allclose(a.Exp, -ref._engine.energy_batch(np.asarray(a.angles))[:, 0], rtol=1e-9, atol=1e-12)


def test_large_landscape_batch_n20():
    """Batched landscape on the tiled path (one launch per pass for all points)."""
    G = nx.random_regular_graph(3, 20, seed=42)
    dev = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(), dtype="complex64")
    dev.sample_cost_landscape({"gamma": [0, 2 * np.pi, 6], "beta": [0, 2 * np.pi, 5]})
    from oracle import qaoa_oracle as orc
    edges = orc.graph_edges(orc.relabel_graph(G))
    colors, table = orc.colour_table(2)
    for (b, g) in ((0, 0), (2, 3), (4, 5)):
        e = orc.maxcut_p1_closed_form(edges, 20, dev.gamma_grid[g], dev.beta_grid[b])
        assert dev.Exp_sampled_p1[b, g] == pytest.approx(-e, rel=1e-4)


def test_hist_and_cvar_on_the_device():
    """hist() / cvar < 1 through the public API on the real engine vs the oracle stand-in."""
    G = nx.random_regular_graph(3, 8, seed=42)
    dev, ref = both(lambda: problems.MaxCut(G), mixers.X, initialstates.Plus, "complex128", cvar=0.4)
    grid = {"gamma": [0, 2 * np.pi, 5], "beta": [0, 2 * np.pi, 4]}
    dev.sample_cost_landscape(grid)
    ref.sample_cost_landscape(grid)
    assert np.allclose(dev.Exp_sampled_p1, ref.Exp_sampled_p1, rtol=1e-9, atol=1e-12)
    dev2, _ = both(lambda: problems.MaxCut(G), mixers.X, initialstates.Plus, "complex128")
    h = dev2.hist(np.array([0.61, 0.39]), 2048)
    assert sum(h.values()) == 2048 and all(len(k) == 8 for k in h)
    best = max(h, key=h.get)
    assert dev2.problem.cost(best) >= 8          # README graph: max cut 10, <cost> ~ 8 at these angles


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-9), ("complex64", 1e-4)])
def test_per_term_gammas_free_and_orbit(dtype, tol):
    """MaxCutFree + XMultiAngle and MaxCutOrbit + XOrbit on the device vs the oracle stand-in (per-term phase
    separator of the single-CTA kernel), n = 5 (house) and n = 10 (Petersen: one edge orbit, one node orbit)."""
    house = nx.Graph([(0, 1), (1, 2), (2, 3), (3, 0), (2, 4), (3, 4)])
    rng = np.random.default_rng(4)
    for G, mk_p, mk_m in ((house, problems.MaxCutFree, lambda g: mixers.XMultiAngle()),
                          (house, problems.MaxCutOrbit, lambda g: mixers.XOrbit(g)),
                          (nx.petersen_graph(), problems.MaxCutOrbit, lambda g: mixers.XOrbit(g)),
                          (nx.random_regular_graph(3, 12, seed=5), problems.MaxCutFree, lambda g: mixers.X())):
        dev, ref = both(lambda: mk_p(G), lambda: mk_m(G), initialstates.Plus, dtype)
        dev.createParameterizedCircuit(2)
        ref.createParameterizedCircuit(2)
        per = dev.n_gamma + dev.n_beta
        assert per == ref.n_gamma + ref.n_beta
        rows = rng.uniform(-1.5, 1.5, size=(3, 2 * per))
        a = dev._evaluate(rows)
        b = ref._evaluate(rows)
        assert np.allclose(a[:, 0], b[:, 0], rtol=tol, atol=tol), (type(dev.problem).__name__, a[:, 0], b[:, 0])
        assert np.allclose(a[:, 1], b[:, 1], rtol=2 * tol, atol=2 * tol)


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-9), ("complex64", 1e-4)])
def test_per_term_gammas_and_plus_parameterized_beyond_one_cta(dtype, tol):
    """Registers that do not fit one CTA (n = 16, 18): per-term gammas and PlusParameterized run layer by layer on the
    tile passes (engine.cu run_layered_point: on-the-fly edge phases, mixer passes with gamma = 0) -- vs the oracle."""
    rng = np.random.default_rng(11)
    G16 = nx.random_regular_graph(3, 16, seed=1)
    G18 = nx.random_regular_graph(3, 18, seed=3)
    cases = ((lambda: problems.MaxCutFree(G16), mixers.X, initialstates.Plus, 3),
             (lambda: problems.MaxCutFree(G16), mixers.XMultiAngle, initialstates.Plus, 2),
             (lambda: problems.MaxCutOrbit(G18), lambda: mixers.XOrbit(G18), initialstates.Plus, 2),
             (lambda: problems.MaxCut(G16), mixers.X, lambda: initialstates.PlusParameterized(num_phases=9), 3),
             (lambda: problems.MaxCutFree(G16), mixers.X, lambda: initialstates.PlusParameterized(num_phases=16), 2))
    for mk_p, mk_m, mk_i, depth in cases:
        dev, ref = both(mk_p, mk_m, mk_i, dtype)
        dev.createParameterizedCircuit(depth)
        ref.createParameterizedCircuit(depth)
        per = dev.n_gamma + dev.n_beta
        rows = rng.uniform(-1.5, 1.5, size=(2, dev.n_init + depth * per))
        a = dev._evaluate(rows)
        b = ref._evaluate(rows)
        assert np.allclose(a[:, 0], b[:, 0], rtol=tol, atol=tol), (type(dev.problem).__name__, a[:, 0], b[:, 0])
        assert np.allclose(a[:, 1], b[:, 1], rtol=2 * tol, atol=2 * tol)
    # the sampler follows the layered path too (kept final state)
    h = dev.hist(rows[0], 256)
    assert sum(h.values()) == 256


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-9), ("complex64", 1e-4)])
def test_plus_parameterized_on_the_device(dtype, tol):
    """PlusParameterized: the leading init phases of the angle vector reach the single-CTA kernel."""
    G = nx.random_regular_graph(3, 10, seed=2)
    dev, ref = both(lambda: problems.MaxCut(G), mixers.X, lambda: initialstates.PlusParameterized(num_phases=7), dtype)
    dev.createParameterizedCircuit(2)
    ref.createParameterizedCircuit(2)
    rows = np.random.default_rng(9).uniform(-2, 2, size=(4, 7 + 4))
    a, b = dev._evaluate(rows), ref._evaluate(rows)
    assert np.allclose(a[:, 0], b[:, 0], rtol=tol, atol=tol) and np.allclose(a[:, 1], b[:, 1], rtol=2 * tol, atol=2 * tol)
    h = dev.hist(rows[0], 512)
    assert sum(h.values()) == 512


def test_get_optimal_solutions_on_the_device():
    """qaoa_extreme_solutions behind QAOA.get_optimal_solutions: the optimal cuts of the README graph."""
    G = nx.random_regular_graph(3, 8, seed=42)
    dev = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(), dtype="complex128",
               optimizer=[COBYLA, {"maxiter": 15}])
    dev.optimize(depth=1, angles={"gamma": [0, np.pi, 4], "beta": [0, np.pi, 4]})
    sols = list(dev.get_optimal_solutions())
    strings = [format(x, "08b")[::-1] for x in range(256)]
    best = max(dev.problem.cost(s) for s in strings)
    assert sorted(sols) == sorted(s for s in strings if dev.problem.cost(s) == best)
    # a Dicke + XY state only has weight-k strings in its support
    q = QAOA(problem=problems.QUBO(Q=np.random.RandomState(1).rand(10, 10)), mixer=mixers.XY(case="ring"),
             initialstate=initialstates.Dicke(3), dtype="complex128")
    eng = q._get_engine()
    idx, cost = eng.extreme_solutions([0.02, 0.7], want_max=True)
    assert len(idx) >= 1 and all(bin(int(x)).count("1") == 3 for x in idx)
    feas = [x for x in range(1 << 10) if bin(x).count("1") == 3]
    d = eng.read_cost_diagonal()
    assert cost == pytest.approx(max(d[x] for x in feas))
