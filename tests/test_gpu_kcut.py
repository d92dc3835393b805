"""GPU: the k-cut "lego" components on the engine: MaxKCutBinaryFullH costs,
MaxKCutFeasible / LessThanK / Dicke1_2 / Tensor initial states, MaxKCutGrover in both forms (one Grover mixer over the
tensor state; tensor product of node-level Grover mixers = the engine's node Grover mixer) -- against the oracle."""
import networkx as nx
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
from qaoa_b200 import QAOA, initialstates, mixers, problems
import parity_helpers as H

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-4, "complex128": 1e-9}


def _setup(k, enc, num_v, dtype, seed=1):
    from qaoa_b200 import engine as eng

    G = nx.gnp_random_graph(num_v, 0.6, seed=seed)
    for i in range(num_v - 1):
        G.add_edge(i, i + 1)
    for (u, v) in G.edges():
        G[u][v]["weight"] = 1.0
    prob_enc = "max_balanced" if enc in ("Dicke1_2", "max_balanced") else "LessThanK"
    problem = problems.MaxKCutBinaryFullH(G, k, prob_enc)
    b, n = problem.N_qubits_per_node, problem.N_qubits
    colors, table = orc.colour_table(k, prob_enc)
    edges = orc.graph_edges(orc.relabel_graph(G))
    diag = orc.maxkcut_diag(edges, num_v, b, table, colors)
    node = orc.feasible_node_state(k, "Dicke1_2" if prob_enc == "max_balanced" else "LessThanK")
    E = eng.Engine(n, dtype=dtype)
    problem.lower_cost(E)
    assert np.array_equal(E.read_cost_diagonal(), diag)
    init = initialstates.MaxKCutFeasible(k, "binary", color_encoding="Dicke1_2" if prob_enc == "max_balanced" else "LessThanK")
    init.setNumQubits(n)
    init.lower_initial(E)
    return eng, E, problem, init, diag, node, n, b


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("k,enc,num_v", [(3, "LessThanK", 4), (5, "LessThanK", 3), (6, "Dicke1_2", 4), (7, "LessThanK", 4),
                                         (3, "LessThanK", 9), (6, "max_balanced", 6), (5, "LessThanK", 7)])
@pytest.mark.parametrize("tensorized", [True, False])
def test_maxkcut_grover_matches_oracle(k, enc, num_v, tensorized, dtype):
    """registers that fit one CTA (single fused kernel) and larger ones (element-wise passes, one sweep per node)"""
    eng, E, problem, init, diag, node, n, b = _setup(k, enc, num_v, dtype)
    mixer = mixers.MaxKCutGrover(k, problem_encoding="binary", color_encoding=enc, tensorized=tensorized)
    mixer.setNumQubits(n)
    mixer.lower_mixer(E)
    full = orc.tensor_state(node, num_v)
    spec = orc.Spec(n, diag, mixer=("nodegrover", node) if tensorized else ("grover", full), init=full)
    rng = np.random.default_rng(k * 100 + num_v)
    for p in (1, 3):
        ang = rng.uniform(-2.0, 2.0, size=2 * p)
        ref = orc.energy(spec, ang)
        got = E.energy(ang)
        tol = TOL[dtype]
        assert H.rel_err(got[0], ref[0]) < tol and H.rel_err(got[1], ref[1]) < 2 * tol, (p, got, ref)
        assert got[2] == ref[2] and got[3] == ref[3]
        assert abs(got[4] - 1) < (1e-4 if dtype == "complex64" else 1e-11)
    rows = rng.uniform(-2.0, 2.0, size=(5, 4))
    got = E.energy_batch(rows)
    for r, g in zip(rows, got):
        assert H.rel_err(g[0], orc.energy(spec, r)[0]) < TOL[dtype]


@pytest.mark.parametrize("k,enc,num_v,tensorized", [(3, "LessThanK", 4, True), (5, "LessThanK", 3, False),
                                                    (6, "Dicke1_2", 5, True), (7, "LessThanK", 5, True)])
def test_mixer_keeps_the_state_feasible(k, enc, num_v, tensorized):
    """reference unittests/test_maxkcut_mixers.py:24-62: initial state, then the mixer at beta = 0.5912847 -- no
    infeasible colour string may carry weight (here: the exact final state instead of 100000 shots)."""
    eng, E, problem, init, diag, node, n, b = _setup(k, enc, num_v, "complex128")
    mixer = mixers.MaxKCutGrover(k, problem_encoding="binary", color_encoding=enc, tensorized=tensorized)
    mixer.setNumQubits(n)
    mixer.lower_mixer(E)
    E.set_option("keep_state", 1)
    E.energy([0.37, 0.5912847])
    psi = E.read_statevector()
    full = orc.tensor_state(node, num_v)
    spec = orc.Spec(n, diag, mixer=("nodegrover", node) if tensorized else ("grover", full), init=full)
    ref = orc.evolve(spec, np.array([0.37, 0.5912847]))
    assert np.abs(psi - ref).max() < 1e-12
    bad = set(init.infeasible)
    for x in np.flatnonzero(np.abs(psi) > 1e-13):
        s = orc.index_to_string(int(x), n)
        assert not ({s[v * b:(v + 1) * b] for v in range(num_v)} & bad), s


def test_kcut_through_the_reference_facing_api():
    from qaoa_b200.utils import COBYLA
    from test_host_api import make

    G = nx.Graph([(0, 1), (1, 2), (0, 2), (2, 3), (3, 4), (4, 0)])
    mk = lambda: (problems.MaxKCutBinaryFullH(G, 3, "LessThanK"),
                  mixers.MaxKCutGrover(3, problem_encoding="binary", color_encoding="LessThanK", tensorized=True),
                  initialstates.MaxKCutFeasible(3, "binary"))
    pr, mx, st = mk()
    dev = QAOA(problem=pr, mixer=mx, initialstate=st, dtype="complex128", optimizer=[COBYLA, {"maxiter": 40}])
    pr, mx, st = mk()
    ref, fake = make(pr, mixer=mx, init=st, optimizer=[COBYLA, {"maxiter": 40}])
    grid = {"gamma": [0, 2 * np.pi, 8], "beta": [0, 2 * np.pi, 8]}
    dev.optimize(depth=2, angles=grid)
    ref.sample_cost_landscape(grid)
    assert np.allclose(dev.Exp_sampled_p1, ref.Exp_sampled_p1, rtol=1e-9, atol=1e-9)
    for d in (1, 2):
        a = dev.optimization_results[d]
        assert np.allclose(a.Exp, -fake.energy_batch(np.asarray(a.angles))[:, 0], rtol=1e-9, atol=1e-9)
    hist = dev.hist(dev.get_angles(2), shots=512)
    bad = set(st.infeasible)
    for s in hist:
        assert not ({s[v * 2:(v + 1) * 2] for v in range(5)} & bad), s


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("num_v,k,mixer_kind", [(3, 3, "xytensor"), (4, 3, "xytensor"), (5, 3, "xytensor"), (4, 4, "xytensor"),
                                                (3, 3, "grover"), (5, 3, "grover_tensorized"), (4, 2, "xytensor")])
def test_onehot_maxkcut_matches_oracle(num_v, k, mixer_kind, dtype):
    """MaxKCutOneHot + Dicke(1,k)-per-node initial state + XYTensor / MaxKCutGrover(onehot) through the driver: the
    device gets 2 gamma (the reference's one-hot phase-separator circuit), all three register sizes of the XY path."""
    from test_host_api import make

    G = nx.gnp_random_graph(num_v, 0.7, seed=num_v)
    for i in range(num_v - 1):
        G.add_edge(i, i + 1)
    rng = np.random.default_rng(7)
    for (u, v) in G.edges():
        G[u][v]["weight"] = float(rng.integers(1, 4))
    def parts():
        mixer = {"xytensor": lambda: mixers.XYTensor(k),
                 "grover": lambda: mixers.MaxKCutGrover(k, "onehot", "max_balanced", False),
                 "grover_tensorized": lambda: mixers.MaxKCutGrover(k, "onehot", "max_balanced", True)}[mixer_kind]()
        return problems.MaxKCutOneHot(G, k), mixer, initialstates.MaxKCutFeasible(k, "onehot")
    pr, mx, st = parts()
    dev = QAOA(problem=pr, mixer=mx, initialstate=st, dtype=dtype)
    pr, mx, st = parts()
    ref, fake = make(pr, mixer=mx, init=st)
    rows = rng.uniform(-1.5, 1.5, size=(4, 4))
    got = dev._evaluate(rows)
    want = ref._evaluate(rows)
    tol = TOL[dtype]
    assert np.allclose(got[:, 0], want[:, 0], rtol=tol, atol=tol), (got[:, 0], want[:, 0])
    assert np.allclose(got[:, 2:4], want[:, 2:4])
    assert np.array_equal(dev._get_engine().read_cost_diagonal(), pr.cost_diagonal())
