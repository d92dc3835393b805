"""CPU: one-hot Max-k-Cut (MaxKCutOneHot) and the XYTensor mixer.  Cost tables are the reference's own known answers
(unittests/test_maxkcut_one_hot_problem.py:23-109); the factor 2 between the reference's phase-separator circuit and its
cost function (problems/maxkcut_one_hot_problem.py docstring) is checked by evaluating that circuit's diagonal."""
import networkx as nx
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
from qaoa_b200 import initialstates, mixers, problems
from test_host_api import make


def _graph(edges, n):
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_weighted_edges_from(edges)
    return G


BARBELL = _graph([(0, 1, 1.0)], 2)
THREE = _graph([(0, 1, 1.0), (1, 2, 1.0)], 3)


def test_labels_known_answers():
    prob = problems.MaxKCutOneHot(BARBELL, 2)
    for s, lab in {"0101": "00", "0110": "01", "1010": "11", "1001": "10"}.items():
        assert prob.binstringToLabels(s) == lab
    prob = problems.MaxKCutOneHot(BARBELL, 3)
    for s, lab in {"001001": "00", "001010": "01", "001100": "02", "010001": "10", "010010": "11", "010100": "12",
                   "100001": "20", "100010": "21", "100100": "22"}.items():
        assert prob.binstringToLabels(s) == lab
    with pytest.raises(ValueError):
        prob.binstringToLabels("000100")


@pytest.mark.parametrize("G,k,table", [
    (BARBELL, 2, {"0101": 0, "1001": 1, "0110": 1, "1010": 0}),
    (THREE, 2, {"010101": 0, "100110": 2, "011010": 1, "101010": 0}),
    (THREE, 3, {"001010100": 2, "010100001": 2, "100100010": 1, "001001001": 0, "001100100": 1, "100100100": 0}),
    (THREE, 6, {"000100100000000010": 2, "100000100000100000": 0, "010000100000100000": 1, "000001100000010000": 2}),
])
def test_cost_known_answers_and_diagonal(G, k, table):
    prob = problems.MaxKCutOneHot(G, k)
    diag = prob.cost_diagonal()
    for s, expected in table.items():
        string = s[::-1]                      # the reference's tables are written most significant qubit first
        assert prob.cost(string) == expected
        assert diag[int(s, 2)] == expected
    n = prob.N_qubits
    rng = np.random.default_rng(0)
    for x in rng.integers(0, 1 << n, size=200):
        string = orc.index_to_string(int(x), n)
        if all(string[v * k:(v + 1) * k].count("1") >= 1 for v in range(prob.num_V)):
            assert diag[x] == prob.cost(string)
        else:
            assert diag[x] == 0.0


def test_reference_circuit_is_twice_gamma_on_the_onehot_subspace():
    """CX - RZ(w gamma) - CX per edge and colour (maxkcut_one_hot_problem.py:91-116) = exp(-i w gamma / 2 Z Z) each;
    on one-hot strings the product is exp(+i 2 gamma cost) up to a global phase."""
    G = _graph([(0, 1, 0.7), (1, 2, 1.3), (0, 2, 0.4)], 3)
    k, gamma = 3, 0.37
    prob = problems.MaxKCutOneHot(G, k)
    n = prob.N_qubits
    feas = [x for x in range(1 << n) if prob.isFeasible(orc.index_to_string(x, n))]
    assert len(feas) == k ** 3
    phase = np.ones(len(feas), dtype=complex)
    for (i, j) in G.edges():
        theta = G[i][j]["weight"] * gamma
        for c in range(k):
            a, b = k * i + c, k * j + c
            zz = np.array([1.0 if ((x >> a) & 1) == ((x >> b) & 1) else -1.0 for x in feas])
            phase *= np.exp(-0.5j * theta * zz)
    cost = prob.cost_diagonal()[feas]
    want = np.exp(1j * np.asarray(prob.engine_gammas([gamma]))[0] * cost)
    ratio = phase / want
    assert np.allclose(ratio, ratio[0]) and abs(abs(ratio[0]) - 1) < 1e-12
    assert prob.engine_gammas([gamma]) == [2 * gamma]


@pytest.mark.parametrize("mixer_kind", ["xytensor", "grover"])
def test_onehot_pipeline_stays_feasible_and_uses_two_gamma(mixer_kind):
    G = _graph([(0, 1, 1.0), (1, 2, 1.0), (0, 2, 1.0)], 3)
    k = 3
    prob = problems.MaxKCutOneHot(G, k)
    mixer = mixers.XYTensor(k) if mixer_kind == "xytensor" else mixers.MaxKCutGrover(k, "onehot", "max_balanced", False)
    q, fake = make(prob, mixer=mixer, init=initialstates.MaxKCutFeasible(k, "onehot"),
                   optimizer=[__import__("qaoa_b200").utils.COBYLA, {"maxiter": 20}])
    if mixer_kind == "xytensor":
        assert fake.mixer[0] == "xy" and fake.mixer[1] == [[3 * v + a, 3 * v + b] for v in range(3)
                                                           for a, b in ([0, 1], [1, 2], [2, 0])]
    row = np.array([0.41, 0.77])
    assert np.allclose(q._engine_row(row), [0.82, 0.77])
    e_dev = q._evaluate([row])[0, 0]
    spec = orc.Spec(9, prob.cost_diagonal(), mixer=fake.mixer, init=fake.init)
    assert e_dev == pytest.approx(orc.energy(spec, [0.82, 0.77])[0], rel=1e-12)
    pr = fake._prob(q._engine_row(row))
    for x in np.flatnonzero(pr > 1e-14):
        assert prob.isFeasible(orc.index_to_string(int(x), 9))
    q.optimize(depth=1, angles={"gamma": [0, np.pi, 6], "beta": [0, np.pi, 6]})
    assert q.get_Exp(1) <= -2.0 + 1e-9          # uniform colouring of a triangle cuts 2 edges on average
