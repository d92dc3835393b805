"""CPU oracle for the QAOA energy-evaluation hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy (complex128) restatement of what the reference
(OpenQuantumComputing/QAOA) computes on the path
``QAOA.loss -> backend.run -> counts -> problem.cost -> Statistic`` in the
*exact-statevector limit* (shots -> infinity).  It is the checker for the CUDA
engine; nothing in ``qaoa_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it.

PARITY PINNING.  The arithmetic of the reference lives in un-vendored
third-party code (qiskit >=2.3.0, qiskit-aer >=0.17.0, ``setup.py:22-24``, no
lock file) that is not installable here, and the reference's own Aer-backed
tests assert no energies.  **Energies are therefore "parity unpinned" against
Aer itself.**  What IS pinned (tests/test_oracle.py):
  * cost functions against the reference's unit tables
    (unittests/test_problems.py:30-38,100-123) and against fixtures generated
    by importing the reference's loadable fragments (tests/golden/, script
    tests/golden/make_golden.py);
  * the phase sign ``diag(U_P) = exp(+i t cost)`` (qaoa/utils/validation.py:43);
  * Dicke amplitudes (unittests/test_maxkcut_feasible_initialstate.py:47-64);
  * Statistic semantics (unittests/test_statistic_utility.py:24-92);
  * the whole Dicke+Grover+QUBO pipeline against the 8192-shot histograms of
    examples/ExactCover/data/exact_cover_path_problem_example.json (chi^2);
  * the closed-form p=1 MaxCut energy (independent derivation);
  * the optimised energies the reference's real pipeline printed into its example notebooks (house graph
    depth 1-3, ToyExample depth 1, the three example GML graphs incl. fix_one_node: qiskit-aer, 1024 shots) against
    the oracle's exact optima within shot noise, and the notebooks' quoted minimum costs bit-exactly.

Conventions (SURVEY.md cheat-sheet): bit i of the statevector index x is the
measured value of qubit i and equals ``string[i]`` of the string handed to
``cost`` (qaoa/qaoa.py:605-609).
"""

from __future__ import annotations

import math
from itertools import combinations

import numpy as np

# --------------------------------------------------------------------------
# graph relabelling  (qaoa/utils/graphutils.py:49-72, 115-137)
# --------------------------------------------------------------------------


def relabel_graph(G):
    """Restates GraphHandler.__ensure_integer_labels__ + maxdegree-last swap.

    Returns a networkx graph with nodes 0..n-1 where the first node of maximum
    degree (in ``G.degree()`` iteration order, stable sort) has been swapped
    with node n-1.
    """
    import networkx as nx

    n = G.number_of_nodes()
    if set(G.nodes) != set(range(n)):
        mapping = {node: i for i, node in enumerate(G.nodes)}
        G = nx.relabel_nodes(G, mapping, copy=True)
    # sorted(..., reverse=True) is stable: first max-degree node in iteration order
    j = sorted(G.degree(), key=lambda kv: kv[1], reverse=True)[0][0]
    if j != n - 1:
        G = nx.relabel_nodes(G, {j: n - 1, n - 1: j}, copy=True)
    return G


def graph_edges(G):
    """(u, v, w) triples in ``G.edges()`` order, weight default 1
    (qaoa/problems/graph_problem.py:171-176)."""
    return [(int(u), int(v), G[u][v].get("weight", 1)) for u, v in G.edges()]


# --------------------------------------------------------------------------
# colour tables  (maxkcut_binary_powertwo.py:111-142, maxkcut_binary_fullH.py:119-192)
# --------------------------------------------------------------------------


def colour_table(k_cuts, color_encoding=None):
    """label string -> colour id, exactly the dictionaries of the reference."""
    if k_cuts == 2:
        colors = {"color1": ["0"], "color2": ["1"]}
    elif k_cuts == 4:
        colors = {"color1": ["00"], "color2": ["01"], "color3": ["10"], "color4": ["11"]}
    elif k_cuts == 8:
        colors = {
            "color%d" % (i + 1): [format(i, "03b")] for i in range(8)
        }
    elif k_cuts == 3:
        if color_encoding != "LessThanK":
            raise ValueError("invalid or unspecified color_encoding")
        colors = {"color1": ["00"], "color2": ["01"], "color3": ["10", "11"]}
    elif k_cuts == 5:
        if color_encoding == "LessThanK":
            colors = {
                "color1": ["000"], "color2": ["001"], "color3": ["010"], "color4": ["011"],
                "color5": ["100", "101", "110", "111"],
            }
        elif color_encoding == "max_balanced":
            colors = {
                "color1": ["000", "001"], "color2": ["010"], "color3": ["011"],
                "color4": ["100", "101"], "color5": ["110", "111"],
            }
        else:
            raise ValueError("invalid or unspecified color_encoding")
    elif k_cuts == 6:
        if color_encoding == "LessThanK":
            colors = {
                "color1": ["000"], "color2": ["001"], "color3": ["010"], "color4": ["011"],
                "color5": ["100"], "color6": ["101", "110", "111"],
            }
        elif color_encoding == "max_balanced":
            colors = {
                "color1": ["000", "001"], "color2": ["010"], "color3": ["011"],
                "color4": ["100", "101"], "color5": ["110"], "color6": ["111"],
            }
        else:
            raise ValueError("invalid or unspecified color_encoding")
    elif k_cuts == 7:
        if color_encoding != "LessThanK":
            raise ValueError("invalid or unspecified color_encoding")
        colors = {
            "color1": ["000"], "color2": ["001"], "color3": ["010"], "color4": ["011"],
            "color5": ["100"], "color6": ["101"], "color7": ["110", "111"],
        }
    else:
        raise ValueError("unsupported k_cuts")
    table = {}
    for key, labels in colors.items():
        for lab in labels:
            table[lab] = key
    return colors, table


# --------------------------------------------------------------------------
# cost functions on strings (slow, literal restatements)
# --------------------------------------------------------------------------


def maxkcut_cost_string(string, edges, num_V, bits_per_node, table, colors, fix_one_node=False):
    """GraphProblem.cost (graph_problem.py:152-176) with slice_string (:133-150)."""
    n_q = (num_V - int(fix_one_node)) * bits_per_node
    if len(string) != n_q:
        raise ValueError("wrong string length")
    k = bits_per_node
    labels = [string[v * k:(v + 1) * k] for v in range(num_V - int(fix_one_node))]
    if fix_one_node:
        labels.append(colors["color1"][0])
    return sum(w for (u, v, w) in edges if table.get(labels[u]) != table.get(labels[v]))


def qubo_cost_string(string, Q, c, b):
    """QUBO.qubo_cost (qubo_problem.py:101-108)."""
    x = np.array(list(map(int, string)))
    return -(x.T @ Q @ x + c.T @ x + b)


def portfolio_qubo(risk, budget, cov_matrix, exp_return, penalty=0):
    """PortfolioOptimization.__init__ QUBO mapping (portfolio_problem.py:48-54)."""
    Q = risk * cov_matrix + penalty * np.ones_like(cov_matrix)
    c = -exp_return - 2 * penalty * budget * np.ones_like(exp_return)
    b = penalty * budget * budget
    return Q, c, b


def portfolio_cost_string(string, risk, budget, cov_matrix, exp_return, penalty=0):
    """PortfolioOptimization.cost (portfolio_problem.py:56-71)."""
    x = np.array(list(map(int, string)))
    cost = risk * (x.T @ cov_matrix @ x) - exp_return.T @ x
    cost += penalty * (x.sum() - budget) ** 2
    return -cost


def exactcover_setup(columns, weights=None, penalty_factor=None, hamming_weight=None):
    """ExactCover.__init__ without scaling (exactcover_problem.py:31-87)."""
    col_size, num_col = columns.shape
    if weights is None:
        weights = np.zeros(num_col)
    if penalty_factor is None:
        if np.all(weights == 0):
            penalty_factor = 1
        elif hamming_weight:
            sw = np.sort(weights)
            penalty_factor = np.sum(sw[-hamming_weight:]) - np.sum(sw[:hamming_weight])
        else:
            penalty_factor = np.sum(weights[weights > 0])
    c = weights - 2 * penalty_factor * (np.ones(col_size) @ columns)
    Q = penalty_factor * (columns.T @ columns)
    b = penalty_factor * col_size
    return weights, penalty_factor, Q, c, b


def exactcover_cost_string(string, columns, weights, penalty_factor):
    """ExactCover.cost (exactcover_problem.py:89-100, 114-121)."""
    x = np.array(list(map(int, string)))
    c_e = np.sum((1 - (columns @ x)) ** 2)
    return -(weights @ x + penalty_factor * c_e)


# --------------------------------------------------------------------------
# vectorised cost diagonals (index x, bit i == string[i])
# --------------------------------------------------------------------------


def index_to_string(x, n):
    return "".join("1" if (x >> i) & 1 else "0" for i in range(n))


def diag_from_string_cost(cost_fn, n):
    """Brute-force diagonal through a string cost function (small n only)."""
    return np.array([cost_fn(index_to_string(x, n)) for x in range(1 << n)])


def maxkcut_diag(edges, num_V, bits_per_node, table, colors, fix_one_node=False):
    """Vectorised GraphProblem.cost over all 2^n indices.

    The label of node v is the string ``string[v*b:(v+1)*b]`` whose FIRST
    character is qubit v*b (the lowest), so the colour LUT is indexed by the
    string, not by the integer value of the bits.
    """
    b = bits_per_node
    n_free = num_V - int(fix_one_node)
    n = n_free * b
    N = 1 << n
    x = np.arange(N, dtype=np.int64)
    # integer colour ids
    names = sorted(set(table.values()), key=lambda s: int(s[5:]))
    cid = {name: i for i, name in enumerate(names)}
    lut = np.zeros(1 << b, dtype=np.int64)
    for val in range(1 << b):
        lab = "".join("1" if (val >> j) & 1 else "0" for j in range(b))
        lut[val] = cid[table[lab]]
    fixed_colour = cid[table[colors["color1"][0]]]

    def colour_of(v):
        if v >= n_free:
            return np.full(N, fixed_colour, dtype=np.int64)
        return lut[(x >> (v * b)) & ((1 << b) - 1)]

    integer = all(float(w).is_integer() for _, _, w in edges)
    diag = np.zeros(N, dtype=np.int64 if integer else np.float64)
    for (u, v, w) in edges:
        diag += (colour_of(u) != colour_of(v)) * (int(w) if integer else w)
    return diag


def qubo_diag(Q, c, b):
    """-(x^T Q x + c^T x + b) for all x (qubo_problem.py:107-108), chunked."""
    n = Q.shape[0]
    N = 1 << n
    out = np.empty(N)
    chunk = 1 << min(n, 16)
    bits = np.arange(n)
    for s in range(0, N, chunk):
        xs = ((np.arange(s, s + chunk)[:, None] >> bits) & 1).astype(np.float64)
        out[s:s + chunk] = -(np.einsum("ij,jk,ik->i", xs, Q, xs) + xs @ c + b)
    return out


# --------------------------------------------------------------------------
# initial states
# --------------------------------------------------------------------------


def plus_state(n):
    """plus_initialstate.py:19-25."""
    return np.full(1 << n, 2.0 ** (-n / 2), dtype=np.complex128)


def dicke_state(n, k):
    """|D^n_k>: uniform real positive superposition of all weight-k strings
    (dicke_initialstate.py:38-56 realises this; pinned for k=1 by
    unittests/test_maxkcut_feasible_initialstate.py:47-64)."""
    psi = np.zeros(1 << n, dtype=np.complex128)
    amp = 1.0 / math.sqrt(math.comb(n, k))
    for ones in combinations(range(n), k):
        x = 0
        for i in ones:
            x |= 1 << i
        psi[x] = amp
    return psi


def statevector_state(vec):
    """statevector_initialstate.py:27-33 (little-endian index)."""
    return np.asarray(vec, dtype=np.complex128).copy()


# --------------------------------------------------------------------------
# mixers
# --------------------------------------------------------------------------


def apply_rx(psi, q, beta):
    """RX(-2 beta) on qubit q = exp(+i beta X) (x_mixer.py:35)."""
    n = int(math.log2(psi.size))
    v = psi.reshape(1 << (n - q - 1), 2, 1 << q)
    c, s = math.cos(beta), math.sin(beta)
    a = v[:, 0, :].copy()
    b = v[:, 1, :].copy()
    v[:, 0, :] = c * a + 1j * s * b
    v[:, 1, :] = 1j * s * a + c * b
    return psi


def apply_x_mixer(psi, betas):
    """X (scalar beta) or XMultiAngle (beta per qubit, x_multiangle_mixer.py:40-50)."""
    n = int(math.log2(psi.size))
    betas = np.broadcast_to(np.asarray(betas, dtype=float), (n,))
    for q in range(n):
        apply_rx(psi, q, betas[q])
    return psi


def xy_pairs(n, case="ring"):
    """XY.generate_pairs (xy_mixer.py:58-79)."""
    assert case in ("ring", "chain")
    if n < 2:
        return []
    pairs = [[i, i + 1] for i in range(n - 1)]
    if case == "ring":
        pairs.append([n - 1, 0])
    return pairs


def apply_xy_gate(psi, q0, q1, beta):
    """XXPlusYYGate(theta=0.5*beta, 0) on (q0, q1) (xy_mixer.py:54-56).

    [3p] qiskit XXPlusYYGate(theta, 0): on span{|01>,|10>} it is
    [[cos(theta/2), -i sin(theta/2)], [-i sin(theta/2), cos(theta/2)]].
    """
    idx = np.arange(psi.size)
    m0, m1 = 1 << q0, 1 << q1
    sel = ((idx & m0) != 0) & ((idx & m1) == 0)  # q0=1,q1=0
    i10 = idx[sel]
    i01 = i10 ^ m0 ^ m1
    ang = 0.25 * beta
    c, s = math.cos(ang), math.sin(ang)
    a = psi[i10].copy()
    b = psi[i01].copy()
    psi[i10] = c * a - 1j * s * b
    psi[i01] = -1j * s * a + c * b
    return psi


def apply_xy_mixer(psi, pairs, beta):
    for (a, b) in pairs:  # order is semantics: gates do not commute
        apply_xy_gate(psi, a, b, beta)
    return psi


def apply_grover(psi, s_state, beta):
    """U_S^dag X^n C^{n-1}P(-beta) X^n U_S = I + (e^{-i beta} - 1)|s><s|
    (grover_mixer.py:43-82, phase gate :75-77)."""
    ov = np.vdot(s_state, psi)
    psi += (np.exp(-1j * beta) - 1.0) * ov * s_state
    return psi


def apply_node_grover(psi, s_node, beta):
    """MaxKCutGrover(tensorized=True) (maxkcut_grover_mixer.py:116-124): Tensor(Grover(node state), num_V) -- the
    Grover mixer of grover_mixer.py:43-82 on every b-qubit node register [v b, (v+1) b), one shared beta."""
    s_node = np.asarray(s_node, dtype=np.complex128)
    K = s_node.size
    b = K.bit_length() - 1
    n = psi.size.bit_length() - 1
    U = np.eye(K, dtype=np.complex128) + (np.exp(-1j * beta) - 1.0) * np.outer(s_node, s_node.conj())
    for v in range(n // b):
        t = psi.reshape(1 << (n - (v + 1) * b), K, 1 << (v * b))   # axis 1 = the node's bits
        psi[:] = np.einsum("jk,akc->ajc", U, t).reshape(-1)
    return psi


def feasible_node_state(k_cuts, color_encoding="LessThanK"):
    """Uniform superposition of the feasible colour strings of one node register in the binary encodings, i.e. what
    LessThanK (lessthank_initialstate.py:69-152) and Dicke1_2 (dicke1_2_initialstate.py:24-41) are documented to
    prepare; the feasible set is the complement of MaxKCutFeasible.infeasible (maxkcut_feasible_initialstate.py:53-68;
    character i of a string = qubit i of the register)."""
    b = int(np.ceil(np.log2(k_cuts)))
    if k_cuts == 3:
        bad = ["11"]
    elif k_cuts == 5:
        bad = ["100", "111", "101"] if color_encoding == "max_balanced" else ["101", "110", "111"]
    elif k_cuts == 6:
        bad = ["000", "111"] if color_encoding in ("Dicke1_2", "max_balanced") else ["110", "111"]
    elif k_cuts == 7:
        bad = ["111"]
    else:
        bad = []
    v = np.zeros(1 << b, dtype=np.complex128)
    for x in range(1 << b):
        if "".join("1" if (x >> i) & 1 else "0" for i in range(b)) not in bad:
            v[x] = 1.0
    return v / np.linalg.norm(v)


def tensor_state(node, num):
    """Tensor(sub, num) (tensor_initialstate.py:44-57): copy i on qubits [i m, (i+1) m)"""
    out = np.ones(1, dtype=np.complex128)
    for _ in range(num):
        out = np.kron(node, out)
    return out


# --------------------------------------------------------------------------
# the path:  init -> [phase(gamma_d) -> mixer(beta_d)/trotter x trotter] -> stats
# --------------------------------------------------------------------------


class Spec:
    """Bundle of (cost diagonal, mixer, initial state) — the oracle's 'plan'.

    mixer: ("x",) | ("xmulti",) | ("xy", pairs) | ("grover", s_state) | ("nodegrover", node_state)
    n_init is always 0 here (Plus / Dicke / StateVector carry no parameters).
    """

    def __init__(self, n, diag, mixer=("x",), init=None, trotter=1, terms=None):
        """terms: optional list of K diagonals with sum(terms) == diag; the phase separator then takes one
        gamma per term, exp(i sum_k gamma_k terms_k) (MaxCutFree: one term per edge, maxcut_free_problem.py:31-99;
        MaxCutOrbit: one per edge orbit, maxcut_orbit_problem.py:51-181)."""
        self.n = n
        self.terms = None if terms is None else [np.asarray(t, dtype=np.float64) for t in terms]
        self.diag = np.asarray(diag, dtype=np.float64)
        self.mixer = mixer
        self.init = plus_state(n) if init is None else np.asarray(init, dtype=np.complex128)
        self.trotter = trotter

    @property
    def n_gamma(self):
        return 1 if self.terms is None else len(self.terms)

    @property
    def n_beta(self):
        return self.n if self.mixer[0] == "xmulti" else 1


def evolve(spec, angles, dtype=np.complex128):
    """State after the circuit of qaoa/qaoa.py:418-520 for the flat angle
    vector of qaoa/qaoa.py:937-990."""
    per = spec.n_gamma + spec.n_beta
    angles = np.asarray(angles, dtype=np.float64)
    assert angles.size % per == 0
    p = angles.size // per
    psi = spec.init.astype(np.complex128).copy()
    for d in range(p):
        gammas = angles[d * per:d * per + spec.n_gamma]
        betas = angles[d * per + spec.n_gamma:(d + 1) * per]
        if spec.terms is None:
            psi *= np.exp(1j * gammas[0] * spec.diag)  # validation.py:43 sign
        else:
            psi *= np.exp(1j * sum(g * t for g, t in zip(gammas, spec.terms)))
        for _ in range(spec.trotter):
            bt = betas / spec.trotter  # qaoa.py:494-503
            kind = spec.mixer[0]
            if kind == "x":
                apply_x_mixer(psi, bt[0])
            elif kind == "xmulti":
                apply_x_mixer(psi, bt)
            elif kind == "xy":
                apply_xy_mixer(psi, spec.mixer[1], bt[0])
            elif kind == "grover":
                apply_grover(psi, spec.mixer[1], bt[0])
            elif kind == "nodegrover":
                apply_node_grover(psi, spec.mixer[1], bt[0])
            else:
                raise ValueError(kind)
        if dtype == np.complex64:
            psi = psi.astype(np.complex64).astype(np.complex128)
    return psi


def stats(spec, psi):
    """Exact-limit analogue of Statistic (statistic.py:57-142):
    returns (E[cost], E[cost^2], min cost over support, max cost over support)."""
    prob = np.abs(psi) ** 2
    c = spec.diag
    sup = prob > 0
    return (float(prob @ c), float(prob @ (c * c)), float(c[sup].min()), float(c[sup].max()))


def energy(spec, angles):
    """out[4] = {E[cost], E[cost^2], min, max}; Exp = -out[0] (qaoa.py:87, 935)."""
    return stats(spec, evolve(spec, angles))


def landscape(spec, gamma_spec, beta_spec):
    """sample_cost_landscape (qaoa.py:559-562, 616-628): arrays shaped
    (n_beta, n_gamma): Exp=-E, Var (population), MaxCost=-max, MinCost=-min."""
    gam = np.linspace(gamma_spec[0], gamma_spec[1], gamma_spec[2], endpoint=False)
    bet = np.linspace(beta_spec[0], beta_spec[1], beta_spec[2], endpoint=False)
    Exp = np.zeros((len(bet), len(gam)))
    Var = np.zeros_like(Exp)
    Mx = np.zeros_like(Exp)
    Mn = np.zeros_like(Exp)
    for bi, b in enumerate(bet):
        for gi, g in enumerate(gam):
            ang = [g] * spec.n_gamma + [b] * spec.n_beta
            e, e2, mn, mx = energy(spec, ang)
            Exp[bi, gi] = -e
            Var[bi, gi] = e2 - e * e
            Mx[bi, gi] = -mx
            Mn[bi, gi] = -mn
    return gam, bet, Exp, Var, Mx, Mn


# --------------------------------------------------------------------------
# independent cross-check: closed-form p=1 MaxCut energy (unit weights)
# --------------------------------------------------------------------------


def maxcut_p1_closed_form(edges, num_V, gamma, beta):
    """<cost> at p=1 for unweighted MaxCut with the reference's sign conventions
    (SURVEY.md §8c K1)."""
    nbrs = {v: set() for v in range(num_V)}
    for (u, v, _w) in edges:
        nbrs[u].add(v)
        nbrs[v].add(u)
    tot = 0.0
    for (u, v, _w) in edges:
        d = len(nbrs[u]) - 1
        e = len(nbrs[v]) - 1
        f = len(nbrs[u] & nbrs[v])
        tot += 0.5 + 0.25 * math.sin(4 * beta) * math.sin(gamma) * (
            math.cos(gamma) ** d + math.cos(gamma) ** e
        ) - 0.25 * math.sin(2 * beta) ** 2 * math.cos(gamma) ** (d + e - 2 * f) * (
            1 - math.cos(2 * gamma) ** f
        )
    return tot


# --------------------------------------------------------------------------
# Statistic restatement (statistic.py:44-142) — used to pin semantics
# --------------------------------------------------------------------------


class StatisticOracle:
    def __init__(self, cvar=1):
        self.cvar = cvar
        self.reset()

    def reset(self):
        self.W = 0
        self.maxval = float("-inf")
        self.minval = float("inf")
        self.E = 0
        self.S = 0
        self.all_values = np.array([])

    def add_sample(self, value, weight, string=None):
        self.W += weight
        tmp = self.E
        self.maxval = max(value, self.maxval)
        self.minval = min(value, self.minval)
        self.E += weight / self.W * (value - self.E)
        self.S += weight * (value - tmp) * (value - self.E)
        if self.cvar < 1:
            idx = np.searchsorted(self.all_values, value)
            self.all_values = np.insert(self.all_values, idx, np.ones(int(weight)) * value)

    def get_Variance(self):
        return self.S / (self.W - 1)

    def get_CVaR(self):
        if self.cvar < 1:
            K = int(np.round(self.cvar * len(self.all_values)))
            return np.sum(self.all_values[-K:]) / K
        return self.E


def exact_cvar(diag, prob, alpha):
    """Shots->infinity limit of Statistic.get_CVaR (statistic.py:132-142):
    mean of the best (largest-cost) alpha-fraction of the distribution."""
    order = np.argsort(-diag, kind="stable")
    c = diag[order]
    p = prob[order]
    cum = np.cumsum(p)
    take = np.clip(alpha - (cum - p), 0.0, p)
    return float(take @ c) / alpha


def edge_terms(edges, num_V, term_of_edge, n_terms, fix_one_node=False):
    """Per-term MaxCut diagonals: terms[k](x) = sum over edges e with term_of_edge[e] == k of w_e [x_u != x_v]
    (the fixed node, if any, is node num_V-1 with bit 0).  maxcut_free_problem.py:66-73 / maxcut_orbit_problem.py:139-144."""
    n = num_V - (1 if fix_one_node else 0)
    x = np.arange(1 << n, dtype=np.int64)
    terms = [np.zeros(1 << n) for _ in range(n_terms)]
    for (u, v, w), k in zip(edges, term_of_edge):
        bu = (x >> u) & 1 if u < n else 0
        bv = (x >> v) & 1 if v < n else 0
        terms[k] += w * (bu ^ bv)
    return terms


def edge_orbits(G):
    """Orbits of the edges of G (in G.edges() order) under Aut(G), ordered by their first edge
    (maxcut_orbit_problem.py:61-107).  Returns term_of_edge."""
    from networkx.algorithms.isomorphism import GraphMatcher
    edges = list(G.edges())
    idx = {}
    for i, (u, v) in enumerate(edges):
        idx[(u, v)] = i
        idx[(v, u)] = i
    parent = list(range(len(edges)))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for auto in GraphMatcher(G, G).isomorphisms_iter():
        for i, (u, v) in enumerate(edges):
            j = idx.get((auto[u], auto[v]))
            if j is not None:
                ri, rj = find(i), find(j)
                if ri != rj:
                    parent[ri] = rj
        if len({find(i) for i in range(len(edges))}) == 1:
            break
    order = {}
    for i in range(len(edges)):
        order.setdefault(find(i), len(order))
    return [order[find(i)] for i in range(len(edges))]


def phased_plus(psi0, phases):
    """RZ(phi_q) = diag(e^{-i phi/2}, e^{+i phi/2}) on qubit q for q < len(phases), applied to psi0
    (plus_parameterized_initialstate.py:55-64: H on every qubit, then rz(param_i, q[i]))."""
    phases = np.asarray(phases, dtype=np.float64)
    if phases.size == 0:
        return psi0
    x = np.arange(psi0.size, dtype=np.int64)
    theta = np.zeros(psi0.size)
    for q, phi in enumerate(phases):
        theta += phi * (((x >> q) & 1) - 0.5)
    return np.asarray(psi0, dtype=np.complex128) * np.exp(1j * theta)
