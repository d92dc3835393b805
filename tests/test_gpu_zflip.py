"""GPU parity of X gates between layers -- the reference's flip=True (qaoa/qaoa.py:505-509, qaoa/utils/flip.py:76-88).

The device runs the circuit segment by segment: layers up to a flip with the state kept in HBM, the X gates as an index
permutation of that state (qaoa_continue_from_kept_state), the remaining layers from there.  Checked against the numpy
oracle, which applies psi'(x) = psi(x ^ mask) between the same layers.  Tolerances as everywhere: 1e-9 relative in
complex128, 1e-4 in complex64.
"""
import networkx as nx
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
import parity_helpers as H

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-4, "complex128": 1e-9}


def _oracle_segments(n, diag, mixer, init, angles, per, cuts):
    """cuts: [(layers done so far, mask)], increasing; returns the final state"""
    psi, start, x = init, 0, np.arange(1 << n)
    for layers, mask in cuts:
        psi = orc.evolve(orc.Spec(n, diag, mixer=mixer, init=psi), angles[start * per: layers * per])
        psi = psi[x ^ mask]
        start = layers
    return orc.evolve(orc.Spec(n, diag, mixer=mixer, init=psi), angles[start * per:])


def _device_segments(E, angles, per, cuts):
    E.set_option("keep_state", 1)
    start = 0
    for layers, mask in cuts:
        E.energy(angles[start * per: layers * per])
        E.continue_from_kept_state(mask)
        start = layers
    E.set_option("keep_state", 0)
    return E.energy(angles[start * per:])


def _maxcut(n):
    edges, colors, table = H.regular_maxcut(n + (n % 2))
    edges = [(u, v, w) for (u, v, w) in edges if u < n and v < n]
    x = np.arange(1 << n)
    diag = np.zeros(1 << n)
    for (u, v, w) in edges:
        diag += (((x >> u) ^ (x >> v)) & 1) * w
    return diag


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [8, 13, 16, 19])
def test_x_mixer_with_flips_between_layers(n, dtype):
    """single-CTA path (n = 8; 13 in complex64) and tile path (13 in complex128, 16, 19); flips after layers 1 and 3 of 4"""
    from qaoa_b200 import engine as eng

    E = eng.Engine(n, dtype=dtype)
    diag = _maxcut(n)
    E.set_cost_diagonal(diag)
    rng = np.random.default_rng(n)
    ang = rng.uniform(0, 2 * np.pi, size=8)
    ang[1::2] = rng.uniform(-1.3, 1.3, size=4)
    cuts = [(1, int(rng.integers(1, 1 << n))), (3, (1 << (n - 1)) | 1)]
    ref = orc.stats(orc.Spec(n, diag), _oracle_segments(n, diag, ("x",), orc.plus_state(n), ang, 2, cuts))
    out = _device_segments(E, ang, 2, cuts)
    assert H.rel_err(out[0], ref[0]) < TOL[dtype] and H.rel_err(out[1], ref[1]) < 2 * TOL[dtype], (out, ref)
    assert abs(out[4] - 1.0) < (1e-4 if dtype == "complex64" else 1e-11)
    # the X gates matter, and a zero mask is the plain circuit
    plain = orc.energy(orc.Spec(n, diag), ang)
    assert abs(plain[0] - ref[0]) > 1e-3
    E.set_initial(eng.INIT_PLUS)
    out0 = _device_segments(E, ang, 2, [(2, 0)])
    assert H.rel_err(out0[0], plain[0]) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_grover_dicke_and_multiangle_with_flips(dtype):
    from qaoa_b200 import engine as eng

    n, k = 15, 5
    Q, c, _ = H.random_qubo(n)
    diag = orc.qubo_diag(Q, c, 0.0)
    E = eng.Engine(n, dtype=dtype)
    E.set_cost_qubo(Q, c, 0.0)
    E.set_initial(eng.INIT_DICKE, k=k)
    E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
    s = orc.dicke_state(n, k)
    ang = np.array([0.31, 0.8, 0.12, 1.7, 0.2, 0.4])
    cuts = [(1, 0b101), (2, 1 << 14)]
    ref = orc.stats(orc.Spec(n, diag), _oracle_segments(n, diag, ("grover", s), s, ang, 2, cuts))
    out = _device_segments(E, ang, 2, cuts)
    tol = 20 * TOL[dtype]      # real-valued costs of range ~100: U32FX diagonal in complex64
    assert H.rel_err(out[0], ref[0]) < tol and H.rel_err(out[1], ref[1]) < 2 * tol, (out, ref)

    E = eng.Engine(n, dtype=dtype)
    E.set_cost_qubo(Q, c, 0.0)
    E.set_mixer(eng.MIXER_XMULTI)
    rng = np.random.default_rng(3)
    ang = np.concatenate([np.concatenate([[g], rng.uniform(-1.2, 1.2, size=n)]) for g in (0.05, 0.02, 0.03)])
    cuts = [(2, 0b110000000000011)]
    ref = orc.stats(orc.Spec(n, diag), _oracle_segments(n, diag, ("xmulti",), orc.plus_state(n), ang, n + 1, cuts))
    out = _device_segments(E, ang, n + 1, cuts)
    assert H.rel_err(out[0], ref[0]) < tol and H.rel_err(out[1], ref[1]) < 2 * tol, (out, ref)


def test_flip_through_the_reference_facing_api():
    """QAOA(flip=True).optimize: bit-flip patterns are learnt depth by depth, the deeper circuits carry the X gates"""
    from qaoa_b200 import QAOA, initialstates, mixers, problems
    from qaoa_b200.utils import COBYLA

    np.random.seed(3)
    G = nx.random_regular_graph(3, 8, seed=42)
    q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(), flip=True, shots=256,
             optimizer=[COBYLA, {"maxiter": 30}])
    q.optimize(depth=3)
    assert len(q.bitflips) == 3 and len(q.get_Exp()) == 3
    diag = q._get_engine().read_cost_diagonal()
    q.bitflips = [[0, 0, 1, 0, 0, 1, 1, 0], [1, 0, 0, 0, 0, 0, 0, 1]]
    masks = q._flip_masks()
    ang = np.array([0.4, 0.3, 0.7, 0.2, 0.5, 0.6])
    q.createParameterizedCircuit(3)
    psi = _oracle_segments(8, diag, ("x",), orc.plus_state(8), ang, 2, [(1, masks[0]), (2, masks[1])])
    assert q._eval_cost(ang) == pytest.approx(-orc.stats(orc.Spec(8, diag), psi)[0], rel=1e-9)
    assert q._eval_cost(ang[:2]) == pytest.approx(-orc.energy(orc.Spec(8, diag), ang[:2])[0], rel=1e-9)
    h = q.hist(ang, 2000)
    assert sum(h.values()) == 2000
    pr = np.abs(psi) ** 2
    emp = np.zeros(256)
    for sbits, cnt in h.items():
        emp[int(sbits[::-1], 2)] = cnt / 2000.0
    assert np.abs(emp - pr).max() < 0.05
