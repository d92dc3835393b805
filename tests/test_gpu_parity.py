"""GPU parity: CUDA engine (through the C ABI) vs the numpy oracle on identical inputs.

Tolerances (BASELINE.json north_star): cost diagonals bit-exact for integer weights;
energies 1e-9 relative in complex128, 1e-4 relative in complex64.
"""
import numpy as np
import pytest

from oracle import qaoa_oracle as orc
import parity_helpers as H

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-4, "complex128": 1e-9}


def _engine(n, dtype):
    from qaoa_b200 import engine as eng
    return eng, eng.Engine(n, dtype=dtype)


def _check(out, spec, angles, dtype, what="", ref=None):
    e, e2, mn, mx = orc.energy(spec, angles) if ref is None else ref
    tol = TOL[dtype]
    assert H.rel_err(out[0], e) < tol, (what, out, e)
    assert H.rel_err(out[1], e2) < 2 * tol, (what, out, e2)
    assert abs(out[2] - mn) <= 1e-6 * max(1, abs(mn)) and abs(out[3] - mx) <= 1e-6 * max(1, abs(mx)), (what, out, mn, mx)
    assert abs(out[4] - 1.0) < (1e-4 if dtype == "complex64" else 1e-11), (what, out)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [4, 8, 12])
def test_small_x_mixer_maxcut(n, dtype):
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = H.maxcut_diag(n, edges, colors, table)
    assert np.array_equal(E.read_cost_diagonal(), diag.astype(np.float64))  # bit-exact
    spec = orc.Spec(n, diag)
    rng = np.random.default_rng(1234)
    for p in (1, 2, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        _check(E.energy(ang), spec, ang, dtype, "p=%d" % p)


def test_readme_known_answers():
    """README graph (n=8, seed 42) energies."""
    eng, E = _engine(8, "complex128")
    edges, colors, table = H.regular_maxcut(8)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(8, edges, 1, lut)
    assert abs(-E.energy([0.3, 0.2])[0] - (-7.124615209734253)) < 1e-9
    assert abs(-E.energy([0.4, 0.6, 0.7, 0.35, 0.9, 0.15])[0] - (-8.995719311366326)) < 1e-9


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [14, 15, 18, 20])
def test_tile_x_mixer_maxcut(n, dtype):
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = H.maxcut_diag(n, edges, colors, table)
    assert np.array_equal(E.read_cost_diagonal(), diag.astype(np.float64))
    spec = orc.Spec(n, diag)
    rng = np.random.default_rng(99 + n)
    for p in (1, 2, 3, 5):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        _check(E.energy(ang), spec, ang, dtype, "n=%d p=%d" % (n, p))


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_tile_vs_small_same_n(dtype):
    """n where both code paths exist: force each and compare with the oracle."""
    n = 14 if dtype == "complex64" else 13
    edges, colors, table = H.regular_maxcut(n if n % 2 == 0 else n + 1)
    edges = [(u, v, w) for (u, v, w) in edges if u < n and v < n]
    diag = np.zeros(1 << n)
    x = np.arange(1 << n)
    for (u, v, w) in edges:
        diag += (((x >> u) ^ (x >> v)) & 1) * w
    spec = orc.Spec(n, diag)
    ang = np.array([0.37, 1.21, 2.2, 0.45, 1.1, 2.9])
    ref = orc.energy(spec, ang)
    outs = []
    for small_max in (n, n - 1):
        eng, E = _engine(n, dtype)
        E.set_option("small_max_qubits", small_max)
        E.set_cost_diagonal(diag)
        out = E.energy(ang)
        outs.append(out)
        assert H.rel_err(out[0], ref[0]) < TOL[dtype]
    assert H.rel_err(outs[0][0], outs[1][0]) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [10, 16])
def test_multiangle_x(n, dtype):
    eng, E = _engine(n, dtype)
    Q, c, b = H.random_qubo(n)
    E.set_cost_qubo(Q, c, b)
    diag = orc.qubo_diag(Q, c, b)
    got = E.read_cost_diagonal()
    scale = np.abs(diag).max()
    assert np.abs(got - diag).max() <= (1e-12 if dtype == "complex128" else 2e-9) * scale
    E.set_mixer(eng.MIXER_XMULTI)
    spec = orc.Spec(n, diag, mixer=("xmulti",))
    rng = np.random.default_rng(5)
    for p in (1, 3):
        ang = rng.uniform(0, 2 * np.pi, size=p * (1 + n))
        ang[::(1 + n)] *= 0.05  # gamma range sensible for costs O(100)
        _check(E.energy(ang), spec, ang, dtype, "xmulti p=%d" % p)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n,k", [(8, 3), (16, 5)])
def test_grover_dicke(n, k, dtype):
    eng, E = _engine(n, dtype)
    Q, c, b = H.random_qubo(n, seed=7)
    E.set_cost_qubo(Q, c, b)
    diag = orc.qubo_diag(Q, c, b)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
    s = orc.dicke_state(n, k)
    spec = orc.Spec(n, diag, mixer=("grover", s), init=s)
    rng = np.random.default_rng(11)
    for p in (1, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[::2] *= 0.05
        out = E.energy(ang)
        _check(out, spec, ang, dtype, "grover p=%d" % p)
        if n > 13:
            # Dicke + Grover(Dicke) runs in the C(n,k)-dimensional feasible subspace; the full register must agree
            from math import comb
            assert E.counter("compact_size") == comb(n, k)
            E.set_option("feasible_subspace", 0)
            full = E.energy(ang)
            E.set_option("feasible_subspace", 1)
            assert H.rel_err(out[0], full[0]) < TOL[dtype] * 0.1 and out[2] == full[2] and out[3] == full[3]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n,k,case", [(8, 3, "ring"), (15, 4, "chain"), (16, 5, "ring")])
def test_xy_dicke(n, k, case, dtype):
    eng, E = _engine(n, dtype)
    Q, c, b = H.random_qubo(n, seed=3)
    E.set_cost_qubo(Q, c, b)
    diag = orc.qubo_diag(Q, c, b)
    pairs = orc.xy_pairs(n, case)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_XY, pairs=pairs, trotter=2)
    spec = orc.Spec(n, diag, mixer=("xy", pairs), init=orc.dicke_state(n, k), trotter=2)
    rng = np.random.default_rng(21)
    for p in (1, 2):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[::2] *= 0.05
        out = E.energy(ang)
        _check(out, spec, ang, dtype, "xy p=%d" % p)
        if n > 13:
            # XY + Dicke runs in the C(n,k)-dimensional feasible subspace (xy_compact_gate_kernel); the tiled
            # full-register kernel must agree
            from math import comb
            assert E.counter("compact_size") == comb(n, k)
            E.set_option("feasible_subspace", 0)
            full = E.energy(ang)
            E.set_option("feasible_subspace", 1)
            assert H.rel_err(out[0], full[0]) < TOL[dtype] * 0.1 and H.rel_err(out[1], full[1]) < TOL[dtype]
            assert abs(out[2] - full[2]) <= 1e-9 * abs(full[2]) and abs(out[3] - full[3]) <= 1e-9 * abs(full[3])


@pytest.mark.parametrize("n,k,pairs", [(20, 7, "full"), (22, 3, "ring"), (34, 2, "chain")])
def test_xy_compact_register_arbitrary_pairs(n, k, pairs):
    """compact XY register: non-adjacent pairs (the rank of the partner changes through the set bits in between),
    long-range ring closure, and n > 32 (64-bit strings) -- against the oracle where numpy can hold it, else against
    exact invariants (norm, feasible extremes)"""
    eng, E = _engine(n, "complex64")
    if pairs == "full":
        pr = [[i, j] for i in range(0, n, 3) for j in range(i + 2, n, 5)]
    else:
        pr = orc.xy_pairs(n, pairs)
    rng = np.random.default_rng(n)
    ang = rng.uniform(0, 2 * np.pi, size=4)
    if n <= 22:
        Q, c, b = H.random_qubo(n, seed=3)
        E.set_cost_qubo(Q, c, b)
        ang[::2] *= 0.05
        E.set_initial(eng.INIT_DICKE, k)
        E.set_mixer(eng.MIXER_XY, pairs=pr)
        out = E.energy(ang)
        from math import comb
        assert E.counter("compact_size") == comb(n, k)
        spec = orc.Spec(n, orc.qubo_diag(Q, c, b), mixer=("xy", pr), init=orc.dicke_state(n, k))
        _check(out, spec, ang, "complex64", "compact xy %s" % pairs)
    else:
        edges = [(i, (i + 1) % n, 1.0) for i in range(n)]
        E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
        E.set_initial(eng.INIT_DICKE, k)
        E.set_mixer(eng.MIXER_XY, pairs=pr)
        out = E.energy(ang)
        assert abs(out[4] - 1.0) < 1e-5
        assert out[2] == 2.0 and out[3] == 4.0      # ring, two ones: adjacent (cut 2) or apart (cut 4)
    E.close()


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_multiangle_angles_near_half_pi_stay_on_the_compiled_programs(dtype):
    """Multi-angle X mixer with single betas at and next to pi/2 (|tan beta| up to 1e16): the layer's total growth fits the
    exponent budget, so every qubit keeps the lifted normal form and no pass falls back to the interpreter kernel
    (option layer_budget, default on); with the option off the same angles take the cot form on the interpreter."""
    n = 18
    eng, E = _engine(n, dtype)
    Q, c, b = H.random_qubo(n, seed=2)
    E.set_cost_qubo(Q, c, b)
    E.set_mixer(eng.MIXER_XMULTI)
    spec = orc.Spec(n, orc.qubo_diag(Q, c, b), mixer=("xmulti",))
    rng = np.random.default_rng(8)
    ang = rng.uniform(-1.0, 1.0, size=3 * (1 + n))
    ang[0::1 + n] *= 0.05
    ang[1 + 3] = np.pi / 2                 # layer 0, qubit 3: exactly pi/2
    ang[(1 + n) + 1 + 17] = np.pi / 2 + 1e-9    # layer 1, top qubit
    ang[2 * (1 + n) + 1 + 9] = -np.pi / 2 + 3e-5
    ref = orc.energy(spec, ang)
    E.counter("reset")
    fast = E.energy(ang)
    assert E.counter("prog_launches") == E.counter("launches") - 2 or E.counter("prog_launches") >= 3   # compiled programs only
    all_prog = E.counter("prog_launches")
    _check(fast, spec, ang, dtype, "layer budget", ref=ref)
    E.set_option("layer_budget", 0)
    E.counter("reset")
    slow = E.energy(ang)
    assert E.counter("prog_launches") < all_prog       # some passes needed the interpreter (cot form)
    _check(slow, spec, ang, dtype, "cot form", ref=ref)
    E.close()


@pytest.mark.parametrize("case", ["u8-c64", "u32fx-c64", "f64-c128", "u8-c128", "degenerate"])
def test_device_cvar_weighted_quantile_select(case):
    """qaoa_cvar (digit-by-digit weighted quantile select over the kept final state) against the host evaluation of
    the same definition -- mean cost over the best alpha of the probability mass, statistic.py:132-142 in the
    shots -> infinity limit -- from the device's own |psi|^2 and cost diagonal, and against the oracle's state."""
    n = 16
    dtype = "complex128" if case.endswith("c128") else "complex64"
    eng, E = _engine(n, dtype)
    if case.startswith("u8") or case == "degenerate":
        edges, colors, table = H.regular_maxcut(n)
        lut, _ = H.lut_from_table(table, 1)
        E.set_cost_maxkcut(n, edges, 1, lut)
        gs = 1.0
    else:
        Q, c, b = H.random_qubo(n, seed=17)
        E.set_cost_qubo(Q, c, b)
        gs = 0.05
    diag = E.read_cost_diagonal()
    rng = np.random.default_rng(3)
    ang = rng.uniform(0, 2 * np.pi, size=4)
    ang[0::2] *= gs
    if case == "degenerate":
        ang[:] = 0.0            # uniform distribution over few cost values: the threshold bin is split
    probs = np.asarray(E.final_probabilities(ang), dtype=np.float64)
    probs /= probs.sum()
    order = np.argsort(-diag, kind="stable")
    cs, ps = diag[order], probs[order]
    cum = np.cumsum(ps)
    psi = orc.evolve(orc.Spec(n, diag), ang)
    for alpha in (1.0, 0.5, 0.1, 0.013, 1e-4):
        take = np.clip(alpha - (cum - ps), 0.0, ps)
        want = float(take @ cs) / alpha
        got = E.cvar(ang, alpha)
        tol = 1e-9 if dtype == "complex128" else 2e-5
        assert abs(got - want) <= tol * max(1.0, abs(want)), (case, alpha, got, want)
        wo = orc.exact_cvar(diag, np.abs(psi) ** 2, alpha)
        assert abs(got - wo) <= (1e-8 if dtype == "complex128" else 1e-4) * max(1.0, abs(wo)), (case, alpha, got, wo)
    assert abs(E.cvar(ang, 1.0) - E.energy(ang)[0]) <= (1e-9 if dtype == "complex128" else 1e-5) * abs(E.energy(ang)[0]) + 1e-9
    # one evaluation for the statistics AND the CVaR (qaoa_cvar_stats: what QAOA.loss uses when cvar < 1)
    st, en = E.cvar_stats(ang, 0.3), E.energy(ang)
    # (the kept-state evaluation runs another plan -- full register, storing final pass --: complex64 tolerance)
    assert np.allclose(st[:5], en, rtol=1e-4 if dtype == "complex64" else 1e-9, atol=1e-9), (st, en)
    assert st[5] == pytest.approx(E.cvar(ang, 0.3), rel=1e-12)
    E.close()


@pytest.mark.parametrize("kind", ["u8", "u32fx"])
@pytest.mark.parametrize("n,k", [(16, 0), (17, 6), (18, 5)])
def test_specialised_grover_passes_vs_generic_kernel(n, k, kind):
    """grover_pass_c64_kernel (16-byte accesses, fixed-point phase reduction) against the generic element-wise kernel
    (option force_generic) and the oracle: |+> (k = 0), Dicke on the full register, Dicke in the compact register"""
    eng, E = _engine(n, "complex64")
    if kind == "u8":
        edges, colors, table = H.regular_maxcut(n + (n % 2))
        edges = [e for e in edges if e[0] < n and e[1] < n]
        lut, _ = H.lut_from_table(table, 1)
        E.set_cost_maxkcut(n, edges, 1, lut)
        diag = E.read_cost_diagonal()
        gs = 1.0
    else:
        Q, c, b = H.random_qubo(n, seed=13)
        E.set_cost_qubo(Q, c, b)
        diag = orc.qubo_diag(Q, c, b)
        gs = 0.05
    if k:
        E.set_initial(eng.INIT_DICKE, k)
        E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
        s = orc.dicke_state(n, k)
    else:
        E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_PLUS)
        s = orc.plus_state(n)
    spec = orc.Spec(n, diag, mixer=("grover", s), init=s)
    rng = np.random.default_rng(n + k)
    for p in (1, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[::2] *= gs
        for fs in ((1, 0) if k else (1,)):
            E.set_option("feasible_subspace", fs)
            fast = E.energy(ang)
            _check(fast, spec, ang, "complex64", "fast grover p=%d fs=%d" % (p, fs))
            E.set_option("force_generic", 1)
            slow = E.energy(ang)
            E.set_option("force_generic", 0)
            assert H.rel_err(fast[0], slow[0]) < 1e-5 and H.rel_err(fast[1], slow[1]) < 1e-5, (fast, slow)
            assert fast[2] == slow[2] and fast[3] == slow[3]
    E.close()


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [9, 15])
def test_statevector_init(n, dtype):
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n + (n % 2))
    edges = [(u, v, w) for (u, v, w) in edges if u < n and v < n]
    x = np.arange(1 << n)
    diag = np.zeros(1 << n)
    for (u, v, w) in edges:
        diag += (((x >> u) ^ (x >> v)) & 1) * w
    E.set_cost_diagonal(diag)
    rng = np.random.default_rng(8)
    vec = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    vec /= np.linalg.norm(vec)
    E.set_initial(eng.INIT_STATEVECTOR, vec=vec)
    spec = orc.Spec(n, diag, init=vec)
    ang = np.array([0.7, 0.3, 1.9, 2.4])
    _check(E.energy(ang), spec, ang, dtype, "statevector")


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_kept_state_matches_oracle(dtype):
    n = 15
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n + 1)
    edges = [(u, v, w) for (u, v, w) in edges if u < n and v < n]
    x = np.arange(1 << n)
    diag = np.zeros(1 << n)
    for (u, v, w) in edges:
        diag += (((x >> u) ^ (x >> v)) & 1) * w
    E.set_cost_diagonal(diag)
    E.set_option("keep_state", 1)
    ang = np.array([0.7, 2.3, 1.9, 0.4, 0.2, 1.3])   # betas on both sides of |tan| = 1
    E.energy(ang)
    psi = E.read_statevector()
    ref = orc.evolve(orc.Spec(n, diag), ang)
    # global phase is not physical
    ph = np.vdot(ref, psi)
    ph /= abs(ph)
    assert np.abs(psi - ph * ref).max() < (2e-6 if dtype == "complex64" else 1e-13)


def test_batch_landscape_matches_oracle():
    n = 14
    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = H.maxcut_diag(n, edges, colors, table)
    spec = orc.Spec(n, diag)
    gam = np.linspace(0, 2 * np.pi, 5, endpoint=False)
    bet = np.linspace(0, 2 * np.pi, 4, endpoint=False)
    ang = np.array([[g, b] for b in bet for g in gam])
    out = E.energy_batch(ang)
    for i, a in enumerate(ang):
        e = orc.energy(spec, a)[0]
        assert H.rel_err(out[i, 0], e) < 1e-4


def test_errors_are_reported():
    eng, E = _engine(6, "complex64")
    with pytest.raises(eng.EngineError):
        E.energy([0.1, 0.2])  # no cost set
    with pytest.raises(eng.EngineError):
        E.set_cost_qubo(np.eye(5))


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_canonical_programs_vs_interpreter_vs_oracle(dtype):
    """n=24: the specialised straight-line pass kernels and the interpreter kernel must agree
    with each other and with the oracle (p=1: FIRST+FINAL; p=3: FIRST, INTERIOR/BOUNDARY, FINAL)."""
    n = 24
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = H.maxcut_diag(n, edges, colors, table)
    spec = orc.Spec(n, diag)
    rng = np.random.default_rng(2024)
    for p in (1, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)      # keep |tan beta| moderate: normal form
        E.set_option("force_generic", 0)
        E.counter("reset")
        a = E.energy(ang)
        assert E.counter("prog_launches") >= 2, "specialised kernels were not used"
        E.set_option("force_generic", 1)
        b = E.energy(ang)
        ref = orc.energy(spec, ang)
        assert H.rel_err(a[0], ref[0]) < TOL[dtype] and H.rel_err(b[0], ref[0]) < TOL[dtype], (a, b, ref)
        assert H.rel_err(a[0], b[0]) < TOL[dtype] * 0.1


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_beta_at_and_near_half_pi(dtype):
    """|tan beta| huge or infinite: the host must switch to the cot form (interpreter kernel)."""
    n = 16
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    spec = orc.Spec(n, H.maxcut_diag(n, edges, colors, table))
    for betas in ([np.pi / 2, 0.3], [np.pi / 2 - 1e-9, 3 * np.pi / 2], [1.5, 1.6], [0.0, np.pi]):
        ang = np.array([0.8, betas[0], 0.35, betas[1]])
        _check(E.energy(ang), spec, ang, dtype, "betas=%r" % (betas,))


@pytest.mark.parametrize("n,ps,low_bits,dtype", [(25, (1, 2, 3), 0, "complex64"), (26, (2,), 0, "complex64"),
                                                  (25, (1, 3), 4, "complex64"), (26, (2,), 4, "complex64"),
                                                  (25, (1, 2, 3), 0, "complex128"), (26, (2,), 0, "complex128")])
def test_balanced_schedule_vs_classic_vs_oracle(n, ps, low_bits, dtype):
    """complex64, n >= 25 (5 low bits, 3 tile shapes): the balanced schedule (MID + BOUND2 passes through
    frame C, FINAL2) must agree with the classic schedule (INTERIOR + BOUNDARY), with the interpreter
    running the same balanced plan, and with the oracle (C twin of the numpy oracle, pinned to it in
    tests/test_oracle.py -- the numpy one needs minutes at this size).  n=26 pads the top tile with lower bits.
    low_bits=4 forces the tile shapes n = 33 / 34 use (BOUND2L4 / FINAL2L4 programs)."""
    from oracle import c_oracle
    eng, E = _engine(n, dtype)
    E.set_option("symmetric", 0)   # the full register (tests/test_gpu_symmetric.py covers the half register)
    if low_bits:   # 4 low bits (what n = 33, 34 choose): high-tile passes run in frames A4 / C4 with 8-byte accesses
        E.set_option("low_bits", low_bits)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = c_oracle.maxcut_diag(n, edges)
    got_diag = E.read_cost_diagonal(0, 1 << 16)
    assert np.array_equal(got_diag, diag[: 1 << 16])
    rng = np.random.default_rng(77)
    for p in ps:
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)
        ref = c_oracle.energy(n, diag, ang, "complex128")
        E.set_option("schedule", 0)
        E.set_option("force_generic", 0)
        E.counter("reset")
        a = E.energy(ang)
        assert E.counter("prog_launches") >= 2
        npass_bal = E.counter("n_passes")
        E.set_option("force_generic", 1)
        b = E.energy(ang)
        E.set_option("force_generic", 0)
        E.set_option("schedule", 1)
        c = E.energy(ang)
        assert E.counter("n_passes") == npass_bal, "same number of HBM sweeps in both schedules"
        for got in (a, b, c):
            _check(got, None, ang, dtype, "n=%d p=%d" % (n, p), ref=ref)
        assert H.rel_err(a[0], b[0]) < TOL[dtype] * 0.1 and H.rel_err(a[0], c[0]) < TOL[dtype] * 0.1, (a, b, c, ref)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_balanced_schedule_real_valued_costs_and_multiangle(dtype):
    """Balanced schedule with a real-valued diagonal (QUBO: U32FX in complex64, F64 in complex128) and per-qubit
    betas (XMultiAngle)."""
    from oracle import c_oracle
    n, p = 25, 2
    eng, E = _engine(n, dtype)
    Q, c, b = H.random_qubo(n)
    E.set_cost_qubo(Q, c, b)
    E.set_mixer(eng.MIXER_XMULTI)
    diag = c_oracle.qubo_diag(Q, c, b)
    rng = np.random.default_rng(5)
    ang = rng.uniform(-1.0, 1.0, size=p * (1 + n))
    ang[0::(1 + n)] *= 0.05     # gammas: QUBO costs are O(100)
    ref = c_oracle.energy(n, diag, ang, "complex128", n_beta=n)
    E.counter("reset")
    out = E.energy(ang)
    assert E.counter("prog_launches") >= 2
    _check(out, None, ang, dtype, "qubo n=25", ref=ref)


@pytest.mark.parametrize("n,dtype", [(10, "complex128"), (10, "complex64"), (18, "complex64"), (15, "complex128")])
def test_cost_distribution_and_sampler(n, dtype):
    """qaoa_cost_distribution (exact CVaR input) and qaoa_sample (QAOA.hist) against the oracle's final state:
    single-CTA path (n=10) and tiled path with the state kept in HBM (n=15, 18)."""
    eng, E = _engine(n, dtype)
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = H.maxcut_diag(n, edges, colors, table)
    spec = orc.Spec(n, diag)
    ang = np.array([0.45, 0.3, 0.7, -0.4])
    prob = np.abs(orc.evolve(spec, ang)) ** 2
    vals, pr = E.cost_distribution(ang)
    ref = np.array([prob[diag == v].sum() for v in vals])
    tol = 1e-10 if dtype == "complex128" else 2e-5
    assert set(vals) <= set(np.unique(diag)) and abs(pr.sum() - 1) < 1e-12
    assert np.abs(pr - ref).max() < tol, np.abs(pr - ref).max()
    assert orc.exact_cvar(diag, prob, 0.25) == pytest.approx(
        float(np.clip(0.25 - (np.cumsum(pr[::-1]) - pr[::-1]), 0, pr[::-1]) @ vals[::-1]) / 0.25, rel=10 * tol)
    # sampler: deterministic for a seed, and distributed like |psi|^2 (chi-square over cost classes)
    shots = 40000
    idx = E.sample(ang, shots, seed=7)
    assert np.array_equal(idx, E.sample(ang, shots, seed=7))
    assert idx.max() < (1 << n) and prob[idx.astype(np.int64)].min() > 0
    counts = np.array([(diag[idx.astype(np.int64)] == v).sum() for v in vals], dtype=float)
    keep = ref * shots > 5
    chi2 = (((counts - ref * shots) ** 2) / (ref * shots))[keep].sum()
    assert chi2 < 3 * keep.sum() + 20, (chi2, keep.sum())
    # evaluation afterwards is unaffected (keep_state was switched on only for the duration of the call)
    _check(E.energy(ang), spec, ang, dtype, "after sampling")


@pytest.mark.parametrize("seed,trials,nlo,nhi", [(20261019, 60, 4, 18), (777, 16, 18, 23)])
def test_randomised_configurations_against_oracle(seed, trials, nlo, nhi):
    """Seeded sweep over the supported combinations (problem x mixer x initial state x precision x depth x size,
    including arbitrary XY pair lists, trotter steps, weighted graphs, k-cut colour tables and fixed nodes):
    every one must agree with the numpy oracle within the stated tolerance.  The second sweep stays on the
    large-register kernels (tiled X passes, XY tiles, element-wise Grover)."""
    from qaoa_b200 import engine as eng
    rng = np.random.default_rng(seed)
    checked = 0
    for trial in range(trials):
        n = int(rng.integers(nlo, nhi))
        dtype = ["complex64", "complex128"][trial % 2]
        p = int(rng.integers(1, 4))
        E = eng.Engine(n, dtype=dtype)
        kind = ["maxcut", "wmaxcut", "qubo", "diag", "kcut"][int(rng.integers(0, 5))]
        if kind == "kcut" and n % 2:
            kind = "maxcut"
        if kind in ("maxcut", "wmaxcut"):
            G = orc.relabel_graph(__import__("networkx").gnp_random_graph(n, 0.4, seed=int(rng.integers(1 << 30))))
            edges = [(u, v, (1.0 if kind == "maxcut" else float(rng.uniform(-1, 2)))) for (u, v, _w) in orc.graph_edges(G)]
            if not edges:
                edges = [(0, 1, 1.0)]
            colors, table = orc.colour_table(2)
            lut, _ = H.lut_from_table(table, 1)
            E.set_cost_maxkcut(n, edges, 1, lut)
            diag = orc.maxkcut_diag(edges, n, 1, table, colors)
        elif kind == "kcut":      # k = 4 colours, two qubits per node
            nv = n // 2
            G = orc.relabel_graph(__import__("networkx").gnp_random_graph(nv, 0.6, seed=int(rng.integers(1 << 30))))
            edges = orc.graph_edges(G) or [(0, 1, 1.0)]
            colors, table = orc.colour_table(4)
            lut, _ = H.lut_from_table(table, 2)
            E.set_cost_maxkcut(nv, edges, 2, lut)
            diag = orc.maxkcut_diag(edges, nv, 2, table, colors)
        elif kind == "qubo":
            Q, c, b = H.random_qubo(n, seed=int(rng.integers(1 << 30)))
            E.set_cost_qubo(Q, c, b)
            diag = orc.qubo_diag(Q, c, b)
        else:
            diag = np.round(rng.uniform(-3, 9, size=1 << n), 3)
            E.set_cost_diagonal(diag)
        got = E.read_cost_diagonal()
        scale = max(1.0, np.abs(diag).max())
        assert np.abs(got - diag).max() <= (1e-12 if dtype == "complex128" else 3e-9) * scale, (trial, kind)
        k = int(rng.integers(1, n))
        init_kind = ["plus", "dicke", "vec"][int(rng.integers(0, 3))]
        mixer_kind = ["x", "xmulti", "xy", "grover"][int(rng.integers(0, 4))]
        if mixer_kind in ("xy",) and init_kind == "plus":
            init_kind = "dicke"
        if init_kind == "plus":
            init = orc.plus_state(n)
            E.set_initial(eng.INIT_PLUS)
        elif init_kind == "dicke":
            init = orc.dicke_state(n, k)
            E.set_initial(eng.INIT_DICKE, k)
        else:
            init = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
            init /= np.linalg.norm(init)
            E.set_initial(eng.INIT_STATEVECTOR, vec=init)
        trotter, nb = 1, 1
        if mixer_kind == "x":
            mixer = ("x",)
            E.set_mixer(eng.MIXER_X)
        elif mixer_kind == "xmulti":
            mixer, nb = ("xmulti",), n
            E.set_mixer(eng.MIXER_XMULTI)
        elif mixer_kind == "xy":
            trotter = int(rng.integers(1, 3))
            if rng.random() < 0.5:
                pairs = orc.xy_pairs(n, ["ring", "chain"][int(rng.integers(0, 2))])
            else:       # arbitrary ordered pair list
                pairs = [list(rng.choice(n, size=2, replace=False)) for _ in range(int(rng.integers(1, 2 * n)))]
                pairs = [[int(a), int(b)] for a, b in pairs]
            mixer = ("xy", pairs)
            E.set_mixer(eng.MIXER_XY, pairs=pairs, trotter=trotter)
        else:
            s = orc.dicke_state(n, k) if init_kind != "plus" else orc.plus_state(n)
            mixer = ("grover", s)
            E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE if init_kind != "plus" else eng.INIT_PLUS, sub_k=k)
        spec = orc.Spec(n, diag, mixer=mixer, init=init, trotter=trotter)
        ang = rng.uniform(-np.pi, np.pi, size=p * (1 + nb))
        ang[0::(1 + nb)] *= 0.3 / max(1.0, np.abs(diag).max() / 10)
        out = E.energy(ang)
        e, e2, mn, mx = orc.energy(spec, ang)
        tol = TOL[dtype]
        ref_scale = max(abs(e), 1e-2 * scale)
        assert abs(out[0] - e) < tol * ref_scale, (trial, n, dtype, kind, init_kind, mixer_kind, p, out, e)
        assert abs(out[1] - e2) < 2 * tol * max(abs(e2), 1e-2 * scale * scale), (trial, out, e2)
        assert abs(out[4] - 1.0) < (2e-4 if dtype == "complex64" else 1e-10), (trial, out)
        if dtype == "complex128":
            assert abs(out[2] - mn) < 1e-9 * scale and abs(out[3] - mx) < 1e-9 * scale, (trial, out, mn, mx)
        E.close()
        checked += 1
    assert checked == trials


def test_balanced_schedule_batched_points():
    """Several points per launch on the balanced schedule (n=25): the persistent CTAs reload the per-point residents
    (round coefficients, phase table) and flush their reduction accumulators whenever they move to another point."""
    from oracle import c_oracle
    n, p = 25, 2
    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    diag = c_oracle.maxcut_diag(n, edges)
    rng = np.random.default_rng(99)
    rows = rng.uniform(-1.2, 1.2, size=(5, 2 * p))
    E.set_option("batch_points", 3)          # 5 points = one chunk of 3 and one of 2
    E.counter("reset")
    got = E.energy_batch(rows)
    assert E.counter("prog_launches") >= 2
    for i in (0, 2, 4):
        ref = c_oracle.energy(n, diag, rows[i], "complex128")
        _check(got[i], None, rows[i], "complex64", "point %d" % i, ref=ref)
    E.set_option("batch_points", 1)
    one = np.stack([E.energy(r) for r in rows])
    assert np.allclose(got[:, :2], one[:, :2], rtol=2e-6)
    assert np.array_equal(got[:, 2:4], one[:, 2:4])


def test_full_size_properties_n32():
    """BASELINE size (MaxCut 3-regular n=32, complex64, 32 GiB state) through size-independent properties:
    the closed-form p=1 energy, the (gamma, beta) -> (-gamma, -beta) symmetry of the energy,
    unit norm after p=5, extremes of the cost over the support, and bit-exact samples of the cost diagonal."""
    import torch
    if torch.cuda.mem_get_info()[0] < 60 * 2 ** 30:
        pytest.skip("needs 60 GiB of free device memory")
    n = 32
    eng, E = _engine(n, "complex64")
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    E.set_cost_maxkcut(n, edges, 1, lut)
    rng = np.random.default_rng(3)
    for off in (0, (1 << 31) + 12345, (1 << 32) - 4096):       # diagonal spot checks against cost(x)
        got = E.read_cost_diagonal(off, 4096)
        x = off + np.arange(4096, dtype=np.int64)
        ref = np.zeros(4096)
        for (u, v, w) in edges:
            ref += (((x >> u) ^ (x >> v)) & 1) * w
        assert np.array_equal(got, ref)
    for _ in range(2):
        g, b = rng.uniform(0.1, 1.2), rng.uniform(0.1, 1.2)
        out = E.energy([g, b])
        assert out[0] == pytest.approx(orc.maxcut_p1_closed_form(edges, n, g, b), rel=1e-4)
        assert abs(out[4] - 1) < 1e-4 and out[2] == 0 and 0 < out[3] <= len(edges)
    ang = rng.uniform(-1.0, 1.0, size=10)
    a, b2 = E.energy(ang), E.energy(-ang)
    assert a[0] == pytest.approx(b2[0], rel=1e-5) and abs(a[4] - 1) < 1e-4
    lo, hi = E.cost_minmax()
    assert a[2] >= lo and a[3] <= hi and lo == 0


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_uploads_are_ordered_before_the_kernels_that_use_them(dtype):
    """Regression: host buffers (StateVector amplitudes, host diagonals) are uploaded on the engine's own stream.
    A plain cudaMemcpy from pageable memory may return with its DMA in flight, and the engine's non-blocking stream
    does not wait for it -- fresh engines, large vectors, results must be reproducible and match the oracle."""
    n = 18
    rng = np.random.default_rng(5)
    diag = np.round(rng.uniform(0, 40, size=1 << n))
    vec = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    vec /= np.linalg.norm(vec)
    s = orc.dicke_state(n, 6)
    spec = orc.Spec(n, diag, mixer=("grover", s), init=vec)
    ang = np.array([0.05, 0.8, 0.03, -0.4])
    ref = orc.energy(spec, ang)
    for _ in range(4):
        eng, E = _engine(n, dtype)
        E.set_cost_diagonal(diag)
        E.set_initial(eng.INIT_STATEVECTOR, vec=vec)
        E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=6)
        out = E.energy(ang)
        assert H.rel_err(out[0], ref[0]) < TOL[dtype], (out, ref)
        assert np.array_equal(E.read_cost_diagonal(), diag)
        E.close()
