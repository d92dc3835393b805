"""GPU: the sharded register with every rank's engine inside ONE process on ONE GPU.

The cross passes address the peers' shards through the same peer table a multi-GPU run uses (here the
"peers" are other buffers of the same device), and passes of different ranks never overlap unless the
protocol allows it -- so this checks the sharded kernels, the planner and the reduction without
needing several GPUs.  The multi-process path (CUDA IPC + torch.distributed) is exercised by
scripts/run_sharded.py on a multi-GPU box and by tests/test_sharded_protocol.py (gloo, CPU)."""
import numpy as np
import pytest

import parity_helpers as H

pytestmark = pytest.mark.gpu

TOL = {"complex64": 1e-4, "complex128": 1e-9}


def _setup(n, world, dtype):
    from oracle import c_oracle
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    grp = LocalShardGroup(n, world, dtype=dtype)
    grp.set_cost_maxkcut(n, edges, 1, lut)
    diag = c_oracle.maxcut_diag(n, edges)
    return eng, grp, c_oracle, diag, edges, lut


# (26, 4) and (25, 2): 24 local bits, where the cross tile carries a spare local bit to keep 256-byte rows
@pytest.mark.parametrize("n,world,dtype", [(24, 2, "complex64"), (25, 4, "complex64"), (26, 8, "complex64"),
                                           (23, 2, "complex128"), (20, 4, "complex64"), (26, 2, "complex128"),
                                           (26, 4, "complex64"), (25, 2, "complex64")])
def test_sharded_energy_matches_oracle_and_single_engine(n, world, dtype):
    eng, grp, c_oracle, diag, edges, lut = _setup(n, world, dtype)
    assert np.array_equal(grp.read_cost_diagonal(), diag), "sharded cost diagonal is not bit-exact"
    single = eng.Engine(n, dtype=dtype)
    single.set_cost_maxkcut(n, edges, 1, lut)
    rng = np.random.default_rng(11)
    for p in (1, 2, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)
        ref = c_oracle.energy(n, diag, ang, "complex128")
        got = grp.energy(ang)
        one = single.energy(ang)
        assert any(grp.last_cross_flags), "no pass touched peer memory"
        tol = TOL[dtype]
        assert H.rel_err(got[0], ref[0]) < tol and H.rel_err(got[1], ref[1]) < 2 * tol, (p, got, ref)
        assert got[2] == ref[2] and got[3] == ref[3]
        assert abs(got[4] - 1) < (1e-4 if dtype == "complex64" else 1e-11)
        assert H.rel_err(got[0], one[0]) < tol * 0.1, (got, one)
    grp.close()


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("n,world,k", [(18, 2, 6), (20, 4, 0), (21, 8, 5)])
def test_sharded_grover_mixer(n, world, k, dtype):
    """Grover mixer on a sharded register (qaoa_shard_grover_step): |s> = Dicke(k) (k > 0) or |+>, generated per shard
    with the global popcount; the only cross-rank traffic is the partial dot product of every layer.  Against the C
    oracle and the single engine's full-register path."""
    from oracle import c_oracle
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    Q, c, b = H.random_qubo(n, seed=n)
    grp = LocalShardGroup(n, world, dtype=dtype)
    grp.set_cost_qubo(Q, c, b)
    single = eng.Engine(n, dtype=dtype)
    single.set_option("feasible_subspace", 0)
    single.set_cost_qubo(Q, c, b)
    if k:
        for E in (grp, single):
            E.set_initial(eng.INIT_DICKE, k)
            E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
        st = ("dicke", k)
    else:
        for E in (grp, single):
            E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_PLUS)
        st = ("plus", 0)
    diag = c_oracle.qubo_diag(Q, c, b)
    rng = np.random.default_rng(5)
    for p in (1, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[0::2] *= 0.05
        ref = c_oracle.run(n, diag, ang, "complex128", mixer="grover", init=st, sub=st)
        got = grp.energy(ang)
        one = single.energy(ang)
        tol = TOL[dtype]
        assert H.rel_err(got[0], ref[0]) < tol and H.rel_err(got[1], ref[1]) < 2 * tol, (p, got, ref)
        assert abs(got[2] - ref[2]) <= 1e-6 * abs(ref[2]) and abs(got[3] - ref[3]) <= 1e-6 * abs(ref[3]), (got, ref)
        assert abs(got[4] - 1) < (1e-4 if dtype == "complex64" else 1e-11)
        assert H.rel_err(got[0], one[0]) < tol * 0.1, (got, one)
    grp.close()
    single.close()


@pytest.mark.parametrize("n,world,symmetric,dtype", [(28, 2, True, "complex64"), (26, 2, True, "complex64"),
                                                    (28, 8, True, "complex64"), (28, 4, True, "complex64"),
                                                    (27, 2, False, "complex64"), (28, 4, False, "complex64"),
                                                    (26, 2, False, "complex128"), (24, 2, False, "complex128")])
def test_chunk_pipeline_chunks_are_independent(n, world, symmetric, dtype):
    """The chunk pipeline (engine.cu plan_chunks, sharded.py pipeline_segments): closed passes launched chunk by chunk,
    and inside every window chunk c runs pre -> cross -> post BEFORE chunk c+1 starts -- if a closed pass coupled two
    chunks, or the tile remap / partial-sum slots of a chunked launch were off, the result would differ from the
    whole-pass evaluation and the oracle.  (One process, one GPU: no overlap here, only the data flow.)"""
    from oracle import c_oracle
    from qaoa_b200.sharded import LocalShardGroup, pipeline_segments

    edges, colors, table = H.regular_maxcut(n + (n & 1))
    edges = [e for e in edges if e[0] < n and e[1] < n]
    lut, _ = H.lut_from_table(table, 1)
    grp = LocalShardGroup(n, world, dtype=dtype, symmetric=symmetric)
    if (n, world, symmetric) == (28, 2, True):
        grp.set_option("persistent_ctas", 32)       # fewer CTAs: the planner may cut the shard into eight chunks
    grp.set_cost_maxkcut(n, edges, 1, lut)
    d8 = c_oracle.maxcut_diag_u8(n, edges)
    rng = np.random.default_rng(n + world)
    seen_chunks = 0
    tol = TOL[dtype]
    for p in (1, 2, 3, 5):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)
        grp.chunk_major = False
        whole = grp.energy(ang)
        grp.chunk_major = True
        got = grp.energy(ang)
        seen_chunks = max(seen_chunks, grp.last_chunks)
        ref = c_oracle.run(n, d8, ang, "complex128", blocked=True)
        assert H.rel_err(got[0], ref[0]) < tol and H.rel_err(got[1], ref[1]) < 2 * tol, (p, got, ref)
        assert got[2] == ref[2] and got[3] == ref[3] and abs(got[4] - 1) < (1e-4 if dtype == "complex64" else 1e-11)
        assert H.rel_err(got[0], whole[0]) < tol * 0.01 and H.rel_err(got[1], whole[1]) < tol * 0.01, (p, got, whole)
    assert seen_chunks >= 2 or (n, dtype) == (24, "complex128"), "no plan of this configuration had chunks"
    if (n, world, symmetric) == (28, 2, True):
        assert seen_chunks == 8
    grp.close()


@pytest.mark.parametrize("case", ["x-u8-c64", "x-qubo-c128", "grover-qubo-c64"])
def test_sampling_and_cvar_on_a_sharded_register(case):
    """QAOA.hist / cvar < 1 on a sharded (full) register: every rank keeps its shard of the final state, the per-chunk
    masses / per-digit histograms are combined across the shards (sharded.py deal_shots, cvar_select).  Same uniform
    numbers => the same shots as the single engine's sampler; CVaR equal to the single engine's."""
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    n, world = 20, 4
    dtype = "complex128" if case.endswith("c128") else "complex64"
    grp = LocalShardGroup(n, world, dtype=dtype)
    one = eng.Engine(n, dtype=dtype)
    if "u8" in case:
        edges, colors, table = H.regular_maxcut(n)
        lut, _ = H.lut_from_table(table, 1)
        for E in (grp, one):
            E.set_cost_maxkcut(n, edges, 1, lut)
        gs = 1.0
    else:
        Q, c, b = H.random_qubo(n, seed=9)
        for E in (grp, one):
            E.set_cost_qubo(Q, c, b)
        gs = 0.05
    if case.startswith("grover"):
        one.set_option("feasible_subspace", 0)
        for E in (grp, one):
            E.set_initial(eng.INIT_DICKE, 7)
            E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=7)
    ang = np.random.default_rng(2).uniform(0, 2 * np.pi, size=6)
    ang[0::2] *= gs
    a = grp.sample(ang, 4096, seed=123)
    b_ = one.sample(ang, 4096, seed=123)
    assert a.shape == b_.shape and (a == b_).mean() > 0.995, (a[:8], b_[:8])
    if case.startswith("grover"):
        assert all(bin(int(x)).count("1") == 7 for x in a[:256])      # samples stay in the feasible subspace
    tol = 1e-9 if dtype == "complex128" else 2e-5
    for alpha in (1.0, 0.3, 0.02):
        got, want = grp.cvar(ang, alpha), one.cvar(ang, alpha)
        assert abs(got - want) <= tol * max(1.0, abs(want)), (case, alpha, got, want)
    e1, e2 = grp.energy(ang), one.energy(ang)          # back to the ordinary (not kept) evaluation
    assert H.rel_err(e1[0], e2[0]) < TOL[dtype]
    grp.close()
    one.close()


def test_sharded_multiangle_dicke_and_cot_form():
    """Per-qubit betas follow the logical qubit across the shard boundary; Dicke initial state generated
    per shard with the global popcount; beta = pi/2 exercises the interpreter's cross variant."""
    from oracle import c_oracle
    from oracle import qaoa_oracle as orc
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    n, world = 20, 4
    Q, c, b = H.random_qubo(n)
    grp = LocalShardGroup(n, world, dtype="complex128")
    grp.set_cost_qubo(Q, c, b)
    grp.set_mixer(eng.MIXER_XMULTI)
    diag = c_oracle.qubo_diag(Q, c, b)
    assert np.abs(grp.read_cost_diagonal() - diag).max() <= 1e-12 * np.abs(diag).max()
    rng = np.random.default_rng(3)
    ang = rng.uniform(-1.0, 1.0, size=2 * (1 + n))
    ang[0::(1 + n)] *= 0.05
    ang[1 + n - 1] = np.pi / 2          # top (global) qubit, layer 0: cot form
    ref = c_oracle.energy(n, diag, ang, "complex128", n_beta=n)
    got = grp.energy(ang)
    assert H.rel_err(got[0], ref[0]) < 1e-9 and H.rel_err(got[1], ref[1]) < 2e-9, (got, ref)
    grp.close()

    n, world, k = 18, 2, 7
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    grp = LocalShardGroup(n, world, dtype="complex128")
    grp.set_cost_maxkcut(n, edges, 1, lut)
    grp.set_initial(eng.INIT_DICKE, k)
    diag = H.maxcut_diag(n, edges, colors, table)
    spec = orc.Spec(n, diag, init=orc.dicke_state(n, k))
    ang = np.array([0.4, 0.3, 0.7, -0.5])
    ref = orc.energy(spec, ang)
    got = grp.energy(ang)
    assert H.rel_err(got[0], ref[0]) < 1e-9, (got, ref)
    assert got[2] == ref[2] and got[3] == ref[3]
    grp.close()


def test_randomised_sharded_configurations_against_the_single_engine():
    """Seeded sweep: size, number of ranks, precision, cost kind, X / multi-angle X, Plus / Dicke, depth -- the
    sharded register must reproduce the single-engine result (itself pinned to the oracle by test_gpu_parity)."""
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    rng = np.random.default_rng(4242)
    for trial in range(14):
        world = int(rng.choice([2, 4, 8]))
        g = world.bit_length() - 1
        dtype = ["complex64", "complex128"][trial % 2]
        n = int(rng.integers(16 + g, 25))
        p = int(rng.integers(1, 4))
        grp = LocalShardGroup(n, world, dtype=dtype)
        one = eng.Engine(n, dtype=dtype)
        if rng.random() < 0.5:
            edges, colors, table = H.regular_maxcut(n, seed=int(rng.integers(1 << 20)))
            lut, _ = H.lut_from_table(table, 1)
            for e in (grp, one):
                e.set_cost_maxkcut(n, edges, 1, lut)
        else:
            Q, c, b = H.random_qubo(n, seed=int(rng.integers(1 << 20)))
            for e in (grp, one):
                e.set_cost_qubo(Q, c, b)
        assert np.array_equal(grp.read_cost_diagonal(), one.read_cost_diagonal())
        nb = 1
        if rng.random() < 0.5:
            nb = n
            for e in (grp, one):
                e.set_mixer(eng.MIXER_XMULTI)
        if rng.random() < 0.4:
            k = int(rng.integers(1, n))
            for e in (grp, one):
                e.set_initial(eng.INIT_DICKE, k)
        lo, hi = one.cost_minmax()
        ang = rng.uniform(-np.pi, np.pi, size=p * (1 + nb))
        ang[0::(1 + nb)] *= 3.0 / max(1.0, hi - lo)
        a, b = grp.energy(ang), one.energy(ang)
        tol = TOL[dtype]
        scale = max(abs(b[0]), 1e-2 * max(abs(lo), abs(hi), 1.0))
        assert abs(a[0] - b[0]) < tol * scale, (trial, n, world, dtype, nb, a, b)
        assert abs(a[4] - 1) < (2e-4 if dtype == "complex64" else 1e-10)
        if dtype == "complex128":
            assert a[2] == b[2] and a[3] == b[3]
        grp.close()
        one.close()


def _twisted_index(z, rank, el, g):
    """stored basis index x' of local index z on `rank` in twisted coordinates (csrc/tile_kernel.cuh)"""
    mask = (1 << g) - 1
    r = np.where(z & 1, rank ^ mask, rank)
    return z | (r.astype(np.int64) << el)


@pytest.mark.parametrize("n,world", [(20, 2), (22, 4), (22, 8), (24, 2), (26, 8), (26, 4)])
def test_sharded_symmetric_register(n, world):
    """MaxCut + X + |+>, n even: each rank stores 2^(n-1-g) amplitudes of the half register in twisted coordinates;
    the mirror pairing stays inside the shard, qubit 0 is mixed in the cross passes.  Checked against the C oracle,
    the plain sharded register and the single engine; diagonal bit-exact under the coordinate map."""
    eng, grp_full, c_oracle, diag, edges, lut = _setup(n, world, "complex64")
    from qaoa_b200.sharded import LocalShardGroup

    g = world.bit_length() - 1
    grp = LocalShardGroup(n, world, dtype="complex64", symmetric=True)
    grp.set_cost_maxkcut(n, edges, 1, lut)
    el = n - 1 - g
    for r, e in enumerate(grp.engines):
        assert e.local_qubits == el
        z = np.arange(1 << el, dtype=np.int64)
        assert np.array_equal(e.read_cost_diagonal(), diag[_twisted_index(z, r, el, g)]), "twisted diagonal differs"
    rng = np.random.default_rng(5)
    for p in (1, 2, 3, 4):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        ang[1::2] = rng.uniform(-1.2, 1.2, size=p)
        ref = c_oracle.energy(n, diag, ang, "complex128")
        got = grp.energy(ang)
        full = grp_full.energy(ang)
        assert any(grp.last_cross_flags)
        assert H.rel_err(got[0], ref[0]) < 1e-4 and H.rel_err(got[1], ref[1]) < 2e-4, (p, got, ref)
        assert got[2] == ref[2] and got[3] == ref[3]
        assert abs(got[4] - 1) < 1e-4
        assert H.rel_err(got[0], full[0]) < 2e-5, (got, full)
    assert grp.engines[0].counter("state_bytes") == 8 << el
    if (n, world) == (26, 4):   # 23 stored local bits: the cross tile carries one of them, the local tiles keep 5 low bits
        assert grp.engines[0].counter("cross_spare_bits") == 1
        assert grp_full.engines[0].counter("cross_spare_bits") == 1
    grp.close()
    grp_full.close()


@pytest.mark.parametrize("n,world", [(20, 4), (26, 4)])   # (26, 4): the cross tile carries a spare local bit
def test_sharded_symmetric_multiangle_cot_form_and_interpreter(n, world):
    """Per-qubit betas (qubit 0 and the rank qubits included), beta = pi/2 on qubit 0 / a rank qubit / the top
    qubit (cot form -> interpreter kernels), and the interpreter forced for every pass."""
    from oracle import c_oracle
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    edges, colors, table = H.regular_maxcut(n, seed=7)
    lut, _ = H.lut_from_table(table, 1)
    diag = c_oracle.maxcut_diag(n, edges)
    rng = np.random.default_rng(9)
    for variant in range(4):
        grp = LocalShardGroup(n, world, dtype="complex64", symmetric=True)
        grp.set_cost_maxkcut(n, edges, 1, lut)
        grp.set_mixer(eng.MIXER_XMULTI)
        p = 3
        ang = rng.uniform(-1.0, 1.0, size=p * (1 + n))
        ang[0::(1 + n)] = rng.uniform(0, 2 * np.pi, size=p)
        if variant == 1:
            ang[1 + 0] = np.pi / 2                    # qubit 0, layer 0
            ang[(1 + n) + 1 + (n - 2)] = np.pi / 2    # a rank qubit, layer 1
        if variant == 2:
            ang[2 * (1 + n) + 1 + (n - 1)] = np.pi / 2   # the virtual top qubit, layer 2
        if variant == 3:
            grp.set_option("force_generic", 1)
        ref = c_oracle.energy(n, diag, ang, "complex128", n_beta=n)
        got = grp.energy(ang)
        assert H.rel_err(got[0], ref[0]) < 1e-4 and H.rel_err(got[1], ref[1]) < 2e-4, (variant, got, ref)
        assert abs(got[4] - 1) < 1e-4
        grp.close()


def test_sharded_symmetric_rejects_what_it_cannot_hold():
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    with pytest.raises(eng.EngineError):
        LocalShardGroup(21, 2, symmetric=True)                       # odd n
    with pytest.raises(eng.EngineError):
        LocalShardGroup(20, 2, dtype="complex128", symmetric=True)   # complex64 only
    grp = LocalShardGroup(20, 2, symmetric=True)
    Q, c, b = H.random_qubo(20)
    with pytest.raises(eng.EngineError):
        grp.set_cost_qubo(Q, c, b)
    edges, colors, table = H.regular_maxcut(20)
    lut, _ = H.lut_from_table(table, 1)
    grp.set_cost_maxkcut(20, edges, 1, lut)
    grp.set_initial(eng.INIT_DICKE, 3)
    with pytest.raises(eng.EngineError):
        grp.energy([0.3, 0.2])
    grp.close()


@pytest.mark.parametrize("n,world,symmetric,k", [(24, 2, True, 3), (24, 4, True, 2), (24, 2, False, 2), (23, 4, False, 2),
                                                 (24, 8, True, 1), (22, 2, True, 0)])
def test_cross_tile_carries_spare_local_bits(n, world, symmetric, k):
    """The cross tile mixes the rank qubits in register slots of frame A; its spare slots may carry the top k local
    qubits (mixed for both layers of every cross pass, needed at 33+ local bits to keep 256-byte rows).  Forced here
    at small sizes (option cross_spare = 10 + k), multi-angle X so that every qubit's angle is its own."""
    from oracle import c_oracle
    from qaoa_b200 import engine as eng
    from qaoa_b200.sharded import LocalShardGroup

    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    diag = c_oracle.maxcut_diag(n, edges)
    grp = LocalShardGroup(n, world, dtype="complex64", symmetric=symmetric)
    grp.set_option("cross_spare", 10 + k)
    grp.set_cost_maxkcut(n, edges, 1, lut)
    grp.set_mixer(eng.MIXER_XMULTI)
    rng = np.random.default_rng(17 + k)
    for p in (1, 2, 3, 4):
        ang = rng.uniform(-1.0, 1.0, size=p * (1 + n))
        ang[0::(1 + n)] = rng.uniform(0, 2 * np.pi, size=p)
        ref = c_oracle.energy(n, diag, ang, "complex128", n_beta=n)
        got = grp.energy(ang)
        assert H.rel_err(got[0], ref[0]) < 1e-4 and H.rel_err(got[1], ref[1]) < 2e-4, (p, got, ref)
        assert abs(got[4] - 1) < 1e-4
    assert grp.engines[0].counter("cross_spare_bits") == k
    grp.close()
