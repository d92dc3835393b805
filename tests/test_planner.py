"""CPU: invariants of the tile planner (csrc/engine.cu build_plan / build_balanced) at sizes no test can execute --
BASELINE configs 3 and 5 included -- through the device-free planner dump (qaoa_debug_plan).

A plan is correct when, in launch order, every qubit is mixed exactly once per layer, layer d's phase separator comes after
ALL of layer d-1's butterflies and before ALL of layer d's, the first pass generates the state, the last one reduces
it, and every tile shape is addressable (14 distinct bits, free bits in at most two runs).  For sharded symmetric
registers qubit 0 must be mixed in cross passes only (twisted coordinates, csrc/tile_kernel.cuh)."""
import itertools

import pytest

from qaoa_b200 import engine as eng

LOAD, STORE, INIT, PHASE, ROUND, WT, BT, REDUCE, XT, XTI = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10


def _reg_tile_bit(frame, after_wt, slot):
    """tile-local bit held by register slot `slot` (frame table of csrc/tile_kernel.cuh)"""
    if after_wt:
        return 1 + slot
    if frame == 2:
        return 5 + slot
    if frame == 3:
        return 9 + slot
    if frame == 4:
        return 4 + slot
    return 0 if slot == 0 else (9 if frame == 0 else 5) + slot


def _plan(**kw):
    try:
        return eng.debug_plan(**kw)
    except eng.EngineError:
        return None


def check_plan(pl, n, p, world, symmetric, dtype):
    g = world.bit_length() - 1
    qs, E, el = pl["qshift"], pl["E"], pl["el"]
    assert E == n + qs
    qubits = list(range(qs, E))
    twisted = bool(symmetric and world > 1)
    # ---- shapes
    for i, sh in enumerate(pl["shapes"]):
        pos = sh["pos"]
        assert len(set(pos)) == 14 and pos == sorted(pos), pos
        assert pos[0] == 0, "every tile carries the lowest element bits (coalescing)"
        local = [b for b in pos if b < el]
        if sh["zshape"]:
            assert i == 0 and pos[13] == E - 1 and local == list(range(13))
        elif sh["cross"]:
            assert [b for b in pos if b >= el] == list(range(el, el + g)), "a cross tile holds all rank bits"
            assert len(local) == 14 - g
        else:
            assert len(local) == 14
        free = [b for b in sh["free"] if b < el]
        assert sorted(free + local) == list(range(el)) or sh["zshape"]
        runs = 1 + sum(1 for a, b in zip(free, free[1:]) if b != a + 1) if free else 0
        assert runs <= 2, ("free bits need more than two runs", free)
    assert sum(1 for sh in pl["shapes"] if sh["cross"]) == (1 if world > 1 else 0)
    assert bool(pl["z2"]) == bool(symmetric and n % 2 == 0 and (world == 1 or dtype == "complex64"))
    # ---- events in launch order
    rounds, phases = [], {}
    for ip, ps in enumerate(pl["passes"]):
        sh = pl["shapes"][ps["shape"]]
        assert ps["cross"] == sh["cross"]
        types = [op[0] for op in ps["ops"]]
        assert types[0] == (INIT if ip == 0 else LOAD)
        if ip == len(pl["passes"]) - 1:
            assert types[-1] == REDUCE and ps["reduce"] and STORE not in types
        else:
            assert types[-1] == STORE and not ps["reduce"]
        frame, wt = ps["ops"][0][1], False
        for io, op in enumerate(ps["ops"]):
            if op[0] == WT:
                wt = True
            elif op[0] == BT:
                frame, wt = frame ^ 1, False
            elif op[0] == XT:
                frame, wt = (2 if frame == 0 else 4), False
            elif op[0] == XTI:
                frame, wt = (0 if frame == 2 else 3), False
            elif op[0] in (STORE, PHASE, REDUCE):
                assert op[1] == frame and not wt, ("step expects another frame than the registers are in", op, frame)
            if op[0] == ROUND:
                r = ps["rounds"][op[2]]
                for slot, eb in enumerate(r["ebit"]):
                    if eb >= 0:
                        assert (op[3] >> slot) & 1 or (twisted and ps["cross"] and slot == 0), "slot outside the round's mask"
                        assert eb == sh["pos"][_reg_tile_bit(frame, wt, slot)], "qubit is not in this register slot here"
                        rounds.append(((ip, io), r["layer"], eb, ps))
                        if twisted and eb == qs:
                            assert ps["cross"] and slot == 0, "twisted shards mix qubit 0 in the cross passes (slot 0)"
            elif op[0] == PHASE:
                layer = ps["phases"][op[2]]
                assert layer not in phases, "a layer's phase separator is applied once"
                phases[layer] = (ip, io)
    assert sorted(phases) == list(range(p))
    for d in range(p):
        mixed = sorted(eb for (_k, layer, eb, _ps) in rounds if layer == d)
        assert mixed == qubits, (d, "every qubit exactly once per layer", mixed)
        for key, layer, _eb, _ps in rounds:
            if layer == d:
                assert key > phases[d], ("butterfly of layer %d before its phase separator" % d, key, phases[d])
            if layer == d - 1:
                assert key < phases[d], ("butterfly of layer %d after the next phase separator" % (d - 1), key, phases[d])
    check_chunks(pl, world)
    return len(pl["passes"]), sum(1 for ps in pl["passes"] if ps["cross"])


def check_chunks(pl, world):
    """Chunk pipeline of sharded registers (engine.cu plan_chunks, sharded.py pipeline_segments): the chunk bits are mixed
    by exactly one local shape; every pass flagged closed tiles over a shape that has all of them as FREE bits (so a chunk
    is a set of whole tiles of that shape, and the pass never couples two chunks -- the mirror pairing of the symmetric
    low tile complements all chunk bits, leaving their XORs alone); in the cross shape they lie below the bits that deal
    the tiles out over the ranks; the driver's segments replay every pass exactly once, in launch order."""
    from qaoa_b200.sharded import pipeline_segments

    ch = pl["chunks"]
    g = world.bit_length() - 1
    flags = [ps["cross"] for ps in pl["passes"]]
    closed = [ps["closed"] for ps in pl["passes"]]
    if ch["m"] == 0:
        assert not any(closed)
    else:
        assert world > 1 and ch["m"] in (1, 2, 3)
        bits = [ch["u"]] + ch["v"][: ch["m"]]
        assert len(set(bits)) == len(bits) and all(pl["qshift"] <= b < pl["el"] for b in bits)
        assert ch["v"][: ch["m"]] == sorted(ch["v"][: ch["m"]])
        open_shape = pl["shapes"][ch["open_shape"]]
        assert not open_shape["cross"] and not open_shape["zshape"] and all(b in open_shape["pos"] for b in bits)
        for ps in pl["passes"]:
            sh = pl["shapes"][ps["shape"]]
            if not ps["closed"]:
                continue
            assert ps["shape"] != ch["open_shape"] and ps["prog"] != "Generic"
            assert all(b in sh["free"] for b in bits), "chunk bit is not a free bit of a closed pass' shape"
            assert not any(eb in bits for r in ps["rounds"] for eb in r["ebit"])
            if sh["cross"]:
                assert ps["prog"] in ("BoundX", "FinalX")
                assert all(sh["free"].index(b) < len(sh["free"]) - g for b in bits), "chunk bit deals cross tiles out over the ranks"
    segs = pipeline_segments(flags, closed)
    order = []
    for seg in segs:
        if seg[0] == "pass":
            order.append(seg[1])
        else:
            _, pre, x, post = seg
            assert flags[x] and closed[x] and all(closed[i] and not flags[i] for i in pre + post)
            order += pre + [x] + post
    assert order == list(range(len(flags))), "segments must replay the passes once, in launch order"


def test_baseline_configurations():
    # config 3: n = 32, p = 5 on one GPU: 11 launches, symmetric half register
    pl = eng.debug_plan(32, p=5)
    assert check_plan(pl, 32, 5, 1, True, "complex64") == (11, 0)
    assert pl["z2"] and [ps["prog"] for ps in pl["passes"]] == ["FirstInit", "Interior"] + ["Bound2", "Mid"] * 4 + ["Final2"]
    pl = eng.debug_plan(32, p=5, symmetric=False)
    assert check_plan(pl, 32, 5, 1, False, "complex64") == (11, 0) and not pl["z2"]
    # config 5: n = 36, p = 3 on 8 and 4 GPUs, symmetric register in twisted coordinates, spare slots of the cross tile
    for world, kx in ((8, 1), (4, 2)):
        for rank in (0, world - 1):
            pl = eng.debug_plan(36, p=3, world=world, rank=rank)
            assert check_plan(pl, 36, 3, world, True, "complex64") == (10, 2)
            assert pl["kx"] == kx and pl["el"] == 35 - (world.bit_length() - 1)
            progs = [ps["prog"] for ps in pl["passes"]]
            assert progs == ["FirstInit", "Interior", "Interior", "BoundX", "Mid", "Interior", "Bound2", "Mid", "Interior", "FinalX"]
            assert all(sh["pos"][4] == 4 for sh in pl["shapes"]), "256-byte rows everywhere"
    pl = eng.debug_plan(36, p=3, world=8, symmetric=False)
    assert check_plan(pl, 36, 3, 8, False, "complex64")[1] == 2 and pl["kx"] == 1
    pl = eng.debug_plan(36, p=3, world=8, cross_spare=False)
    assert pl["kx"] == 0 and "Bound2L4" in [ps["prog"] for ps in pl["passes"]]
    assert eng.debug_plan(34, p=3, world=2)["kx"] == 1
    # chunk pipeline: four chunks at every BASELINE multi-GPU configuration, windows around every cross pass
    from qaoa_b200.sharded import pipeline_segments
    for n, p, world in ((32, 5, 2), (32, 5, 4), (32, 5, 8), (36, 3, 8), (36, 3, 4), (34, 3, 2)):
        pl = eng.debug_plan(n, p=p, world=world)
        assert pl["chunks"]["m"] in (2, 3)
        segs = pipeline_segments([ps["cross"] for ps in pl["passes"]], [ps["closed"] for ps in pl["passes"]])
        wins = [sg for sg in segs if sg[0] == "window"]
        assert len(wins) == (p + 1) // 2 and all(len(w[1]) >= 1 for w in wins), segs
        assert all(len(w[3]) == 2 for w in wins[:-1]) and wins[-1][3] == []


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("symmetric", [True, False])
def test_planner_invariants_over_sizes(world, symmetric):
    checked = 0
    for n, p, dtype in itertools.product(range(15, 41), (1, 2, 3, 4, 5, 6), ("complex64", "complex128")):
        if n > 39 or (dtype == "complex128" and n > 36):
            continue
        pl = _plan(n_qubits=n, dtype=dtype, world=world, rank=world - 1, symmetric=symmetric, p=p,
                   mixer=eng.MIXER_XMULTI if (n + p) % 3 == 0 else eng.MIXER_X)
        if pl is None:       # configuration the engine rejects (shard too small, odd n / complex128 on a symmetric shard)
            g = world.bit_length() - 1
            shard_bits = n - g - (1 if symmetric and world > 1 else 0) + (1 if dtype == "complex128" else 0)
            assert shard_bits < 15 or (symmetric and world > 1 and (n % 2 or dtype == "complex128")), (n, p, dtype)
            continue
        npass, ncross = check_plan(pl, n, p, world, symmetric, dtype)
        checked += 1
        if world > 1 and dtype == "complex64":
            assert ncross == (p + 1) // 2, "one cross pass per two layers (balanced schedule)"
        if dtype == "complex64" and len(pl["shapes"]) >= 3:
            assert all(ps["prog"] != "Generic" for ps in pl["passes"]), (n, p, [ps["prog"] for ps in pl["passes"]])
    assert checked > 40


@pytest.mark.parametrize("schedule,low_bits", [(1, 0), (2, 4), (2, 5), (1, 4)])
def test_planner_invariants_with_forced_options(schedule, low_bits):
    for n, p, world, sym in itertools.product((20, 24, 27, 30, 33), (1, 2, 3, 4), (1, 2, 8), (True, False)):
        if sym and world > 1 and schedule == 1:
            continue            # sharded symmetric registers run the balanced schedule only
        pl = _plan(n_qubits=n, world=world, rank=0, symmetric=sym, p=p, schedule=schedule, low_bits=low_bits)
        if pl is not None:
            check_plan(pl, n, p, world, sym, "complex64")
