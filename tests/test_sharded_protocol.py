"""CPU, world_size 2 (gloo): the host-side protocol of the sharded register (qaoa_b200/sharded.py).

DistShardedEngine is run over a stand-in for the CUDA engine that keeps each rank's shard in POSIX
shared memory -- the CPU analogue of the CUDA-IPC peer mapping -- and implements the same pass protocol
(shard_begin / shard_pass_is_cross / shard_run_pass / shard_finish) in numpy: local passes touch the own
shard only, cross passes read and write the peers' shards.  The stand-in checks, at the start of every
pass, that the barriers the protocol requires have actually happened (rank 1 is slowed down so that a
missing barrier is caught), and the combined result must equal the oracle's energy."""
import os
import socket
import time
from multiprocessing import shared_memory

import numpy as np
import pytest

from oracle import qaoa_oracle as orc
import parity_helpers as H


class ShmShardEngine:
    """numpy stand-in with the shard_* surface of qaoa_b200.engine.Engine (complex128)."""

    def __init__(self, n, rank, world, diag, slow=0.0):
        self.n, self.rank, self.world = n, rank, world
        self.g = world.bit_length() - 1
        self.nl = n - self.g
        self.n_beta = 1
        self.diag = np.asarray(diag, dtype=np.float64)
        self.slow = slow
        # shared segment: [done_pass (int64), pad, shard (2^nl complex128)]
        self.shm = shared_memory.SharedMemory(create=True, size=64 + (16 << self.nl))
        self.peers = {rank: self.shm}
        self._view(rank)[0][0] = -1

    def _view(self, r):
        buf = self.peers[r].buf
        return np.ndarray((1,), dtype=np.int64, buffer=buf, offset=0), \
            np.ndarray((1 << self.nl,), dtype=np.complex128, buffer=buf, offset=64)

    # -- peer mapping ---------------------------------------------------------------------------
    def shard_export(self, which):
        return self.shm.name.encode().ljust(64, b"\0") if which == 0 else b"\0" * 64

    def shard_import(self, which, handles):
        if which != 0:
            return
        for r, hd in enumerate(handles):
            if r != self.rank:
                self.peers[r] = shared_memory.SharedMemory(name=hd.rstrip(b"\0").decode())

    def shard_sync(self):
        pass

    def set_cost(self):
        pass

    # -- pass protocol ----------------------------------------------------------------------------
    def shard_begin(self, angles):
        self.ang = np.asarray(angles, dtype=np.float64)
        p = self.ang.size // 2
        self.passes = []
        for d in range(p):
            self.passes += [("local", d), ("cross", d)]
        self._view(self.rank)[0][0] = -1
        return len(self.passes)

    def shard_pass_is_cross(self, i):
        return self.passes[i][0] == "cross"

    def _check_barrier(self, i):
        from qaoa_b200.sharded import needs_barrier

        flags = [k == "cross" for k, _ in self.passes]
        if needs_barrier(flags, i):
            for r in self.peers:
                done = int(self._view(r)[0][0])
                assert done >= i - 1, "rank %d started pass %d while rank %d had only finished pass %d" % (
                    self.rank, i, r, done)

    def shard_run_pass(self, i):
        kind, d = self.passes[i]
        self._check_barrier(i)
        if self.slow:
            time.sleep(self.slow)
        gamma, beta = self.ang[2 * d], self.ang[2 * d + 1]
        c, s = np.cos(beta), np.sin(beta)
        _, mine = self._view(self.rank)
        if kind == "local":
            if d == 0:
                mine[:] = 2.0 ** (-0.5 * self.n)
            lo = self.rank << self.nl
            mine *= np.exp(1j * gamma * self.diag[lo: lo + (1 << self.nl)])
            psi = mine.reshape([2] * self.nl)
            for q in range(self.nl):   # RX(-2 beta) = [[c, i s], [i s, c]] on local qubit q (axis nl-1-q)
                ax = self.nl - 1 - q
                a, b = np.take(psi, 0, axis=ax).copy(), np.take(psi, 1, axis=ax).copy()
                idx0 = [slice(None)] * self.nl
                idx1 = [slice(None)] * self.nl
                idx0[ax], idx1[ax] = 0, 1
                psi[tuple(idx0)] = c * a + 1j * s * b
                psi[tuple(idx1)] = 1j * s * a + c * b
        else:
            # this rank's share of the cross tiles: the slice of LOCAL indices [rank/world, (rank+1)/world)
            # of EVERY shard; the global qubits are the rank index bits
            w = (1 << self.nl) // self.world
            sl = slice(self.rank * w, (self.rank + 1) * w)
            shards = [self._view(r)[1] for r in range(self.world)]
            block = np.stack([sh[sl] for sh in shards])          # (world, w): axis 0 = global qubits
            block = block.reshape([2] * self.g + [w])
            for q in range(self.g):
                ax = self.g - 1 - q
                a, b = np.take(block, 0, axis=ax).copy(), np.take(block, 1, axis=ax).copy()
                idx0 = [slice(None)] * (self.g + 1)
                idx1 = [slice(None)] * (self.g + 1)
                idx0[ax], idx1[ax] = 0, 1
                block[tuple(idx0)] = c * a + 1j * s * b
                block[tuple(idx1)] = 1j * s * a + c * b
            block = block.reshape(self.world, w)
            if i == len(self.passes) - 1:
                # like the engine's FINAL pass: reduce the tiles this rank processed, store nothing
                d = np.stack([self.diag[(r << self.nl) + sl.start: (r << self.nl) + sl.stop] for r in range(self.world)])
                pr = np.abs(block) ** 2
                sup = pr > 0
                self.partial = np.array([pr.sum(), (pr * d).sum(), (pr * d * d).sum(), d[sup].min(), d[sup].max()])
            else:
                for r in range(self.world):
                    shards[r][sl] = block[r]
        self._view(self.rank)[0][0] = i

    def shard_finish(self):
        return self.partial

    def close(self):
        for r, sh in self.peers.items():
            sh.close()
        self.shm.unlink()


class TwistedShmShardEngine(ShmShardEngine):
    """Stand-in for a SHARDED SYMMETRIC register (csrc/tile_kernel.cuh "twisted coordinates"): rank r
    holds phi~(r, z), z = the n-1-g low bits of the stored index x', r_j = x'[n-1-g+j] XOR x'[0], in the S frame.  Local
    passes: phase separator, the local qubits 1.., the mirror pairing of the virtual top qubit (same shard!).  Cross
    passes: qubit 0 -- the pairing (z_0 = 0, r) <-> (z_0 = 1, ~r) -- and the rank qubits with their sign twist, on this
    rank's slice of every shard.  Same algebra as tests/test_twisted_coordinates.py, spread over processes."""

    def __init__(self, n, rank, world, diag, slow=0.0):
        super().__init__(n, rank, world, diag, slow)
        self.nl = n - 1 - self.g                      # stored local qubits (the segment is larger than needed)
        z = np.arange(1 << self.nl)
        mask = world - 1
        self.z = z
        self.x = [z | (np.where(z & 1, r ^ mask, r) << self.nl) for r in range(world)]   # x'(r, z)
        pc = np.array([bin(int(v)).count("1") for v in self.x[rank]])
        self.pc = pc
        self.cost = [self.diag[x] for x in self.x]
        self.sigma = np.real((1j) ** ((2 - n) % 4)) * (-1.0) ** pc

    def shard_run_pass(self, i):
        kind, d = self.passes[i]
        self._check_barrier(i)
        if self.slow:
            time.sleep(self.slow)
        gamma, beta = self.ang[2 * d], self.ang[2 * d + 1]
        c, s = np.cos(beta), np.sin(beta)
        z, el, mask = self.z, self.nl, self.world - 1
        _, mine = self._view(self.rank)
        if kind == "local":
            if d == 0:
                mine[:] = np.sqrt(2.0) * 2.0 ** (-0.5 * self.n) * (1j) ** (self.pc % 4)
            mine *= np.exp(1j * gamma * self.cost[self.rank])
            for q in range(1, el):
                lo = z[(z >> q) & 1 == 0]
                a, b = mine[lo].copy(), mine[lo | (1 << q)].copy()
                mine[lo], mine[lo | (1 << q)] = c * a + s * b, c * b - s * a
            mine[:] = c * mine + s * self.sigma * mine[z ^ ((1 << el) - 1)]
        else:
            w = (1 << (el - 1)) // self.world            # this rank's share of the (z >> 1) values, both z_0
            rest = np.arange(self.rank * w, (self.rank + 1) * w)
            ev, od = rest << 1, (rest << 1) | 1
            shards = [self._view(r)[1] for r in range(self.world)]
            blk = {(r, z0): shards[r][ev if z0 == 0 else od].copy() for r in range(self.world) for z0 in (0, 1)}
            new = dict(blk)
            for r in range(self.world):                   # qubit 0
                a, b = blk[(r, 0)], blk[(r ^ mask, 1)]
                new[(r, 0)], new[(r ^ mask, 1)] = c * a + s * b, c * b - s * a
            blk = new
            for j in range(self.g):                       # rank qubits: the qubit is 0 where r_j == z_0
                new = dict(blk)
                for r in range(self.world):
                    if (r >> j) & 1:
                        continue
                    r1 = r | (1 << j)
                    for z0, (ra, rb) in ((0, (r, r1)), (1, (r1, r))):
                        a, b = blk[(ra, z0)], blk[(rb, z0)]
                        new[(ra, z0)], new[(rb, z0)] = c * a + s * b, c * b - s * a
                blk = new
            if i == len(self.passes) - 1:
                tot = np.zeros(3)
                lo, hi = np.inf, -np.inf
                for (r, z0), v in blk.items():
                    cst = self.cost[r][ev if z0 == 0 else od]
                    pr = np.abs(v) ** 2
                    tot += [pr.sum(), (pr * cst).sum(), (pr * cst * cst).sum()]
                    sup = pr > 0
                    lo, hi = min(lo, cst[sup].min()), max(hi, cst[sup].max())
                self.partial = np.array([tot[0], tot[1], tot[2], lo, hi])
            else:
                for (r, z0), v in blk.items():
                    shards[r][ev if z0 == 0 else od] = v
        self._view(self.rank)[0][0] = i


def _twisted_worker(rank, world, port, n, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qaoa_b200 import sharded

        edges, colors, table = H.regular_maxcut(n)
        diag = H.maxcut_diag(n, edges, colors, table)
        eng = sharded.DistShardedEngine(
            n, symmetric=True,
            engine_factory=lambda r, w: TwistedShmShardEngine(n, r, w, diag, slow=0.03 if r == 1 else 0.0))
        assert eng.symmetric
        ang = np.array([0.3, 0.2, 0.5, 1.1, 1.4, -0.7])
        got = eng.energy(ang)
        ref = np.array(orc.energy(orc.Spec(n, diag), ang))
        np.save(os.path.join(out_dir, "t%d.npy" % rank), np.concatenate([got, ref]))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_twisted_symmetric_shards_over_processes_gloo(world, tmp_path):
    """the sharded symmetric register's data flow under torch.distributed (gloo): the mirror pairing never leaves a
    shard, qubit 0 and the rank qubits are mixed in the cross passes, every rank ends with the oracle's energy"""
    import torch.multiprocessing as mp

    n = 10
    mp.spawn(_twisted_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        row = np.load(tmp_path / ("t%d.npy" % r))
        got, ref = row[:5], row[5:9]
        assert np.allclose(got[:4], ref, rtol=1e-12), (r, got, ref)
        assert abs(got[4] - 1.0) < 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qaoa_b200 import sharded

        edges, colors, table = H.regular_maxcut(n)
        diag = H.maxcut_diag(n, edges, colors, table)
        eng = sharded.DistShardedEngine(
            n, engine_factory=lambda r, w: ShmShardEngine(n, r, w, diag, slow=0.05 if r == 1 else 0.0))
        ang = np.array([0.3, 0.2, 0.5, 0.1, 1.4, -0.7])
        status, got = "ok", None
        try:
            got = eng.energy(ang)
        except AssertionError as exc:
            status = "violation: %s" % exc
        ref = np.array(orc.energy(orc.Spec(n, diag), ang))
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([[1.0 if status == "ok" else 0.0],
                                                                          got if got is not None else np.zeros(5), ref]))
        assert eng.barriers >= 2 * (ang.size // 2)
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_dist_sharded_protocol_gloo(world, tmp_path):
    import torch.multiprocessing as mp

    n = 10
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        row = np.load(tmp_path / ("r%d.npy" % r))
        assert row[0] == 1.0
        got, ref = row[1:6], row[6:10]
        assert np.allclose(got[:4], ref, rtol=1e-12), (r, got, ref)
        assert abs(got[4] - 1.0) < 1e-12


def test_combine_partials_and_barrier_rule():
    from qaoa_b200.sharded import combine_partials, needs_barrier

    parts = [[0.25, 1.0, 5.0, 1.0, 4.0], [0.75, 2.0, 7.0, 0.0, 3.0]]
    out = combine_partials(parts)
    assert np.allclose(out, [3.0, 12.0, 0.0, 4.0, 1.0])
    flags = [False, False, True, False, False, True]
    assert [needs_barrier(flags, i) for i in range(6)] == [False, False, True, True, False, True]


def _split_worker(rank, world, port, out_dir):
    import networkx as nx
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qaoa_b200 import QAOA, initialstates, mixers, problems
        from qaoa_b200.qaoa import B200Backend
        from test_host_api import OracleEngine

        G = nx.random_regular_graph(3, 8, seed=42)
        prob = problems.MaxCut(G)
        q = QAOA(problem=prob, mixer=mixers.X(), initialstate=initialstates.Plus(),
                 backend=B200Backend(split_batches=True))
        fake = OracleEngine(prob.N_qubits)
        prob.lower_cost(fake)
        q.initialstate.lower_initial(fake)
        q.mixer.lower_mixer(fake, trotter=1)
        q._engine = fake
        q.sample_cost_landscape({"gamma": [0, 2 * np.pi, 5], "beta": [0, 2 * np.pi, 3]})
        np.save(os.path.join(out_dir, "land%d.npy" % rank), np.concatenate([q.Exp_sampled_p1.ravel(), [fake.calls]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_landscape_split_across_ranks_gloo(tmp_path):
    """split_batches: the 15 landscape points are dealt out over 2 ranks (8 + 7 evaluations), gathered back in grid
    order, and both ranks see the landscape a single process computes."""
    import networkx as nx
    import torch.multiprocessing as mp

    from qaoa_b200 import initialstates, mixers, problems
    from test_host_api import make

    mp.spawn(_split_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ref, _ = make(problems.MaxCut(nx.random_regular_graph(3, 8, seed=42)), mixer=mixers.X(), init=initialstates.Plus())
    ref.sample_cost_landscape({"gamma": [0, 2 * np.pi, 5], "beta": [0, 2 * np.pi, 3]})
    calls = []
    for r in range(2):
        row = np.load(tmp_path / ("land%d.npy" % r))
        assert np.allclose(row[:-1], ref.Exp_sampled_p1.ravel(), rtol=1e-13)
        calls.append(int(row[-1]))
    assert sorted(calls) == [7, 8]


class GroverShmEngine(ShmShardEngine):
    """Stand-in for the sharded Grover protocol (Engine.shard_grover_step): every rank sweeps its own shard, the only
    thing that crosses the ranks is the partial dot product <s|v> of every layer (all-reduced by DistShardedEngine)."""

    def __init__(self, n, rank, world, diag, s_state):
        super().__init__(n, rank, world, diag)
        lo = rank << self.nl
        self.s_loc = np.asarray(s_state, dtype=np.complex128)[lo: lo + (1 << self.nl)]
        self.c_loc = self.diag[lo: lo + (1 << self.nl)]
        self.v = None
        self.steps = []

    def set_mixer(self, kind, *a, **k):
        self.n_beta = 1

    def shard_grover_step(self, step, angles, dot_global=None):
        ang = np.asarray(angles, dtype=np.float64)
        p = ang.size // 2
        self.steps.append(step)
        if step == 0:
            assert dot_global is None
            self.v = self.s_loc.copy()
        else:
            dot = complex(dot_global[0], dot_global[1])
            self.v = self.v + (np.exp(-1j * ang[2 * (step - 1) + 1]) - 1.0) * dot * self.s_loc
        if step < p:
            self.v = self.v * np.exp(1j * ang[2 * step] * self.c_loc)
            d = np.vdot(self.s_loc, self.v)
            return np.array([d.real, d.imag])
        pr = np.abs(self.v) ** 2
        sup = pr > 0
        lo, hi = (self.c_loc[sup].min(), self.c_loc[sup].max()) if sup.any() else (1e300, -1e300)
        return np.array([pr.sum(), (pr * self.c_loc).sum(), (pr * self.c_loc ** 2).sum(), lo, hi])


def _grover_worker(rank, world, port, n, k, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qaoa_b200 import engine as eng_mod
        from qaoa_b200 import sharded

        Q, c, b = H.random_qubo(n, seed=4)
        diag = orc.qubo_diag(Q, c, b)
        s_state = orc.dicke_state(n, k)
        eng = sharded.DistShardedEngine(n, engine_factory=lambda r, w: GroverShmEngine(n, r, w, diag, s_state))
        eng.set_mixer(eng_mod.MIXER_GROVER, sub_kind=eng_mod.INIT_DICKE, sub_k=k)
        ang = np.array([0.03, 0.7, 0.05, 1.9, 0.02, -0.4])
        got = eng.energy(ang)
        assert eng.eng.steps == [0, 1, 2, 3]
        ref = np.array(orc.energy(orc.Spec(n, diag, mixer=("grover", s_state), init=s_state), ang))
        np.save(os.path.join(out_dir, "g%d.npy" % rank), np.concatenate([got, ref]))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_grover_protocol_gloo(world, tmp_path):
    """Grover mixer on a sharded register under torch.distributed (gloo): step d's local sweep, ONE all-reduced complex
    scalar per layer, the statistics combined at the end -- every rank ends with the oracle's energy"""
    import torch.multiprocessing as mp

    n, k = 10, 4
    mp.spawn(_grover_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        row = np.load(tmp_path / ("g%d.npy" % r))
        got, ref = row[:5], row[5:9]
        assert np.allclose(got[:4], ref, rtol=1e-12), (r, got, ref)
        assert abs(got[4] - 1.0) < 1e-12


class SamplingShmEngine(GroverShmEngine):
    """adds the local halves of sampling / CVaR (Engine.shard_chunk_probs / shard_locate / shard_cvar_digit) in numpy:
    chunks of 16 amplitudes, F64 cost keys (sign-flipped bit pattern, 8 digits)"""

    CH = 4

    def __init__(self, n, rank, world, diag, s_state):
        super().__init__(n, rank, world, diag, s_state)
        self.local_qubits = self.nl
        self.options = {}

    def set_option(self, key, value):
        self.options[key] = value

    def counter(self, key):
        return {"diag_kind": 3.0, "diag_c0": 0.0, "diag_delta": 1.0}[key]

    def _probs(self):
        assert self.options.get("keep_state") == 1, "sampling / CVaR need the kept final state"
        return np.abs(self.v) ** 2

    def shard_chunk_probs(self):
        return self._probs().reshape(-1, 1 << self.CH).sum(axis=1)

    def shard_locate(self, chunk_of, resid):
        p = self._probs().reshape(-1, 1 << self.CH)
        out = np.zeros(len(chunk_of), dtype=np.uint64)
        for i, (c, r) in enumerate(zip(chunk_of, resid)):
            cum = np.cumsum(p[int(c)])
            out[i] = (int(c) << self.CH) + min(int(np.searchsorted(cum, r, side="right")), (1 << self.CH) - 1)
        return out

    def shard_cvar_key_bits(self):
        return 64

    def shard_cvar_digit(self, prefix, shift, check):
        bits = self.c_loc.view(np.uint64)
        keys = np.where(bits >> np.uint64(63), ~bits, bits | np.uint64(1 << 63))
        p = self._probs()
        sel = ((keys >> np.uint64(shift)) >> np.uint64(8)) == np.uint64(prefix) if check else np.ones(len(keys), dtype=bool)
        d = ((keys >> np.uint64(shift)) & np.uint64(255)).astype(int)[sel]
        h = np.zeros(512)
        np.add.at(h, d, p[sel])
        np.add.at(h, 256 + d, (p * self.c_loc)[sel])
        return h


def _sampling_worker(rank, world, port, n, k, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qaoa_b200 import engine as eng_mod
        from qaoa_b200 import sharded

        Q, c, b = H.random_qubo(n, seed=4)
        diag = np.ascontiguousarray(orc.qubo_diag(Q, c, b))
        s_state = orc.dicke_state(n, k)
        eng = sharded.DistShardedEngine(n, engine_factory=lambda r, w: SamplingShmEngine(n, r, w, diag, s_state))
        eng.set_mixer(eng_mod.MIXER_GROVER, sub_kind=eng_mod.INIT_DICKE, sub_k=k)
        ang = np.array([0.03, 0.7, 0.05, 1.9])
        idx = eng.sample(ang, 512, seed=3 + rank)        # rank 0's uniform numbers are broadcast: identical everywhere
        cv = eng.cvar(ang, 0.3)
        assert eng.eng.options.get("keep_state") == 0       # kept states are switched off again
        psi = orc.evolve(orc.Spec(n, diag, mixer=("grover", s_state), init=s_state), ang)
        want = orc.exact_cvar(diag, np.abs(psi) ** 2, 0.3)
        np.save(os.path.join(out_dir, "s%d.npy" % rank), np.concatenate([idx.astype(np.float64), [cv, want]]))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_sampling_and_cvar_protocol_gloo(world, tmp_path):
    """hist() / cvar < 1 on a sharded register under torch.distributed (gloo): broadcast uniform numbers, gathered chunk
    masses, shots located on the owning rank and summed back, all-reduced digit histograms of the quantile select"""
    import torch.multiprocessing as mp

    n, k = 10, 4
    mp.spawn(_sampling_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / ("s%d.npy" % r)) for r in range(world)]
    for row in rows:
        assert np.array_equal(row[:512], rows[0][:512])                 # the same shots on every rank
        assert abs(row[512] - row[513]) <= 1e-10 * max(1.0, abs(row[513]))   # CVaR = the oracle's
    idx = rows[0][:512].astype(np.int64)
    assert all(bin(int(x)).count("1") == k for x in idx)                # Grover over Dicke: samples stay feasible
