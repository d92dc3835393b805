"""One QAOA register sharded over several B200s (include/qaoa_b200.h "Sharded registers").

The register is split on its top log2(world) qubits, one shard per GPU.  Mixer gates on those
"global" qubits are executed by CROSS passes -- ordinary tile passes of the engine whose tile
contains the rank bits, so their loads and stores travel over NVLink to the peers' HBM.  Nothing is
swapped or staged: the exchange is fused into the pass.  This module is the host-side plumbing:

* :class:`DistShardedEngine` -- one process per GPU under ``torch.distributed`` (NCCL or gloo): CUDA IPC
  handles are exchanged with ``all_gather_object``, passes that touch peer memory are bracketed by
  ``dist.barrier()``, the five partial sums are combined with a scalar all-reduce.
* :class:`LocalShardGroup` -- every rank's engine inside ONE process (one or several GPUs); used to
  check the sharded kernels on a single GPU: passes of different ranks only ever run back to back.

Both expose the :class:`qaoa_b200.engine.Engine` surface the problem / mixer / initial-state lowering
uses (``set_cost_*``, ``set_mixer``, ``set_initial``, ``energy``, ``energy_batch``), so ``QAOA`` can sit
on top of either.  The reference has no multi-device path.
"""
from __future__ import annotations

import numpy as np

from qaoa_b200 import engine as _eng


def default_device():
    """The CUDA device of this process: LOCAL_RANK under a launcher (torchrun exports it), else torch's current device.
    One process per GPU -- ranks of one node must never share a device (all shards would sit in one GPU's HBM and the
    'NVLink' cross passes would never leave it)."""
    import os

    lr = os.environ.get("LOCAL_RANK")
    if lr is not None:
        return int(lr)
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:   # noqa: BLE001
        pass
    return 0


def combine_partials(parts):
    """parts: (world, 5) rows {sum p, sum p c, sum p c^2, min c, max c} -> (E, E2, min, max, norm)."""
    parts = np.asarray(parts, dtype=np.float64).reshape(-1, 5)
    s0, s1, s2 = parts[:, 0].sum(), parts[:, 1].sum(), parts[:, 2].sum()
    return np.array([s1 / s0, s2 / s0, parts[:, 3].min(), parts[:, 4].max(), s0])


def needs_barrier(cross_flags, i):
    """A barrier across ranks separates pass i from pass i-1 whenever either of them touches peer memory."""
    return bool(cross_flags[i] or (i > 0 and cross_flags[i - 1]))


def pipeline_segments(cross_flags, closed):
    """Splits the pass list of a sharded evaluation into segments for the chunk pipeline: ("pass", i) = the whole pass,
    ("window", pre, x, post) = cross pass x with the chunk-closed local passes right before and after it.  Inside a
    window chunk c runs pre -> x -> post independently of the other chunks (engine.cu plan_chunks), so the cross pass of
    one chunk can overlap the local passes of another."""
    n = len(cross_flags)
    segs, i = [], 0
    while i < n:
        # a window starts at i if the closed local passes from i on lead straight to a closed cross pass
        j = i
        while j < n and closed[j] and not cross_flags[j]:
            j += 1
        if j < n and cross_flags[j] and closed[j]:
            pre, x = list(range(i, j)), j
            k = x + 1
            while k < n and closed[k] and not cross_flags[k]:
                k += 1
            post = list(range(x + 1, k))
            # leave the closed passes that lead to the NEXT cross pass to that window if they are all there is between
            if k < n and cross_flags[k] and closed[k]:
                post = []
                k = x + 1
            segs.append(("window", pre, x, post))
            i = k
        elif j > i and not (j < n and cross_flags[j]):
            segs.extend(("pass", q) for q in range(i, j))
            i = j
        else:
            segs.append(("pass", i))
            i += 1
    return segs


def cvar_select(alpha, key_bits, digit_hist, decode):
    """Weighted quantile select over an order-preserving integer key of the cost, most significant 8-bit digit first
    (the host half of qaoa_cvar, csrc/engine.cu): digit_hist(prefix, shift, check) returns the GLOBAL 512-entry histogram
    {mass[256], mass x cost[256]} of that digit among the keys matching the prefix; decode(key) -> cost."""
    prefix, above_mass, above_pc, total, thr_mass = 0, 0.0, 0.0, None, 0.0
    for shift in range(key_bits - 8, -1, -8):
        check = shift != key_bits - 8
        hist = np.asarray(digit_hist(prefix, shift, check), dtype=np.float64)
        if not check:
            total = float(hist[:256].sum())
            if not total > 0.0:
                raise _eng.EngineError("state has zero norm")
        want = alpha * total
        am, ap, dstar = above_mass, above_pc, 0
        for d in range(255, -1, -1):
            if am + hist[d] >= want or d == 0:
                dstar = d
                break
            am += hist[d]
            ap += hist[256 + d]
        if hist[dstar] <= 0.0:
            nz = np.flatnonzero(hist[:256] > 0.0)
            if nz.size:
                dstar = int(nz[0])
        above_mass, above_pc, thr_mass = am, ap, float(hist[dstar])
        prefix = (prefix << 8) | dstar
    cthr = decode(prefix)
    take = min(max(alpha * total - above_mass, 0.0), thr_mass)
    return (above_pc + take * cthr) / (above_mass + take)


def decode_cost_key(kind, c0, delta):
    """key -> cost for the diagonal kinds of the engine (1: U8, 2: U32FX, 3: F64 sign-flipped bit pattern)"""
    if int(kind) == 3:
        def dec(key):
            bits = (key & 0x7FFFFFFFFFFFFFFF) if key >> 63 else (~key) & 0xFFFFFFFFFFFFFFFF
            return float(np.array([bits], dtype=np.uint64).view(np.float64)[0])
        return dec
    return lambda key: c0 + delta * float(key)


def deal_shots(u, chunk_sums):
    """u: uniform numbers in [0, 1); chunk_sums[r]: probability mass of every chunk of rank r's shard.  Returns per rank
    (shot indices, local chunk ids, residual mass inside the chunk) -- the host half of qaoa_sample over several shards."""
    sizes = [len(c) for c in chunk_sums]
    cum = np.cumsum(np.concatenate(chunk_sums))
    total = cum[-1]
    if not total > 0.0:
        raise _eng.EngineError("state has zero norm")
    t = np.clip(np.asarray(u, dtype=np.float64), 0.0, 1.0) * total
    c = np.minimum(np.searchsorted(cum, t, side="right"), len(cum) - 1)
    mass = np.diff(np.concatenate([[0.0], cum]))
    for i in range(len(c)):                   # never land in an empty chunk
        while c[i] > 0 and mass[c[i]] <= 0.0:
            c[i] -= 1
    resid = t - np.where(c > 0, cum[np.maximum(c - 1, 0)], 0.0)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    out = []
    for r in range(len(chunk_sums)):
        mine = np.flatnonzero((c >= starts[r]) & (c < starts[r + 1]))
        out.append((mine, (c[mine] - starts[r]).astype(np.uint32), resid[mine]))
    return out


class _ShardedSurface:
    """Engine-like methods shared by the two drivers (fan a call out to the local engine(s))."""

    n_beta = 1
    n_gamma = 1     # one gamma per layer: per-term gammas (MaxCutFree / MaxCutOrbit) are single-GPU only
    n_init = 0      # no parameterised initial states on sharded registers

    def _engines(self):
        raise NotImplementedError

    def _diag_changed(self):
        raise NotImplementedError

    def _before_diag_change(self):
        """Called before the engines free their cost diagonal (a peer may still have it mapped)."""

    def set_option(self, key, value):
        for e in self._engines():
            e.set_option(key, value)

    def set_cost_maxkcut(self, *a, **k):
        self._before_diag_change()
        for e in self._engines():
            e.set_cost_maxkcut(*a, **k)
        self._diag_changed()

    def set_cost_qubo(self, *a, **k):
        self._before_diag_change()
        for e in self._engines():
            e.set_cost_qubo(*a, **k)
        self._diag_changed()

    def set_cost_diagonal(self, diag):
        raise NotImplementedError("sharded engine: host-provided cost diagonals are not supported")

    def set_cost_maxkcut_terms(self, *a, **k):
        raise NotImplementedError("sharded engine: per-term gammas (MaxCutFree / MaxCutOrbit) are single-GPU only")

    def cost_minmax(self):
        """min / max of the cost over ALL basis states (every shard's device reduction, combined)."""
        lo, hi = zip(*[e.cost_minmax() for e in self._engines()])
        return self._combine_minmax(min(lo), max(hi))

    def _combine_minmax(self, lo, hi):
        return lo, hi

    mixer_kind = _eng.MIXER_X

    def set_mixer(self, kind, *a, **k):
        if kind == _eng.MIXER_GROVER and getattr(self, "symmetric", False):
            raise _eng.EngineError("sharded symmetric register: X-type mixers only (use symmetric=False for the Grover mixer)")
        for e in self._engines():
            e.set_mixer(kind, *a, **k)
        self.n_beta = self._engines()[0].n_beta
        self.mixer_kind = kind

    def set_initial(self, kind, *a, **k):
        for e in self._engines():
            e.set_initial(kind, *a, **k)

    def energy_batch(self, angles):
        angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
        return np.stack([self.energy(row) for row in angles])


class LocalShardGroup(_ShardedSurface):
    """All ``world`` ranks inside this process (``devices[r]`` = CUDA device of rank r, default all 0)."""

    def __init__(self, n_qubits, world, dtype="complex64", devices=None, engine_factory=None, symmetric=False):
        devices = [0] * world if devices is None else list(devices)
        make = engine_factory or (lambda r: _eng.Engine(n_qubits, dtype=dtype, device=devices[r], rank=r, world=world,
                                                        shard_symmetric=symmetric))
        self.symmetric = bool(symmetric)
        self.n_qubits, self.world = int(n_qubits), int(world)
        self.engines = [make(r) for r in range(world)]
        ptrs = [e.shard_local_ptr(0) for e in self.engines]
        for e in self.engines:
            e.shard_set_peers(0, ptrs)
        self.last_cross_flags = []

    def _engines(self):
        return self.engines

    def _diag_changed(self):
        ptrs = [e.shard_local_ptr(1) for e in self.engines]
        for e in self.engines:
            e.shard_set_peers(1, ptrs)

    def _sync_all(self):
        for e in self.engines:
            e.shard_sync()

    def read_cost_diagonal(self):
        return np.concatenate([e.read_cost_diagonal() for e in self.engines])

    def _energy_grover(self, angles):
        """Grover mixer: no state exchange -- per layer every shard's partial <s|v> is summed and handed back"""
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        p = angles.size // 2
        dot = None
        for step in range(p):
            dot = np.sum([e.shard_grover_step(step, angles, dot) for e in self.engines], axis=0)
        return combine_partials([e.shard_grover_step(p, angles, dot) for e in self.engines])

    chunk_major = False      # tests: run the windows of the chunk pipeline chunk by chunk (pre -> cross -> post per chunk)
    last_chunks = 1

    def energy(self, angles):
        if self.mixer_kind == _eng.MIXER_GROVER:
            return self._energy_grover(angles)
        counts = [e.shard_begin(angles) for e in self.engines]
        assert len(set(counts)) == 1
        self._sync_all()   # (first call: the diagonal streams of every rank are built from peer memory)
        flags = [self.engines[0].shard_pass_is_cross(i) for i in range(counts[0])]
        self.last_cross_flags = flags
        nch, closed = (1, [False] * counts[0])
        if self.chunk_major and hasattr(self.engines[0], "shard_chunk_info"):
            nch, closed = self.engines[0].shard_chunk_info(counts[0])
        self.last_chunks = nch
        if nch > 1:
            # every rank's launches of one step are separated from the next step by a device-wide sync (one process, one
            # GPU: no overlap here -- this path checks that the chunks really are independent of each other)
            for seg in pipeline_segments(flags, closed):
                if seg[0] == "pass":
                    self._sync_all()
                    for e in self.engines:
                        e.shard_run_pass(seg[1])
                    continue
                _, pre, x, post = seg
                for c in range(nch):
                    for i in pre + [x] + post:
                        self._sync_all()
                        for e in self.engines:
                            e.shard_run_pass_chunk(i, c)
            self._sync_all()
            return combine_partials([e.shard_finish() for e in self.engines])
        for i in range(counts[0]):
            if needs_barrier(flags, i):
                self._sync_all()
            for e in self.engines:
                e.shard_run_pass(i)
        return combine_partials([e.shard_finish() for e in self.engines])

    def _evaluate_keeping_state(self, angles):
        if self.symmetric:
            raise NotImplementedError("sharded symmetric register: the final state cannot be kept; use symmetric=False")
        self.set_option("keep_state", 1)
        try:
            return self.energy(angles)
        except Exception:
            self.set_option("keep_state", 0)
            raise

    def sample(self, angles, shots, seed=None):
        """``shots`` basis indices drawn from |psi(angles)|^2 of the WHOLE register (QAOA.hist, qaoa/qaoa.py:1134-1172)"""
        self._evaluate_keeping_state(angles)
        try:
            u = np.random.default_rng(seed).random(int(shots))
            deal = deal_shots(u, [e.shard_chunk_probs() for e in self.engines])
            out = np.zeros(int(shots), dtype=np.uint64)
            for r, (e, (mine, chunk, resid)) in enumerate(zip(self.engines, deal)):
                out[mine] = e.shard_locate(chunk, resid) + (np.uint64(r) << np.uint64(e.local_qubits))
            return out
        finally:
            self.set_option("keep_state", 0)

    def cvar(self, angles, alpha):
        """exact CVaR_alpha of the final state's cost distribution over all shards (statistic.py:132-142)"""
        self._evaluate_keeping_state(angles)
        try:
            e0 = self.engines[0]
            dec = decode_cost_key(e0.counter("diag_kind"), e0.counter("diag_c0"), e0.counter("diag_delta"))
            return cvar_select(float(alpha), e0.shard_cvar_key_bits(),
                               lambda pre, sh, ck: np.sum([e.shard_cvar_digit(pre, sh, ck) for e in self.engines], axis=0), dec)
        finally:
            self.set_option("keep_state", 0)

    def close(self):
        for e in self.engines:
            e.shard_close_peers()
        for e in self.engines:
            e.close()


class DistShardedEngine(_ShardedSurface):
    """This process' shard of a register spread over ``torch.distributed`` ranks (one process per GPU)."""

    def __init__(self, n_qubits, dtype="complex64", device=None, group=None, engine_factory=None, symmetric=False):
        import torch.distributed as dist

        self.symmetric = bool(symmetric)
        self._dist = dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if device is None and engine_factory is None:
            device = default_device()
        make = engine_factory or (lambda r, w: _eng.Engine(n_qubits, dtype=dtype, device=device, rank=r, world=w,
                                                           shard_symmetric=symmetric))
        self.n_qubits = int(n_qubits)
        self.eng = make(self.rank, self.world)
        self.barriers = 0
        self._diag_exchanged = False
        # NCCL + a real engine: cross-rank barriers and the final combination are ordered on the engine's own CUDA
        # stream (an NCCL collective between two passes), so the host never waits inside an evaluation
        self._ext = None
        if engine_factory is None and dist.get_backend(group) == "nccl" and hasattr(self.eng, "stream_handle"):
            import torch
            dev = torch.device("cuda", int(device))
            self._ext = torch.cuda.ExternalStream(self.eng.stream_handle(), device=dev)
            self._bar = torch.zeros(1, dtype=torch.int32, device=dev)
            self._part = torch.zeros(5, dtype=torch.float64, device=dev)
            self._parts_all = torch.zeros(5 * self.world, dtype=torch.float64, device=dev)
            self._ext2 = torch.cuda.ExternalStream(self.eng.stream2_handle(), device=dev)
            self._sms = torch.cuda.get_device_properties(dev).multi_processor_count
        import os
        # chunk pipeline (cross pass of one chunk beside the local passes of another); QAOA_CHUNK_PIPELINE=0 turns it off
        self.pipeline = os.environ.get("QAOA_CHUNK_PIPELINE", "1") != "0"
        # SMs a cross pass gets inside a pipeline window (NVLink-bound: alone, 24 CTAs keep the links busy, scripts/cross_ctas.py; beside local passes 32 is the measured optimum)
        self.cross_ctas = int(os.environ.get("QAOA_CROSS_CTAS", "32"))
        self.last_chunks = 1
        self._exchange(0)

    def _engines(self):
        return [self.eng]

    def _barrier(self):
        self.barriers += 1
        self._dist.barrier(group=self.group)

    def _exchange(self, which):
        """every rank maps every other rank's buffer (state: which=0, cost diagonal: which=1)"""
        mine = self.eng.shard_export(which)
        handles = [None] * self.world
        self._dist.all_gather_object(handles, mine, group=self.group)
        self.eng.shard_import(which, handles)
        self.eng.shard_sync()
        self._barrier()

    def _before_diag_change(self):
        """Replacing the cost function is COLLECTIVE: the old diagonal is IPC-mapped by the peers, and an exporter must
        not free memory a peer still maps -- every rank unmaps first (barrier, close, barrier), then the engines free."""
        if self._diag_exchanged:
            self.eng.shard_sync()
            self._barrier()
            self.eng.shard_close_peers()      # unmaps state and diagonal peers
            self._barrier()
            self._diag_exchanged = False
            self._exchange(0)                 # the state buffers are unchanged: map them again

    def _diag_changed(self):
        self._exchange(1)
        self._diag_exchanged = True

    def _combine_minmax(self, lo, hi):
        import torch

        dist = self._dist
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([-lo, hi], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        t = t.cpu()
        return -float(t[0]), float(t[1])

    def read_cost_diagonal(self):
        return self.eng.read_cost_diagonal()

    def _all_reduce(self, part):
        import torch

        dist = self._dist
        backend = dist.get_backend(self.group)
        dev = "cuda" if backend == "nccl" else "cpu"
        sums = torch.tensor(part[:3], dtype=torch.float64, device=dev)
        lo = torch.tensor(part[3:4], dtype=torch.float64, device=dev)
        hi = torch.tensor(part[4:5], dtype=torch.float64, device=dev)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        s = sums.cpu().numpy()
        return np.array([s[1] / s[0], s[2] / s[0], float(lo.cpu()[0]), float(hi.cpu()[0]), s[0]])

    def _energy_grover(self, angles):
        """Grover mixer (grover_mixer.py:43-82): the state never crosses the ranks; per layer ONE complex scalar, the
        partial <s|v> of every shard, is all-reduced, and the five statistics at the end."""
        import torch

        dist = self._dist
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        angles = np.ascontiguousarray(angles, dtype=np.float64).reshape(-1)
        p = angles.size // 2
        dot = None
        for step in range(p):
            t = torch.tensor(self.eng.shard_grover_step(step, angles, dot), dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            dot = t.cpu().numpy()
        return self._all_reduce(self.eng.shard_grover_step(p, angles, dot))

    def energy(self, angles):
        """(E[cost], E[cost^2], min, max, norm) of the WHOLE register; identical on every rank."""
        if self.mixer_kind == _eng.MIXER_GROVER:
            return self._energy_grover(angles)
        eng = self.eng
        n_passes = eng.shard_begin(angles)
        flags = [eng.shard_pass_is_cross(i) for i in range(n_passes)]
        if self._ext is not None:
            return self._energy_stream_ordered(n_passes, flags)
        for i in range(n_passes):
            if needs_barrier(flags, i):
                eng.shard_sync()
                self._barrier()
            eng.shard_run_pass(i)
        # the all-reduce doubles as the barrier that keeps the next evaluation's first pass from
        # overwriting a shard some peer is still reading in its last (cross) pass
        return self._all_reduce(eng.shard_finish())

    def _energy_stream_ordered(self, n_passes, flags):
        """The same protocol with every cross-rank barrier an NCCL all-reduce of one word ON THE ENGINE'S STREAM: it
        completes on a rank only after every rank has reached it, i.e. after every rank's earlier passes, and the later
        passes of this rank queue behind it -- no stream synchronise + host barrier bubbles around the cross passes.  The
        five partial sums go device to device into ONE all-gather (which also fences the next evaluation)."""
        import torch

        dist, eng = self._dist, self.eng
        nch, closed = eng.shard_chunk_info(n_passes) if self.pipeline else (1, [False] * n_passes)
        self.last_chunks = nch
        A, B = self._ext, self._ext2

        def barrier():
            self.barriers += 1
            dist.all_reduce(self._bar, group=self.group)

        with torch.cuda.stream(A):
            if nch == 1:
                for i in range(n_passes):
                    if needs_barrier(flags, i):
                        barrier()
                    eng.shard_run_pass(i)
            else:
                # Chunk pipeline.  Main stream A: local passes (inside a window on SMs - cross_ctas CTAs); side stream B:
                # the window's cross pass, chunk by chunk, on cross_ctas CTAs, every chunk between two barriers.  Events
                # order the two streams per chunk; all ranks enqueue the same sequence, so the collectives match up.
                local_ctas = max(1, self._sms - self.cross_ctas)
                prev_cross = False
                for seg in pipeline_segments(flags, closed):
                    if seg[0] == "pass":
                        i = seg[1]
                        if flags[i] or prev_cross:
                            barrier()
                        eng.shard_run_pass(i)
                        prev_cross = bool(flags[i])
                        continue
                    _, pre, x, post = seg
                    if prev_cross:
                        barrier()
                    ev_a = [torch.cuda.Event() for _ in range(nch)]
                    ev_b = [torch.cuda.Event() for _ in range(nch)]
                    for c in range(nch):
                        for i in pre:
                            eng.shard_run_pass_chunk(i, c, 0, local_ctas)
                        ev_a[c].record(A)
                    with torch.cuda.stream(B):
                        for c in range(nch):
                            B.wait_event(ev_a[c])
                            barrier()                 # every rank has finished chunk c's passes before the cross pass
                            eng.shard_run_pass_chunk(x, c, 1, self.cross_ctas)
                            barrier()                 # every rank's cross pass of chunk c is complete
                            ev_b[c].record(B)
                    for c in range(nch):
                        A.wait_event(ev_b[c])
                        for i in post:
                            eng.shard_run_pass_chunk(i, c, 0, local_ctas)
                    prev_cross = False
            eng.shard_finish_device(self._part.data_ptr())
            dist.all_gather_into_tensor(self._parts_all, self._part, group=self.group)
            parts = self._parts_all.cpu().numpy().reshape(self.world, 5)
        eng.shard_sync()          # stream is idle: collects the per-pass event timings
        return combine_partials(parts)

    def _host_tensor(self, arr, dtype):
        import torch
        dev = "cuda" if self._dist.get_backend(self.group) == "nccl" else "cpu"
        return torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype).to(dev)

    def _evaluate_keeping_state(self, angles):
        if self.symmetric:
            raise NotImplementedError("sharded symmetric register: the final state cannot be kept; use symmetric=False")
        self.eng.set_option("keep_state", 1)
        try:
            return self.energy(angles)
        except Exception:
            self.eng.set_option("keep_state", 0)
            raise

    def sample(self, angles, shots, seed=None):
        """``shots`` basis indices drawn from |psi(angles)|^2 of the WHOLE register (QAOA.hist, qaoa/qaoa.py:1134-1172);
        collective, identical on every rank: rank 0's uniform numbers are broadcast, the per-chunk masses of all shards
        gathered, every rank locates the shots that fall into its own shard, one all-reduce assembles the indices."""
        import torch

        dist = self._dist
        self._evaluate_keeping_state(angles)
        try:
            u = self._host_tensor(np.random.default_rng(seed).random(int(shots)), torch.float64)
            dist.broadcast(u, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            mine_sums = self._host_tensor(self.eng.shard_chunk_probs(), torch.float64)
            sums = [torch.empty_like(mine_sums) for _ in range(self.world)]
            dist.all_gather(sums, mine_sums, group=self.group)
            deal = deal_shots(u.cpu().numpy(), [t.cpu().numpy() for t in sums])
            mine, chunk, resid = deal[self.rank]
            out = np.zeros(int(shots), dtype=np.int64)
            out[mine] = (self.eng.shard_locate(chunk, resid) + (np.uint64(self.rank) << np.uint64(self.eng.local_qubits))).astype(np.int64)
            t = self._host_tensor(out, torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return t.cpu().numpy().astype(np.uint64)
        finally:
            self.eng.set_option("keep_state", 0)

    def cvar(self, angles, alpha):
        """exact CVaR_alpha over all shards: the per-digit histograms of the weighted quantile select are all-reduced"""
        import torch

        dist = self._dist
        self._evaluate_keeping_state(angles)
        try:
            eng = self.eng

            def digit(prefix, shift, check):
                t = self._host_tensor(eng.shard_cvar_digit(prefix, shift, check), torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                return t.cpu().numpy()

            dec = decode_cost_key(eng.counter("diag_kind"), eng.counter("diag_c0"), eng.counter("diag_delta"))
            return cvar_select(float(alpha), eng.shard_cvar_key_bits(), digit, dec)
        finally:
            self.eng.set_option("keep_state", 0)

    def close(self):
        """collective: every rank unmaps its peers before any rank frees the buffers they map"""
        self._barrier()
        if hasattr(self.eng, "shard_close_peers"):
            self.eng.shard_close_peers()
        self._barrier()
        self.eng.close()
