"""QAOA driver with the reference's public API, evaluated on the B200 engine.

``QAOA(problem=..., mixer=..., initialstate=...)``, ``sample_cost_landscape()``,
``optimize(depth)``, ``get_Exp/get_Var/get_gamma/get_beta/get_angles`` keep the signatures and
semantics of the reference driver (reference qaoa/qaoa.py:157-1218).  What changes is the
evaluation path: where the reference builds, transpiles and binds a qiskit circuit, runs it on
qiskit-aer for ``shots`` samples and averages ``problem.cost`` over the counts
(qaoa/qaoa.py:418-520, 592-614, 653-661, 691-740, 912-922), this class lowers the three
components onto a device-resident statevector engine once and asks it for the EXACT moments
``E[cost]``, ``E[cost^2]``, ``min``, ``max`` of the final state.

Semantics in the exact engine:
  * ``Exp = -E[cost]`` (cvar = 1) is the shots -> infinity limit of the reference's estimate;
  * ``Var`` is the population variance (the reference's S/(W-1) is undefined without samples);
  * ``BestCost/WorstCost`` are -max/-min of the cost over the support of |psi|^2;
  * ``flip=True`` (X gates between layers from the boosted most-frequent string, qaoa.py:505-509, 840-848) is
    honoured on a single GPU: the histogram it needs is sampled on the device at the depth's best angles;
  * ``noisemodel``, ``precision``, ``memorysize > 0``, QNSPSA and foreign
    ``backend`` objects cannot be honoured by an exact statevector engine and are REJECTED
    (NotImplementedError) — nothing is ever routed to a CPU simulator.
"""

from __future__ import annotations

import logging
import time

import numpy as np

from qaoa_b200.initialstates import InitialState
from qaoa_b200.mixers import Mixer
from qaoa_b200.problems import Problem
from qaoa_b200.utils import COBYLA, ExactStatistic

LOG = logging.getLogger(__name__)


class B200Backend:
    """Stands where the reference keeps its qiskit backend (``backend.name``,
    ``backend.configuration().local`` are the only members the reference reads,
    qaoa/qaoa.py:564 and qaoa/utils/qaoaIO.py:255)."""

    name = "b200_statevector"

    def __init__(self, device=None, dtype="complex128", sharded=False, split_batches=False, shard_symmetric=True):
        """Multi-GPU modes (under ``torch.distributed``, one process per GPU, every rank driving the same
        QAOA object and seeing identical results):
        ``sharded=True``: ONE register split on its top qubits over all ranks (qaoa_b200/sharded.py); with
        ``shard_symmetric`` (default) MaxCut + X-type mixer + |+> on an even number of qubits in complex64 stores only
        the half register with the top qubit = 0 (same results, half the memory, HBM and NVLink traffic);
        ``split_batches=True``: every rank holds a full replica and batched evaluations (landscape grids, layer
        grid searches) are dealt out round-robin, the results gathered -- no data-path collective."""
        # None: resolved when the engine is built -- LOCAL_RANK under torchrun, else torch's current device, else 0
        # (so that sharded=True / split_batches=True put every rank on its OWN GPU without the caller passing it)
        self.device = device
        self.dtype = dtype
        self.sharded = sharded
        self.shard_symmetric = shard_symmetric
        self.split_batches = split_batches

    class _Conf:
        local = True

    def configuration(self):
        return self._Conf()


class OptResult:
    """Per-depth optimisation record (same fields as reference qaoa/qaoa.py:30-154)."""

    def __init__(self, depth):
        self.depth = depth
        self.angles = []
        self.Exp = []
        self.Var = []
        self.WorstCost = []
        self.BestCost = []
        self.BestSols = []
        self.shots = []
        self.index_Exp_min = -1
        self.opt_time = -1.0

    def add_iteration(self, angles, stat, shots):
        self.angles.append(angles)
        self.Exp.append(-stat.get_CVaR())
        self.Var.append(stat.get_Variance())
        self.BestCost.append(-stat.get_max())
        self.WorstCost.append(-stat.get_min())
        self.BestSols.append(stat.get_max_sols())
        self.shots.append(shots)

    def compute_best_index(self):
        self.index_Exp_min = int(np.argmin(self.Exp))

    def get_best_Exp(self):
        return self.Exp[self.index_Exp_min]

    def get_best_Var(self):
        return self.Var[self.index_Exp_min]

    def get_best_angles(self):
        return self.angles[self.index_Exp_min]

    def num_fval(self):
        return len(self.Exp)

    def num_shots(self):
        return sum(self.shots)

    def get_best_solution(self):
        lowest = np.min(self.BestCost)
        sols = []
        for i in np.where(np.asarray(self.BestCost) == lowest)[0]:
            sols.extend(self.BestSols[i])
        return np.unique(sols), lowest


class _DepthHistogram(dict):
    """QAOA.samplecount_hists[depth]: in the reference the raw counts of the optimiser's last evaluation at that depth
    (qaoa/qaoa.py:830), consumed by utils.post_process_all_depths and the plot routines.  An exact engine has no such
    by-product, so the histogram is drawn by the device sampler -- `shots` strings at the depth's best angles, keys in
    the sampler's printing order (qubit 0 last) -- the first time anybody looks at it."""

    def __init__(self, owner, depth):
        super().__init__()
        self._owner, self._depth, self._drawn = owner, depth, False

    def _draw(self):
        if not self._drawn:
            self._drawn = True
            own = self._owner
            best = own.optimization_results[self._depth].get_best_angles()
            depth_before = own.parametrized_circuit_depth
            super().update({s[::-1]: c for s, c in own.hist(best, own.shots).items()})
            # looking at a histogram must not move the circuit depth under the caller (hist() re-parameterises the
            # circuit like the reference's, qaoa.py:1134-1172): put the depth back where it was
            if depth_before and depth_before != own.parametrized_circuit_depth:
                own.createParameterizedCircuit(depth_before)

    def __reduce__(self):      # pickles as a plain dict: what has been drawn so far (never samples as a side effect)
        return (dict, (dict(dict.items(self)),))


def _lazy(name):
    def method(self, *a, **kw):
        self._draw()
        return getattr(dict, name)(self, *a, **kw)
    method.__name__ = name
    return method


for _name in ("__getitem__", "__iter__", "__len__", "__contains__", "__repr__", "__eq__", "keys", "values", "items", "get",
              "copy"):
    setattr(_DepthHistogram, _name, _lazy(_name))


def _reject(what):
    raise NotImplementedError(
        what + " cannot be honoured by the exact B200 statevector engine and is rejected "
        "(the engine never falls back to a sampling/CPU simulator)"
    )


class QAOA:
    def __init__(
        self,
        problem,
        mixer,
        initialstate,
        backend=None,
        noisemodel=None,
        optimizer=None,
        precision=None,
        shots=1024,
        cvar=1,
        memorysize=-1,
        interpolate=True,
        flip=False,
        post=False,
        number_trottersteps_mixer=1,
        sequential=False,
        dtype=None,
        device=0,
    ) -> None:
        assert issubclass(type(problem), Problem)
        assert issubclass(type(mixer), Mixer)
        assert issubclass(type(initialstate), InitialState)

        if backend is None:
            backend = B200Backend(device=device, dtype=dtype or "complex128")
        if not isinstance(backend, B200Backend):
            _reject("backend=%r" % (backend,))
        if dtype is not None:
            backend.dtype = dtype
        if noisemodel is not None:
            _reject("noisemodel")
        if precision is not None:
            _reject("precision (shot-adaptive estimation)")
        if memorysize is not None and memorysize > 0:
            _reject("memorysize > 0 (per-shot memory)")
        if flip and getattr(backend, "sharded", False):
            _reject("flip=True on a sharded register")
        if getattr(problem, "N_ancilla_qubits", 0):
            _reject("problems with ancilla qubits")
        if not (0 < cvar <= 1):
            raise ValueError("cvar must be in (0, 1]")

        self.problem = problem
        self.mixer = mixer
        self.initialstate = initialstate
        self.initialstate.setNumQubits(self.problem.N_qubits)
        self.mixer.setNumQubits(self.problem.N_qubits)

        self.parameterized_circuit = None
        self.parameterized_circuit_depth = 0
        self.parametrized_circuit_depth = 0  # the reference writes this spelling (qaoa.py:520)

        self.backend = backend
        self.optimizer = [COBYLA, {}] if optimizer is None else optimizer
        if getattr(self.optimizer[0], "__name__", "") == "QNSPSA":
            _reject("QNSPSA (needs circuit fidelities)")
        self.noisemodel = None
        self.shots = shots
        self.precision = None
        self.stat = ExactStatistic(cvar=cvar)
        self.cvar = cvar
        self.memorysize = memorysize
        self.memory = False
        self.interpolate = interpolate
        self.sequential = sequential
        self.usebarrier = False
        self.isQNSPSA = False

        self.current_depth = 0
        self.gamma_params = None
        self.beta_params = None
        self.init_params = None
        self.n_gamma = 1
        self.n_beta = 1
        self.n_init = 0

        self.Exp_sampled_p1 = None
        self.landscape_p1_angles = {}
        self.Var_sampled_p1 = None
        self.MaxCost_sampled_p1 = None
        self.MinCost_sampled_p1 = None

        self.optimization_results = {}
        self.memory_lists = []
        # bit-flip boosting between layers (qaoa.py:505-509, 840-848): after depth d has been optimised, the most frequent
        # sampled string is improved by greedy classical bit flips and the XOR pattern becomes X gates after layer d of
        # every deeper circuit.  The device applies them as an index permutation of the kept state (_device_call).
        self.flip = bool(flip)
        from qaoa_b200.utils.flip import BitFlip
        self.flipper = BitFlip(self.problem.N_qubits)
        self.bitflips = []
        # classical post-processing of the final histogram (qaoa.py:852-856): K = `post` passes of greedy bit flips per
        # sampled string; the histogram is drawn by the device sampler at the best angles of the last depth
        self.post = post
        self.Exp_post_processed = None
        self.Var_post_processed = None
        self.samplecount_hists = {}
        self.last_hist = {}
        self.number_trottersteps_mixer = number_trottersteps_mixer

        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engine"] = None  # device handles are re-created lazily after unpickling
        state["_engine_full"] = None
        state.pop("_cvar_sorted", None)
        return state

    def _get_engine(self):
        """Builds the device plan (what createParameterizedCircuit + transpile were, qaoa.py:418-520)."""
        if self._engine is None:
            from qaoa_b200.engine import Engine

            if self.problem.get_num_parameters() != 1 and not hasattr(self.problem, "term_of_edge"):
                _reject("problems with more than one gamma per layer (other than MaxCutFree / MaxCutOrbit)")
            if self.initialstate.get_num_parameters() != 0 and type(self.initialstate).__name__ != "PlusParameterized":
                _reject("parameterised initial states (other than PlusParameterized)")
            # plugins that cannot be lowered are rejected BEFORE a device is touched (same error with or without a GPU)
            mixer_cls, state_cls = type(self.mixer), type(self.initialstate)
            if mixer_cls.lower_mixer is Mixer.lower_mixer or getattr(self.mixer, "not_lowerable", False):
                self.mixer.lower_mixer(None, trotter=self.number_trottersteps_mixer)        # raises NotImplementedError
            if state_cls.state_descriptor is InitialState.state_descriptor and state_cls.lower_initial is InitialState.lower_initial:
                self.initialstate.state_descriptor()                                        # raises NotImplementedError
            device = getattr(self.backend, "device", None)
            if device is None:
                from qaoa_b200.sharded import default_device
                device = default_device()
            if getattr(self.backend, "sharded", False):
                from qaoa_b200.sharded import DistShardedEngine
                if hasattr(self.problem, "term_of_edge") and self.problem.get_num_parameters() != 1:
                    _reject("per-term gammas (MaxCutFree / MaxCutOrbit) on a sharded register")
                eng = DistShardedEngine(self.problem.N_qubits, dtype=self.backend.dtype, device=device,
                                        symmetric=self._shard_symmetric())
            else:
                eng = Engine(self.problem.N_qubits, dtype=self.backend.dtype, device=device)
            self.problem.lower_cost(eng)
            self.initialstate.lower_initial(eng)
            self.mixer.lower_mixer(eng, trotter=self.number_trottersteps_mixer)
            self._engine = eng
        return self._engine

    def _full_sharded_engine(self):
        """the same circuit on the FULL sharded register (kept final states: hist(), cvar < 1); collective like _get_engine"""
        if getattr(self, "_engine_full", None) is None:
            from qaoa_b200.sharded import DistShardedEngine, default_device

            device = getattr(self.backend, "device", None)
            eng = DistShardedEngine(self.problem.N_qubits, dtype=self.backend.dtype,
                                    device=default_device() if device is None else device, symmetric=False)
            self.problem.lower_cost(eng)
            self.initialstate.lower_initial(eng)
            self.mixer.lower_mixer(eng, trotter=self.number_trottersteps_mixer)
            self._engine_full = eng
        return self._engine_full

    def _shard_symmetric(self):
        """MaxCut proper + X-type mixer + |+>, n even, complex64: cost(x) = cost(~x) and the state keeps that symmetry,
        so a sharded register stores the half with the top qubit = 0 only (csrc/tile_kernel.cuh, twisted coordinates)."""
        import torch.distributed as dist

        pr, n = self.problem, self.problem.N_qubits
        g = dist.get_world_size().bit_length() - 1 if dist.is_available() and dist.is_initialized() else 0
        return bool(getattr(self.backend, "shard_symmetric", True) and g > 0
                    and getattr(pr, "N_qubits_per_node", 0) == 1 and getattr(pr, "k_cuts", 0) == 2
                    and not getattr(pr, "fix_one_node", False) and pr.get_num_parameters() == 1
                    and type(self.mixer).__name__ in ("X", "XMultiAngle", "XOrbit")
                    and type(self.initialstate).__name__ == "Plus"
                    and str(self.backend.dtype) in ("complex64", "0") and n % 2 == 0 and n - 1 - g >= 15)

    def _evaluate(self, angle_rows):
        """(B, n_angles) FULL flat angle vectors (initial-state parameters first) -> (B, 6) = E, E2, min, max, norm, CVaR.

        CVaR (= E when cvar == 1) is the shots -> infinity limit of the reference's
        ``Statistic.get_CVaR`` (qaoa/utils/statistic.py:132-142): the mean of the best (largest-cost)
        ``cvar`` fraction of the distribution, computed from the device's cost-value distribution."""
        rows = np.atleast_2d(np.asarray(angle_rows, dtype=np.float64))
        eng = self._get_engine()
        if self._rows_need_lowering():
            # shared mixer parameters (XOrbit) / rescaled cost parameters: the angles the device takes
            rows = np.asarray([self._engine_row(r) for r in rows], dtype=np.float64)
        world, rank, dist = 1, 0, None
        if getattr(self.backend, "split_batches", False) and not getattr(self.backend, "sharded", False):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
        mine = rows[rank::world] if world > 1 and len(rows) >= world else rows
        if self.cvar < 1 and hasattr(eng, "cvar_stats") and not getattr(self.backend, "sharded", False):
            # one kept-state evaluation per row gives the statistics AND the CVaR (qaoa_cvar_stats)
            out = np.array([self._device_call("cvar_stats", r, self.cvar) for r in mine])
        else:
            if self._flips_active(mine[0]):
                res = np.array([self._device_call("energy", r) for r in mine])
            else:
                res = eng.energy_batch(mine)
            if self.cvar < 1:
                cv = np.array([self._cvar_of(r) for r in mine])
            else:
                cv = res[:, 0]
            out = np.column_stack([res, cv])
        if mine is not rows:      # landscape / grid-search points were dealt out over the ranks: put them back in order
            parts = [None] * world
            dist.all_gather_object(parts, out)
            full = np.empty((len(rows), out.shape[1]))
            for r, part in enumerate(parts):
                full[r::world] = part
            out = full
        return out

    def _flip_masks(self):
        """bit masks (bit q = qubit q) of the X gates that follow layers 0, 1, ... (reference qaoa.py:505-509: layer d of
        a depth-D circuit, d != D - 1, is followed by flipper.create_circuit(bitflips[d]))"""
        masks = []
        for pattern in self.bitflips:
            self.flipper.create_circuit(pattern)
            masks.append(self.flipper.flip_mask)
        return masks

    def _flips_active(self, eng_row):
        if not self.flip or not self.bitflips:
            return False
        eng = self._get_engine()
        depth = (len(eng_row) - eng.n_init) // (eng.n_gamma + eng.n_beta)
        return any(self._flip_masks()[: depth - 1])

    def _device_call(self, name, eng_row, *args, **kw):
        """``engine.<name>(angles, ...)`` on the final state of the circuit at ``eng_row`` (device angle layout).  With
        X gates between layers the circuit runs segment by segment on the device: layers up to a flip with the state
        kept in HBM, the flip as an index permutation of that state (Engine.continue_from_kept_state), the next layers
        from there; ``name`` is called for the last segment, whose result is the whole circuit's."""
        eng = self._get_engine()
        row = np.asarray(eng_row, dtype=np.float64)
        if not hasattr(eng, name):
            _reject("%s on a sharded register (single-GPU engines only)" % name)
        if name in ("sample", "cvar") and getattr(eng, "symmetric", False) and getattr(self.backend, "sharded", False):
            # the sharded SYMMETRIC register cannot keep its final state; sampling and CVaR of such a run use a second,
            # full sharded register lowered from the same problem (built on first use; twice the stored amplitudes)
            eng = self._full_sharded_engine()
        if not self._flips_active(row):
            return getattr(eng, name)(row, *args, **kw)
        n_init, per = eng.n_init, eng.n_gamma + eng.n_beta
        depth = (len(row) - n_init) // per
        masks = self._flip_masks()[: depth - 1]
        eng.set_option("keep_state", 1)
        try:
            start = 0                                   # first row entry of the segment in flight
            for d, mask in enumerate(masks):
                if mask:
                    stop = n_init + (d + 1) * per
                    eng.energy(row[start:stop])
                    eng.continue_from_kept_state(mask)
                    start = stop
            eng.set_option("keep_state", 0)
            return getattr(eng, name)(row[start:], *args, **kw)
        finally:
            eng.set_option("keep_state", 0)
            self.initialstate.lower_initial(eng)        # back to the circuit's own initial state

    _CVAR_HOST_MAX_QUBITS = 26

    def _cvar_of(self, eng_row):
        """exact CVaR of the final state's cost distribution (shots -> infinity limit of statistic.py:132-142).  On the
        device engine: ``qaoa_cvar`` (digit-by-digit weighted quantile select, integer AND real-valued costs -- weighted
        graphs, scaled ExactCover, Portfolio: the reference's CVaR.ipynb and CompareGroverXY.ipynb cases).  Engines
        without it (the oracle stand-in of the CPU tests, sharded registers) keep the host evaluation: cost-value
        histogram for integer costs, else |psi|^2 read back and the tail mean over the sorted diagonal (<= 26 qubits)."""
        from qaoa_b200.engine import EngineError

        eng = self._get_engine()
        if hasattr(eng, "cvar"):
            # the device engine: weighted quantile select over the kept final state (qaoa_cvar), every cost
            # representation, no read-back of |psi|^2, no qubit limit
            return self._device_call("cvar", eng_row, self.cvar)
        if getattr(self, "_cvar_on_host", None) is None:
            try:
                costs, probs = self._device_call("cost_distribution", eng_row)
                self._cvar_on_host = False
                return self._exact_cvar(costs, probs)
            except EngineError as exc:
                if "integer-valued" not in str(exc):
                    raise
                self._cvar_on_host = True
        if not self._cvar_on_host:
            return self._exact_cvar(*self._device_call("cost_distribution", eng_row))
        if self.problem.N_qubits > self._CVAR_HOST_MAX_QUBITS:
            _reject("cvar < 1 with real-valued costs on more than %d qubits" % self._CVAR_HOST_MAX_QUBITS)
        if getattr(self, "_cvar_sorted", None) is None:
            diag = np.asarray(eng.read_cost_diagonal(), dtype=np.float64)
            order = np.argsort(-diag, kind="stable")
            self._cvar_sorted = (order, diag[order])
        order, costs = self._cvar_sorted
        probs = np.asarray(self._device_call("final_probabilities", eng_row))[order]
        cum = np.cumsum(probs)
        take = np.clip(self.cvar - (cum - probs), 0.0, probs)
        return float(take @ costs) / self.cvar

    def _exact_cvar(self, costs, probs):
        order = np.argsort(-costs, kind="stable")
        c, p = costs[order], probs[order]
        cum = np.cumsum(p)
        take = np.clip(self.cvar - (cum - p), 0.0, p)
        return float(take @ c) / self.cvar

    # ------------------------------------------------------------------ accessors
    def exp_landscape(self):
        return self.Exp_sampled_p1

    def var_landscape(self):
        return self.Var_sampled_p1

    def get_Exp(self, depth=None):
        if not depth:
            return [self.optimization_results[d].get_best_Exp() for d in range(1, self.current_depth + 1)]
        if depth > self.current_depth + 1:
            raise ValueError
        return self.optimization_results[depth].get_best_Exp()

    def get_Var(self, depth):
        if depth > self.current_depth + 1:
            raise ValueError
        return self.optimization_results[depth].get_best_Var()

    def get_beta(self, depth):
        if depth > self.current_depth + 1:
            raise ValueError
        return self.optimization_results[depth].get_best_angles()[1::2]

    def get_gamma(self, depth):
        if depth > self.current_depth + 1:
            raise ValueError
        return self.optimization_results[depth].get_best_angles()[::2]

    def get_angles(self, depth):
        if depth > self.current_depth + 1:
            raise ValueError
        return self.optimization_results[depth].get_best_angles()

    def get_memory_lists(self):
        return self.memory_lists

    # ------------------------------------------------------------------ plan ("circuit")
    def createParameterizedCircuit(self, depth):
        """Fixes the parameter layout for ``depth`` layers.  No circuit is built: the engine applies
        init -> [phase(gamma_d) -> mixer(beta_d)] x depth directly on the statevector."""
        self.problem.create_circuit()
        self.mixer.create_circuit()
        self.initialstate.create_circuit()
        self.n_gamma = self.problem.get_num_parameters()
        self.n_beta = self.mixer.get_num_parameters()
        self.n_init = self.initialstate.get_num_parameters()
        self.gamma_params = [["gamma_{}_{}".format(d, i) for i in range(self.n_gamma)] for d in range(depth)]
        self.beta_params = [["beta_{}_{}".format(d, i) for i in range(self.n_beta)] for d in range(depth)]
        self.init_params = ["init_{}".format(i) for i in range(self.n_init)]
        self.parameterized_circuit_depth = depth
        self.parametrized_circuit_depth = depth
        self._get_engine()

    def getParametersToBind(self, angles, depth, asList=False):
        """Name -> value view of the flat angle vector (layout of reference qaoa.py:937-990)."""
        per_layer = self.n_gamma + self.n_beta
        assert len(angles) == self.n_init + depth * per_layer
        wrap = (lambda v: [v]) if asList else (lambda v: v)
        bound = {}
        for i in range(self.n_init):
            bound["init_{}".format(i)] = wrap(angles[i])
        at = self.n_init
        for d in range(depth):
            for i in range(self.n_gamma):
                bound["gamma_{}_{}".format(d, i)] = wrap(angles[at + i])
            at += self.n_gamma
            for i in range(self.n_beta):
                bound["beta_{}_{}".format(d, i)] = wrap(angles[at + i])
            at += self.n_beta
        return bound

    # ------------------------------------------------------------------ landscape
    def _grid(self, angles):
        g, b = angles["gamma"], angles["beta"]
        return (np.linspace(g[0], g[1], g[2], endpoint=False), np.linspace(b[0], b[1], b[2], endpoint=False))

    def _vanilla_rows(self, gamma_grid, beta_grid, prefix=()):
        """All grid points in the symmetric subspace, beta-major like the reference loops
        (qaoa.py:637-649): every gamma of the layer = grid gamma, every beta = grid beta."""
        prefix = list(prefix)
        rows = [
            prefix + [g] * self.n_gamma + [b] * self.n_beta
            for b in beta_grid
            for g in gamma_grid
        ]
        return np.asarray(rows, dtype=np.float64)

    def sample_cost_landscape(self, angles={"gamma": [0, 2 * np.pi, 20], "beta": [0, 2 * np.pi, 20]}):
        """p = 1 landscape on the (gamma, beta) grid: all points go to the device in one batched call
        (the reference's batch branch, qaoa.py:630-662; ``sequential`` gives identical numbers here)."""
        self.landscape_p1_angles = angles
        LOG.info("Calculating energy landscape for depth p=1...")
        if not self.backend.configuration().local:
            raise NotImplementedError
        self.createParameterizedCircuit(1)
        self.gamma_grid, self.beta_grid = self._grid(angles)
        rows = self._vanilla_rows(self.gamma_grid, self.beta_grid, prefix=[0.0] * self.n_init)
        LOG.info("Executing sample_cost_landscape")
        LOG.info("circuits: %d", len(rows))
        res = self._evaluate(rows)
        shape = (angles["beta"][2], angles["gamma"][2])
        self.Exp_sampled_p1 = -res[:, 5].reshape(shape)
        self.Var_sampled_p1 = np.maximum(res[:, 1] - res[:, 0] ** 2, 0.0).reshape(shape)
        self.MaxCost_sampled_p1 = -res[:, 3].reshape(shape)
        self.MinCost_sampled_p1 = -res[:, 2].reshape(shape)
        self.last_hist = {}
        LOG.info("Calculating Energy landscape done")

    def measurementStatistics(self, job):
        """Statistics of SAMPLED counts, with the reference's semantics (qaoa.py:669-740): ``job`` is a histogram
        {string: count} in the sampler's printing order (qubit 0 LAST) -- e.g. ``hist()`` with its keys reversed --, a list of
        such histograms for a landscape grid in (beta, gamma) order, or any object whose ``result().get_counts()`` returns
        either.  A single histogram is accumulated into ``self.stat`` (CVaR, Bessel-corrected variance, extremes among the
        sampled strings); a list fills ``Exp/Var/MaxCost/MinCost_sampled_p1`` on the grid of ``landscape_p1_angles``.
        The exact engine never calls this itself; it serves code that post-processes sampled histograms."""
        counts_list = job.result().get_counts() if hasattr(job, "result") else job
        self.last_hist = counts_list

        def accumulate(counts):
            for string, weight in counts.items():
                self.stat.add_sample(self.problem.cost(string[::-1]), weight, string[::-1])

        if isinstance(counts_list, list):
            rows = []
            for counts in counts_list:
                self.stat.reset()
                accumulate(counts)
                rows.append((self.stat.get_CVaR(), self.stat.get_Variance(), self.stat.get_max(), self.stat.get_min()))
            rows = np.asarray(rows, dtype=np.float64)
            shape = (self.landscape_p1_angles["beta"][2], self.landscape_p1_angles["gamma"][2])
            self.Exp_sampled_p1 = -rows[:, 0].reshape(shape)
            self.Var_sampled_p1 = rows[:, 1].reshape(shape)
            self.MaxCost_sampled_p1 = -rows[:, 2].reshape(shape)
            self.MinCost_sampled_p1 = -rows[:, 3].reshape(shape)
        else:
            accumulate(counts_list)

    @staticmethod
    def _argmin_first(values):
        """np.argmin with a deterministic tie-break: landscapes of symmetric problems have exactly
        degenerate minima (16-fold for MaxCut on [0, 2pi)^2) that differ only by rounding noise, so
        values are quantised to 1e-9 of the landscape's scale before the first minimum is taken."""
        v = np.asarray(values, dtype=np.float64)
        scale = max(np.max(np.abs(v)), 1e-300)
        q = np.round(v / (1e-9 * scale))
        return int(np.argmin(q, axis=None))

    # ------------------------------------------------------------------ optimisation
    def optimize(self, depth, angles={"gamma": [0, 2 * np.pi, 20], "beta": [0, 2 * np.pi, 20]}):
        """Depth-by-depth local optimisation (control flow of reference qaoa.py:742-856)."""
        while self.current_depth < depth:
            n_gamma = self.problem.get_num_parameters()
            n_beta = self.mixer.get_num_parameters()
            n_init = self.initialstate.get_num_parameters()
            t0 = time.perf_counter()
            if self.current_depth == 0:
                if self.Exp_sampled_p1 is None:
                    self.sample_cost_landscape(angles=angles)
                flat = self._argmin_first(self.Exp_sampled_p1)
                ib, ig = np.unravel_index(flat, self.Exp_sampled_p1.shape)
                angles0 = np.array([0.0] * n_init + [self.gamma_grid[ig]] * n_gamma + [self.beta_grid[ib]] * n_beta)
            else:
                best = self.get_angles(self.current_depth)
                angles0 = self.interp(best) if self.interpolate else self._grid_search_layer(best, angles)

            target = self.current_depth + 1
            self.optimization_results[target] = OptResult(target)
            self.createParameterizedCircuit(int((len(angles0) - n_init) / (n_gamma + n_beta)))
            res = self.local_opt(angles0)
            self.samplecount_hists[target] = _DepthHistogram(self, target)
            self.optimization_results[target].compute_best_index()
            self.optimization_results[target].opt_time = time.perf_counter() - t0
            LOG.info("cost(depth %d = %s", target, res.fun)
            if self.flip:
                # qaoa.py:840-848.  (The reference looks at the raw counts of the optimiser's LAST evaluation; here the
                # histogram is sampled on the device at the depth's best angles.)
                counts = self.samplecount_hists[target]
                string = max(counts, key=counts.get)
                boosted = self.flipper.boost_samples(problem=self.problem, string=string)
                self.bitflips.append(self.flipper.xor(string, boosted))
            self.current_depth += 1

        if self.post:
            from qaoa_b200.utils.post import post_processing

            best = self.optimization_results[self.current_depth].get_best_angles()
            # sampler's printing order (qubit 0 last), as the reference stores its raw counts
            samples = {s[::-1]: c for s, c in self.hist(best, self.shots).items()}
            self.samplecount_hists[self.current_depth] = samples
            post_processing(self, samples=samples, K=self.post)
            self.Exp_post_processed = -self.stat.get_CVaR()
            self.Var_post_processed = self.stat.get_Variance()

    def local_opt(self, angles0):
        opt = self.optimizer[0](**self.optimizer[1])
        return opt.minimize(self.loss, x0=angles0)

    def loss(self, angles):
        """One energy evaluation, recorded in the current depth's OptResult (reference qaoa.py:885-935)."""
        if not self.backend.configuration().local:
            raise NotImplementedError
        angles = np.asarray(angles, dtype=np.float64)
        assert len(angles) == self.n_init + self.parametrized_circuit_depth * (self.n_gamma + self.n_beta)
        e, e2, cmin, cmax, _norm, cv = self._evaluate(angles)[0]
        self.stat.reset()
        self.stat.set_exact(e, e2, cmin, cmax, cvar_value=cv)
        self.optimization_results[self.current_depth + 1].add_iteration(angles.copy(), self.stat, self.shots)
        return -self.stat.get_CVaR()

    def interp(self, angles):
        """INTERP warm start p -> p+1, independently for every parameter index (reference
        qaoa.py:992-1027, https://doi.org/10.1103/PhysRevX.10.021067)."""
        per_layer = self.n_gamma + self.n_beta
        depth = (len(angles) - self.n_init) // per_layer
        head = np.asarray(angles[: self.n_init], dtype=np.float64)
        old = np.asarray(angles[self.n_init:], dtype=np.float64).reshape(depth, per_layer)
        padded = np.zeros((depth + 2, per_layer))
        padded[1:-1] = old
        w = np.arange(depth + 1, dtype=np.float64)[:, None]
        new = w / depth * padded[:-1] + (depth - w) / depth * padded[1:]
        return np.concatenate([head, new.reshape(-1)])

    def _eval_cost(self, angle_array):
        """Unrecorded evaluation (reference qaoa.py:1029-1066)."""
        if not self.backend.configuration().local:
            raise NotImplementedError
        angle_array = np.asarray(angle_array, dtype=np.float64)
        e, e2, cmin, cmax, _, cv = self._evaluate(angle_array)[0]
        self.stat.reset()
        self.stat.set_exact(e, e2, cmin, cmax, cvar_value=cv)
        return -self.stat.get_CVaR()

    def _grid_search_layer(self, prev_angles, angles):
        """Warm start for depth p from a 2-D grid over the NEW layer with layers 1..p-1 locked
        (reference qaoa.py:1068-1132).  The whole grid is one batched device call; the first
        strictly smallest value wins, as in the reference's sequential scan."""
        per_layer = self.n_gamma + self.n_beta
        new_depth = (len(prev_angles) - self.n_init) // per_layer + 1
        self.createParameterizedCircuit(new_depth)
        gamma_grid, beta_grid = self._grid(angles)
        LOG.info("Layer grid search at depth %d: %dx%d points", new_depth, len(gamma_grid), len(beta_grid))
        rows = self._vanilla_rows(gamma_grid, beta_grid, prefix=list(np.asarray(prev_angles, dtype=np.float64)))
        res = self._evaluate(rows)
        costs = -res[:, 5]
        k = int(np.argmin(costs))
        LOG.info("Layer grid search done, best cost: %.6f", -costs[k])
        return rows[k].copy()

    def hist(self, angles, shots):
        """Sampled histogram {bit-string (string[i] = qubit i): count} of the state at ``angles``
        (reference qaoa.py:1134-1172), drawn on the device from |psi|^2."""
        n_gamma = self.problem.get_num_parameters()
        n_beta = self.mixer.get_num_parameters()
        n_init = self.initialstate.get_num_parameters()
        depth = int((len(angles) - n_init) / (n_gamma + n_beta))
        self.createParameterizedCircuit(depth)
        eng = self._get_engine()
        seed = getattr(self, "sample_seed", None)       # set QAOA.sample_seed for reproducible histograms
        if seed is not None:
            self._sample_calls = getattr(self, "_sample_calls", 0) + 1
            idx = self._device_call("sample", self._engine_row(angles), int(shots), seed=[int(seed), self._sample_calls])
        else:
            idx = self._device_call("sample", self._engine_row(angles), int(shots))
        n = self.problem.N_qubits
        out = {}
        values, counts = np.unique(idx, return_counts=True)
        for x, c in zip(values.tolist(), counts.tolist()):
            out["".join("1" if (x >> i) & 1 else "0" for i in range(n))] = int(c)
        return out

    def random_init(self, gamma_bounds, beta_bounds, depth):
        gam = np.random.uniform(gamma_bounds[0], gamma_bounds[1], size=depth)
        bet = np.random.uniform(beta_bounds[0], beta_bounds[1], size=depth)
        out = np.empty(2 * depth, dtype=gam.dtype)
        out[0::2], out[1::2] = gam, bet
        return out

    def _fill_best_sols(self, result):
        """OptResult.BestSols is filled lazily: the exact engine reports the extreme COSTS with every evaluation, the
        bit-strings attaining them (Statistic.get_max_sols in the reference, qaoa/qaoa.py:89-91) cost one more
        evaluation with the state kept, so they are fetched only for the iterations that matter."""
        if not result.BestCost:
            return
        lowest = np.min(result.BestCost)
        n = self.problem.N_qubits
        eng = self._get_engine()
        for i in np.where(np.asarray(result.BestCost) == lowest)[0][:1]:
            if not result.BestSols[i] and hasattr(eng, "extreme_solutions"):
                idx, _cost = self._device_call("extreme_solutions", self._engine_row(result.angles[i]), want_max=True)
                result.BestSols[i] = ["".join("1" if (int(x) >> q) & 1 else "0" for q in range(n)) for x in idx]

    def _rows_need_lowering(self):
        from qaoa_b200.problems.base_problem import Problem

        gam = getattr(type(self.problem), "engine_gammas", Problem.engine_gammas)
        return type(self.mixer).engine_betas is not Mixer.engine_betas or gam is not Problem.engine_gammas

    def _engine_row(self, angles):
        """full flat angle vector -> what the device takes (shared mixer parameters expanded, cost parameters rescaled)"""
        row = np.asarray(angles, dtype=np.float64)
        if not self._rows_need_lowering():
            return row
        n_gamma, n_beta = self.problem.get_num_parameters(), self.mixer.get_num_parameters()
        n_init, per = self.initialstate.get_num_parameters(), n_gamma + n_beta
        parts = [row[:n_init]]
        for d in range((len(row) - n_init) // per):
            at = n_init + d * per
            gam = row[at: at + n_gamma]
            if hasattr(self.problem, "engine_gammas"):
                gam = np.asarray(self.problem.engine_gammas(gam), dtype=np.float64)
            parts += [gam, self.mixer.engine_betas(row[at + n_gamma: at + per])]
        return np.concatenate(parts)

    def get_optimal_solutions(self):
        sols, costs = [], []
        for d in self.optimization_results:
            self._fill_best_sols(self.optimization_results[d])
            s, c = self.optimization_results[d].get_best_solution()
            sols.append(s)
            costs.append(c)
        keep = np.where(np.asarray(costs) == np.min(costs))[0]
        return np.unique([x for i in keep for x in sols[i]])

    def validate_circuit(self, t=1, flip=True, atol=1e-8, rtol=1e-8):
        """The reference compares the circuit's diagonal with exp(i t cost) (qaoa.py:1217-1218 -> validation.py:16-79).
        Here the phase separator IS the diagonal exp(i t c_dev) of the device's cost diagonal; it is compared with
        exp(i t problem.cost(string)) for every bit-string, same arguments and report keys (<= 16 qubits)."""
        from qaoa_b200.utils import validation

        return validation.check_phase_separator_exact_qaoa(self, t=t, flip=flip, atol=atol, rtol=rtol)
