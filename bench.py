#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 QAOA engine (contract in the task statement).

Workload (BASELINE.json metric "ms per QAOA energy eval (MaxCut n=32,p=5)"): MaxCut on
nx.random_regular_graph(3, n, seed=42), X mixer, |+> initial state, depth p=5, complex64
statevector.  One step = one energy evaluation <gamma,beta|H_P|gamma,beta>.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 32] [--p 5] [--impl reference]

N = 1 : n = 32 on one B200.  MaxCut + X mixer + |+> keeps psi(x) = psi(~x), so the engine stores the symmetric
        half of the register (2^31 amplitudes = 16 GiB, far larger than the 126 MB L2: nothing to flush) and pairs
        every element with its mirror image for the top qubit (tile_kernel.cuh "symmetric registers"); the same
        evaluation on the full 32 GiB register (option symmetric=0) is reported under "full_register".  Side legs:
        "config4" (BASELINE config 4, n=30 p=3: Grover+Dicke, multi-angle X, XY), "landscape" (config 2), "optimize"
        (config 3 through the public API), "cpu_baseline".
N > 1 : launched under torchrun, one rank per GPU; STRONG scaling of the SAME workload: the one n = 32, p = 5 register
        sharded on its top log2(N) qubits over the N GPUs (qaoa_b200/sharded.py): mixer gates on the global qubits as
        cross passes over NVLink peer memory, energy by a scalar all-reduce; value = ms per evaluation of that register,
        max over ranks.  "parity" compares the sharded path with the C oracle (n = 28) in the same job; "config5" is
        BASELINE config 5 (n = 36, p = 3 on 4 / 8 GPUs; n = 34 on 2) with the closed-form p = 1 energy as its check.
--impl reference : times the CPU restatement of the reference path (oracle/c/qaoa_ref.c: OpenMP over the host cores,
        cache-blocked gate groups like a gate-fusing simulator) on a bounded sample (smaller n), scaled to the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ms_per_energy_eval_maxcut_n32_p5"
UNIT = "ms"


def workload(n):
    import networkx as nx
    from qaoa_b200.utils.graphutils import GraphHandler

    G = nx.random_regular_graph(3 if n % 2 == 0 else 4, n, seed=42)   # odd n: no 3-regular graph exists
    gh = GraphHandler(G, check_isomorphism=False)
    return gh.edge_list()


def angles_for(p, seed):
    return np.random.default_rng(seed).uniform(0, 2 * np.pi, size=2 * p)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def cpu_time_one(c_oracle, ns, p, ang, reps):
    """seconds per evaluation of the C restatement (cache-blocked X layers, u8 diagonal, complex64) at n = ns"""
    d8 = c_oracle.maxcut_diag_u8(ns, workload(ns))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        c_oracle.run(ns, d8, ang, "complex64", blocked=True)
        ts.append(time.perf_counter() - t0)
    return ts


def pick_cpu_n(c_oracle, p, n_target, budget_s, evals, requested):
    """Largest sample size <= n_target whose `evals` evaluations fit the CPU time budget and the host memory
    (calibrated at n = 24; work ~ 2^n * n; memory 9 bytes per amplitude: complex64 state + u8 diagonal)."""
    if requested > 0:
        return requested
    ns = min(24, n_target)
    ang = angles_for(p, 1234)
    cpu_time_one(c_oracle, ns, p, ang, 1)
    t = min(cpu_time_one(c_oracle, ns, p, ang, 2))
    mem = c_oracle.host_mem_available_gib() * 2.0 ** 30
    best = ns
    for cand in range(ns + 1, n_target + 1):
        fits = 9.5 * 2.0 ** cand < 0.8 * mem
        if fits and t * (2.0 ** cand * cand) / (2.0 ** ns * ns) * evals <= budget_s:
            best = cand
    return best


def cpu_sample(c_oracle, ns, p, n, reps, ang):
    """times `reps` evaluations at n = ns and scales to n: (list of scaled ms, cpu_baseline object)"""
    ts = cpu_time_one(c_oracle, ns, p, ang, reps)
    ms_s = 1e3 * sum(ts) / len(ts)
    scale = (2.0 ** n * n) / (2.0 ** ns * ns)       # 2^n amplitudes x n gates per layer
    cores = c_oracle.num_threads()
    if ns == n:
        sample = "the workload itself: n=%d p=%d complex64, %d evaluation(s), %.0f ms each, no scaling" % (n, p, reps, ms_s)
    else:
        sample = ("n=%d (1/%d of the amplitudes) p=%d complex64: %.1f ms/eval measured, scaled by 2^n*n to n=%d"
                  % (ns, 2 ** (n - ns), p, ms_s, n))
    sample += ("; oracle/c/qaoa_ref.c, cache-blocked gate groups (14 low qubits per block, then <= 5 qubits per sweep: the "
               "grouping a gate-fusing simulator applies), OpenMP on %d threads; qiskit-aer itself is not installable here" % cores)
    return [1e3 * t * scale for t in ts], {"value": ms_s * scale, "unit": UNIT, "cores": cores, "kind": "port",
                                            "sample": sample, "same_config": ns == n}


def run_reference(args, rank):
    """CPU arm: the C restatement of the reference path on the host cores; the workload itself when K evaluations fit
    the time budget (n = 32 complex64 needs 36 GiB of host memory), else a bounded sample (smaller n) scaled to it."""
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.build()
    c_oracle.use_all_cores()
    ang = angles_for(args.p, 1234)
    ns = pick_cpu_n(c_oracle, args.p, args.n, 200.0, args.steps + min(args.warmup, 1), args.cpu_n)
    if min(args.warmup, 1):
        cpu_time_one(c_oracle, ns, args.p, ang, 1)
    ms, base = cpu_sample(c_oracle, ns, args.p, args.n, args.steps, ang)
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": "MaxCut %d-regular n=%d seed=42, X mixer, Plus, p=%d, complex64; 1 step = 1 energy evaluation"
                               % (3 if args.n % 2 == 0 else 4, args.n, args.p),
                   "reference_arm": "CPU restatement of the reference path (kind: port), all %d host threads" % base["cores"]},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def landscape_leg(device, world=1):
    """BASELINE config 2: MaxCut 3-regular n=24, X mixer, p=1, 100x100 (gamma, beta) grid through the public
    QAOA.sample_cost_landscape call (host grid in, four host arrays out).  With several GPUs every rank holds a
    replica and the grid points are dealt out over the ranks (B200Backend(split_batches=True)); all ranks call this."""
    import networkx as nx
    from qaoa_b200 import QAOA, initialstates, mixers, problems
    from qaoa_b200.qaoa import B200Backend

    n, g = 24, 100
    G = nx.random_regular_graph(3, n, seed=42)
    q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(),
             backend=B200Backend(device=device, dtype="complex64", split_batches=world > 1))
    q.sample_cost_landscape({"gamma": [0, 2 * np.pi, 8], "beta": [0, 2 * np.pi, 8]})   # warm-up: plan, streams
    grid = {"gamma": [0, 2 * np.pi, g], "beta": [0, 2 * np.pi, g]}
    t0 = time.perf_counter()
    q.sample_cost_landscape(grid)
    dt = time.perf_counter() - t0
    sym = q._get_engine().counter("symmetric_plans") >= 1
    # FIRST writes the stored register, FINAL reads it, one diagonal byte per amplitude each
    bytes_per_point = 2 * (8.0 + 1.0) * 2.0 ** (n - 1 if sym else n)
    return {"workload": "MaxCut 3-regular n=%d seed=42, X, Plus, p=1, %dx%d grid, complex64%s" % (
                n, g, g, "" if world == 1 else ", grid points dealt out over %d GPUs" % world),
            "register": "symmetric half (2^%d amplitudes per point)" % (n - 1) if sym else "full",
            "points_per_s": g * g / dt, "seconds": dt, "min_exp": float(q.Exp_sampled_p1.min()),
            "algorithmic_gb_per_point": bytes_per_point / 1e9,
            "achieved_gbs": bytes_per_point * g * g / dt / 1e9}


def optimize_leg(device, n=32, depth=5, maxiter=40, grid=12):
    """BASELINE config 3 as the reference states it: MaxCut 3-regular n=32, X mixer, Plus, `optimize(depth=5)`, complex64,
    through the public QAOA API (landscape seed, COBYLA per depth -- capped at `maxiter` evaluations so that the leg takes
    seconds; the reference default is 1000 -- INTERP between depths).  ms per evaluation is what the optimiser sees: host
    angles in, host statistics out, Python bookkeeping and scipy's COBYLA step included."""
    import networkx as nx
    from qaoa_b200 import QAOA, initialstates, mixers, problems
    from qaoa_b200.qaoa import B200Backend
    from qaoa_b200.utils import COBYLA

    G = nx.random_regular_graph(3, n, seed=42)
    t_all = time.perf_counter()
    q = QAOA(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus(),
             backend=B200Backend(device=device, dtype="complex64"), optimizer=[COBYLA, {"maxiter": maxiter}])
    spec = {"gamma": [0, 2 * np.pi, grid], "beta": [0, 2 * np.pi, grid]}
    t0 = time.perf_counter()
    q.sample_cost_landscape(spec)            # includes lowering, the plan and the diagonal streams
    t_land = time.perf_counter() - t0
    per_depth = []
    for d in range(1, depth + 1):
        t0 = time.perf_counter()
        q.optimize(depth=d, angles=spec)
        dt = time.perf_counter() - t0
        r = q.optimization_results[d]
        per_depth.append({"depth": d, "evaluations": r.num_fval(), "seconds": dt,
                          "ms_per_evaluation": 1e3 * dt / max(1, r.num_fval()), "best_exp": float(q.get_Exp(d))})
    out = {"workload": "MaxCut 3-regular n=%d seed=42, X, Plus, complex64: %dx%d landscape, then optimize(depth=%d) with "
                       "COBYLA(maxiter=%d) through the public QAOA API" % (n, grid, grid, depth, maxiter),
           "total_seconds": time.perf_counter() - t_all, "landscape_seconds": t_land, "landscape_points": grid * grid,
           "per_depth": per_depth}
    eng_obj = q._get_engine()
    if hasattr(eng_obj, "close"):
        eng_obj.close()
    return out


PASS_NAMES = ["interpreter", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2", "Final2",
              "BoundX", "FinalX", "Bound2L4", "Final2L4"]


def kernel_name(i):
    return "tile_pass_kernel (interpreter)" if i == 0 else "tile_prog_kernel<Prog%s>" % PASS_NAMES[i]


def closed_form_p1(n, edges, gamma, beta):
    """<cost> of p = 1 MaxCut QAOA on an unweighted graph (reference sign conventions): the
    size-independent check of the registers no CPU oracle can hold (n = 34, 36)."""
    adj = [set() for _ in range(n)]
    for u, v, _ in edges:
        adj[u].add(v)
        adj[v].add(u)
    tot = 0.0
    for u, v, _ in edges:
        d, e = len(adj[u]) - 1, len(adj[v]) - 1
        f = len(adj[u] & adj[v])
        tot += 0.5 + 0.25 * np.sin(4 * beta) * np.sin(gamma) * (np.cos(gamma) ** d + np.cos(gamma) ** e) \
            - 0.25 * np.sin(2 * beta) ** 2 * np.cos(gamma) ** (d + e - 2 * f) * (1 - np.cos(2 * gamma) ** f)
    return tot


def sharded_leg(n, p, sym, steps, warmup, local_rank, world, dist, torch, oracle_check=False):
    """ONE n-qubit register sharded on its top log2(world) qubits over the GPUs (qaoa_b200/sharded.py): `steps` timed
    evaluations between barriers, max over ranks; per-pass CUDA events; the closed-form p = 1 energy as a check; with
    oracle_check the same angles through the C oracle on rank 0 (sizes the host can hold)."""
    from qaoa_b200.sharded import DistShardedEngine

    edges = workload(n)
    E = DistShardedEngine(n, dtype="complex64", device=local_rank, symmetric=sym)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    E.set_option("time_passes", 1)
    ang = angles_for(p, 1234)
    a1 = np.array([0.4, 0.3])
    got1 = E.energy(a1)[0]
    cf = closed_form_p1(n, edges, a1[0], a1[1])
    for _ in range(warmup):
        out = E.energy(ang)
    E.eng.counter("reset")
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = E.energy(ang)
    torch.cuda.synchronize()
    dist.barrier()
    ms_wall = 1e3 * (time.perf_counter() - t0) / steps
    stats = {nm: (E.eng.counter("pass_ms_%d" % i), E.eng.counter("pass_count_%d" % i), E.eng.counter("pass_bytes_%d" % i))
             for i, nm in enumerate(PASS_NAMES)}
    ms_kernels = sum(v[0] for v in stats.values()) / steps
    launches = E.eng.counter("launches") / steps
    t = torch.tensor([ms_wall, ms_kernels], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_wall, ms_kernels = float(t[0]), float(t[1])
    lq = E.eng.local_qubits
    shard_bytes = 8.0 * 2.0 ** lq
    local = {k: v for k, v in stats.items() if v[1] and k not in ("BoundX", "FinalX")}
    dom = max(local, key=lambda k: local[k][0])
    d_ms, d_cnt, d_bytes = local[dom]
    peak, peak_src = peaks()
    achieved = (d_bytes / d_cnt) / (d_ms / d_cnt * 1e-3) / 1e9
    cross_ms = (stats["BoundX"][0] + stats["FinalX"][0]) / steps
    nvlink = None
    chunks = int(getattr(E, "last_chunks", 1))
    if stats["BoundX"][1]:
        # a whole BOUNDX pass = `chunks` launches in the chunk pipeline; read-modify-write: loads + stores, each (1 - 1/N) remote
        ms = stats["BoundX"][0] / stats["BoundX"][1] * chunks
        per_dir = 2.0 * (1.0 - 1.0 / world) * shard_bytes
        nvlink = {"kernel": "tile_prog_kernel<ProgBoundX>", "ms_per_pass": round(ms, 3),
                  "gbs_per_direction": round(per_dir / (ms * 1e-3) / 1e9, 1), "peer_copy_reference_gbs": 770.0}
    res = {"n": n, "p": p, "gpus": world, "register": "symmetric half, twisted shards" if sym else "full",
           "stored_gib_per_gpu": shard_bytes / 2 ** 30, "ms_per_eval": round(ms_wall, 3), "kernels_ms": round(ms_kernels, 3),
           "cross_pass_ms": round(cross_ms, 3), "launches": int(launches), "pipeline_chunks": chunks, "energy": float(-out[0]), "norm": float(out[4]),
           "p1_check": {"engine": float(got1), "closed_form": float(cf), "rel_err": float(abs(got1 - cf) / abs(cf))},
           "dominant_local_pass": {"kernel": "tile_prog_kernel<Prog%s>" % dom, "ms_per_launch": round(d_ms / d_cnt, 3),
                                   "achieved_gbs": round(achieved, 1), "frac": round(achieved / peak, 4)},
           "nvlink": nvlink,
           "passes_ms": {k: [round(v[0] / steps, 3), int(v[1] / steps)] for k, v in stats.items() if v[1]}}
    if oracle_check:
        ref = None
        if dist.get_rank() == 0:
            from oracle import c_oracle
            c_oracle.build()
            c_oracle.use_all_cores()
            ref = c_oracle.run(n, c_oracle.maxcut_diag_u8(n, edges), ang, "complex128", blocked=True)
        dist.barrier()
        if ref is not None:
            res["oracle"] = {"energy": float(-ref[0]), "rel_err": float(abs(out[0] - ref[0]) / abs(ref[0])),
                             "rel_err_E2": float(abs(out[1] - ref[1]) / abs(ref[1])),
                             "extremes_equal": bool(out[2] == ref[2] and out[3] == ref[3]), "tolerance": 1e-4,
                             "checker": "oracle/c/qaoa_ref.c complex128"}
    E.close()
    return res, ang, (achieved, peak, peak_src, dom, d_ms, d_cnt)


def run_multi(args, rank, world, local_rank):
    """N > 1 (torchrun, one rank per GPU): STRONG scaling of the N = 1 workload -- the same n-qubit, depth-p register
    sharded over the N GPUs; parity of the sharded path against the C oracle in the same job; BASELINE config 5."""
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, p = args.n, args.p
    sym = bool(args.symmetric) and n % 2 == 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    head, ang, roof = sharded_leg(n, p, sym, args.steps, args.warmup, local_rank, world, dist, torch)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # parity of the multi-process path (CUDA IPC peers, NVLink cross passes, all-reduce) against the C oracle
    parity = []
    n_par = min(28, n - 4) if n >= 30 else max(n - 4, 20)
    n_par -= n_par & 1
    for s in ([True, False] if sym else [False]):
        r, _, _ = sharded_leg(n_par, p, s, 1, 1, local_rank, world, dist, torch, oracle_check=True)
        if rank == 0:
            o = r["oracle"]
            parity.append({"n": n_par, "p": p, "register": "symmetric" if s else "full", "energy": r["energy"],
                           "oracle": o["energy"], "rel_err": o["rel_err"], "extremes_equal": o["extremes_equal"],
                           "p1_closed_form_rel_err": r["p1_check"]["rel_err"],
                           "ok": bool(o["rel_err"] < 1e-4 and o["rel_err_E2"] < 2e-4 and o["extremes_equal"]
                                      and r["p1_check"]["rel_err"] < 1e-4)})
    config5 = None
    if n == 32 and not args.no_config5:
        n5 = 34 if world == 2 else 36        # n = 36 needs 128 GiB of state + streams per GPU on 2 GPUs
        config5, _, _ = sharded_leg(n5, 3, True, max(2, min(args.steps, 5)), 2, local_rank, world, dist, torch)
        config5["workload"] = "BASELINE config 5: MaxCut 3-regular n=%d seed=42, X, Plus, p=3, complex64, sharded over %d GPUs%s" % (
            n5, world, "" if n5 == 36 else " (n=36 does not fit 2 GPUs)")
    landscape = None
    if not args.no_landscape:
        dist.barrier()
        landscape = landscape_leg(local_rank, world)
        dist.barrier()
    if rank == 0:
        achieved, peak, peak_src, dom, d_ms, d_cnt = roof
        gb = world.bit_length() - 1
        line = {
            "metric": METRIC, "value": head["ms_per_eval"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_eval"], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "config": {"workload": "MaxCut 3-regular n=%d seed=42, X mixer, Plus, p=%d, complex64: the N=1 register sharded on its "
                                   "top %d qubits over %d GPUs; 1 step = 1 energy evaluation" % (n, p, gb, world),
                       "register": head["register"], "stored_gib_per_gpu": head["stored_gib_per_gpu"],
                       "l2": "shard >> 126 MB L2, no flush needed",
                       "value_is": "wall ms per evaluation through DistShardedEngine.energy between barriers, max over ranks"},
            "e2e": {"value": head["ms_per_eval"], "unit": UNIT, "h2d_bytes_per_step": int(ang.nbytes + 4096) * world,
                    "d2h_bytes_per_step": 40 * world, "api": "qaoa_shard_* (C ABI) under torch.distributed; same call as value"},
            "gpu_launches": head["launches"] * world,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "tile_prog_kernel<Prog%s>" % dom, "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src, "traffic": None,
                         "launches_timed": int(d_cnt), "ms_per_launch": round(d_ms / d_cnt, 3),
                         "note": "dominant LOCAL pass on rank 0 (chunk launches share the SMs with a cross pass); cross passes are NVLink-bound"},
            "sharded": {k: head[k] for k in ("kernels_ms", "cross_pass_ms", "pipeline_chunks", "nvlink", "passes_ms")},
        }
        if landscape is not None:
            line["landscape"] = {k: landscape[k] for k in ("workload", "points_per_s")}
        if config5 is not None:
            line["config5"] = {k: config5[k] for k in ("workload", "ms_per_eval", "kernels_ms", "cross_pass_ms", "pipeline_chunks",
                                                       "stored_gib_per_gpu", "nvlink", "energy", "norm", "p1_check")}
        # the checks go LAST so that a tail of the line shows them: energy of this run (equal to the N=1 line's), the
        # closed-form p=1 energy, and the sharded path against the C oracle in this very job
        line["energy"] = head["energy"]
        line["norm"] = head["norm"]
        line["p1_check"] = head["p1_check"]
        line["parity"] = parity
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def qubo_instance(n):
    rs = np.random.RandomState(42)
    return rs.rand(n, n), 3 * (rs.rand(n) - 0.5)


def portfolio_instance(n):
    """the reference's PortfolioOptimization cost (portfolio_problem.py:48-71) as a QUBO: risk 0.5, budget n/3, penalty 0"""
    rng = np.random.RandomState(0)
    A = rng.randn(n, 101)
    cov, mu = np.cov(A), A.mean(axis=1)
    return 0.5 * cov, -mu, 0.0, n // 3


def config4_leg(device, n=30, p=3, reps=3):
    """BASELINE config 4 (n=30, p=3, complex64) through the C ABI: Portfolio + Grover(Dicke) + Dicke in the compact
    feasible-subspace register and on the full register, QUBO + multi-angle X + Plus, QUBO + XY ring + Dicke.
    Per case: CUDA-event ms per evaluation, algorithmic bytes (engine counter), GB/s against the measured HBM peak."""
    from qaoa_b200 import engine as eng

    peak, _ = peaks()
    out = []

    def timed(label, kernel, E, ang):
        E.set_option("time_passes", 1)
        E.energy(ang)
        E.energy(ang)
        E.counter("reset")
        ms = []
        for _ in range(reps):
            E.energy_batch_resident(ang[None, :])
            res = E.fetch_results(1)[0]
            ms.append(E.counter("last_device_ms"))
        gb = E.counter("bytes_alg") / reps / 1e9
        m = float(np.mean(ms))
        out.append({"case": label, "kernel": kernel, "ms_per_eval": round(m, 3), "launches": int(E.counter("launches") / reps),
                    "algorithmic_gb": round(gb, 3), "achieved_gbs": round(gb / m * 1e3, 1), "frac": round(gb / m * 1e3 / peak, 4),
                    "energy": float(res[0]), "norm": float(res[4])})
        ps = {PASS_NAMES[i]: (E.counter("pass_ms_%d" % i), E.counter("pass_count_%d" % i), E.counter("pass_bytes_%d" % i))
              for i in range(len(PASS_NAMES))}
        if any(v[1] for v in ps.values()):
            out[-1]["passes"] = {k: {"ms_per_launch": round(v[0] / v[1], 3), "launches_per_eval": v[1] / reps,
                                     "frac": round(v[2] / v[0] / 1e6 / peak, 4)} for k, v in ps.items() if v[1]}
        E.close()

    Qp, cp, bp, k = portfolio_instance(n)
    rng = np.random.default_rng(1234)
    ang2 = rng.uniform(0, 2 * np.pi, size=2 * p)
    for fs in (1, 0):
        E = eng.Engine(n, "complex64", device=device)
        E.set_option("feasible_subspace", fs)
        E.set_cost_qubo(Qp, cp, bp)
        E.set_initial(eng.INIT_DICKE, k)
        E.set_mixer(eng.MIXER_GROVER, sub_kind=eng.INIT_DICKE, sub_k=k)
        a = ang2.copy()
        a[0::2] *= 3.0
        timed("Portfolio + Grover(Dicke %d) + Dicke %d, %s" % (k, k, "compact C(n,k) register" if fs else "full 2^n register"),
              "grover_pass_c64_kernel", E, a)
    Q, c = qubo_instance(n)
    E = eng.Engine(n, "complex64", device=device)
    E.set_cost_qubo(Q, c, 0.0)
    E.set_mixer(eng.MIXER_XMULTI)
    am = rng.uniform(0, 2 * np.pi, size=p * (1 + n))
    am[0::1 + n] *= 0.02
    timed("QUBO + XMultiAngle + Plus", "tile_prog_kernel (U32FX diagonal)", E, am)
    E = eng.Engine(n, "complex64", device=device)
    E.set_cost_qubo(Q, c, 0.0)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_XY, pairs=[[i, (i + 1) % n] for i in range(n)])
    a = ang2.copy()
    a[0::2] *= 0.02
    timed("QUBO + XY ring + Dicke %d, compact C(n,k) register" % k, "xy_compact_gate_kernel", E, a)
    E = eng.Engine(n, "complex64", device=device)
    E.set_option("feasible_subspace", 0)
    E.set_cost_qubo(Q, c, 0.0)
    E.set_initial(eng.INIT_DICKE, k)
    E.set_mixer(eng.MIXER_XY, pairs=[[i, (i + 1) % n] for i in range(n)])
    timed("QUBO + XY ring + Dicke %d, full 2^n register" % k, "xy_tile_kernel", E, a)
    return out


def run_single(args, local_rank):
    import torch

    torch.cuda.set_device(local_rank)
    from qaoa_b200 import engine as eng

    n, p = args.n, args.p
    edges = workload(n)
    E = eng.Engine(n, "complex64", device=local_rank)
    E.set_option("symmetric", args.symmetric)
    if args.column >= 0:
        E.set_option("column_kernels", args.column)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    ang = angles_for(p, 1234)
    E.set_option("time_passes", 1)

    # ---- device-resident metric: CUDA-event time of the evaluation's kernels -------------------
    for _ in range(args.warmup):
        E.energy(ang)
    E.counter("reset")
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    dev_ms = []
    for _ in range(args.steps):
        E.energy_batch_resident(ang[None, :])
        dev_ms.append(E.counter("last_device_ms"))
    torch.cuda.synchronize()
    launches = E.counter("launches") / args.steps
    pass_stats = {i: (E.counter("pass_ms_%d" % i), E.counter("pass_count_%d" % i), E.counter("pass_bytes_%d" % i))
                  for i in range(len(PASS_NAMES))}
    out = E.fetch_results(1)[0]
    sym = E.counter("symmetric_plans") >= 1
    stored_amps = 2.0 ** (n - 1) if sym else 2.0 ** n

    # ---- end to end through the C ABI: host angles in, host result out ------------------------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        E.energy(ang)
    torch.cuda.synchronize()
    wall_e2e = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    a1 = np.array([0.4, 0.3])
    p1 = E.energy(a1)[0]
    cf = closed_form_p1(n, edges, a1[0], a1[1]) if n % 2 == 0 else None

    full_register = None
    if sym:
        E.close()
        E = eng.Engine(n, "complex64", device=local_rank)
        E.set_option("symmetric", 0)
        E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
        for _ in range(2):
            E.energy(ang)
        ms_full = []
        for _ in range(3):
            E.energy_batch_resident(ang[None, :])
            out_full = E.fetch_results(1)[0]
            ms_full.append(E.counter("last_device_ms"))
        full_register = {"value": float(np.mean(ms_full)), "unit": UNIT, "energy": float(-out_full[0]),
                         "note": "same evaluation with option symmetric=0: all 2^%d amplitudes stored (%d GiB)" % (n, 8 * 2 ** n >> 30)}
    E.close()

    def side(fn, *a, **k):
        try:
            return fn(*a, **k)
        except Exception as exc:  # noqa: BLE001  (a side leg must never cost the bench line)
            return {"failed": repr(exc)}

    def parity_leg(n_par=28):
        """the engine (symmetric and full register) against the C oracle in this very job, same angles, n = 28"""
        from oracle import c_oracle
        c_oracle.build()
        c_oracle.use_all_cores()
        e28 = workload(n_par)
        ref = c_oracle.run(n_par, c_oracle.maxcut_diag_u8(n_par, e28), ang, "complex128", blocked=True)
        res = []
        for symm in (1, 0):
            Ep = eng.Engine(n_par, "complex64", device=local_rank)
            Ep.set_option("symmetric", symm)
            Ep.set_cost_maxkcut(n_par, e28, 1, np.array([0, 1], dtype=np.uint8))
            got = Ep.energy(ang)
            Ep.close()
            rel = float(abs(got[0] - ref[0]) / abs(ref[0]))
            res.append({"n": n_par, "p": p, "register": "symmetric" if symm else "full", "energy": float(-got[0]),
                        "oracle": float(-ref[0]), "rel_err": rel, "extremes_equal": bool(got[2] == ref[2] and got[3] == ref[3]),
                        "ok": bool(rel < 1e-4 and abs(got[1] - ref[1]) < 2e-4 * abs(ref[1]) and got[2] == ref[2] and got[3] == ref[3])})
        return res

    parity = None if args.no_cpu_baseline else side(parity_leg)
    landscape = None if args.no_landscape else side(landscape_leg, local_rank, 1)
    config4 = None if args.no_config4 else side(config4_leg, local_rank)
    optimize = None if args.no_optimize else side(optimize_leg, local_rank, n=n)

    ms_dev = float(np.mean(dev_ms))
    ms_e2e = 1e3 * wall_e2e / args.steps
    peak, peak_src = peaks()
    dom = max(pass_stats, key=lambda i: pass_stats[i][0])
    d_ms, d_cnt, d_bytes = pass_stats[dom]
    achieved = (d_bytes / d_cnt) / (d_ms / d_cnt * 1e-3) / 1e9 if d_cnt else 0.0
    total_pass_ms = sum(v[0] for v in pass_stats.values())
    total_bytes = sum(v[2] for v in pass_stats.values())
    P = max(1, -(-(n - 4) // 10))            # SURVEY 8d pass model: s * 2^n * (2 p P - 2), P = ceil((n-4)/10)
    survey_bytes = 8.0 * 2.0 ** n * (2 * p * P - 2)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        ent = tj["kernels"].get(PASS_NAMES[dom])
        if ent and d_cnt:
            if ent.get("n") == n and ent.get("symmetric") == bool(sym):
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"] + " (measured at this size)"
            else:   # a ratio measured at a smaller register, scaled: say so
                traffic = ent["dram_over_algorithmic"] * d_bytes / d_cnt
                traffic_src = ent["source"] + " (ncu dram bytes / algorithmic bytes at n=%d, times this launch's algorithmic bytes)" % ent["n"]
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": METRIC, "value": ms_dev, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": "MaxCut 3-regular n=%d seed=42, X mixer, Plus, p=%d, complex64; 1 step = 1 energy evaluation" % (n, p),
                   "register": ("symmetric half register: psi(x) = psi(~x) for MaxCut + X + Plus, 2^%d stored amplitudes, "
                                "the top qubit's rotation pairs every element with its mirror image" % (n - 1)) if sym
                               else "full register, 2^%d stored amplitudes" % n,
                   "l2": "state is %d GiB >> 126 MB L2, no flush needed" % (int(8 * stored_amps) >> 30),
                   "energy": float(-out[0]), "norm": float(out[4]),
                   "p1_check": None if cf is None else {"engine": float(p1), "closed_form": float(cf),
                                                        "rel_err": float(abs(p1 - cf) / abs(cf))},
                   "oracle_parity_at_this_size": "tests/test_gpu_fullsize.py (n=32 p=5 vs the C oracle, 1e-4)"},
        "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": int(ang.nbytes), "d2h_bytes_per_step": 40,
                "api": "qaoa_energy (C ABI via ctypes): host angle vector in, host {E,E2,min,max,norm} out"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": kernel_name(dom), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "launches_timed": int(d_cnt), "ms_per_launch": d_ms / d_cnt if d_cnt else None,
                     "kernel_share_of_step": d_ms / total_pass_ms if total_pass_ms else None,
                     "whole_eval_actual_gbs": total_bytes / (total_pass_ms * 1e-3) / 1e9 if total_pass_ms else None,
                     "algorithmic_bytes_per_eval": total_bytes / args.steps,
                     "whole_eval_frac_of_peak": total_bytes / (total_pass_ms * 1e-3) / 1e9 / peak if total_pass_ms else None,
                     "survey_model_bytes_per_eval": survey_bytes},
        "passes": {kernel_name(i): {"ms_per_launch": round(v[0] / v[1], 3), "launches_per_eval": v[1] / args.steps,
                                    "gbs": round(v[2] / v[0] / 1e6, 1), "frac": round(v[2] / v[0] / 1e6 / peak, 4)}
                   for i, v in pass_stats.items() if v[1]},
    }
    if full_register is not None:
        line["full_register"] = full_register
    if landscape is not None:
        line["landscape"] = landscape
    if config4 is not None:
        line["config4"] = config4
    if optimize is not None:
        line["optimize"] = optimize
    if not args.no_cpu_baseline:
        from oracle import c_oracle
        try:
            c_oracle.build()
            c_oracle.use_all_cores()
            ns = pick_cpu_n(c_oracle, p, n, 20.0, 4, args.cpu_n)
            line["cpu_baseline"] = cpu_sample(c_oracle, ns, p, n, 3, ang)[1]
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % exc}
    if parity is not None:
        line["parity"] = parity          # last, so that a tail of the line shows it
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--p", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-n", type=int, default=0, dest="cpu_n",
                    help="sample size of the CPU legs (0: the largest that fits the time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-landscape", action="store_true")
    ap.add_argument("--no-optimize", action="store_true", help="skip the BASELINE config 3 leg (optimize to depth 5)")
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE config 4 leg (n=30 Grover / multi-angle / XY)")
    ap.add_argument("--no-config5", action="store_true", help="N>1: skip the BASELINE config 5 leg (n=36, p=3)")
    ap.add_argument("--symmetric", type=int, default=1, help="0: always store the full register")
    ap.add_argument("--column", type=int, default=-1, help="engine option column_kernels (-1: the engine's default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
    elif world > 1:
        run_multi(args, rank, world, local_rank)
    else:
        run_single(args, local_rank)


if __name__ == "__main__":
    main()
