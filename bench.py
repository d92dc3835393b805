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


def pick_cpu_n(c_oracle, p, n_target, budget_s, evals, requested):
    """Largest sample size whose `evals` evaluations fit the CPU time budget: calibrate at n=22, work ~ 2^n * n."""
    if requested > 0:
        return requested
    ns = 22
    diag = c_oracle.maxcut_diag(ns, workload(ns))
    ang = angles_for(p, 1234)
    c_oracle.energy(ns, diag, ang, "complex64")
    t0 = time.perf_counter()
    c_oracle.energy(ns, diag, ang, "complex64")
    t = time.perf_counter() - t0
    best = ns
    for cand in range(ns + 1, min(n_target, 30) + 1):       # n = 30: 8 GiB state + 8 GiB diagonal on the host
        if t * (2.0 ** cand * cand) / (2.0 ** ns * ns) * evals <= budget_s:
            best = cand
    return best


def run_reference(args, rank):
    """CPU arm: oracle/c/qaoa_ref.c on the host cores, bounded sample, scaled to the workload."""
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.build()
    c_oracle.use_all_cores()
    ns = pick_cpu_n(c_oracle, args.p, args.n, 60.0, args.steps + 1, args.cpu_n)
    edges = workload(ns)
    diag = c_oracle.maxcut_diag(ns, edges)
    ang = angles_for(args.p, 1234)
    for _ in range(max(1, min(args.warmup, 1))):
        c_oracle.energy(ns, diag, ang, "complex64")
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        c_oracle.energy(ns, diag, ang, "complex64")
        ts.append(time.perf_counter() - t0)
    ms_sample = 1e3 * sum(ts) / len(ts)
    scale = (2.0 ** args.n * args.n) / (2.0 ** ns * ns)   # work ~ 2^n amplitudes x (n + 1) sweeps per layer
    value = ms_sample * scale
    cores = c_oracle.num_threads()
    sample = ("n=%d (1/%d of the amplitudes), p=%d, complex64, %.1f ms/eval measured, scaled by 2^n*n to n=%d "
              "(n=%d needs 2 x 32 GiB of host memory)" % (ns, 2 ** (args.n - ns), args.p, ms_sample, args.n, args.n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_sample, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": "MaxCut %d-regular n=%d seed=42, X mixer, Plus, p=%d, complex64"
                               % (3 if args.n % 2 == 0 else 4, args.n, args.p),
                   "reference_arm": "CPU restatement of the reference path (qiskit-aer itself is not installable: "
                                    "no network); OpenMP over %d host threads" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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


def run_sharded(args, rank, world, local_rank, dist, torch):
    """N > 1: ONE register of n + log2(N) qubits sharded over the N GPUs (weak scaling, 2^n amplitudes per GPU)."""
    from qaoa_b200.sharded import DistShardedEngine

    gb = world.bit_length() - 1
    n, p = args.n + gb, args.p
    sym = bool(args.symmetric)
    if sym:
        # symmetric register (MaxCut + X + |+>, n even): 2^(n-1) stored amplitudes over the N GPUs.  N = 4, 8: BASELINE
        # config 5 itself (n = 36, p = 3); N = 2: n = 34 (n = 36 needs 128 GiB of state + 64 GiB of diagonal streams per GPU)
        if args.n == 32:
            n, p = (34 if world == 2 else 36), 3
        else:   # reduced sizes (quick checks): the next even n
            n = n + (n & 1)
    edges = workload(n)
    E = DistShardedEngine(n, dtype="complex64", device=local_rank, symmetric=sym)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    E.set_option("time_passes", 1)
    ang = angles_for(p, 1234)
    for _ in range(args.warmup):
        out = E.energy(ang)
    E.eng.counter("reset")
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = E.energy(ang)
    torch.cuda.synchronize()
    dist.barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    names = ["interpreter", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2", "Final2",
             "BoundX", "FinalX", "Bound2L4", "Final2L4"]
    stats = {nm: (E.eng.counter("pass_ms_%d" % i), E.eng.counter("pass_count_%d" % i), E.eng.counter("pass_bytes_%d" % i))
             for i, nm in enumerate(names)}
    ms_dev = sum(v[0] for v in stats.values()) / args.steps
    launches = E.eng.counter("launches") / args.steps
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    peak, peak_src = peaks()
    local = {k: v for k, v in stats.items() if v[1] and k not in ("BoundX", "FinalX")}
    dom = max(local, key=lambda k: local[k][0])
    d_ms, d_cnt, d_bytes = local[dom]
    achieved = (d_bytes / d_cnt) / (d_ms / d_cnt * 1e-3) / 1e9
    lq = E.eng.local_qubits   # stored qubits per shard
    shard_bytes = 8.0 * 2.0 ** lq
    cross = {}
    if stats["BoundX"][1]:
        ms = stats["BoundX"][0] / stats["BoundX"][1]
        # a read-modify-write cross pass moves (1 - 1/N) of the shard over NVLink in EACH direction twice (loads + stores)
        per_dir = 2.0 * (1.0 - 1.0 / world) * shard_bytes
        cross = {"kernel": "tile_prog_kernel<ProgBoundX>", "ms_per_launch": ms, "launches_per_eval": stats["BoundX"][1] / args.steps,
                 "bytes_per_direction_per_launch": per_dir, "achieved_gbs_per_direction": per_dir / (ms * 1e-3) / 1e9,
                 "reference_gbs_per_direction": 770.0, "reference_source": "B200_PROFILING.md measured peer copy"}
    E.close()
    return {
        "value": ms_dev, "ms_per_step": ms_dev, "n_gpus": world, "scaling": "weak",
        "config": {"workload": "MaxCut %d-regular n=%d seed=42, X mixer, Plus, p=%d, complex64, ONE register sharded on its top "
                               "%d qubits over %d GPUs (2^%d amplitudes = %d GiB per GPU); 1 step = 1 energy evaluation"
                               % (3 if n % 2 == 0 else 4, n, p, gb, world, lq, int(shard_bytes) >> 30),
                   "register": ("symmetric half register in twisted shard coordinates: 2^%d stored amplitudes, the mirror "
                                "pairing of the top qubit stays on the owning GPU, qubit 0 is mixed in the cross passes" % (n - 1))
                               if sym else "full register",
                   "l2": "shard is %d GiB >> 126 MB L2, no flush needed" % (int(shard_bytes) >> 30),
                   "energy": float(-out[0]), "norm": float(out[4]),
                   "value_is": "sum of the CUDA-event times of the pass kernels of one evaluation, max over ranks",
                   "e2e_is": "wall time per evaluation through DistShardedEngine.energy (host angles in, host result out, "
                             "barriers and the scalar all-reduce included), max over ranks"},
        "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": int(ang.nbytes) + 4096, "d2h_bytes_per_step": 40,
                "api": "qaoa_shard_begin / qaoa_shard_run_pass / qaoa_shard_finish (C ABI via ctypes) under torch.distributed"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "tile_prog_kernel<Prog%s>" % dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": None, "launches_timed": int(d_cnt),
                     "ms_per_launch": d_ms / d_cnt, "note": "dominant LOCAL (HBM-bound) pass on rank 0; the cross passes are NVLink-bound, see nvlink"},
        "nvlink": cross,
        "passes": {k: {"ms_total": v[0], "launches": int(v[1])} for k, v in stats.items() if v[1]},
    }


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
    ap.add_argument("--symmetric", type=int, default=1, help="0: always store the full register")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from qaoa_b200 import engine as eng

    n, p = args.n, args.p
    sharded = None
    if world > 1:
        sharded = run_sharded(args, rank, world, local_rank, dist, torch)
    edges = workload(n)
    E = eng.Engine(n, "complex64", device=local_rank)
    E.set_option("symmetric", args.symmetric)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    ang = angles_for(p, 1234 + rank)
    E.set_option("time_passes", 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident metric: CUDA-event time of the evaluation's kernels -------------------
    for _ in range(args.warmup):
        E.energy(ang)
    E.counter("reset")
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        E.energy_batch_resident(ang[None, :])
        dev_ms.append(E.counter("last_device_ms"))
    barrier()
    wall_resident = time.perf_counter() - t0
    launches = E.counter("launches")
    pass_stats = {i: (E.counter("pass_ms_%d" % i), E.counter("pass_count_%d" % i), E.counter("pass_bytes_%d" % i))
                  for i in range(13)}
    out = E.fetch_results(1)[0]
    sym = E.counter("symmetric_plans") >= 1
    stored_amps = 2.0 ** (n - 1) if sym else 2.0 ** n

    # ---- end to end through the public API: host angles in, host result out ----------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = E.energy(ang)
    barrier()
    wall_e2e = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    full_register = None
    if sym and world == 1:
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

    landscape = None
    if not args.no_landscape:
        barrier()
        landscape = landscape_leg(local_rank, world)
        barrier()

    optimize = None
    if not args.no_optimize and world == 1 and rank == 0:
        try:
            optimize = optimize_leg(local_rank, n=n)
        except Exception as exc:  # noqa: BLE001  (a side leg must never cost the bench line)
            optimize = {"failed": repr(exc)}

    ms_dev = float(np.mean(dev_ms))
    ms_e2e = 1e3 * wall_e2e / args.steps
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = peaks()
        names = {0: "tile_pass_kernel (interpreter)", 1: "tile_prog_kernel<ProgFirstInit>", 2: "tile_prog_kernel<ProgFirstLoad>",
                 3: "tile_prog_kernel<ProgInterior>", 4: "tile_prog_kernel<ProgBoundary>", 5: "tile_prog_kernel<ProgFinal>",
                 6: "tile_prog_kernel<ProgMid>", 7: "tile_prog_kernel<ProgBound2>", 8: "tile_prog_kernel<ProgFinal2>",
                 9: "tile_prog_kernel<ProgBoundX>", 10: "tile_prog_kernel<ProgFinalX>",
                 11: "tile_prog_kernel<ProgBound2L4>", 12: "tile_prog_kernel<ProgFinal2L4>"}
        dom = max(pass_stats, key=lambda i: pass_stats[i][0])
        d_ms, d_cnt, d_bytes = pass_stats[dom]
        achieved = (d_bytes / d_cnt) / (d_ms / d_cnt * 1e-3) / 1e9 if d_cnt else 0.0
        total_pass_ms = sum(v[0] for v in pass_stats.values())
        total_bytes = sum(v[2] for v in pass_stats.values())
        # the survey's pass model: s * 2^n * (2 p P - 2), P = ceil((n-4)/10)
        P = max(1, -(-(n - 4) // 10))
        survey_bytes = 8.0 * 2.0 ** n * (2 * p * P - 2)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01b_traffic.json" if sym else "r01_traffic.json")) as f:
                ratios = json.load(f)["traffic_over_algorithmic"]
            key = [k for k in ratios if k in names[dom]]
            if key and d_cnt:
                # DRAM bytes per launch = (ncu dram read+write / algorithmic bytes, measured at n=28) x algorithmic bytes
                traffic = ratios[key[0]] * d_bytes / d_cnt
        except Exception:
            pass
        # whole-job value: a step evaluates one n-qubit register per GPU (independent units, no collective);
        # ms per evaluation = step time (max over ranks) / number of GPUs
        line = {
            "metric": METRIC, "value": ms_dev / world, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "c64", "data": "synthetic",
            "config": {"workload": "MaxCut 3-regular n=%d seed=42, X mixer, Plus, p=%d, complex64; 1 step = 1 energy "
                                   "evaluation per GPU (%d independent evaluations per step, each rank its own angle vector: "
                                   "the landscape / optimiser workloads shard over evaluations with no data-path collective); "
                                   "value = step time / %d; the sharded single-register path is reported under 'sharded'"
                                   % (n, p, world, world),
                       "register": ("symmetric half register: psi(x) = psi(~x) for MaxCut + X + Plus, 2^%d stored amplitudes, "
                                    "the top qubit's rotation pairs every element with its mirror image" % (n - 1)) if sym
                                   else "full register, 2^%d stored amplitudes" % n,
                       "l2": "state is %d GiB >> 126 MB L2, no flush needed" % (int(8 * stored_amps) >> 30),
                       "energy": float(-out[0]), "norm": float(out[4])},
            "e2e": {"value": ms_e2e / world, "unit": UNIT, "h2d_bytes_per_step": int(ang.nbytes) * world,
                    "d2h_bytes_per_step": 40 * world,
                    "api": "qaoa_energy (C ABI via ctypes): host angle vector in, host {E,E2,min,max,norm} out"},
            "gpu_launches": int(launches) * world,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": "profiles/%s ratio x algorithmic bytes" % ("r01b_symmetric_n28.md" if sym else "r01_tile_kernels_n28.md"),
                         "launches_timed": int(d_cnt), "ms_per_launch": d_ms / d_cnt if d_cnt else None,
                         "kernel_share_of_step": d_ms / total_pass_ms if total_pass_ms else None,
                         "whole_eval_actual_gbs": total_bytes / (total_pass_ms * 1e-3) / 1e9 if total_pass_ms else None,
                         "algorithmic_bytes_per_eval": total_bytes / args.steps,
                         "whole_eval_frac_of_peak": total_bytes / (total_pass_ms * 1e-3) / 1e9 / peak if total_pass_ms else None,
                         "survey_model_bytes_per_eval": survey_bytes,
                         "survey_model_note": "SURVEY 8d model: 28 sweeps of the FULL 2^n register; the engine does 20 sweeps "
                                              "of the stored register (half of it when symmetric)"},
            "passes": {names[i]: {"ms_total": v[0], "launches": int(v[1])} for i, v in pass_stats.items() if v[1]},
        }
        if full_register is not None:
            line["full_register"] = full_register
        if landscape is not None:
            line["landscape"] = landscape
        if optimize is not None:
            line["optimize"] = optimize
        if sharded is not None:
            # ONE register of n + log2(N) qubits sharded over the N GPUs (same 2^n amplitudes per GPU):
            # cross passes over NVLink peer memory + scalar all-reduce; its own value / e2e / roofline / nvlink
            line["sharded"] = sharded
        if not args.no_cpu_baseline and world == 1:
            from oracle import c_oracle
            try:
                c_oracle.build()
                c_oracle.use_all_cores()
                ns = pick_cpu_n(c_oracle, p, n, 20.0, 4, args.cpu_n)
                cdiag = c_oracle.maxcut_diag(ns, workload(ns))
                c_oracle.energy(ns, cdiag, ang, "complex64")
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    c_oracle.energy(ns, cdiag, ang, "complex64")
                ms_s = 1e3 * (time.perf_counter() - t0) / reps
                scale = (2.0 ** n * n) / (2.0 ** ns * ns)
                line["cpu_baseline"] = {
                    "value": ms_s * scale, "unit": UNIT, "cores": c_oracle.num_threads(), "kind": "port",
                    "sample": "n=%d p=%d complex64: %.1f ms/eval measured on the host cores, scaled by 2^n*n to n=%d"
                              % (ns, p, ms_s, n)}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % exc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
