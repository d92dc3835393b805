"""Worker of tests/test_gpu_multigpu.py: one process per GPU under torchrun (NCCL).  Every rank drives
DistShardedEngine -- the CUDA-IPC / torch.distributed path that bench.py's sharded leg uses -- for the FULL sharded
register and for the sharded SYMMETRIC register; rank 0 compares each energy with the C oracle (complex128 reference,
tolerance 1e-4 relative for the complex64 engine, 1e-9 for complex128) and with the closed-form p = 1 MaxCut energy, and
checks that every rank received identical results.  Exit code 0 = all checks passed on every rank."""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def closed_form_p1(n, edges, gamma, beta):
    """<cost> of p = 1 MaxCut QAOA on any unweighted graph (reference sign conventions)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", dest="n", type=int, default=26)   # ("--n" is ambiguous to torchrun's own parser)
    ap.add_argument("--dtype", default="complex64")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert torch.cuda.device_count() >= world, "one GPU per rank"

    import parity_helpers as H
    from qaoa_b200.sharded import DistShardedEngine

    n, dtype = args.n, args.dtype
    tol = 1e-4 if dtype == "complex64" else 1e-9
    edges, colors, table = H.regular_maxcut(n)
    lut, _ = H.lut_from_table(table, 1)
    rng = np.random.default_rng(77)
    cases = []
    for p in (1, 2, 3):
        ang = rng.uniform(0, 2 * np.pi, size=2 * p)
        cases.append(ang)
    refs = None
    if rank == 0:
        from oracle import c_oracle
        c_oracle.build()
        c_oracle.use_all_cores()
        d8 = c_oracle.maxcut_diag_u8(n, edges)
        refs = [c_oracle.run(n, d8, ang, "complex128", blocked=True) for ang in cases]
    report = {"n": n, "world": world, "dtype": dtype, "checks": []}
    ok = True
    variants = [False] + ([True] if dtype == "complex64" and n % 2 == 0 else [])
    for sym in variants:
        E = DistShardedEngine(n, dtype=dtype, device=local_rank, symmetric=sym)
        assert int(E.eng.counter("device")) == local_rank, "every rank must open its own GPU"
        E.set_cost_maxkcut(n, edges, 1, lut)
        for i, ang in enumerate(cases):
            got = E.energy(ang)
            # identical on every rank
            t = torch.tensor(got, dtype=torch.float64, device="cuda")
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo, hi))
            entry = {"symmetric": sym, "p": len(ang) // 2, "energy": float(got[0]), "same_on_all_ranks": same}
            if rank == 0:
                ref = refs[i]
                entry["oracle"] = float(ref[0])
                entry["rel_err"] = abs(got[0] - ref[0]) / abs(ref[0])
                good = entry["rel_err"] < tol and abs(got[1] - ref[1]) < 2 * tol * abs(ref[1]) and \
                    got[2] == ref[2] and got[3] == ref[3] and abs(got[4] - 1) < (1e-4 if dtype == "complex64" else 1e-11)
                if len(ang) == 2:
                    cf = closed_form_p1(n, edges, ang[0], ang[1])
                    entry["closed_form_rel_err"] = abs(got[0] - cf) / abs(cf)
                    good = good and entry["closed_form_rel_err"] < tol
                entry["ok"] = bool(good and same)
                ok = ok and entry["ok"]
            else:
                ok = ok and same
            report["checks"].append(entry)
        E.close()
    # Replacing the cost function is collective (peers unmap the old diagonal before its owner frees it): lower a second
    # graph onto the same sharded engine, then the first one again -- the energies must follow
    E = DistShardedEngine(n, dtype=dtype, device=local_rank, symmetric=False)
    E.set_cost_maxkcut(n, edges, 1, lut)
    first = E.energy(cases[1])
    edges2 = [e for i, e in enumerate(edges) if i % 5]
    E.set_cost_maxkcut(n, edges2, 1, lut)
    other = E.energy(cases[1])
    E.set_cost_maxkcut(n, edges, 1, lut)
    again = E.energy(cases[1])
    entry = {"case": "cost replaced twice", "energy": float(again[0]), "same_on_all_ranks": True}
    if rank == 0:
        ref2 = c_oracle.run(n, c_oracle.maxcut_diag_u8(n, edges2), cases[1], "complex128", blocked=True)
        entry["oracle"] = float(ref2[0])
        entry["rel_err"] = max(abs(other[0] - ref2[0]) / abs(ref2[0]), abs(again[0] - refs[1][0]) / abs(refs[1][0]))
        entry["ok"] = bool(entry["rel_err"] < tol and again[0] == first[0])
        ok = ok and entry["ok"]
    report["checks"].append(entry)
    try:
        E.set_cost_maxkcut_terms(n, edges, list(range(len(edges))), len(edges), 1, lut)
        ok = False                                        # per-term gammas must be rejected on a sharded register
    except NotImplementedError:
        pass
    E.close()
    # Grover mixer over a Dicke state on the sharded FULL register: no state exchange, one all-reduced scalar per layer
    ng, kg = min(n, 24), 7
    Q, cq, bq = H.random_qubo(ng, seed=ng)
    E = DistShardedEngine(ng, dtype=dtype, device=local_rank, symmetric=False)
    E.set_cost_qubo(Q, cq, bq)
    from qaoa_b200 import engine as eng_mod
    E.set_initial(eng_mod.INIT_DICKE, kg)
    E.set_mixer(eng_mod.MIXER_GROVER, sub_kind=eng_mod.INIT_DICKE, sub_k=kg)
    ang = rng.uniform(0, 2 * np.pi, size=6)
    ang[0::2] *= 0.05
    got = E.energy(ang)
    t = torch.tensor(got, dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    entry = {"case": "grover+dicke", "n": ng, "p": 3, "energy": float(got[0]), "same_on_all_ranks": same}
    if rank == 0:
        ref = c_oracle.run(ng, c_oracle.qubo_diag(Q, cq, bq), ang, "complex128", mixer="grover", init=("dicke", kg), sub=("dicke", kg))
        entry["oracle"] = float(ref[0])
        entry["rel_err"] = abs(got[0] - ref[0]) / abs(ref[0])
        entry["ok"] = bool(entry["rel_err"] < tol and abs(got[4] - 1) < 1e-4 and same)
        ok = ok and entry["ok"]
    else:
        ok = ok and same
    report["checks"].append(entry)
    E.close()
    # sampling (QAOA.hist) and exact CVaR on the sharded full register: kept shards, gathered chunk masses, all-reduced
    # per-digit histograms -- identical on every rank, CVaR against the numpy oracle's state, shots against its probabilities
    ns = 22
    es, _, tab_s = H.regular_maxcut(ns)
    E = DistShardedEngine(ns, dtype=dtype, device=local_rank, symmetric=False)
    E.set_cost_maxkcut(ns, es, 1, lut)
    ang = cases[2]
    idx = E.sample(ang, 4096, seed=5)
    cv = E.cvar(ang, 0.2)
    t = torch.tensor(np.concatenate([idx.astype(np.float64), [cv]]), dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    entry = {"case": "sample + cvar", "n": ns, "cvar": float(cv), "same_on_all_ranks": same}
    if rank == 0:
        from oracle import qaoa_oracle as orc
        dg = c_oracle.maxcut_diag_u8(ns, es).astype(np.float64)
        pr = np.abs(orc.evolve(orc.Spec(ns, dg), ang)) ** 2
        want = orc.exact_cvar(dg, pr, 0.2)
        mean, var = float(pr @ dg), float(pr @ dg ** 2 - (pr @ dg) ** 2)
        z = abs(dg[idx.astype(np.int64)].mean() - mean) / np.sqrt(var / len(idx))
        entry["oracle"] = float(want)
        entry["rel_err"] = abs(cv - want) / abs(want)
        entry["sample_mean_z"] = float(z)
        entry["ok"] = bool(entry["rel_err"] < tol and z < 5.0 and same and int(idx.max()) < (1 << ns))
        ok = ok and entry["ok"]
    else:
        ok = ok and same
    report["checks"].append(entry)
    E.close()
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        report["ok"] = int(flag[0]) == 0
        print("MULTIGPU_REPORT " + json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(int(flag[0]))


if __name__ == "__main__":
    main()
