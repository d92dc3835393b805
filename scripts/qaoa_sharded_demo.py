"""The reference-facing QAOA class on a register sharded over the GPUs of a node (torchrun, one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
      scripts/qaoa_sharded_demo.py --qubits 26

Every rank drives the same QAOA object (landscape, then optimize to depth 2) and sees identical numbers; rank 0
compares them with the same run on a single-GPU engine."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=26)
    ap.add_argument("--mixer", default="x", help="x: MaxCut + X + Plus; grover: QUBO + Grover(Dicke) + Dicke (full sharded register)")
    args = ap.parse_args()
    import networkx as nx
    import torch
    import torch.distributed as dist

    from qaoa_b200 import QAOA, initialstates, mixers, problems
    from qaoa_b200.qaoa import B200Backend
    from qaoa_b200.utils import COBYLA

    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.qubits
    G = nx.random_regular_graph(3, n, seed=42)
    opt = [COBYLA, {"maxiter": 12, "tol": 1e-3}]
    grid = {"gamma": [0, np.pi, 4], "beta": [0, np.pi, 4]}
    if args.mixer == "grover":
        rs = np.random.RandomState(7)
        Qm, cv, k = rs.rand(n, n) / n, rs.rand(n) - 0.5, n // 3

        def parts():
            return dict(problem=problems.QUBO(Q=Qm, c=cv, b=0.0), mixer=mixers.Grover(initialstates.Dicke(k)),
                        initialstate=initialstates.Dicke(k))
        grid = {"gamma": [0, 2.0, 4], "beta": [0, np.pi, 4]}
    else:
        def parts():
            return dict(problem=problems.MaxCut(G), mixer=mixers.X(), initialstate=initialstates.Plus())
    q = QAOA(optimizer=opt, backend=B200Backend(device=local_rank, dtype="complex64", sharded=True), **parts())
    q.optimize(depth=2, angles=grid)
    if rank == 0:
        print("sharded register: %s, 2^%d stored amplitudes per GPU" % (
            "symmetric half (twisted coordinates)" if q._get_engine().symmetric else "full", q._get_engine().eng.local_qubits))
        # same landscape on a single-GPU engine, and every iterate the sharded optimiser visited re-evaluated there
        # (comparing two optimiser RUNS is ill-posed: the seed is a symmetric grid point and rounding noise picks the branch)
        one = QAOA(optimizer=opt, backend=B200Backend(device=local_rank, dtype="complex64"), **parts())
        one.sample_cost_landscape(grid)
        err = np.abs(q.Exp_sampled_p1 - one.Exp_sampled_p1).max() / np.abs(one.Exp_sampled_p1).max()
        print("landscape max rel diff sharded vs single GPU: %.2e" % err)
        worst = 0.0
        for d in (1, 2):
            r = q.optimization_results[d]
            ref = -one._get_engine().energy_batch(np.asarray(r.angles))[:, 0]
            worst = max(worst, float(np.abs(np.asarray(r.Exp) - ref).max() / np.abs(ref).max()))
            print("depth %d: %d iterates, best Exp %.6f" % (d, r.num_fval(), q.get_Exp(d)))
        print("optimiser iterates, max rel diff sharded vs single GPU: %.2e" % worst)
        assert err < 1e-4 and worst < 1e-4
    # hist() and cvar < 1 on the sharded register (collective calls: every rank takes part; MaxCut runs use a second,
    # full sharded register for them because the symmetric one cannot keep its final state)
    best = q.optimization_results[2].get_best_angles()
    q.sample_seed = 11
    h = q.hist(best, 2048)
    q.cvar = 0.25
    cv = q._cvar_of(q._engine_row(best))
    q.cvar = 1
    if rank == 0:
        assert sum(h.values()) == 2048
        mean = sum(one.problem.cost(s) * c for s, c in h.items()) / 2048.0
        e = one._get_engine().energy(np.asarray(one._engine_row(best)))
        exact, var = e[0], e[1] - e[0] ** 2
        z = abs(mean - exact) / np.sqrt(max(var, 1e-12) / 2048.0)
        one.cvar = 0.25
        cv1 = one._cvar_of(one._engine_row(best))
        print("hist: sampled mean cost %.4f vs exact %.4f (z = %.2f); CVaR(0.25) sharded %.6f vs single GPU %.6f" % (mean, exact, z, cv, cv1))
        assert z < 5.0 and abs(cv - cv1) < 1e-4 * max(1.0, abs(cv1))
        print("OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
