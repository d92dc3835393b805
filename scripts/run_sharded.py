"""One QAOA register sharded over the GPUs of a node, one process per GPU (torchrun).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
      scripts/run_sharded.py --qubits 33 --depth 5 [--check] [--reps 3]

Prints one JSON line on rank 0: ms per energy evaluation (max over ranks, device-synchronised wall time
between barriers), the time spent in cross (NVLink) and local passes, and -- with --check, n <= 28 --
the energy of the C/OpenMP oracle for the same graph and angles."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", dest="n", type=int, default=33)
    ap.add_argument("--depth", dest="p", type=int, default=5)
    ap.add_argument("--dtype", default="complex64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--low-bits", type=int, default=0, dest="low_bits")
    ap.add_argument("--symmetric", action="store_true",
                    help="sharded symmetric register (MaxCut + X + |+>, n even): 2^(n-1) stored amplitudes")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import networkx as nx
    from qaoa_b200.sharded import DistShardedEngine
    from qaoa_b200.utils.graphutils import GraphHandler

    n, p = args.n, args.p
    G = nx.random_regular_graph(3 if n % 2 == 0 else 4, n, seed=42)   # odd n: no 3-regular graph exists
    edges = GraphHandler(G, check_isomorphism=False).edge_list()
    E = DistShardedEngine(n, dtype=args.dtype, device=local_rank, symmetric=args.symmetric)
    if args.low_bits:
        E.set_option("low_bits", args.low_bits)
    E.set_cost_maxkcut(n, edges, 1, np.array([0, 1], dtype=np.uint8))
    E.set_option("time_passes", 1)
    ang = np.random.default_rng(1234).uniform(0, 2 * np.pi, size=2 * p)
    for _ in range(args.warmup):
        out = E.energy(ang)
    E.eng.counter("reset")
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = E.energy(ang)
    torch.cuda.synchronize()
    dist.barrier()
    ms = 1e3 * (time.perf_counter() - t0) / args.reps
    names = ["interpreter", "FirstInit", "FirstLoad", "Interior", "Boundary", "Final", "Mid", "Bound2", "Final2", "BoundX", "FinalX", "Bound2L4", "Final2L4"]
    passes = {}
    for i, nm in enumerate(names):
        cnt = E.eng.counter("pass_count_%d" % i)
        if cnt:
            passes[nm] = {"ms_per_launch": E.eng.counter("pass_ms_%d" % i) / cnt, "launches_per_eval": cnt / args.reps,
                          "gb_per_launch": E.eng.counter("pass_bytes_%d" % i) / cnt / 1e9}
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line = {"n": n, "p": p, "world": world, "dtype": args.dtype, "symmetric": bool(args.symmetric),
            "stored_amplitudes_per_gpu": 1 << E.eng.local_qubits, "ms_per_eval": float(t[0]), "energy": float(-out[0]),
            "norm": float(out[4]), "min": float(out[2]), "max": float(out[3]), "barriers_per_eval": E.barriers / (args.reps + args.warmup),
            "passes_rank0": passes}
    if args.check and rank == 0:
        from oracle import c_oracle
        c_oracle.build()
        diag = c_oracle.maxcut_diag(n, edges)
        ref = c_oracle.energy(n, diag, ang, "complex128")
        line["oracle_energy"] = float(-ref[0])
        line["rel_err"] = abs(out[0] - ref[0]) / abs(ref[0])
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    E.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
