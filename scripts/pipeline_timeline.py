"""torchrun, >= 2 GPUs: timeline of ONE sharded evaluation with the chunk pipeline -- every launch of rank 0 with its stream
(main: local passes, side: cross passes), chunk, start and end (CUDA events, ms after the first launch), and how much of
the cross passes' time runs under local passes.  Evidence for DESIGN section 6 ("chunk pipeline")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from qaoa_b200.sharded import DistShardedEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
p = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
E = DistShardedEngine(n, dtype="complex64", device=lr, symmetric=True)
E.set_cost_maxkcut(n, bench.workload(n), 1, np.array([0, 1], dtype=np.uint8))
E.set_option("time_passes", 1)
ang = bench.angles_for(p, 1234)
for _ in range(3):
    E.energy(ang)
E.eng.counter("reset")
E.energy(ang)
tl = E.eng.timeline()
if rank == 0:
    names = bench.PASS_NAMES
    print("# n=%d p=%d, %d GPUs, %d pipeline chunks, cross passes on %d CTAs; rank 0; ms after the first launch" % (
        n, p, world, E.last_chunks, E.cross_ctas))
    print("# stream  program    chunk   start     end   duration")
    for prog, chunk, side, t0, t1 in tl:
        print("%-7s %-10s %5d %8.3f %8.3f %8.3f" % ("side" if side else "main", names[int(prog)], int(chunk), t0, t1, t1 - t0))
    cross = [(t0, t1) for prog, chunk, side, t0, t1 in tl if side]
    local = [(t0, t1) for prog, chunk, side, t0, t1 in tl if not side]
    def overlap(a, b):
        return max(0.0, min(a[1], b[1]) - max(a[0], b[0]))
    cross_total = sum(t1 - t0 for t0, t1 in cross)
    under = sum(overlap(c, l) for c in cross for l in local)
    print("# cross passes: %.3f ms in total, %.3f ms (%.0f %%) of it concurrent with a local pass; evaluation span %.3f ms" % (
        cross_total, under, 100.0 * under / max(cross_total, 1e-9), max(t1 for *_x, t1 in tl)))
E.close()
dist.destroy_process_group()
