"""2+ GPUs (torchrun): per-pass times of the sharded n-qubit register against the number of persistent CTAs -- how many
SMs a cross pass needs to keep NVLink busy, and how the local passes slow down on fewer SMs."""
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
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
E = DistShardedEngine(n, dtype="complex64", device=lr, symmetric=True)
E.set_cost_maxkcut(n, bench.workload(n), 1, np.array([0, 1], dtype=np.uint8))
E.set_option("time_passes", 1)
ang = bench.angles_for(p, 1234)
E.energy(ang)
for ctas in (148, 132, 116, 96, 64, 48, 32, 24, 16):
    E.set_option("persistent_ctas", ctas)
    E.energy(ang)
    E.eng.counter("reset")
    for _ in range(3):
        out = E.energy(ang)
    st = {nm: round(E.eng.counter("pass_ms_%d" % i) / max(1, E.eng.counter("pass_count_%d" % i)), 3)
          for i, nm in enumerate(bench.PASS_NAMES) if E.eng.counter("pass_count_%d" % i)}
    if rank == 0:
        print(json.dumps({"ctas": ctas, "ms_per_launch": st, "energy": float(out[0])}), flush=True)
E.close()
dist.destroy_process_group()
