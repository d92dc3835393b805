"""GPU, >= 2 devices: the REAL multi-process sharded path -- one process per GPU under torchrun (NCCL), CUDA IPC peer
mappings, cross passes over NVLink, scalar all-reduce (qaoa_b200/sharded.py DistShardedEngine) -- against the C oracle,
the closed-form p = 1 energy and itself across ranks (tests/multigpu_worker.py does the work).  Skipped on a box with a
single GPU; there tests/test_gpu_sharded.py covers the same kernels with every rank's engine inside one process."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(world, n, dtype, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "multigpu_worker.py"),
           "--qubits", str(n), "--dtype", dtype]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    rep = None
    for line in proc.stdout.splitlines():
        if line.startswith("MULTIGPU_REPORT "):
            rep = json.loads(line[len("MULTIGPU_REPORT "):])
    if rep is not None:
        for c in rep["checks"]:
            print("CHECK", json.dumps(c))
    print(proc.stderr[-1500:])
    assert proc.returncode == 0 and rep is not None and rep["ok"], (proc.stdout[-300:], proc.stderr[-300:])
    return rep


@pytest.mark.parametrize("world,n,dtype", [(2, 26, "complex64"), (2, 24, "complex128"), (4, 28, "complex64"),
                                           (8, 28, "complex64")])
def test_dist_sharded_engine_against_the_oracle(world, n, dtype):
    have = _gpus()
    if have < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, have))
    rep = _run(world, n, dtype, 29600 + world)
    assert len(rep["checks"]) >= 3
    for c in rep["checks"]:
        assert c["ok"] and c["rel_err"] < (1e-4 if dtype == "complex64" else 1e-9), c


@pytest.mark.parametrize("mixer", ["x", "grover"])
def test_public_api_on_a_sharded_register(mixer):
    """QAOA(..., backend=B200Backend(sharded=True)) under torchrun on 2 GPUs (scripts/qaoa_sharded_demo.py): landscape and
    optimiser iterates of the sharded register equal the single-GPU engine's -- MaxCut + X + Plus on the symmetric
    sharded register, QUBO + Grover(Dicke) + Dicke on the full one"""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641" if mixer == "x" else "29642", os.path.join(os.path.dirname(HERE), "scripts", "qaoa_sharded_demo.py"),
           "--qubits", "24", "--mixer", mixer]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(proc.stdout[-1500:])
    print(proc.stderr[-1500:])
    assert proc.returncode == 0 and "OK" in proc.stdout
