"""bench.py's contract, the parts that run without a GPU: the reference arm (`--impl reference`: the C restatement of the
reference path timed on the host cores) prints ONE JSON line with the keys the driver reads, rank 0 alone works under a
multi-rank launch, and the helpers that size the CPU sample and check the p = 1 energy behave."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_bench(extra, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run_bench(["--impl", "reference", "--n", "20", "--p", "2", "--cpu-n", "18", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["higher_is_better"] is False and d["unit"] == "ms" and d["dtype"] == "c64" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["value"] > 0 and d["value"] == d["ms_per_step"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "n=18" in cb["sample"] and cb["same_config"] is False          # a bounded sample, scaled, and it says so
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench(["--impl", "reference", "--n", "20", "--p", "1", "--cpu-n", "16", "--steps", "1", "--warmup", "0"],
                  {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_reference_arm_runs_the_workload_itself_when_it_fits():
    r = run_bench(["--impl", "reference", "--n", "16", "--p", "1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["same_config"] is True and "no scaling" in d["cpu_baseline"]["sample"]


def test_closed_form_p1_matches_the_oracle():
    import bench
    from oracle import qaoa_oracle as orc
    n = 10
    edges = bench.workload(n)
    diag = np.zeros(1 << n)
    x = np.arange(1 << n)
    for u, v, w in edges:
        diag += w * (((x >> u) ^ (x >> v)) & 1)
    for g, b in ((0.4, 0.3), (1.1, -0.7)):
        psi = orc.evolve(orc.Spec(n, diag), np.array([g, b]))
        want = float((np.abs(psi) ** 2) @ diag)
        assert abs(bench.closed_form_p1(n, edges, g, b) - want) < 1e-10


def test_workload_is_the_seeded_regular_graph():
    import bench
    for n in (16, 24, 32):
        e = bench.workload(n)
        deg = np.zeros(n, dtype=int)
        for u, v, w in e:
            assert w == 1 and u != v
            deg[u] += 1
            deg[v] += 1
        assert (deg == 3).all() and len(e) == 3 * n // 2
        assert e == bench.workload(n)                                     # deterministic: seed 42
    a = bench.angles_for(5, 1234)
    assert a.shape == (10,) and np.array_equal(a, bench.angles_for(5, 1234))
