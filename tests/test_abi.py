"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qaoa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qaoa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from qaoa_b200 import engine

    lib = ctypes.CDLL(engine.library_path())
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), "libqaoa_b200.so does not export %s" % name


def test_binding_covers_the_header():
    from qaoa_b200 import engine

    lib = engine.load_library()
    for name in declared_symbols():
        assert getattr(lib, name).argtypes is not None or name == "qaoa_global_error", name


def test_no_cpu_fallback_without_gpu():
    """Creating an engine without a CUDA device must fail loudly (never route to a CPU path)."""
    import pytest

    from qaoa_b200 import engine

    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is visible")
    with pytest.raises(engine.EngineError):
        engine.Engine(8)


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "qaoa_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("qaoa_oracle", "oracle") or "import oracle" not in text
                assert "from oracle" not in text and "import oracle" not in text, fn


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU restatement on the host cores, bounded sample): one JSON line with the keys the
    driver reads; needs no GPU.  Under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-n", "14"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["metric"] == "ms_per_energy_eval_maxcut_n32_p5" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "n=14" in line["cpu_baseline"]["sample"] and "workload" in line["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT,
                           env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert other.returncode == 0 and not [l for l in other.stdout.splitlines() if l.startswith("{")]
