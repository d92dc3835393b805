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
