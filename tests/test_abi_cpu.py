"""CPU tests: the C-ABI library loads and exports every symbol include/knowhere_b200.h declares;
without a GPU every entry point fails loudly with knowhere::Status::cuda_runtime_error (22)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "knowhere_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(kb):
    L = kb.lib()
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/knowhere_b200.h but not exported"


def test_sass_is_sm100a_only():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    from knowhere_b200 import LIB
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(kb):
    if kb.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(kb.KnowhereError) as e:
        kb.Index("FLAT", "L2", 8)
    assert e.value.status == 22
    with pytest.raises(kb.KnowhereError) as e:
        kb.brute_force_search(np.zeros((4, 8), np.float32), np.zeros((1, 8), np.float32), 1)
    assert e.value.status == 22
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    # the product path must not route through oracle/ (checked textually over the package sources)
    pkg = os.path.join(ROOT, "knowhere_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("never includes anything under oracle/", "") or f.endswith(".cuh"), f
                assert "import oracle" not in txt and "from oracle" not in txt, f
