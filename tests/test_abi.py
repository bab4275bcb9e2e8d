"""CPU checks of the drop-in boundary: libsketchy_b200.so builds for sm_100a, loads, and exports exactly the
symbols include/sketchy_b200.h declares. No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sketchy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from sketchy_b200 import _lib, build
    build.build()
    lib = _lib.load_library()
    declared = _header_functions()
    assert len(declared) >= 30
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.skb_version().decode().startswith("sketchy_b200")


def test_library_is_sm100a_only_with_bulk_copy():
    import shutil
    import subprocess
    from sketchy_b200 import _lib, build
    build.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sketchy_b200 import _lib
    with pytest.raises(_lib.SkbError) as ei:
        _lib.Context(0)
    assert ei.value.code == -3  # SKB_ERR_NO_DEVICE: no CPU fallback exists


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under sketchy_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "sketchy_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f
