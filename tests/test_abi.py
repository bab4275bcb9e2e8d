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


def test_avx2_block_packer_matches_the_bytewise_rule():
    """Host-side packer fast path (pack_avx2.cpp): codes/mask per 32-byte block must follow the same rule as the
    byte-wise packer (needletail normalize(false): ACGT/acgt/Uu -> 0..3, other kept bytes invalid with code 0), and a
    block holding a removed byte (blank, tab, CR, LF) must be left alone."""
    import ctypes as C
    import numpy as np
    if "avx2" not in open("/proc/cpuinfo").read():
        pytest.skip("no AVX2 on this host")
    from sketchy_b200 import _lib, build
    build.build()
    lib = C.CDLL(_lib.SO_PATH)
    fn = lib.skb_pack_blocks_avx2
    fn.restype = C.c_uint64
    fn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"ACGTacgtUuNnRYKM-.*xX@\x00\xff\xc1\xe1", dtype=np.uint8)
    p = np.array([20] * 4 + [6] * 4 + [2, 2] + [1] * (alphabet.size - 10), dtype=np.float64)
    s = rng.choice(alphabet, size=32 * 200 + 17, p=p / p.sum()).astype(np.uint8)
    codes = np.zeros(2 * 201, dtype=np.uint32)
    mask = np.zeros(201, dtype=np.uint32)
    n = fn(s.ctypes.data, s.size, codes.ctypes.data, mask.ctypes.data)
    assert n == 32 * 200
    code_of = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2,
               ord("T"): 3, ord("t"): 3, ord("U"): 3, ord("u"): 3}
    for blk in range(200):
        e_codes = [0, 0]
        e_mask = 0
        for j in range(32):
            b = int(s[32 * blk + j])
            if b in code_of:
                e_codes[j // 16] |= code_of[b] << (2 * (j % 16))
            else:
                e_mask |= 1 << j
        assert int(codes[2 * blk]) == e_codes[0] and int(codes[2 * blk + 1]) == e_codes[1], blk
        assert int(mask[blk]) == e_mask, blk
    # stops in front of the first block with a removed byte
    for ws in b" \t\r\n":
        t = s.copy()
        t[32 * 7 + 5] = ws
        assert fn(t.ctypes.data, t.size, codes.ctypes.data, mask.ctypes.data) == 32 * 7
    # every byte value, at every position of a block
    for v in range(256):
        for pos in (v % 32, (v * 7 + 3) % 32):
            blk = np.full(32, ord("G"), dtype=np.uint8)
            blk[pos] = v
            c2 = np.zeros(2, dtype=np.uint32)
            m2 = np.zeros(1, dtype=np.uint32)
            n = fn(blk.ctypes.data, 32, c2.ctypes.data, m2.ctypes.data)
            if v in b" \t\r\n":
                assert n == 0, v
                continue
            assert n == 32, v
            exp = [0xAAAAAAAA, 0xAAAAAAAA]
            exp[pos // 16] = (exp[pos // 16] & ~(3 << (2 * (pos % 16)))) | (code_of.get(v, 0) << (2 * (pos % 16)))
            assert [int(c2[0]), int(c2[1])] == exp, v
            assert int(m2[0]) == (0 if v in code_of else 1 << pos), v
