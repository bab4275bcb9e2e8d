"""Build libsketchy_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsketchy_b200.so")
SOURCES = ["api.cu", "kernels_sketch.cu", "kernels_predict.cu"]
HEADERS = ["common.cuh", "kernels.h", "pack_avx2.cpp", "pack_avx512.cpp", os.path.join("..", "..", "include", "sketchy_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsketchy_b200 can only be built with the CUDA toolkit")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(HERE, "host", f) for f in ("main.cpp", "msh.cpp", "msh.hpp", "fastx.hpp", "ingest.hpp", "capnp_lite.hpp")]
    cli = os.path.join(HERE, "bin", "sketchy")
    if not os.path.exists(cli):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = _nvcc()
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    # host-only AVX2 packer: its own object so that only this file is built with -mavx2 (called after a CPU check)
    pobj = os.path.join(HERE, "build", "pack_avx2.o")
    cxx = shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O3", "-mavx2", "-fPIC", "-std=c++17", "-Wall", "-c", os.path.join(CSRC, "pack_avx2.cpp"), "-o", pobj])
    objs.append(pobj)
    pobj5 = os.path.join(HERE, "build", "pack_avx512.o")   # likewise: only this file is built with the AVX-512 flags
    subprocess.check_call([cxx, "-O3", "-mavx512f", "-mavx512bw", "-mavx512vbmi", "-fPIC", "-std=c++17", "-Wall", "-c",
                           os.path.join(CSRC, "pack_avx512.cpp"), "-o", pobj5])
    objs.append(pobj5)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    build_host()
    return SO


HOST = os.path.join(HERE, "host")
CLI = os.path.join(HERE, "bin", "sketchy")


def build_host() -> str:
    """The `sketchy` CLI host (C++): same sub-commands / flags / rows as the reference binary, over the C ABI."""
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", os.path.join(HOST, "main.cpp"), os.path.join(HOST, "msh.cpp"), "-o", CLI,
           "-L" + HERE, "-lsketchy_b200", "-lz", "-ldl", "-pthread", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + HERE]
    subprocess.check_call(cmd)
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
