"""In-tree build of libfastsparse_b200.so for sm_100a (explicit nvcc, no JIT cache).

    python -m libfastsparse_b200.build [--force]

Object files go to libfastsparse_b200/lib/obj/, the library to
libfastsparse_b200/lib/libfastsparse_b200.so (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libfastsparse_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-fopenmp,-Wall,-Wno-unused-function",
              "--expt-relaxed-constexpr", "-I", INCLUDE]

SOURCES = ["fsb_capi.cu", "kernels_csr.cu", "kernels_csr_staged.cu", "kernels_csr_stream.cu", "kernels_cbcsr.cu", "kernels_blocked.cu", "kernels_build.cu",
           "kernels_dense.cu", "fsb_cg.cu", "fsb_comm.cu", "fsb_io.cu", "fsb_hostcopy.cu", "fsb_host.cpp", "fsb_dropin.cpp"]


def _deps_mtime() -> float:
    t = 0.0
    for d in (CSRC, INCLUDE):
        for root, _, files in os.walk(d):
            for f in files:
                if f.endswith((".h", ".cuh")):
                    t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src: str, hdr_time: float, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_time):
        return obj
    cmd = [NVCC, *ARCH, *NVCC_FLAGS, "-c", path, "-o", obj]
    if src.endswith(".cpp"):
        cmd = [NVCC, *NVCC_FLAGS, "-x", "cu", *ARCH, "-c", path, "-o", obj] if False else \
              [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-Wall", "-I", INCLUDE, "-c", path, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {src}\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr.strip():
        print(r.stderr)
    return obj


def build_variant(name: str, defines: list[str]) -> str:
    """Experiment build: lib/libfastsparse_b200_<name>.so with extra -D flags (selected at run time by FSB_LIB)."""
    objdir = os.path.join(LIBDIR, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    out = os.path.join(LIBDIR, f"libfastsparse_b200_{name}.so")
    objs = []

    def one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC, src)
        if src.endswith(".cpp"):
            cmd = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-I", INCLUDE, "-c", path, "-o", obj]
        else:
            cmd = [NVCC, *ARCH, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {src}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    r = subprocess.run([NVCC, *ARCH, "-shared", "-ccbin", HOST_CXX, "-o", out, *objs, "-cudart", "static", "-Xcompiler", "-fopenmp",
                        "-ldl", "-lgomp"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_time = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr_time, force, verbose), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", HOST_CXX, "-o", LIB, *objs, "-cudart", "static", "-Xcompiler", "-fopenmp",
               "-Xlinker", "--no-undefined", "-ldl", "-lgomp"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
