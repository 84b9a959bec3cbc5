"""Builds the C callers of the drop-in headers (include/fastsparse/) into tests/_build/:

  dropin_test        tests/dropin/test_dropin.c -- our own C acceptance test (always)
  time_dropin        tests/dropin/time_dropin.c -- bcsr_A_mul_Bn with malloc'd operands, timed (bench.py e2e.dropin_c)
  sampler_loop       examples/sampler_loop.c -- plain C against include/fsb.h alone (always)
  ref_test_sparse    the reference's test_sparse.c, UNMODIFIED, compiled from where it lies
  ref_bench_csr      ... bench_csr.c
  ref_bench_a_mul_b  ... bench_a_mul_b.c
  ref_preprocess     ... preprocess.c
                     (only when /root/reference exists; the binaries travel to the GPU box)

No reference source is copied: the reference .c files are compiled in place with
-I include/fastsparse first on the include path, so their #include "csr.h" etc. resolve to
the drop-in headers, and linked against libfastsparse_b200.so."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "tests", "_build")
INC = os.path.join(ROOT, "include", "fastsparse")
LIBDIR = os.path.join(ROOT, "libfastsparse_b200", "lib")
REF = "/root/reference"
CC = "/usr/bin/gcc"


def _cc(src, out, verbose, via_stdin=False, extra_inc=None):
    # via_stdin: feed the (reference) source through stdin so that its `#include "csr.h"` cannot
    # find the sibling reference headers in the source's own directory and resolves to -I INC
    cmd = [CC, "-std=gnu99", "-O2", "-g", "-fopenmp", "-Wall", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
           "-Wno-absolute-value", "-I", INC] + (["-I", extra_inc] if extra_inc else []) + (["-x", "c", "-"] if via_stdin else [src]) + \
          ["-o", out, "-L", LIBDIR, "-lfastsparse_b200", "-Wl,-rpath," + "$ORIGIN/../../libfastsparse_b200/lib", "-lm"]
    if verbose:
        print(" ".join(cmd) + (f" < {src}" if via_stdin else ""))
    r = subprocess.run(cmd, capture_output=True, text=True, stdin=open(src) if via_stdin else None, cwd=OUT)
    if r.returncode != 0:
        raise RuntimeError(f"drop-in build failed: {src}\n{r.stderr}")
    return out


def build(verbose: bool = False):
    os.makedirs(OUT, exist_ok=True)
    built = [_cc(os.path.join(HERE, "test_dropin.c"), os.path.join(OUT, "dropin_test"), verbose),
             _cc(os.path.join(HERE, "time_dropin.c"), os.path.join(OUT, "time_dropin"), verbose, extra_inc=os.path.join(ROOT, "include")),
             # plain C against the C ABI alone (include/fsb.h): the resident sampler loop
             _cc(os.path.join(ROOT, "examples", "sampler_loop.c"), os.path.join(OUT, "sampler_loop"), verbose, extra_inc=os.path.join(ROOT, "include"))]
    if os.path.isdir(REF):
        for name in ("test_sparse", "bench_csr", "bench_a_mul_b", "preprocess"):
            built.append(_cc(os.path.join(REF, name + ".c"), os.path.join(OUT, "ref_" + name), verbose, via_stdin=True))
    return built


if __name__ == "__main__":
    for b in build(verbose=True):
        print(b)
