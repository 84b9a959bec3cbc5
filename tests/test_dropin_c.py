"""The C face of the product: programs written against the reference's headers, compiled
against include/fastsparse/ and linked to libfastsparse_b200.so (tests/dropin/build_dropin.py).

  * CPU: the binaries exist / build, and without a GPU they fail loudly (no CPU fallback);
  * GPU: our C acceptance test passes, and the reference's own test_sparse.c -- compiled
    UNMODIFIED against the drop-in headers -- passes all 29 of its tests on the B200; its
    bench_csr / bench_a_mul_b / preprocess drivers run to completion on the fixture."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

BUILD = os.path.join(ROOT, "tests", "_build")


def _ensure_built():
    if not os.path.exists(os.path.join(BUILD, "dropin_test")):
        import runpy
        runpy.run_path(os.path.join(ROOT, "tests", "dropin", "build_dropin.py"))["build"]()


def _run(name, *args, timeout=300):
    exe = os.path.join(BUILD, name)
    if not os.path.exists(exe):
        pytest.skip(f"{name} not built (the reference sources are only available in the build container)")
    return subprocess.run([exe, *args], cwd=GOLDEN, capture_output=True, text=True, timeout=timeout)


def test_dropin_binaries_build_and_refuse_cpu(have_gpu):
    _ensure_built()
    assert os.path.exists(os.path.join(BUILD, "dropin_test"))
    if have_gpu:
        pytest.skip("a GPU is present")
    r = _run("dropin_test")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_dropin_c_acceptance():
    _ensure_built()
    r = _run("dropin_test")
    assert r.returncode == 0 and "DROPIN TEST PASSED" in r.stdout, r.stdout + r.stderr


def _run_env(name, env, *args, timeout=300):
    exe = os.path.join(BUILD, name)
    if not os.path.exists(exe):
        pytest.skip(f"{name} not built")
    return subprocess.run([exe, *args], cwd=GOLDEN, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **env))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["0", "fast", "1"])
def test_dropin_c_acceptance_in_every_cache_mode(mode):
    """FSB_CACHE=0 (no caching: two-matrix calls such as bsbm_cg / bsbm_AtA hold two transient handles until the
    call settles -- round 1 freed the first one early), FSB_CACHE=fast, and the exact default."""
    _ensure_built()
    r = _run_env("dropin_test", {"FSB_CACHE": mode})
    if mode == "fast":      # the sampled fingerprint cannot see the in-place edit: exactly that check fails, nothing else
        assert "in-place edit" in r.stdout and "DROPIN TEST FAILED (2)" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 0 and "DROPIN TEST PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_test_sparse_passes_without_the_cache():
    r = _run_env("ref_test_sparse", {"FSB_CACHE": "0"})
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout and "Tests run: 29" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_time_dropin_small():
    """tests/dropin/time_dropin.c (bench.py's C-caller leg): malloc'd operands through bcsr_A_mul_Bn, self-checked."""
    _ensure_built()
    r = _run_env("time_dropin", {}, "200000", "30000", "4000000", "32", "2")
    assert r.returncode == 0, r.stdout + r.stderr
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["max_abs_err_sampled_rows"] <= 1e-9 and line["ms_per_call"] > 0


@pytest.mark.gpu
def test_c_example_sampler_loop(tmp_path):
    """examples/sampler_loop.c: plain C on the C ABI alone -- file straight into HBM, device noise, block-CG solves."""
    _ensure_built()
    r = _run("sampler_loop", "data/sbm-100-50.data", "8", "3")
    assert r.returncode == 0 and "SAMPLER LOOP OK" in r.stdout and r.stdout.count("iterations") == 3, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_test_sparse_unmodified_passes_on_gpu():
    r = _run("ref_test_sparse")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL TESTS PASSED" in r.stdout and "Tests run: 29" in r.stdout, r.stdout


@pytest.mark.gpu
def test_reference_drivers_unmodified_run_on_gpu(tmp_path):
    r = _run("ref_bench_csr", "-f", "data/sbm-100-50.data")
    assert r.returncode == 0 and "[par B'B x]" in r.stdout, r.stdout + r.stderr
    r = _run("ref_bench_a_mul_b", "-f", "data/sbm-100-50.data", "-b", "8", "-c", "-r")
    assert r.returncode == 0 and "[BlockCG2]\tniter:" in r.stdout and "[cg8**-csr]" in r.stdout, r.stdout + r.stderr
    # preprocess writes <file>.csr.bin next to its input: work on a copy
    import shutil
    shutil.copy(os.path.join(GOLDEN, "data", "sbm-100-50.data"), tmp_path / "m.data")
    exe = os.path.join(BUILD, "ref_preprocess")
    r = subprocess.run([exe, "-f", str(tmp_path / "m.data")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Testing deserialization ... done!" in r.stdout, r.stdout + r.stderr
    assert os.path.getsize(tmp_path / "m.data.csr.bin") == 2537
    r = _run("ref_bench_csr", "-p", "-f", str(tmp_path / "m.data.csr.bin"))
    assert r.returncode == 0 and "[par B'B x]" in r.stdout, r.stdout + r.stderr
