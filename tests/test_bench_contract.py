"""bench.py's contract on a machine without a GPU: the reference arm (the reference's own CPU code on a bounded
sample) prints one well-formed JSON line, and our arm refuses loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "c2_small", "--steps", "1", "--warmup", "1",
                        "--sample-rows", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "spmm_nnz_rhs_per_s" and line["unit"] == "nnz*RHS/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    cb, e2e = line["cpu_baseline"], line["e2e"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "rows" in cb["sample"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_never_loads_the_product_library():
    """The reference arm's inputs come from oracle/'s own generator: neither the python package nor the .so is loaded."""
    code = ("import sys, bench; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c2_small', '--steps', '1', '--warmup', '1', "
            "'--sample-rows', '20000']; rc = bench.main(); maps = open('/proc/self/maps').read(); "
            "assert rc == 0; assert 'libfastsparse_b200' not in sys.modules, 'package imported'; "
            "assert 'libfastsparse_b200.so' not in maps, 'product .so mapped'; assert 'libfsoracle' in maps or 'libfsref' in maps")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_both_arms_print_the_same_config_object():
    import bench
    for n in (1, 2, 8):
        c = bench.config_for("c2", n)
        assert "workload" in c and "model" not in c and ("row-partitioned" in c["parallelism"]) == (n > 1)
    # the reference arm's line carries exactly config_for(workload, gpus)
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--workload", "c2_small", "--steps", "1", "--warmup", "1",
                        "--sample-rows", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["config"] == bench.config_for("c2_small", 2) and line["scaling"] == "strong"


def test_reference_arm_default_is_the_whole_workload():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "c2_small", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["detail"]["whole_workload"] is True and "the full c2_small matrix (1000000 rows, 20000000 nnz)" in line["cpu_baseline"]["sample"]


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use every usable core."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--workload", "c2_small", "--steps", "1", "--warmup", "1",
                        "--sample-rows", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    usable = len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["cores"] == usable, (line["cpu_baseline"], usable)
    assert line["n_gpus"] == 2


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_without_a_gpu(have_gpu):
    if have_gpu:
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
