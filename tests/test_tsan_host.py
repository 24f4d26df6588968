"""ThreadSanitizer run of the library's HOST code on the CPU (no GPU needed): csrc/solver.cu (contexts, caches, point
store, lexicographic chain), csrc/generator.cpp (generator, worker pool with work stealing, cross-rank record exchange,
cooperative workers) and the CPLEX seam, compiled with g++ -fsanitize=thread against the stand-ins of
tests/tsan/cuda_double.cpp (`make -C oracle tsan`) and driven by tests/tsan/tsan_host.cpp with 12 pool workers, W = k
cooperative workers and 6 seam threads.  Any data race report fails the test; every front is compared with the
single-context run inside the driver."""
from __future__ import annotations

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_build", "tsan_host")


@pytest.fixture(scope="module")
def tsan_bin():
    cxx = os.environ.get("TSAN_CXX", "/usr/bin/g++")
    if not (os.path.exists(cxx) or shutil.which(cxx)) or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("needs a g++ with libtsan and the CUDA headers")
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "tsan"], capture_output=True, text=True)
    if r.returncode != 0 and "cannot find -ltsan" in r.stderr:
        pytest.skip("this compiler has no libtsan")
    assert r.returncode == 0, r.stderr[-3000:]
    return BIN


@pytest.mark.parametrize("kind,n,k,seed,reps", [("kp", 14, 3, 5, 25), ("ap", 4, 3, 2, 25), ("kp", 10, 4, 3, 10)])
def test_host_code_is_race_free(tsan_bin, tmp_path, kind, n, k, seed, reps):
    from moip_aira_b200 import instances
    path = str(tmp_path / "m.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0", MOIP_BOX_BUDGET="1")
    r = subprocess.run([tsan_bin, path, str(reps), "12"], capture_output=True, text=True, env=env, timeout=900)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0 and "TSAN_HOST_OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
