"""AddressSanitizer + UndefinedBehaviorSanitizer over the host-only C++ of the boundary (the
reference has no sanitizer run; this repo's device side has compute-sanitizer runs under profiles/).  CPU only.

* the BIF / DSC loaders on every prefix and on thousands of single-byte mutations of valid files: load or throw,
  never an out-of-bounds read, an overflow or an abort (tests/cpp/fuzz_serializer.cpp);
* the repo's own C++ test of the loaders (tests/cpp/test_serializer.cpp) rebuilt with the sanitizers on."""
import os
import shutil
import subprocess

import pytest

from bayesiannetwork_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAN = ["-std=c++11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
       "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp")]
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def _build(tmp_path, source, name):
    exe = str(tmp_path / name)
    r = subprocess.run(["g++", *SAN, "-o", exe, os.path.join(ROOT, "tests", "cpp", source)], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libasan / libubsan not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _run(cmd):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    return r.stdout


def test_network_file_loaders_survive_truncated_and_mutated_files(tmp_path):
    from bayesiannetwork_b200 import _build as libbuild, netfile
    libbuild.build()
    exe = _build(tmp_path, "fuzz_serializer.cpp", "fuzz_serializer")
    out = _run([exe, os.path.join(ROOT, "tests", "golden", "asia.bif"), "1500"])
    assert "refused" in out
    # the same through the DSC loader, on a file written from a network with 2-4 states and up to 3 parents per node
    net = synth.random_dag(12, 3, 2, 4, seed=4)
    dsc = tmp_path / "dag12.dsc"
    dsc.write_text(netfile.dump_dsc(net))
    out = _run([exe, str(dsc), "1200"])
    assert "refused" in out
    bif = tmp_path / "dag12.bif"
    bif.write_text(netfile.dump_bif(net, order="descending"))
    _run([exe, str(bif), "600"])


def test_own_loader_tests_under_the_sanitizers(tmp_path):
    exe = _build(tmp_path, "test_serializer.cpp", "test_serializer_san")
    out = _run([exe])
    assert "No errors detected" in out or "no errors" in out.lower() or out == "" or "passed" in out.lower()
