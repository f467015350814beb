"""ASAN + UBSAN (+ -Wall -Wextra -Werror) build of the C++ CPU numerics and their self-test (SURVEY §5.2:
the reference has undefined shifts and races; here the format core must be clean under the sanitizers)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_cpu_ops_selftest_under_asan_ubsan():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "sanitize_cpu.sh")], capture_output=True, text=True, timeout=580)
    if r.returncode != 0 and ("cannot find -lasan" in r.stderr or "cannot find -lubsan" in r.stderr or "libasan" in r.stderr and "No such file" in r.stderr):
        pytest.skip("sanitizer runtimes not installed")
    assert r.returncode == 0, (r.stdout[-2000:] + r.stderr[-4000:])
    assert "0 failure(s)" in r.stdout
