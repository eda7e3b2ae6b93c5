"""CPU: the host thread pool behind the *_host gathers (csrc/host_pool.cpp) — plain and non-temporal widening, the streamed form
against a producer thread that raises arrival flags out of order, and a producer that dies half way (tests/cpp/test_host_pool.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_pool")


@pytest.fixture(scope="module")
def exe():
    cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_host_pool.cpp"),
           os.path.join(ROOT, "q-learning_b200", "csrc", "host_pool.cpp"), "-o", EXE]
    env = dict(os.environ); env.pop("CC", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return EXE


@pytest.mark.parametrize("threads", ["1", "3", ""])
def test_host_pool(exe, threads):
    env = dict(os.environ)
    if threads:
        env["QLC_HOST_THREADS"] = threads
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "host pool ok" in r.stdout, r.stdout + r.stderr
