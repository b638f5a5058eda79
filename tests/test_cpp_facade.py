"""The C++ facade (include/vdbrt/RayTracer.h): client code in the reference's API shape compiles against it (CPU check)
and reproduces BASELINE config 1 on the GPU (606 028 hit pixels at 1024^2)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def test_facade_client_compiles():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_facade_client_runs_config1():
    r = subprocess.run([EXE, "1024"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "606028 hit pixels" in r.stdout and "facade ok" in r.stdout
