"""The reference-side adapter (integration/aligator/gar/b200-riccati.hpp -- what a maintainer adds to aligator)
compiled against stand-in Eigen / aligator types (tests/cxx/aligator_stub; the image has no Eigen): compiles and
links against the C-ABI library on CPU, runs its parity checks against the oracle on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cxx", "adapter_test.cpp")
EXE = os.path.join(ROOT, "tests", "cxx", "adapter_test")


def _build():
    import __graft_entry__ as g
    g.build()
    libdir = os.path.join(ROOT, "aligator_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", SRC, "-o", EXE,
                           "-I" + os.path.join(ROOT, "tests", "cxx", "aligator_stub"), "-I" + os.path.join(ROOT, "integration"),
                           "-I" + os.path.join(ROOT, "include"),
                           "-L" + libdir, "-laligator_b200_gar", "-Wl,-rpath," + libdir])


def test_adapter_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapter_parity_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ADAPTER OK" in r.stdout
