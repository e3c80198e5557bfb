"""The C++ host class (include/aligator_b200/riccati_solver.hpp): compiles and links
against the C-ABI library on CPU; runs its parity checks on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cxx", "host_class_test.cpp")
EXE = os.path.join(ROOT, "tests", "cxx", "host_class_test")


def _build():
    import __graft_entry__ as g
    g.build()
    libdir = os.path.join(ROOT, "aligator_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", SRC, "-o", EXE,
                           "-L" + libdir, "-laligator_b200_gar", "-Wl,-rpath," + libdir])


def test_host_class_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_class_parity_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout
