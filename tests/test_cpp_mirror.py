"""The C++ mirror of the reference API (plonk_gadgets_b200/host/plonk_gadgets.hpp) replaying the reference's six integration
tests (host/reference_tests.cpp): on the GPU through libpg_b200.so (`-m gpu`), and on the CPU through the test-only host
backend (checks the mirror's own logic: argument order, error mapping, batching of the cases)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "plonk_gadgets_b200", "host")


def test_cpp_mirror_on_host_backend():
    from tests.test_emu_engine import _build
    emu = _build("libpg_emu.so", "engine_emu.cpp")
    out = os.path.join(os.path.dirname(emu), "reference_tests_emu.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(HOST, "reference_tests.cpp"), "-L" + os.path.dirname(emu), "-lpg_emu",
                    "-Wl,-rpath," + os.path.dirname(emu), "-o", out], check=True)
    res = subprocess.run([out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all passed" in res.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu():
    exe = os.path.join(HOST, "reference_tests.bin")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all passed" in res.stdout
