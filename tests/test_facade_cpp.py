"""The C++ facades (include/fastlio_b200/*.hpp) compile against a PointType / state / dyn_share look-alike exactly as
src/laserMapping.cpp would use them; the binary runs the map + callback on the GPU when one is present."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_smoke")


def _build():
    from better_fastlio2_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp"), "-L", libdir, "-lfastlio_b200", f"-Wl,-rpath,{libdir}",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_facades_compile_and_fail_loudly_without_gpu():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.stdout, out.stderr)
    assert "NO_GPU compile-only ok" in out.stdout or "FACADE_OK" in out.stdout


@pytest.mark.gpu
def test_facades_run_on_gpu():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "FACADE_OK" in out.stdout
