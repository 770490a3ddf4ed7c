"""The C++ host adapters (include/tetwild_gpu.hpp) -- the reference-signature layer TetWild's own C++ would include.
CPU tier: the header and the reference-style driver compile warning-free as C++11 and link against the C-ABI library.
GPU tier: the driver (tests/cpp/test_adapters.cpp) runs on the device and checks every answer against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EXE = os.path.join(ROOT, "tests", "_build", "test_adapters")


def _build():
    from tetwild_b200 import build
    import oracle
    build.build()
    oracle.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src = os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp")
    deps = [src, os.path.join(ROOT, "include", "tetwild_gpu.hpp"), os.path.join(ROOT, "include", "tetwild_gpu.h"),
            os.path.join(ROOT, "tetwild_b200", "libtetwild_gpu.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return
    cmd = [cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), src, "-o", EXE,
           "-L" + os.path.join(ROOT, "tetwild_b200"), "-ltetwild_gpu", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
           "-Wl,-rpath," + os.path.join(ROOT, "tetwild_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)


def test_adapters_compile_and_link():
    _build()
    assert os.path.exists(EXE)
    # without a device the adapters must refuse loudly (twg::Error from Context), never compute on the CPU
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_adapters_against_oracle():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok: 0 failures" in r.stdout
