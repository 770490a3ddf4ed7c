import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU and libtetwild_gpu.so (run with -m gpu)")


def unhex(lst, shape=None):
    a = np.array([float.fromhex(x) for x in lst], dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def harness():
    """The product's __host__ __device__ numeric core compiled as plain C++ (tests/host_harness.cpp)."""
    import ctypes as C
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_build", "libhost_harness.so")
    src = os.path.join(here, "host_harness.cpp")
    deps = [src, os.path.join(ROOT, "tetwild_b200", "csrc", "tw_math.cuh"), os.path.join(ROOT, "tetwild_b200", "csrc", "sampling.cuh"),
            os.path.join(ROOT, "tetwild_b200", "csrc", "winding_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", out, src])
    H = C.CDLL(out)
    H.hh_tri_sqdist.restype = C.c_double
    H.hh_sample_triangle.restype = C.c_uint64
    H.hh_amips_energy.restype = C.c_double
    H.hh_angle_sum.restype = C.c_double
    H.hh_norm3.restype = C.c_double
    H.hh_box_d2_lb.restype = C.c_float
    H.hh_tri_bound_lb2.restype = C.c_double
    return H


@pytest.fixture(scope="session")
def ctx():
    """A live GPU context. GPU tests must run on the CUDA library: no skip-on-missing, they fail loudly."""
    import tetwild_b200 as tw
    c = tw.Context(0)
    yield c
    c.close()


def amips_close(T, got, ref, tol=1e-9):
    """AMIPS parity criterion (DESIGN.md "AMIPS tolerance"): per tet, every component of E / J / H agrees within
    tol * the natural magnitude of that tensor: max(|ref|_inf, E/l) for J and max(|ref|_inf, E/l^2) for H, with l the
    rms edge length of the tet (J of a perfectly regular tet is exactly 0, so a pure relative test is meaningless
    there). Returns the worst ratio error/scale for (E, J, H)."""
    T = np.asarray(T)
    X = T.T.reshape(-1, 4, 3)
    ed = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    l2 = sum(((X[:, a] - X[:, b]) ** 2).sum(1) for a, b in ed) / 6.0
    l = np.sqrt(l2)
    Eg, Jg, Hg = got
    Er, Jr, Hr = ref
    worst = []
    worst.append(float((np.abs(Eg - Er) / np.abs(Er)).max()))
    sj = np.maximum(np.abs(Jr).max(1), np.abs(Er) / l)
    worst.append(float((np.abs(Jg - Jr).max(1) / sj).max()))
    sh = np.maximum(np.abs(Hr.reshape(len(Hr), -1)).max(1), np.abs(Er) / l2)
    worst.append(float((np.abs(Hg.reshape(len(Hg), -1) - Hr.reshape(len(Hr), -1)).max(1) / sh).max()))
    return worst
