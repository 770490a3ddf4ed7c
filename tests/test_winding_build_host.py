"""CPU tier: the winding-number hierarchy is built on the HOST (csrc/winding.cu::build_host_tree: kd order, leaf boxes,
exterior-edge caps traced into polylines, on several threads). tools/wbuild_prof.cu compiles that very source for the host
and prints checksums of the three arrays the device receives; the build must be deterministic (thread scheduling must not
show), and a closed manifold surface must have an empty cap at the root (its boundary is empty)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
EXE = os.path.join(ROOT, "tests", "_build", "wbuild_prof")


@pytest.fixture(scope="module")
def tool():
    src = os.path.join(ROOT, "tools", "wbuild_prof.cu")
    csrc = os.path.join(ROOT, "tetwild_b200", "csrc")
    deps = [src, os.path.join(csrc, "winding.cu"), os.path.join(csrc, "winding.cuh"), os.path.join(csrc, "winding_build.cu"), os.path.join(csrc, "common.cuh")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.check_call(["nvcc", "-ccbin", "/usr/bin/g++", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                               "-o", EXE, src, os.path.join(csrc, "ctx.cu"), os.path.join(csrc, "multi.cu"), os.path.join(csrc, "qsort.cu"), os.path.join(csrc, "winding_build.cu")], cwd=os.path.join(ROOT, "tools"))
    return EXE


def run(tool, n):
    out = subprocess.run([tool, str(n)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"checksums nodes (\w+) caps (\w+) tris (\w+)", out.stdout)
    s = re.search(r"facets (\d+) nodes (\d+) cap points (\d+)", out.stdout)
    r = re.search(r"root cap points (\d+), children cap points (\d+) (\d+)", out.stdout)
    return m.groups(), tuple(int(x) for x in s.groups()) + tuple(int(x) for x in r.groups()), out.stderr


@pytest.mark.parametrize("n", [5, 40, 331])
def test_build_is_deterministic(tool, n):
    a, sa, _ = run(tool, n)
    for _ in range(2):
        b, sb, _ = run(tool, n)
        assert a == b and sa == sb
    facets, nodes, caps, root_cap, left_cap, right_cap = sa
    # a closed manifold surface has no boundary: nothing at the root; its two halves share one cut (same loop, reversed;
    # the polylines may be cut at different apex vertices, so the point counts agree only roughly)
    assert root_cap == 0
    if nodes > 4:
        assert left_cap > 0 and right_cap > 0 and abs(left_cap - right_cap) <= 4 + 0.2 * left_cap
    assert facets == 2 * n * (n - 1) and nodes >= 2 and (nodes & (nodes - 1)) == 0
    # caps are boundaries of sub-meshes: a few sqrt(facets) points per node at most, nothing like the facet count
    assert caps < 40 * nodes * max(1.0, (facets / nodes) ** 0.5)


def test_phase_timers_report(tool):
    _, _, err = run(tool, 120)
    for phase in ("1 vertex merge", "2 kd order", "4b half-edge sort", "5 caps traced"):
        assert phase in err
