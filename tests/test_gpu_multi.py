"""GPU tier: the multi-device context (twg_create_multi, csrc/multi.cu) -- handles replicated per device, host-buffer batches
split by contiguous index range, one host thread per device, results written straight into the caller's arrays. Every result
must be bit-identical to the one-device call. On a one-GPU box the split runs over two worker contexts of device 0; with two
or more GPUs it runs over real devices (gpurun --gpus 2)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from tetwild_b200 import synth
import tetwild_b200 as tw

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def device_sets():
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


@pytest.mark.parametrize("devs", device_sets())
def test_multi_context_matches_single_device(devs, ctx, oracle):
    m = tw.Context(devs)
    assert m.num_devices == len(devs)
    V, F = synth.torus_knot(400, 60)
    sd, eps, eps2 = synth.state_eps(2e-3)
    S1, Sm = tw.Surface(ctx, V, F), tw.Surface(m, V, F)
    P = synth.envelope_points(V, F, 700_001, eps, seed=4)       # odd: ragged split
    a = S1.points_out(P, eps2)
    b = Sm.points_out(P, eps2)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:50000], oracle.Surface(V, F).points_out(P[:50000], eps2, threads=8))
    assert np.array_equal(Sm.points_out(P[:1000], eps2), a[:1000])     # small batch: device 0 only
    f1, q1, d1 = S1.nearest(P[:300_000])
    fm, qm, dm = Sm.nearest(P[:300_000])
    # d2 is the exact minimum either way; between equidistant facets (queries on shared edges) the winner depends on how the
    # batch was cut into packets, so ids / points are compared where the ids agree
    same = f1 == fm
    assert np.array_equal(d1, dm) and same.mean() > 0.8 and np.array_equal(q1[same], qm[same])
    assert np.allclose(((P[:300_000] - qm) ** 2).sum(1), dm, rtol=1e-6, atol=1e-18)
    T = synth.face_queries(V, F, 20_001, 0.01, eps, seed=2)
    assert np.array_equal(S1.faces_out(T, sd, eps2), Sm.faces_out(T, sd, eps2))
    # winding: one hierarchy built on the host, uploaded to every device
    Vs, Fs = synth.uv_sphere(120, 120)
    Q = synth.winding_queries(Vs, 400_003, seed=3)
    W1, k1 = tw.Winding(ctx, Vs, Fs).eval(Q)
    Wm, km = tw.Winding(m, Vs, Fs).eval(Q)
    # a warp of 32 Morton-neighbours shares one traversal, so cutting the batch differently regroups the queries and changes the
    # order in which a query's terms are added: W agrees to rounding, the decisions exactly
    assert np.abs(W1 - Wm).max() < 1e-12 and np.array_equal(k1, km)
    keep, retried = m.inout_filter(Vs, Fs[:, [0, 2, 1]], Q[:200_000])      # flip-and-retry through the split path
    assert retried and np.array_equal(keep, k1[:200_000])
    # flat AMIPS batch
    X = synth.random_tets(600_001, seed=5)
    E1, J1, H1 = ctx.amips_ejh_soa(X)
    Em, Jm, Hm = m.amips_ejh_soa(X)
    assert np.array_equal(E1, Em) and np.array_equal(J1, Jm) and np.array_equal(H1, Hm)
    Vt, Tt = synth.grid_tet_mesh(40, 40, 40)
    assert np.array_equal(ctx.amips_quality(Vt, Tt), m.amips_quality(Vt, Tt))
    # the resident mesh lives on device 0 of the multi context
    M = tw.TetMesh(m, Vt[:2000], Tt[(Tt < 2000).all(1)])
    assert np.isfinite(M.quality()).all()
    M.close()
    assert m.launches > 0
    S1.close(); Sm.close()
    m.close()


def test_multi_context_from_cpp():
    """tests/cpp/test_multi.cpp: the C++ adapter layer (twg::Context over several devices) drives the same split"""
    from tetwild_b200 import build
    build.build()
    exe = os.path.join(ROOT, "tests", "_build", "test_multi")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_multi.cpp"), "-o", exe, "-L" + os.path.join(ROOT, "tetwild_b200"), "-ltetwild_gpu",
                           "-Wl,-rpath," + os.path.join(ROOT, "tetwild_b200")])
    n = torch.cuda.device_count()
    r = subprocess.run([exe, str(max(2, min(n, 8))), "1" if n >= 2 else "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok: 0 failures" in r.stdout
