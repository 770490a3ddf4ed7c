"""GPU tier: envelope / nearest-facet kernels through the C ABI against the oracle. Decisions and squared distances
are bit exact; facet ids are compared where the minimum is unique."""
import numpy as np
import pytest

from conftest import load_golden, unhex
from tetwild_b200 import synth
import tetwild_b200 as tw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knot(ctx, oracle):
    V, F = synth.torus_knot(200, 40)
    return V, F, tw.Surface(ctx, V, F), oracle.Surface(V, F)


def test_points_out_exact(knot):
    V, F, S, OS = knot
    for eps_rel in (1e-3, 4e-3):
        sd, eps, eps2 = synth.state_eps(eps_rel)
        P = synth.envelope_points(V, F, 200000, eps, seed=20240501)
        got = S.points_out(P, eps2)
        ref = OS.points_out(P, eps2, threads=4)
        assert np.array_equal(got, ref)
        assert 0.2 < got.mean() < 0.8
    assert np.array_equal(S.points_out(P[:3000], eps2), OS.points_out(P[:3000], eps2, brute=True, threads=4))


def test_tree_golden(ctx):
    g = load_golden("tree_golden.json")
    V, F = synth.torus_knot(60, 12)
    S = tw.Surface(ctx, V, F)
    P = unhex(g["P"], (-1, 3))
    assert np.array_equal(S.points_out(P, float.fromhex(g["eps2"])), np.array(g["out"], dtype=np.uint8))
    f, q, d = S.nearest(P)
    assert np.array_equal(d, unhex(g["nearest_d2"]))
    same = f == np.array(g["nearest_facet_original_ids"], dtype=np.uint32)
    assert same.mean() > 0.9
    assert np.array_equal(q[same], unhex(g["nearest_pt"], (-1, 3))[same])


def test_nearest_exact(knot):
    V, F, S, OS = knot
    sd, eps, eps2 = synth.state_eps(1e-3)
    P = synth.envelope_points(V, F, 100000, eps, seed=5)
    f, q, d = S.nearest(P)
    fr, qr, dr = OS.nearest(P, threads=4)
    assert np.allclose(d, dr, rtol=1e-14, atol=0)      # squared_distance(): equal up to the reference's own tie pruning
    assert np.array_equal(d[:5000], OS.sqdist_brute(P[:5000], threads=4)[0])   # and bit exact against the true minimum
    same = f == fr
    assert same.mean() > 0.7                           # ties (points exactly on shared edges) may pick either facet
    assert np.array_equal(q[same], qr[same])
    assert np.allclose(((P - q) ** 2).sum(1), d, rtol=1e-6, atol=1e-18)
    assert np.array_equal(S.squared_distance(P[:1000]), d[:1000])
    # isPointOutEnvelop goes through squared_distance() > eps_2 (LocalOperations.cpp:1037): same decision
    assert np.array_equal((d > eps2).astype(np.uint8), S.points_out(P, eps2))


def test_edge_cases(ctx, oracle):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    for F in (np.array([[0, 1, 2]]), np.array([[0, 1, 2], [0, 1, 3]]), np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3]])):
        S = tw.Surface(ctx, V, F.astype(np.uint32))
        OS = oracle.Surface(V, F.astype(np.uint32))
        P = np.random.default_rng(0).uniform(-0.5, 1.5, size=(5000, 3))
        P[:100, 2] = 0.0
        for eps2 in (0.0, 1e-4, 0.3):
            assert np.array_equal(S.points_out(P, eps2), OS.points_out(P, eps2))
        d = S.nearest(P)[2]
        assert np.array_equal(d, OS.sqdist_brute(P)[0])          # = min over all facets of the per-facet d2, bit exact
        assert np.allclose(d, OS.nearest(P)[2], rtol=1e-14, atol=0)  # the reference's pruned search can sit 1 ulp above on ties
        assert len(S.points_out(np.zeros((0, 3)), 1e-3)) == 0
        assert np.array_equal(S.points_out(P[:1], 1e-4), OS.points_out(P[:1], 1e-4))
    # degenerate facets (the reference's boundary mesh stores edges as degenerate triangles, Preprocess.cpp:192-197)
    Fd = np.array([[0, 1, 1], [1, 2, 2], [2, 3, 3], [0, 1, 2]], dtype=np.uint32)
    S, OS = tw.Surface(ctx, V, Fd), oracle.Surface(V, Fd)
    assert np.array_equal(S.nearest(P)[2], OS.sqdist_brute(P)[0])
    # an edge stored both as a degenerate facet and as the side of a real triangle evaluates to two d2 values a few ulps
    # apart; the reference's pruned search keeps whichever it meets first (measured 1.8e-14 relative), the device keeps the min
    assert np.allclose(S.nearest(P)[2], OS.nearest(P)[2], rtol=1e-12, atol=0)
    assert np.array_equal(S.points_out(P, 1e-2), OS.points_out(P, 1e-2))


def test_sample_triangle_device_bit_exact(ctx, oracle):
    rng = np.random.default_rng(4)
    for it in range(60):
        tri = rng.normal(size=(3, 3)) * rng.choice([0.0005, 0.003, 0.01, 0.05])
        if it % 5 == 0:
            tri = np.round(tri * 100) / 100
        if it % 7 == 0:
            tri = np.array([[0, 0, 0], [0.05, 0, 0], [0, 0.05, 0]]) + rng.integers(-3, 3, size=3)
        sd = 1e-3 if it % 2 else 2.5e-3
        assert np.array_equal(ctx.sample_triangle(tri, sd), oracle.sample_triangle(tri, sd))


def test_faces_out_exact(knot, ctx, oracle):
    V, F, S, OS = knot
    # small candidate faces (a handful of samples each), eps / sampling_dist through State.cpp:36-41
    for eps_rel, edge in ((2e-3, 0.004), (1e-3, 0.006), (4e-3, 0.01)):
        sd, eps, eps2 = synth.state_eps(eps_rel)
        T = synth.face_queries(V, F, 1500, edge, eps, seed=int(edge * 1e4))
        T[::97] = np.array([0, 0, 5, 1, 1, 6, 2, 2, 7.0])      # collinear -> IN (LocalOperations.cpp:1048)
        got = S.faces_out(T, sd, eps2)
        ref, ns = OS.faces_out(T, sd, eps2, threads=4)
        assert np.array_equal(got, ref)
        assert 0.03 < got.mean() < 0.97
    # faces of the surface itself are IN; pushed far away they are OUT
    tri = V[F[:500].astype(np.int64)].reshape(-1, 9)
    assert not S.faces_out(tri, sd, eps2).any()
    assert S.faces_out(tri + 0.05, sd, eps2).all()
    # BASELINE config-1-shaped calls: 20 480-triangle icosphere, faces of edge ~ diag/20 -> ~1.1k samples each
    Vi, Fi = synth.icosphere(5)
    Vi = synth.normalise_unit_diag(Vi)
    Si, OSi = tw.Surface(ctx, Vi, Fi), oracle.Surface(Vi, Fi)
    for sd, eps, edge, sig in ((1e-3, 3e-3, 0.05, 4e-3), (1e-3, 2e-3, 0.03, 3e-3), (5e-4, 1.5e-3, 0.02, 2e-3)):
        T = synth.face_queries(Vi, Fi, 600, edge, sig, seed=int(edge * 1e4))
        got = Si.faces_out(T, sd, eps * eps)
        ref, ns = OSi.faces_out(T, sd, eps * eps, threads=4)
        assert np.array_equal(got, ref) and 0.1 < got.mean() < 0.6 and ns.mean() > 400
    assert len(Si.faces_out(np.zeros((0, 9)), 1e-3, 1e-6)) == 0


def test_faces_golden(ctx):
    """decisions recorded from the reference-built composition of LocalOperations.cpp:1046-1109 (tests/golden/faces_golden.json)"""
    g = load_golden("faces_golden.json")
    V, F = synth.torus_knot(60, 12)
    S = tw.Surface(ctx, V, F)
    T = unhex(g["tris"], (-1, 9))
    sd = float.fromhex(g["sd"])
    assert list(S.faces_out(T, sd, float.fromhex(g["eps2"]))) == g["out"]
    assert list(S.faces_out(T, sd, float.fromhex(g["preprocess_eps2"]), degenerate_shortcut=False)) == g["preprocess_out"]


def test_preprocess_face_test_has_no_degenerate_shortcut(knot, oracle):
    """Preprocess::isOutEnvelop (Preprocess.cpp:643-747) samples every face of the candidate set, degenerate or not;
    LocalOperations::isFaceOutEnvelop_sampling returns IN for a degenerate face (:1048)."""
    V, F, S, OS = knot
    sd, eps, eps2 = synth.state_eps(1e-3)
    eps2 *= 0.8 * 0.8                                  # Preprocess.cpp:201-205
    rng = np.random.default_rng(12)
    T = synth.face_queries(V, F, 800, 0.006, eps, seed=77)
    # exactly collinear faces: vertices on a 2^-20 grid, third vertex = a + k d with k in {1/2, 2, 3} (exact in double)
    a = np.round(T[::5, 0:3] * 2 ** 20) / 2 ** 20
    d = np.round((T[::5, 3:6] - a) * 2 ** 19) / 2 ** 19
    T[::5, 0:3], T[::5, 3:6] = a, a + d
    T[::5, 6:9] = a + d * rng.choice([0.5, 2.0, 3.0], (len(a), 1))
    T[3::50, 3:6] = T[3::50, 0:3]; T[3::50, 6:9] = T[3::50, 0:3]      # three coincident vertices
    got = S.faces_out(T, sd, eps2, degenerate_shortcut=False)
    ref, _ = OS.faces_out(T, sd, eps2, threads=4, degenerate_shortcut=False)
    assert np.array_equal(got, ref)
    short = S.faces_out(T, sd, eps2)
    assert not short[::5].any() and got[::5].any()                    # the shortcut hides OUT degenerate faces
    assert np.array_equal(short, OS.faces_out(T, sd, eps2, threads=4)[0])


def test_faces_large_and_tiny(ctx, oracle):
    """Face sizes that drive every path of the face kernel: tiny faces (3 samples, Common.cpp:154-158), faces taller than
    one pass of the run table (hundreds of lattice rows), faces whose candidate-facet list overflows (per-sample descents),
    on a planar grid where IN / OUT is known, and against the oracle when tilted."""
    g = 24
    x = np.linspace(0, 1, g + 1)
    X, Y = np.meshgrid(x, x, indexing="ij")
    V = np.stack([X.ravel(), Y.ravel(), 0 * X.ravel()], 1)
    idx = np.arange((g + 1) * (g + 1)).reshape(g + 1, g + 1)
    F = np.concatenate([np.stack([idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:]], -1).reshape(-1, 3),
                        np.stack([idx[:-1, :-1], idx[1:, 1:], idx[:-1, 1:]], -1).reshape(-1, 3)]).astype(np.uint32)
    S, OS = tw.Surface(ctx, V, F), oracle.Surface(V, F)
    sd, eps = 1e-3, 5e-4
    rng = np.random.default_rng(8)
    faces = []
    for edge in (3e-4, 0.02, 0.11, 0.45):          # 3 samples / ~300 / ~8 k / ~120 k samples, 1 .. 390 rows
        for _ in range(6):
            c = rng.uniform(0.3, 0.7, 2)
            a = rng.uniform(0, 2 * np.pi) + np.array([0, 2.1, 4.2])
            t = np.stack([c[0] + edge / 1.7 * np.cos(a), c[1] + edge / 1.7 * np.sin(a), np.zeros(3)], 1)
            faces.append(t)
    T = np.array(faces).reshape(-1, 9)
    assert not S.faces_out(T, sd, eps * eps).any()                     # in the plane: IN, every sample visited
    Tup = T.copy(); Tup[:, 2::3] += 2 * eps
    assert S.faces_out(Tup, sd, eps * eps).all()                       # lifted by 2 eps: OUT
    Ttilt = T.copy(); Ttilt[:, 8] += rng.uniform(0, 4 * eps, len(T))   # one corner lifted: mixed, decided by a few samples
    got = S.faces_out(Ttilt, sd, eps * eps)
    ref, ns = OS.faces_out(Ttilt, sd, eps * eps, threads=4)
    assert np.array_equal(got, ref) and 0 < got.sum() < len(T) and ns.max() > 50000


def test_full_size_config2(ctx, oracle):
    """BASELINE config 2 at full size (200 000 triangles, 10 M points): subsample vs the oracle + properties."""
    V, F = synth.torus_knot(1000, 100)
    assert len(F) == 200000
    S = tw.Surface(ctx, V, F)
    sd, eps, eps2 = synth.state_eps(1e-3)
    n = 10_000_000
    rng = np.random.default_rng(1)
    # cheap full-size generator: facet barycentric samples with normal offsets (half of them exactly on facets)
    tri = V[F.astype(np.int64)]
    f = rng.integers(0, len(F), size=n)
    w = rng.dirichlet([1, 1, 1], size=n)
    P = (tri[f] * w[:, :, None]).sum(1)
    nrm = np.cross(tri[f, 1] - tri[f, 0], tri[f, 2] - tri[f, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    off = rng.normal(0, eps, size=n)
    off[: n // 2] = 0.0
    P = P + nrm * off[:, None]
    out = S.points_out(P, eps2)
    # points within eps*(1-1e-6) of their own facet's plane offset are IN; idempotent; order independent
    assert not out[np.abs(off) < eps * (1 - 1e-6)].any()
    idx = rng.choice(n, 200000, replace=False)
    OS = oracle.Surface(V, F)
    assert np.array_equal(out[idx], OS.points_out(P[idx], eps2, threads=8))
    perm = rng.permutation(n)[:2_000_000]
    assert np.array_equal(S.points_out(P[perm], eps2), out[perm])
