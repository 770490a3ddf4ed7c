"""GPU tier: the one-launch tiny-call paths (arguments in the kernel parameters, results in the mapped slab, completion word raised
by the kernel: points_tiny_kernel, env_faces_tiny_kernel, amips_ring_tiny_kernel, mesh_quality_tiny_kernel) give what the batched
paths give for the same units -- at one unit, at the largest size a path takes, and one past it (the next path)."""
import numpy as np
import pytest

from tetwild_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(ctx):
    import tetwild_b200 as tw
    V, F = synth.icosphere(4)
    V = synth.normalise_unit_diag(V)
    sd, eps, eps2 = synth.state_eps(1e-3)
    S = tw.Surface(ctx, V, F)
    Vm, Tm = synth.grid_tet_mesh(14, 13, 12)
    M = tw.TetMesh(ctx, Vm, Tm)
    M.build_rings()
    yield {"S": S, "M": M, "V": V, "F": F, "Vm": Vm, "Tm": Tm, "sd": sd, "eps": eps, "eps2": eps2}
    M.close()
    S.close()


def test_tiny_points_and_nearest_equal_batched(ctx, scene):
    S, eps, eps2 = scene["S"], scene["eps"], scene["eps2"]
    P = synth.envelope_points(scene["V"], scene["F"], 6000, eps, seed=4)
    ref_out = S.points_out(P, eps2)                    # sorted, batched kernel
    ref_f, ref_n, ref_d = S.nearest(P)
    assert 0.05 < ref_out.mean() < 0.95
    for n in (1, 2, 31, 64, 65):
        for b in (0, 777, 5000):
            sl = slice(b, b + n)
            assert np.array_equal(S.points_out(P[sl], eps2), ref_out[sl])
            f, q, d = S.nearest(P[sl])
            assert np.array_equal(d, ref_d[sl])                                        # exact minimum either way
            same = f == ref_f[sl]
            assert np.array_equal(q[same], ref_n[sl][same])                            # ties may name another facet at the same distance
            r = ((P[sl] - q) ** 2).sum(1)                                              # the returned point realises d2 (up to its own rounding)
            assert (np.abs(r - d) <= 1e-12 * d + 1e-14 * np.sqrt(d) + 1e-18).all()   # 1e-18: the quadratic form of the distance routine cancels to ~1e-20 absolute for points ON a facet


def test_tiny_faces_equal_batched(ctx, scene):
    S, sd, eps, eps2 = scene["S"], scene["sd"], scene["eps"], scene["eps2"]
    T = synth.face_queries(scene["V"], scene["F"], 600, 0.03, eps, seed=8)
    T[4] = np.tile(T[4][:3], 3)                                                        # a degenerate face (all three vertices equal)
    T[5, 6:9] = 0.5 * (T[5, 0:3] + T[5, 3:6])                                          # and a collinear one
    ref = S.faces_out(T, sd, eps2)
    ref_nd = S.faces_out(T, sd, eps2, degenerate_shortcut=False)
    assert 0.05 < ref.mean() < 0.95
    for n in (1, 3, 16, 17):
        for b in (0, 4, 300):
            sl = slice(b, b + n)
            assert np.array_equal(S.faces_out(T[sl], sd, eps2), ref[sl])
            assert np.array_equal(S.faces_out(T[sl], sd, eps2, degenerate_shortcut=False), ref_nd[sl])


def test_tiny_mesh_calls_equal_batched(ctx, scene):
    M, Vm, Tm = scene["M"], scene["Vm"], scene["Tm"]
    rng = np.random.default_rng(6)
    tids = rng.choice(len(Tm), 5000, replace=False).astype(np.int32)
    ref_q = M.quality(tids)
    for n in (1, 7, 256, 257):
        for b in (0, 1234):
            assert np.array_equal(M.quality(tids[b:b + n]), ref_q[b:b + n])
    vids = rng.choice(len(Vm), 2500, replace=False).astype(np.int32)
    ref = M.vertex_ring_ejh(vids)
    X = Vm[vids] + rng.normal(0, 0.01, size=(len(vids), 3))
    ref_e = M.vertex_trial_energy(vids, X)
    for n in (1, 5, 32, 33):
        for b in (0, 900):
            got = M.vertex_ring_ejh(vids[b:b + n])
            for x, y in zip(got, ref):
                assert np.array_equal(x, y[b:b + n], equal_nan=True)
            assert np.array_equal(M.vertex_trial_energy(vids[b:b + n], X[b:b + n]), ref_e[b:b + n])
    # repeated tiny calls reuse the slab and the completion word: a long run stays consistent
    for k in range(300):
        j = k % 2400
        assert M.quality(tids[j:j + 3])[1] == ref_q[j + 1]
        assert M.vertex_trial_energy(vids[j:j + 1], X[j:j + 1])[0] == ref_e[j]
