"""GPU tier: winding-number kernels through the C ABI against the oracle. Decisions (W > 0.5) exact; W within 1e-10."""

import numpy as np
import pytest

from tetwild_b200 import synth
import tetwild_b200 as tw

pytestmark = pytest.mark.gpu


def test_closed_sphere_known_answers(ctx, oracle):
    V, F = synth.uv_sphere(24, 24, noise=0.0)
    r = np.linalg.norm(V, axis=1).max()
    Q = np.array([[0, 0, 0], [0.3 * r, 0.1 * r, -0.2 * r], [2 * r, 0, 0], [0, -3 * r, r]])
    W, keep = ctx.winding_number(V, F, Q)
    assert np.allclose(W, [1, 1, 0, 0], atol=1e-12) and list(keep) == [1, 1, 0, 0]
    W, keep = ctx.winding_number(V, F[:, [0, 2, 1]], Q)
    assert np.allclose(W, [-1, -1, 0, 0], atol=1e-12) and not keep.any()
    W, _ = ctx.winding_number(V, np.concatenate([F, F]), Q)     # repeated faces double W (InoutFiltering.cpp:99-100)
    assert np.allclose(W, [2, 2, 0, 0], atol=1e-12)
    keep, retried = ctx.inout_filter(V, F[:, [0, 2, 1]], Q)      # flip-and-retry (InoutFiltering.cpp:56-75)
    assert retried and list(keep) == [1, 1, 0, 0]
    keep, retried = ctx.inout_filter(V, F, Q)
    assert not retried and list(keep) == [1, 1, 0, 0]


@pytest.mark.parametrize("sort,leaf,device_build", [(1, 32, 1), (1, 64, 1), (0, 64, 1), (1, 16, 1), (1, 32, 0), (0, 256, 0)])
def test_vs_oracle_direct(oracle, sort, leaf, device_build):
    """query order (Morton-sorted or the caller's), leaf block size and where the hierarchy is built (device / host threads)
    are per-context options (twg_set_option); none of them may change a result beyond rounding"""
    c = tw.Context(0)
    c.set_option("winding_sort", sort)
    c.set_option("winding_leaf", leaf)
    c.set_option("winding_device_build", device_build)
    V, F = synth.uv_sphere(90, 90, noise=0.02, seed=3)
    Q = synth.winding_queries(V, 30000, seed=8)
    Wt = tw.Winding(c, V, F)
    W, keep = Wt.eval(Q)
    Wd = oracle.winding_direct(V, F, Q, threads=8)
    assert np.abs(W - Wd).max() < 1e-10
    assert np.array_equal(keep, (Wd > 0.5).astype(np.uint8))
    assert 0.1 < keep.mean() < 0.6
    st = Wt.stats()
    assert st["triangles"] == len(F) and st["cap_segments"] > 0
    Wt.close()
    c.close()


def test_open_and_soup(ctx, oracle):
    # open surface (hemisphere): fractional winding numbers; triangle soup with unmerged duplicate vertices
    V, F = synth.uv_sphere(60, 60, noise=0.0)
    keepf = V[F.astype(np.int64)].mean(1)[:, 2] > 0
    Fo = F[keepf]
    Q = synth.winding_queries(V, 20000, seed=2)
    W, keep = tw.Winding(ctx, V, Fo).eval(Q)
    Wd = oracle.winding_direct(V, Fo, Q, threads=8)
    assert np.abs(W - Wd).max() < 1e-10 and (np.abs(Wd - 0.5) > 1e-9).all()
    assert np.array_equal(keep, (Wd > 0.5).astype(np.uint8))
    Vs = V[F.astype(np.int64)].reshape(-1, 3)                    # every facet owns its 3 vertices
    Fs = np.arange(len(Vs), dtype=np.uint32).reshape(-1, 3)
    Ws, _ = tw.Winding(ctx, Vs, Fs).eval(Q[:5000])
    assert np.abs(Ws - oracle.winding_direct(V, F, Q[:5000], threads=8)).max() < 1e-10
    # tiny inputs and the empty surface
    assert np.allclose(tw.Winding(ctx, V, F[:1]).eval(Q[:100])[0], oracle.winding_direct(V, F[:1], Q[:100]), atol=1e-12)
    W0, k0 = tw.Winding(ctx, V, F[:0]).eval(Q[:10])
    assert not W0.any() and not k0.any()
    assert len(tw.Winding(ctx, V, F).eval(Q[:0])[0]) == 0


def test_self_intersecting_union(ctx, oracle):
    V, F = synth.sphere_union(6, 40, 41, seed=5)
    V2, F2 = synth.uv_sphere(30, 30, radius=0.05, normalise=False)   # plus a small sphere nested at the centroid of sphere 0
    c0 = V[: 2 + 40 * 40].mean(0)
    F = np.concatenate([F, F2 + len(V)]).astype(np.uint32)
    V = np.concatenate([V, V2 + c0])
    Q = np.concatenate([synth.winding_queries(V, 20000, seed=4), c0 + 0.01 * np.random.default_rng(1).normal(size=(200, 3))])
    W, keep = tw.Winding(ctx, V, F).eval(Q)
    Wd = oracle.winding_direct(V, F, Q, threads=8)
    assert np.abs(W - Wd).max() < 1e-10
    assert np.array_equal(keep, (Wd > 0.5).astype(np.uint8))
    assert W.max() > 1.5  # overlapping spheres: winding number 2 and more


def test_full_size_config4_surface(ctx, oracle):
    """BASELINE config 4 surface (1.0 M triangles) with 2 M queries: subsample vs the oracle hierarchy + properties."""
    V, F = synth.uv_sphere(708, 708)
    assert abs(len(F) - 1_000_000) < 2000
    Wt = tw.Winding(ctx, V, F)
    Q = synth.winding_queries(V, 2_000_000, seed=11)
    W, keep = Wt.eval(Q)
    assert np.abs(W - np.round(W)).max() < 1e-9            # closed surface: integer winding numbers
    r = np.linalg.norm(Q, axis=1)
    rmin, rmax = np.linalg.norm(V, axis=1).min(), np.linalg.norm(V, axis=1).max()
    assert keep[r < rmin * 0.999].all() and not keep[r > rmax * 1.001].any()
    idx = np.random.default_rng(0).choice(len(Q), 3000, replace=False)
    OT = oracle.WindingTree(V, F)
    Wo = OT.eval(Q[idx], threads=8)
    assert np.abs(W[idx] - Wo).max() < 1e-9 and np.array_equal(keep[idx], (Wo > 0.5).astype(np.uint8))


@pytest.mark.parametrize("case", ["sphere", "knot_open", "two_spheres", "soup_dups", "tiny", "big"])
def test_device_build_is_bit_identical_to_host_build(case):
    """csrc/winding_build.cu (the whole hierarchy built on the device: vertex merge, kd order, exterior edges, cap polylines) against
    csrc/winding.cu::build_host_tree: the three arrays the kernel reads -- nodes, cap polylines, facets -- byte for byte. Closed,
    open, self-intersecting, duplicated / non-manifold and degenerate inputs, one block to 16 k blocks."""
    rng = np.random.default_rng(4)
    if case == "sphere":
        V, F = synth.uv_sphere(120, 90)
    elif case == "knot_open":
        V, F = synth.torus_knot(300, 40)
        F = F[rng.random(len(F)) < 0.8]                     # holes: open boundary chains
    elif case == "two_spheres":
        V1, F1 = synth.uv_sphere(60, 60)
        V = np.concatenate([V1, V1 * 0.7 + 0.2])
        F = np.concatenate([F1, F1 + len(V1)])
    elif case == "soup_dups":
        V1, F1 = synth.uv_sphere(40, 40, noise=0.0)
        Vs = V1[F1.astype(np.int64)].reshape(-1, 3)         # unmerged duplicate vertices: the build merges them
        Fs = np.arange(len(Vs), dtype=np.uint32).reshape(-1, 3)
        V = np.concatenate([Vs, -Vs[:300]])
        F = np.concatenate([Fs, Fs[:500], Fs[:500], np.array([[0, 0, 1], [5, 5, 5]], dtype=np.uint32), (np.arange(300, dtype=np.uint32) + len(Vs)).reshape(-1, 3)])
    elif case == "tiny":
        V, F = synth.icosphere(0)
    else:
        V, F = synth.uv_sphere(500, 500)
    got = []
    for dev in (1, 0):
        c = tw.Context(0)
        c.set_option("winding_device_build", dev)
        W = tw.Winding(c, V, F)
        got.append((W.stats(), W.download()))
        W.close()
        c.close()
    (sd, (nd, cd, td)), (sh, (nh, ch, th)) = got
    assert sd == sh, (sd, sh)
    assert np.array_equal(td, th), "facet array differs"
    bad = np.where((nd != nh).any(1))[0]
    assert len(bad) == 0, "nodes differ: first at %s of %d" % (bad[:5], len(nd))
    assert np.array_equal(cd, ch), "cap polylines differ"
