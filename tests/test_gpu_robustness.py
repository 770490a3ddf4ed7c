"""GPU tier: the failure modes VERDICT r01 / ADVICE r01 named -- traversal-stack overflow, asynchronous calls on different
caller streams sharing scratch, out-of-range indices reaching a kernel, per-context options."""
import numpy as np
import pytest

from tetwild_b200 import synth
import tetwild_b200 as tw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("front", [32, 64])
def test_envelope_stack_overflow_is_exact(oracle, front):
    """Queries near the centre of a finely tessellated sphere with eps a hair below / above their distance to it: every leaf BOX
    is within eps (the boxes of tilted facets reach inwards) while no FACET is, so the traversal admits the whole tree.
    179 400 facets give an 18-level heap; with a 64-node group frontier (level 6, all admitted) the first 8-wide step below it
    already exceeds the 64-entry per-lane stack (envelope.cu kEnvStack). Overflowing queries are re-decided by the exact
    binary descent: decisions must equal brute force and -- for the 64-node frontier -- the fallback must actually have run
    (round 1 dropped subtrees silently). With the default 32-node frontier this heap cannot overflow (32 + 7 x 4 < 64)."""
    c = tw.Context(0)
    c.set_option("env_front", front)
    assert c.get_option("env_front") == front
    V, F = synth.uv_sphere(300, 300, noise=0.0)
    S, OS = tw.Surface(c, V, F), oracle.Surface(V, F)
    rng = np.random.default_rng(3)
    r = np.linalg.norm(V, axis=1).max()
    P = rng.normal(size=(4300, 3)) * (2e-4 * r)
    d2 = OS.sqdist_brute(P, threads=oracle.max_threads())[0]
    n0 = c.debug_counter(0)
    for eps2 in (0.9995 * d2.min(), float(np.median(d2)), 1.0005 * d2.max()):
        got = S.points_out(P, eps2)
        assert np.array_equal(got, (d2 > eps2).astype(np.uint8)), "eps2=%g: %d decisions differ" % (eps2, int((got != (d2 > eps2)).sum()))
    if front == 64:
        assert c.debug_counter(0) - n0 > 0, "the overflow path was not exercised: make the test harder"
    # large eps on the config-2 surface (VERDICT r01 task 6): every decision equal to the oracle's
    V, F = synth.torus_knot(1000, 100)
    S2, OS2 = tw.Surface(c, V, F), oracle.Surface(V, F)
    lo, hi = V.min(0), V.max(0)
    P = 0.5 * (lo + hi) + 1.6 * (hi - lo) * (rng.random((200000, 3)) - 0.5)
    for eps in (0.2, 0.3):
        got = S2.points_out(P, eps * eps)
        ref = OS2.points_out(P, eps * eps, threads=oracle.max_threads())
        assert np.array_equal(got, ref) and 0.02 < got.mean() < 0.98
    S.close()
    S2.close()
    c.close()


def test_dev_calls_on_two_streams_do_not_share_state(ctx, oracle):
    """_dev calls on different caller streams used one sort scratch and one work counter. Each stream
    now owns a lane; two batches of different size issued back to back on two streams must both be complete and exact."""
    import torch
    V, F = synth.torus_knot(300, 60)
    S, OS = tw.Surface(ctx, V, F), oracle.Surface(V, F)
    sd, eps, eps2 = synth.state_eps(2e-3)
    dev = torch.device("cuda", 0)
    Pa = synth.envelope_points(V, F, 600000, eps, seed=1)
    Pb = synth.envelope_points(V, F, 150000, eps, seed=2)
    ra, rb = OS.points_out(Pa, eps2, threads=8), OS.points_out(Pb, eps2, threads=8)
    dPa, dPb = torch.from_numpy(Pa).to(dev), torch.from_numpy(Pb).to(dev)
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for rep in range(6):
        oa = torch.full((len(Pa),), 7, device=dev, dtype=torch.uint8)
        ob = torch.full((len(Pb),), 7, device=dev, dtype=torch.uint8)
        torch.cuda.synchronize()
        S.points_out_dev(dPa.data_ptr(), len(Pa), eps2, oa.data_ptr(), sa.cuda_stream)
        S.points_out_dev(dPb.data_ptr(), len(Pb), eps2, ob.data_ptr(), sb.cuda_stream)
        S.points_out_dev(dPb.data_ptr(), len(Pb), eps2, ob.data_ptr(), sb.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(oa.cpu().numpy(), ra), "rep %d: stream A" % rep
        assert np.array_equal(ob.cpu().numpy(), rb), "rep %d: stream B" % rep
    # many distinct caller streams: lanes are recycled, results stay exact
    streams = [torch.cuda.Stream(device=dev) for _ in range(20)]
    outs = [torch.full((len(Pb),), 7, device=dev, dtype=torch.uint8) for _ in streams]
    for st, o in zip(streams, outs):
        S.points_out_dev(dPb.data_ptr(), len(Pb), eps2, o.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), rb)
    S.close()


def test_out_of_range_indices_are_refused_not_dereferenced(ctx):
    V, T = synth.grid_tet_mesh(4, 4, 4)
    bad = T.copy()
    bad[5, 2] = len(V) + 1000
    with pytest.raises(tw.TetWildGPUError):
        ctx.amips_quality(V, bad)
    q = ctx.amips_quality(V, T)          # the context survives (no sticky CUDA error)
    assert np.isfinite(q).all()
    off = np.array([0, 4, 8], dtype=np.uint64)
    cen = np.array([int(T[0, 0]), int(T[4, 0])], dtype=np.int32)
    with pytest.raises(tw.TetWildGPUError):
        ctx.amips_ring_ejh(V, bad, off, cen)
    with pytest.raises(tw.TetWildGPUError):
        ctx.amips_ring_ejh(V, T, off, cen, t_ids=np.array([0, 1, 2, 3, 4, 5, 6, len(T) + 5], dtype=np.int32))
    with pytest.raises(tw.TetWildGPUError):
        ctx.amips_ring_ejh(V, T, np.array([0, 9, 4], dtype=np.uint64), cen)
    E = ctx.amips_ring_ejh(V, T, off, cen)[0]
    assert np.isfinite(E).all()
    Fbad = np.array([[0, 1, 2], [0, 1, 99]], dtype=np.uint32)
    with pytest.raises(tw.TetWildGPUError):
        tw.Surface(ctx, np.eye(3), Fbad)
    # removed slots (negative first index) named by a ring of the resident mesh are skipped, not read
    M = tw.TetMesh(ctx, V, T)
    M.set_tets([3], [[-1, -1, -1, -1]])
    E2 = M.ring_ejh(np.arange(8, dtype=np.int32), off, cen)[0]
    assert np.isfinite(E2).all()
    M.close()
    assert np.isfinite(ctx.amips_quality(V, T)).all()


def test_options_are_per_context():
    a, b = tw.Context(0), tw.Context(0)
    a.set_option("env_group", 32)
    a.set_option("sort_bits", 99)          # clamped
    assert a.get_option("env_group") == 32 and b.get_option("env_group") == 128
    assert a.get_option("sort_bits") == 30
    with pytest.raises(tw.TetWildGPUError):
        a.set_option("no_such_option", 1)
    a.close()
    b.close()


@pytest.mark.parametrize("n,bits", [(1, 24), (31, 24), (4096, 24), (4097, 24), (100_003, 24), (1_500_000, 24), (250_000, 30), (250_000, 13), (250_000, 8)])
def test_own_radix_sort_is_a_stable_sort(n, bits):
    """csrc/qsort.cu (no library sort on the query path): the result is a permutation, the top `sort_bits` of the Morton keys are
    non-decreasing, equal keys keep the caller's order (stable => deterministic), and the gathering form returns exactly
    P[perm]. Sizes around the 4096-pair tile, odd sizes, 1..4 passes."""
    c = tw.Context(0)
    c.set_option("sort_bits", bits)
    rng = np.random.default_rng(n + bits)
    P = rng.random((n, 3))
    if n > 1000:
        P[: n // 3] = P[rng.integers(0, 50, size=n // 3)]      # heavy duplicates: equal keys
        P[n // 2] = [np.nan, 0.5, 0.5]                            # non-finite coordinates are clamped, not propagated
        P[n // 2 + 1] = [5.0, -3.0, 0.5]                          # outside the box: clamped to its faces
    box = np.array([0, 0, 0, 1, 1, 1.0])
    perm, keys, srt = c.debug_sort_points(P, box)
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    top = keys >> (30 - bits)
    assert (np.diff(top.astype(np.int64)) >= 0).all()
    same = np.diff(top.astype(np.int64)) == 0
    assert (np.diff(perm.astype(np.int64))[same] > 0).all(), "equal keys must keep the caller's order"
    assert np.array_equal(srt, P[perm], equal_nan=True)
    # the keys are the 30-bit Morton codes of the clamped, quantised points
    u = np.clip(np.nan_to_num(P, nan=0.0), 0, 1)
    qd = (u * 1023.0).astype(np.uint32)

    def spread(v):
        v = v & 0x3ff
        v = (v | (v << 16)) & 0x030000ff
        v = (v | (v << 8)) & 0x0300f00f
        v = (v | (v << 4)) & 0x030c30c3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(qd[:, 0]) | (spread(qd[:, 1]) << 1) | (spread(qd[:, 2]) << 2)
    assert np.array_equal(keys, code[perm])
    perm2, _, _ = c.debug_sort_points(P, box, want_keys=False)
    assert np.array_equal(perm, perm2)
    # without a box the batch's own bounding box is reduced on the device first
    perm3, keys3, _ = c.debug_sort_points(P[np.isfinite(P).all(1)], None, want_sorted=False)
    assert (np.diff((keys3 >> (30 - bits)).astype(np.int64)) >= 0).all()
    c.close()


def test_facet_orders_give_the_same_answers(oracle):
    """Option surface_order picks how the facets are laid along the leaves of the implicit heap -- 0 Z curve, 1 Hilbert curve,
    2 kd (median splits along the longest axis, aligned with the heap's node ranges; default), 3 kd with the axis chosen by child
    surface area. The order changes the boxes, never
    an answer: envelope decisions, face decisions and nearest distances are identical, facet ids stay the caller's."""
    import tetwild_b200 as tw
    V, F = synth.torus_knot(120, 24)
    V = synth.normalise_unit_diag(V)
    sd, eps, eps2 = synth.state_eps(1e-3)
    P = synth.envelope_points(V, F, 60000, eps, seed=12)
    T = synth.face_queries(V, F, 400, 0.004, eps, seed=13)
    res = []
    for order in (0, 1, 2, 3):
        c = tw.Context(0)
        c.set_option("surface_order", order)
        S = tw.Surface(c, V, F)
        out = S.points_out(P, eps2)
        fo = S.faces_out(T, sd, eps2)
        f, q, d = S.nearest(P[:20000])
        few = S.points_out(P[:50], eps2), S.nearest(P[:50])[2]        # the tiny-call kernels walk the same structure
        res.append((out, fo, d, f, few))
        S.close()
        c.close()
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2])
        assert np.array_equal(r[4][0], res[0][0][:50]) and np.array_equal(r[4][1], res[0][2][:50])
        assert (r[3] < len(F)).all()
    ref = oracle.Surface(V, F).points_out(P[:5000], eps2)
    assert np.array_equal(res[2][0][:5000], ref)
    assert 0.05 < res[2][0].mean() < 0.95 and 0 < res[2][1].sum() < len(T)
