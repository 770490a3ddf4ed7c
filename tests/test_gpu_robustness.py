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
    is within eps (the boxes of tilted facets reach inwards) while no FACET is, so the traversal admits the whole tree and
    the 64-entry per-lane stack (envelope.cu kEnvStack) overflows. Overflowing queries are re-decided by the exact binary
    descent: decisions must equal brute force and the fallback must actually have run (round 1 dropped subtrees silently)."""
    c = tw.Context(0)
    c.set_option("env_front", front)
    assert c.get_option("env_front") == front
    V, F = synth.uv_sphere(160, 160, noise=0.0)
    S, OS = tw.Surface(c, V, F), oracle.Surface(V, F)
    rng = np.random.default_rng(3)
    r = np.linalg.norm(V, axis=1).max()
    P = rng.normal(size=(6000, 3)) * (2e-4 * r)
    d2 = OS.sqdist_brute(P, threads=oracle.max_threads())[0]
    n0 = c.debug_counter(0)
    for eps2 in (0.9995 * d2.min(), float(np.median(d2)), 1.0005 * d2.max()):
        got = S.points_out(P, eps2)
        assert np.array_equal(got, (d2 > eps2).astype(np.uint8)), "eps2=%g: %d decisions differ" % (eps2, int((got != (d2 > eps2)).sum()))
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
    a.set_option("env_group", 128)
    a.set_option("sort_bits", 99)          # clamped
    assert a.get_option("env_group") == 128 and b.get_option("env_group") == 64
    assert a.get_option("sort_bits") == 30
    with pytest.raises(tw.TetWildGPUError):
        a.set_option("no_such_option", 1)
    a.close()
    b.close()
