"""GPU tier (pytest -m gpu): AMIPS kernels through the C ABI against the oracle.
Tolerance: 1e-9 of the tensor magnitude (conftest.amips_close); MAX_ENERGY gating decisions exact."""
import numpy as np
import pytest

from conftest import amips_close, load_golden, unhex
from tetwild_b200 import synth

pytestmark = pytest.mark.gpu


def test_ejh_soa_vs_oracle(ctx, oracle):
    T = synth.random_tets(40000, seed=7)
    E, J, H = ctx.amips_ejh_soa(T)
    ref = oracle.amips_ejh_soa(T, threads=4)
    assert max(amips_close(T, (E, J, H), ref)) < 1e-9
    assert np.allclose(ctx.amips_energy_soa(T), E, rtol=1e-14)
    if oracle.ref_available():
        assert max(amips_close(T, (E, J, H), oracle.ref_amips_ejh_soa(T, threads=4))) < 1e-9


def test_ejh_golden(ctx):
    g = load_golden("amips_golden.json")
    n = g["n"]
    T = unhex(g["T_rows_12xn"], (12, n))
    got = ctx.amips_ejh_soa(T)
    assert max(amips_close(T, got, (unhex(g["E"]), unhex(g["J"], (n, 3)), unhex(g["H"], (n, 9))))) < 1e-9
    assert abs(got[0][0] - 3.0) < 1e-13  # regular tet, README.md:141


@pytest.mark.parametrize("n", [1, 2, 3, 127, 128, 129, 255, 257, 1000, 4099])
def test_ejh_ragged_sizes(ctx, oracle, n):
    T = synth.random_tets(n, seed=100 + n)
    got = ctx.amips_ejh_soa(T)
    assert max(amips_close(T, got, oracle.amips_ejh_soa(T))) < 1e-9
    # only some outputs requested, and an unaligned (odd element offset) view of the inputs
    E2, J2, H2 = ctx.amips_ejh_soa(T, want=(False, True, False))
    assert E2 is None and H2 is None and np.array_equal(J2, got[1])
    if n > 2:
        Tu = np.ascontiguousarray(T[:, 1:])
        gu = ctx.amips_ejh_soa(Tu)
        assert np.array_equal(gu[0], got[0][1:]) and np.array_equal(gu[2], got[2][1:])


def test_empty(ctx):
    E, J, H = ctx.amips_ejh_soa(np.zeros((12, 0)))
    assert len(E) == 0 and J.shape == (0, 3)


def test_quality_gates(ctx, oracle):
    T = synth.random_tets(5000, seed=21, scale_lo=0.1, scale_hi=10)
    X = T.T.reshape(-1, 4, 3)
    V = X.reshape(-1, 3).copy()
    tets = np.arange(len(V), dtype=np.int32).reshape(-1, 4)
    tets[::3] = tets[::3][:, [0, 2, 1, 3]]          # inverted -> MAX_ENERGY
    V[tets[5::50, 3]] = V[tets[5::50, 0]]            # two coincident vertices -> orientation ZERO -> MAX_ENERGY
    w = np.array([0.25, 0.5, 0.25])
    V[tets[7::50, 3]] = (V[tets[7::50, :3]] * w[None, :, None]).sum(1)  # (nearly) coplanar: exact predicate decides
    got = ctx.amips_quality(V, tets)
    ref = oracle.amips_quality(V, tets, threads=4)
    assert np.array_equal(got == oracle.MAX_ENERGY, ref == oracle.MAX_ENERGY)
    m = ref != oracle.MAX_ENERGY
    assert m.sum() > 2000 and (~m).sum() > 1500
    # the exact predicate keeps arbitrarily flat POSITIVE tets, whose energy is huge and ill-conditioned: compare
    # the well-shaped ones at 1e-9 and require the flat ones to be huge on both sides
    good = m & (ref < 1e6)
    assert np.abs(got[good] - ref[good]).max() <= 1e-9 * np.abs(ref[good]).max() and (np.abs(got[good] / ref[good] - 1) < 1e-9).all()
    assert (got[m & ~good] > 1e5).all()


def test_ring_ejh(ctx, oracle):
    V, tets, off, center = synth.ring_groups(3000, seed=5, scale_lo=0.1, scale_hi=10)
    E, J, H, ok = ctx.amips_ring_ejh(V, tets, off, center)
    Er, Jr, Hr, okr = oracle.amips_ring_ejh(V, tets, off, center, threads=4)
    assert np.array_equal(ok, okr) and ok.all()
    assert (np.abs(E - Er) / np.abs(Er)).max() < 1e-9
    k = np.diff(off.astype(np.int64))
    # sums of k per-tet tensors: compare against the magnitude of the summed tensor or of one term, whichever is larger
    sj = np.maximum(np.abs(Jr).max(1), np.abs(Er) / k * 1.0)
    assert (np.abs(J - Jr).max(1) / np.maximum(sj, 1e-300)).max() < 1e-8
    assert (np.abs(H - Hr).max(1) / np.abs(Hr).max(1)).max() < 1e-9
    # through a t_ids indirection (the reference passes conn_tets ids), shuffled tet storage
    perm = np.random.default_rng(0).permutation(len(tets))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    E2, J2, H2, ok2 = ctx.amips_ring_ejh(V, tets[perm], off, center, t_ids=inv.astype(np.int32))
    assert np.array_equal(E2, E) and np.array_equal(J2, J) and np.array_equal(H2, H)
    En = ctx.amips_ring_energy(V, tets, off)
    assert (np.abs(En - oracle.amips_ring_energy(V, tets, off)) / En).max() < 1e-9


def test_ring_rejects(ctx, oracle):
    """NewtonsUpdate returns false on NaN energy / non-finite J,H (VertexSmoother.cpp:684-699); getNewEnergy clamps."""
    V, tets, off, center = synth.ring_groups(64, seed=9)
    V = V.copy()
    V[tets[int(off[3]) + 1, (list(tets[int(off[3]) + 1]).index(center[3]) + 1) % 4], 1] = np.nan   # poison one ring vertex of group 3
    big = tets[int(off[7]), (list(tets[int(off[7])]).index(center[7]) + 2) % 4]
    V[big] = 1e200                                                                                   # overflow in group 7
    E, J, H, ok = ctx.amips_ring_ejh(V, tets, off, center)
    Er, Jr, Hr, okr = oracle.amips_ring_ejh(V, tets, off, center)
    assert np.array_equal(ok, okr) and ok[3] == 0 and ok[7] == 0 and ok.sum() == 62
    En, Enr = ctx.amips_ring_energy(V, tets, off), oracle.amips_ring_energy(V, tets, off)
    assert En[3] == oracle.MAX_ENERGY and Enr[3] == oracle.MAX_ENERGY and En[7] == Enr[7] == oracle.MAX_ENERGY


def test_large_batch_properties(ctx, oracle):
    """BASELINE config-3 shape at 4 M tets: subsample against the oracle + size-independent properties."""
    n = 4_000_000
    T = synth.random_tets(n, seed=77)
    E, J, H = ctx.amips_ejh_soa(T)
    assert np.isfinite(E).all() and (E >= 3.0 - 1e-9).all()          # E >= 3 with equality for the regular tet
    assert np.allclose(H[:, 1], H[:, 3]) and np.allclose(H[:, 2], H[:, 6]) and np.allclose(H[:, 5], H[:, 7])
    idx = np.random.default_rng(1).choice(n, 20000, replace=False)
    Ts = np.ascontiguousarray(T[:, idx])
    assert max(amips_close(Ts, (E[idx], J[idx], H[idx]), oracle.amips_ejh_soa(Ts, threads=4))) < 1e-9
    # relabelling vertices 1..3 cyclically leaves E, J, H (w.r.t. vertex 0) unchanged
    T2 = np.concatenate([T[0:3], T[6:9], T[9:12], T[3:6]])
    E2, J2, H2 = ctx.amips_ejh_soa(T2)
    assert max(amips_close(T, (E2, J2, H2), (E, J, H))) < 1e-9


def test_ring_batches_do_not_depend_on_their_composition():
    """A ring's sums depend on the ring alone: the same rings as one batch, as a sub-batch, through the t_ids indirection and
    through the tiny-call path give the same bits. Rings of 0 .. 70 tets (empty rings, rings of more than 32 tets) and a
    rejected ring."""
    V, tets, off, center = synth.ring_groups(5000, seed=21, kmin=3, kmax=70, scale_lo=0.1, scale_hi=10)
    V = V.copy()
    V[tets[int(off[11]) + 2, (list(tets[int(off[11]) + 2]).index(center[11]) + 1) % 4], 2] = np.nan
    cnt = np.diff(off.astype(np.int64))
    keep = np.ones(len(cnt), bool)
    dropped = [100, 2000, 2001, 2002, len(cnt) - 1]
    keep[dropped] = False
    sel = np.repeat(keep, cnt)
    tets = np.ascontiguousarray(tets[sel])
    off = np.concatenate([[0], np.cumsum(np.where(keep, cnt, 0))]).astype(np.uint64)
    perm = np.random.default_rng(2).permutation(len(tets))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    import tetwild_b200 as tw
    c = tw.Context(0)
    a = c.amips_ring_ejh(V, tets, off, center)
    b = c.amips_ring_ejh(V, tets[perm], off, center, t_ids=inv.astype(np.int32))
    sub = c.amips_ring_ejh(V, tets, off[:301], center[:300])
    e = c.amips_ring_energy(V, tets, off)
    c.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    for x, y in zip(a, sub):
        assert np.array_equal(x[:300], y, equal_nan=True)
    okm = a[3]
    assert okm[11] == 0 and okm.sum() == len(center) - 6 and not okm[dropped].any()
    assert (e[dropped] == tw.MAX_ENERGY).all() and (a[0][dropped] == 0).all()
