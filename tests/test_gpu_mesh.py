"""GPU tier (pytest -m gpu): the resident tet mesh (twg_mesh_*) through the C ABI against the oracle.
The mesh mirrors the scheduler's tet_vertices[].posf / tets / t_is_removed / conn_tets (LocalOperations.h:35-45)."""
import numpy as np
import pytest

import tetwild_b200 as tw
from tetwild_b200 import synth

pytestmark = pytest.mark.gpu


def host_rings(nV, tets):
    """conn_tets restated with numpy: vertex -> incident live tets in ascending tet id (CSR)"""
    live = np.nonzero(tets[:, 0] >= 0)[0]
    v = tets[live].ravel()
    t = np.repeat(live, 4)
    o = np.lexsort((t, v))
    off = np.zeros(nV + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.bincount(v, minlength=nV))
    return off, t[o].astype(np.int32)


def ring_close(got, ref, off):
    E, J, H, ok = got
    Er, Jr, Hr, okr = ref
    assert np.array_equal(ok, okr)
    k = np.maximum(np.diff(off.astype(np.int64)), 1)
    assert (np.abs(E - Er) <= 1e-9 * np.abs(Er)).all()
    sj = np.maximum(np.abs(Jr).max(1), np.abs(Er) / k)
    assert (np.abs(J - Jr).max(1) <= 1e-8 * np.maximum(sj, 1e-300)).all()
    assert (np.abs(H - Hr).max(1) <= 1e-9 * np.abs(Hr).max(1)).all()


@pytest.fixture(scope="module")
def grid():
    return synth.grid_tet_mesh(14, 12, 10, seed=3)


def test_quality_and_dihedral_whole_mesh(ctx, oracle, grid):
    V, T = grid
    T = T.copy()
    T[::7] = T[::7][:, [0, 2, 1, 3]]      # inverted tets -> MAX_ENERGY
    T[5::31, 0] = -1                       # removed slots (t_is_removed)
    M = tw.TetMesh(ctx, V, T)
    assert M.num_vertices == len(V) and M.num_tets == len(T)
    q = M.quality()
    Tl = T.copy()
    Tl[T[:, 0] < 0] = 0                    # the oracle has no removed marker: degenerate tet 0,0,0,0 -> MAX_ENERGY too
    qr = oracle.amips_quality(V, Tl, threads=4)
    assert np.array_equal(q == tw.MAX_ENERGY, qr == oracle.MAX_ENERGY)
    m = qr != oracle.MAX_ENERGY
    assert m.sum() > 0.8 * len(T) and (np.abs(q[m] - qr[m]) <= 1e-9 * qr[m]).all()
    # subset by ids, in the caller's order
    ids = np.random.default_rng(0).choice(len(T), 1000, replace=False).astype(np.int32)
    assert np.array_equal(M.quality(ids), q[ids])
    # dihedral angles: same operations as the reference up to acos (libm vs CUDA: a few ulp)
    lo, hi = M.dihedral()
    lor, hir = oracle.tet_dihedral(V, T, threads=4)
    assert np.abs(lo - lor).max() < 1e-13 and np.abs(hi - hir).max() < 1e-13
    assert (lo[T[:, 0] < 0] == 0).all() and (hi[T[:, 0] < 0] == np.pi).all()
    lo2, hi2 = M.dihedral(ids)
    assert np.array_equal(lo2, lo[ids]) and np.array_equal(hi2, hi[ids])
    M.close()


def test_dihedral_known_answers(ctx):
    R = np.array([[0, 0, 0], [1, 0, 0], [.5, 3 ** .5 / 2, 0], [.5, 3 ** .5 / 6, 6 ** .5 / 3]])
    Cn = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]])
    flat = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.]])           # vertex on the opposite plane -> (0, pi)
    dup = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 1.]])            # degenerate plane -> (0, pi)
    V = np.concatenate([R, Cn, flat, dup])
    M = tw.TetMesh(ctx, V, np.arange(16, dtype=np.int32).reshape(4, 4))
    lo, hi = M.dihedral()
    assert abs(lo[0] - np.arccos(1 / 3)) < 1e-14 and abs(hi[0] - np.arccos(1 / 3)) < 1e-14
    assert abs(lo[1] - np.arccos(1 / 3 ** .5)) < 1e-14 and abs(hi[1] - np.pi / 2) < 1e-14
    assert lo[2] == 0 and hi[2] == np.pi and lo[3] == 0 and hi[3] == np.pi


def test_rings_match_conn_tets(ctx, grid):
    V, T = grid
    T = T.copy()
    T[3::17, 0] = -1
    M = tw.TetMesh(ctx, V, T)
    off, t = M.get_rings()
    offr, tr = host_rings(len(V), T)
    assert np.array_equal(off, offr) and np.array_equal(t, tr)


def test_vertex_ring_newton_terms(ctx, oracle, grid):
    V, T = grid
    M = tw.TetMesh(ctx, V, T)
    off, t = host_rings(len(V), T)
    v_ids = np.random.default_rng(1).permutation(len(V)).astype(np.int32)[:1500]
    sel_off = np.zeros(len(v_ids) + 1, dtype=np.uint64)
    sel_off[1:] = np.cumsum(off[v_ids + 1] - off[v_ids])
    sel_t = np.concatenate([t[int(off[v]):int(off[v + 1])] for v in v_ids]).astype(np.int32)
    ref = oracle.amips_ring_ejh(V, T, sel_off, v_ids, t_ids=sel_t, threads=4)
    ring_close(M.vertex_ring_ejh(v_ids), ref, sel_off)
    # the explicit-member form on the resident mesh, and getNewEnergy
    ring_close(M.ring_ejh(sel_t, sel_off, v_ids), ref, sel_off)
    En = M.ring_energy(sel_t, sel_off)
    Enr = oracle.amips_ring_energy(V, T, sel_off, t_ids=sel_t, threads=4)
    assert (np.abs(En - Enr) <= 1e-9 * Enr).all()
    # must agree bit for bit with the ship-everything entry point (same kernel, same order)
    E0, J0, H0, ok0 = ctx.amips_ring_ejh(V, T, sel_off, v_ids, t_ids=sel_t)
    E1, J1, H1, ok1 = M.ring_ejh(sel_t, sel_off, v_ids)
    assert np.array_equal(E0, E1) and np.array_equal(J0, J1) and np.array_equal(H0, H1)


def test_updates_follow_the_scheduler(ctx, oracle, grid):
    """accepted operations: vertices move (smoothing), tets are removed and appended (split / collapse / swap)"""
    V, T = (a.copy() for a in grid)
    M = tw.TetMesh(ctx, V, T)
    rng = np.random.default_rng(5)
    # 1. smoothing moves 200 vertices
    mv = rng.choice(len(V), 200, replace=False).astype(np.int32)
    V[mv] += rng.normal(0, 0.005, (200, 3))
    M.set_vertices(mv, V[mv])
    assert np.array_equal(M.get_vertices(), V)
    # 2. an edge split: one new vertex (centroid of tet 10), tet 10 replaced by 4 tets -> 3 appended slots
    old = T[10].copy()
    c = V[old].mean(0)
    nv = len(V)
    V = np.vstack([V, c])
    new = []
    for k in range(4):
        tt = old.copy()
        tt[k] = nv
        new.append(tt)
    new = np.array(new, dtype=np.int32)
    nT0 = len(T)
    T = np.vstack([T, new[1:]])
    T[10] = new[0]
    M.resize(len(V), len(T))
    M.set_vertices([nv], c[None])
    M.set_tets(np.array([10, nT0, nT0 + 1, nT0 + 2], dtype=np.int32), new)
    # 3. a collapse removes tets 20..24
    rm = np.arange(20, 25, dtype=np.int32)
    T[rm, 0] = -1
    M.set_tets(rm, T[rm])
    off, t = M.get_rings()
    offr, tr = host_rings(len(V), T)
    assert np.array_equal(off, offr) and np.array_equal(t, tr)
    assert off[nv + 1] - off[nv] == 4
    Tl = T.copy()
    Tl[T[:, 0] < 0] = 0
    q, qr = M.quality(), oracle.amips_quality(V, Tl, threads=4)
    assert np.array_equal(q == tw.MAX_ENERGY, qr == oracle.MAX_ENERGY)
    m = qr != oracle.MAX_ENERGY
    assert (np.abs(q[m] - qr[m]) <= 1e-9 * qr[m]).all()
    v_ids = np.concatenate([[nv], mv[:50], old]).astype(np.int32)
    sel_off = np.zeros(len(v_ids) + 1, dtype=np.uint64)
    sel_off[1:] = np.cumsum(offr[v_ids + 1] - offr[v_ids])
    sel_t = np.concatenate([tr[int(offr[v]):int(offr[v + 1])] for v in v_ids]).astype(np.int32)
    ring_close(M.vertex_ring_ejh(v_ids), oracle.amips_ring_ejh(V, T, sel_off, v_ids, t_ids=sel_t), sel_off)


def test_invalid_ids_fail_loudly(ctx, grid):
    V, T = grid
    M = tw.TetMesh(ctx, V, T)
    with pytest.raises(tw.TetWildGPUError):
        M.set_vertices([len(V)], np.zeros((1, 3)))
    with pytest.raises(tw.TetWildGPUError):
        M.quality(np.array([len(T)], dtype=np.int32))
    with pytest.raises(tw.TetWildGPUError):
        M.set_tets([0], np.array([[0, 1, 2, len(V)]], dtype=np.int32))
    with pytest.raises(tw.TetWildGPUError):
        M.resize(len(V) - 1, len(T))
    with pytest.raises(tw.TetWildGPUError):
        M.vertex_ring_ejh(np.array([-1], dtype=np.int32))


def test_out_of_range_indices_are_refused(ctx, grid):
    """host entry points validate indices before anything reaches the device"""
    V, T = grid
    bad = T.copy(); bad[7, 2] = len(V)
    with pytest.raises(tw.TetWildGPUError):
        tw.TetMesh(ctx, V, bad)
    Vs, Fs = synth.uv_sphere(12, 12)
    Fb = Fs.copy(); Fb[3, 1] = len(Vs)
    with pytest.raises(tw.TetWildGPUError):
        tw.Surface(ctx, Vs, Fb)
    with pytest.raises(tw.TetWildGPUError):
        tw.Winding(ctx, Vs, Fb)


def test_empty_mesh(ctx):
    M = tw.TetMesh(ctx, np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int32))
    assert M.num_tets == 0 and len(M.quality()) == 0
    off, t = M.get_rings()
    assert len(off) == 1 and off[0] == 0 and len(t) == 0
