"""GPU tier: one vertex-smoothing pass of the scheduler's shape, driven through the resident tet mesh.

VertexSmoother::smooth (src/tetwild/VertexSmoother.cpp:19-112) visits vertices one by one: NewtonsUpdate over the one-ring
(:627-702), a Newton step with back-tracking (NewtonsMethod :464-541: halve the step while a tet flips or the ring energy
does not decrease, getNewEnergy :544-625), then the accepted position is written back. Vertices that share no tet do not
interact, so the same pass runs as a sequence of INDEPENDENT SETS, each one batched call per stage -- the integration
INTEGRATION.md describes. The test runs the pass twice with identical host logic, once against the device (twg_mesh_*) and
once against the CPU oracle, and compares the trajectories."""
import numpy as np
import pytest

import tetwild_b200 as tw
from tetwild_b200 import synth

pytestmark = pytest.mark.gpu

MAX_IT = 20


def rings_of(nV, T):
    v = T.ravel()
    t = np.repeat(np.arange(len(T)), 4)
    o = np.lexsort((t, v))
    off = np.zeros(nV + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.bincount(v, minlength=nV))
    return off, t[o].astype(np.int32)


def colour_classes(nV, T, movable):
    """greedy colouring: two vertices of one tet never share a colour"""
    off, adj = rings_of(nV, T)
    colour = np.full(nV, -1)
    for v in np.nonzero(movable)[0]:
        used = set(colour[T[adj[off[v]:off[v + 1]]].ravel()])
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    return [np.nonzero(colour == c)[0].astype(np.int32) for c in range(colour.max() + 1)], off, adj


class DeviceBackend:
    def __init__(self, ctx, V, T):
        self.M = tw.TetMesh(ctx, V, T)

    def newton_terms(self, v_ids, V, T, off, adj):
        return self.M.vertex_ring_ejh(v_ids)

    def move(self, v_ids, X):
        self.M.set_vertices(v_ids, X)

    def ring_energy_and_flip(self, v_ids, V, T, off, adj):
        cnt = (off[v_ids + 1] - off[v_ids])
        goff = np.zeros(len(v_ids) + 1, dtype=np.uint64)
        goff[1:] = np.cumsum(cnt)
        tids = np.concatenate([adj[off[v]:off[v + 1]] for v in v_ids]).astype(np.int32)
        E = self.M.ring_energy(tids, goff)
        q = self.M.quality(tids)
        flipped = np.add.reduceat((q == tw.MAX_ENERGY).astype(np.int64), goff[:-1].astype(np.int64)) > 0
        return E, flipped


class OracleBackend:
    def __init__(self, oracle, V, T):
        self.O, self.V = oracle, V.copy()

    def newton_terms(self, v_ids, V, T, off, adj):
        cnt = (off[v_ids + 1] - off[v_ids])
        goff = np.zeros(len(v_ids) + 1, dtype=np.uint64)
        goff[1:] = np.cumsum(cnt)
        tids = np.concatenate([adj[off[v]:off[v + 1]] for v in v_ids]).astype(np.int32)
        return self.O.amips_ring_ejh(self.V, T, goff, v_ids, t_ids=tids, threads=4)

    def move(self, v_ids, X):
        self.V[v_ids] = X

    def ring_energy_and_flip(self, v_ids, V, T, off, adj):
        cnt = (off[v_ids + 1] - off[v_ids])
        goff = np.zeros(len(v_ids) + 1, dtype=np.uint64)
        goff[1:] = np.cumsum(cnt)
        tids = np.concatenate([adj[off[v]:off[v + 1]] for v in v_ids]).astype(np.int32)
        E = self.O.amips_ring_energy(self.V, T, goff, t_ids=tids, threads=4)
        q = self.O.amips_quality(self.V, T[tids], threads=4)
        flipped = np.add.reduceat((q == self.O.MAX_ENERGY).astype(np.int64), goff[:-1].astype(np.int64)) > 0
        return E, flipped


def smoothing_pass(B, V, T, classes, off, adj):
    """one Newton step with back-tracking per movable vertex; returns the new positions and the number of accepted moves"""
    V = V.copy()
    accepted = 0
    for S in classes:
        E0, J, H, ok = B.newton_terms(S, V, T, off, adj)
        todo = np.nonzero(ok)[0]
        a = np.ones(len(S))
        X0 = V[S].copy()
        Hm = H.reshape(-1, 3, 3)
        for _ in range(MAX_IT):
            if len(todo) == 0:
                break
            # X = H^-1 (H X0 - a J)   (VertexSmoother.cpp:489)
            rhs = np.einsum("nij,nj->ni", Hm[todo], X0[todo]) - a[todo, None] * J[todo]
            try:
                X = np.linalg.solve(Hm[todo], rhs[:, :, None])[:, :, 0]
            except np.linalg.LinAlgError:
                break
            fin = np.isfinite(X).all(1)
            X[~fin] = X0[todo][~fin]
            B.move(S[todo], X)
            E1, flipped = B.ring_energy_and_flip(S[todo], V, T, off, adj)
            good = fin & ~flipped & np.isfinite(E1) & (E1 < E0[todo])
            B.move(S[todo[~good]], X0[todo[~good]])          # rejected: restore, halve the step (:497-521)
            V[S[todo[good]]] = X[good]
            accepted += int(good.sum())
            a[todo[~good]] *= 0.5
            todo = todo[~good]
    return V, accepted


def total_energy(oracle, V, T):
    q = oracle.amips_quality(V, T, threads=4)
    assert (q != oracle.MAX_ENERGY).all(), "a tet is inverted"
    return float(q.sum())


def test_one_smoothing_pass_matches_the_cpu_path(ctx, oracle):
    V, T = synth.grid_tet_mesh(9, 8, 7, jitter=0.27, seed=21)
    nV = len(V)
    lo, hi = V.min(0), V.max(0)
    movable = ((V > lo + 0.08) & (V < hi - 0.08)).all(1)        # the hull stays put (no surface projection in this test)
    classes, off, adj = colour_classes(nV, T, movable)
    assert 8 <= len(classes) <= 40 and sum(len(c) for c in classes) == movable.sum() > 200
    e_before = total_energy(oracle, V, T)
    Vg, acc_g = smoothing_pass(DeviceBackend(ctx, V, T), V, T, classes, off, adj)
    Vo, acc_o = smoothing_pass(OracleBackend(oracle, V, T), V, T, classes, off, adj)
    e_g, e_o = total_energy(oracle, Vg, T), total_energy(oracle, Vo, T)
    assert acc_g > 0.9 * movable.sum() and abs(acc_g - acc_o) <= 2
    assert e_g < 0.97 * e_before and abs(e_g - e_o) < 1e-6 * e_o
    # same trajectory: every vertex within 1e-8 (an accept / reject decision on an exact tie could split one; none here)
    d = np.abs(Vg - Vo).max(1)
    assert (d < 1e-8).mean() > 0.999 and np.array_equal(Vg[~movable], V[~movable])
