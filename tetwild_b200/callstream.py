"""A pass-shaped call stream: the hot-path calls ONE pass of the reference's scheduler makes, in the reference's order.

BASELINE.json configs[0] / configs[4] ask for "replay its logged envelope/AMIPS calls" of a real TetWild run; no TetWild binary
can be built here (CGAL, geogram, libigl, Boost, GMP headers absent -- SURVEY.md 8c), so this module GENERATES a stream with the
call mix the scheduler's control flow produces (src/tetwild/MeshRefinement.cpp:120-183: split, collapse, swap, smooth) instead
of recording one:

  per collapse candidate (EdgeCollapser.cpp:311-329, :727-775)
      isPointOutBoundaryEnvelop(v)            1 point                      for surface vertices
      isPointOutEnvelop(v)                    1 point
      calTetQualities(new_tets)               8..30 tets (the tets around the merged vertex)
      isFaceOutEnvelop(tri) per new face      1..20 faces of edge ~ diag/50, stop at the first OUT (:770)
  per smoothing candidate (VertexSmoother.cpp:465-541, :627-702, :544-625, :354-362, :425)
      NewtonsUpdate(conn_tets[v], v)          1 one-ring (~24 tets)
      line search: move v, getNewEnergy, undo 1..4 trial positions (twg_mesh_vertex_trial_energy: the mesh is not modified)
      surface vertices: nearest_facet(p)      1 projection, then isFaceOutEnvelop on the ring's surface faces (2..8 faces)
  per split / swap candidate (EdgeSplitter.cpp:130-147, EdgeRemover)
      calTetQualities(new_tets)               2..12 tets

The same stream is replayed (a) call by call, the way the unchanged sequential scheduler would issue it through the adapters
(every call a host round trip), (b) re-batched by kind -- what the adapters' batched forms allow when candidates are
evaluated speculatively -- and (c) call by call on the CPU oracle, single-threaded, which is how the reference itself runs.
Results of (a), (b), (c) must be identical (decisions) / within 1e-9 (energies).
"""
import time

import numpy as np

from . import synth

KINDS = ("point_out", "boundary_point_out", "faces_out", "quality", "newton", "trial_energy", "nearest")


def make_pass_stream(n_collapse=3000, n_smooth=3000, n_split=2000, seed=1, eps_rel=1e-3, surface_fraction=0.35):
    """-> dict(surface=(V, F), boundary=(Vb, Fb), mesh=(Vm, Tm), sd, eps2, calls=[(kind, payload), ...])"""
    rng = np.random.default_rng(seed)
    V, F = synth.icosphere(5)
    V = synth.normalise_unit_diag(V)
    sd, eps, eps2 = synth.state_eps(eps_rel)
    # boundary mesh: edges stored as degenerate triangles (Preprocess.cpp:192-197) -- a few open-boundary loops of the surface
    loop = np.arange(0, 60)
    Vb = V[F[loop, 0]]
    Fb = np.stack([np.arange(60), (np.arange(60) + 1) % 60, (np.arange(60) + 1) % 60], 1).astype(np.uint32)
    # tet mesh of the scheduler's shape, scaled into the sphere
    Vm, Tm = synth.grid_tet_mesh(14, 14, 14)
    Vm = (Vm - 0.5) * 0.5
    nV, nT = len(Vm), len(Tm)
    # conn_tets
    order = np.argsort(Tm.ravel(), kind="stable")
    vert_sorted = Tm.ravel()[order]
    tet_of = (order // 4).astype(np.int32)
    off = np.searchsorted(vert_sorted, np.arange(nV + 1)).astype(np.int64)
    ring_sizes = np.diff(off)
    interior = np.where(ring_sizes >= 12)[0]
    tri = V[F.astype(np.int64)]

    def surf_point(k=1, sigma=1.0):
        f = rng.integers(0, len(F), size=k)
        w = rng.dirichlet([1, 1, 1], size=k)
        p = (tri[f] * w[:, :, None]).sum(1)
        nrm = np.cross(tri[f, 1] - tri[f, 0], tri[f, 2] - tri[f, 0])
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        return p + nrm * rng.normal(0, sigma * eps, size=(k, 1))

    calls = []
    ops = ["collapse"] * n_collapse + ["smooth"] * n_smooth + ["split"] * n_split
    # the scheduler runs the operations one after the other (all splits, all collapses, all swaps, all smoothing): keep that order
    for op in sorted(ops, key=lambda o: ("split", "collapse", "smooth").index(o)):
        on_surface = rng.random() < surface_fraction
        if op == "split":
            t0 = int(rng.integers(0, nT - 12))
            calls.append(("quality", np.arange(t0, t0 + int(rng.integers(2, 13)), dtype=np.int32)))
        elif op == "collapse":
            v = int(rng.choice(interior))
            if on_surface:
                p = surf_point(1, 0.7)[0]
                calls.append(("boundary_point_out", p))
                calls.append(("point_out", p))
            ring = tet_of[off[v]:off[v + 1]]
            calls.append(("quality", ring[: int(rng.integers(8, min(31, len(ring) + 1)))].astype(np.int32)))
            if on_surface:
                k = int(rng.integers(1, 21))
                calls.append(("faces_out", synth.face_queries(V, F, k, 0.02, eps, seed=int(rng.integers(1 << 30)))))
        else:
            v = int(rng.choice(interior))
            calls.append(("newton", v))
            for _ in range(int(rng.integers(1, 5))):
                calls.append(("trial_energy", (v, Vm[v] + rng.normal(0, 0.004, size=3))))
            if on_surface:
                calls.append(("nearest", surf_point(1, 2.0)[0]))
                k = int(rng.integers(2, 9))
                calls.append(("faces_out", synth.face_queries(V, F, k, 0.02, eps, seed=int(rng.integers(1 << 30)))))
    return {"surface": (V, F), "boundary": (Vb, Fb), "mesh": (Vm, Tm), "conn": (off, tet_of), "sd": sd, "eps2": eps2, "calls": calls}


def mix(stream):
    out = {k: 0 for k in KINDS}
    units = {k: 0 for k in KINDS}
    for kind, pay in stream["calls"]:
        out[kind] += 1
        units[kind] += len(pay) if kind in ("faces_out", "quality") else 1
    return out, units


class GpuReplayer:
    """the product path: C ABI through tetwild_b200.api (resident mesh, surface + boundary structures)"""

    def __init__(self, ctx, stream):
        import tetwild_b200 as tw
        self.s = stream
        self.S = tw.Surface(ctx, *stream["surface"])
        self.B = tw.Surface(ctx, *stream["boundary"])
        self.M = tw.TetMesh(ctx, *stream["mesh"])
        self.M.build_rings()
        self.off, self.tet_of = stream["conn"]

    def close(self):
        self.S.close(); self.B.close(); self.M.close()

    def one(self, kind, pay):
        s = self.s
        if kind == "point_out":
            return self.S.squared_distance(pay[None])[0] > s["eps2"]                 # isPointOutEnvelop: squared_distance() > eps_2 (:1037)
        if kind == "boundary_point_out":
            return self.B.squared_distance(pay[None])[0] > s["eps2"]
        if kind == "faces_out":
            for t in pay:                                                            # the reference stops at the first OUT face (:770)
                if self.S.faces_out(t[None], s["sd"], s["eps2"])[0]:
                    return True
            return False
        if kind == "quality":
            return self.M.quality(pay)
        if kind == "newton":
            return self.M.vertex_ring_ejh([pay])
        if kind == "trial_energy":
            v, p = pay
            return self.M.vertex_trial_energy([v], p[None])[0]
        if kind == "nearest":
            return self.S.nearest(pay[None])
        raise ValueError(kind)

    def call_by_call(self):
        t0 = time.perf_counter()
        res = [self.one(k, p) for k, p in self.s["calls"]]
        return time.perf_counter() - t0, res

    def batched(self):
        """the same calls grouped by kind: one C-ABI call per kind (every call of the stream is stateless, so the grouping is exact)"""
        s = self.s
        calls = s["calls"]
        t0 = time.perf_counter()
        res = [None] * len(calls)
        idx = {k: [i for i, (kk, _) in enumerate(calls) if kk == k] for k in KINDS}
        if idx["point_out"]:
            d = self.S.squared_distance(np.array([calls[i][1] for i in idx["point_out"]]))
            for i, x in zip(idx["point_out"], d > s["eps2"]):
                res[i] = bool(x)
        if idx["boundary_point_out"]:
            d = self.B.squared_distance(np.array([calls[i][1] for i in idx["boundary_point_out"]]))
            for i, x in zip(idx["boundary_point_out"], d > s["eps2"]):
                res[i] = bool(x)
        if idx["faces_out"]:
            T = np.concatenate([calls[i][1] for i in idx["faces_out"]])
            o = self.S.faces_out(T, s["sd"], s["eps2"])
            b = 0
            for i in idx["faces_out"]:
                k = len(calls[i][1])
                res[i] = bool(o[b:b + k].any())
                b += k
        if idx["quality"]:
            ids = np.concatenate([calls[i][1] for i in idx["quality"]])
            q = self.M.quality(ids)
            b = 0
            for i in idx["quality"]:
                k = len(calls[i][1])
                res[i] = q[b:b + k]
                b += k
        if idx["newton"]:
            E, J, H, ok = self.M.vertex_ring_ejh(np.array([calls[i][1] for i in idx["newton"]], dtype=np.int32))
            for j, i in enumerate(idx["newton"]):
                res[i] = (E[j:j + 1], J[j:j + 1], H[j:j + 1], ok[j:j + 1])
        if idx["trial_energy"]:
            e = self.M.vertex_trial_energy(np.array([calls[i][1][0] for i in idx["trial_energy"]], dtype=np.int32),
                                           np.array([calls[i][1][1] for i in idx["trial_energy"]]))
            for i, x in zip(idx["trial_energy"], e):
                res[i] = x
        if idx["nearest"]:
            f, q, d = self.S.nearest(np.array([calls[i][1] for i in idx["nearest"]]))
            for j, i in enumerate(idx["nearest"]):
                res[i] = (f[j:j + 1], q[j:j + 1], d[j:j + 1])
        return time.perf_counter() - t0, res


class CpuReplayer:
    """the reference's way: one call at a time on one core (oracle = the checker, here also the timed CPU arm)"""

    def __init__(self, oracle, stream):
        self.O, self.s = oracle, stream
        self.S = oracle.Surface(*stream["surface"])
        self.B = oracle.Surface(*stream["boundary"])
        self.V = stream["mesh"][0].copy()
        self.T = stream["mesh"][1]
        self.off, self.tet_of = stream["conn"]

    def one(self, kind, pay):
        O, s = self.O, self.s
        if kind == "point_out":
            return self.S.nearest(pay[None])[2][0] > s["eps2"]
        if kind == "boundary_point_out":
            return self.B.nearest(pay[None])[2][0] > s["eps2"]
        if kind == "faces_out":
            for t in pay:
                if self.S.faces_out(t[None], s["sd"], s["eps2"])[0][0]:
                    return True
            return False
        if kind == "quality":
            return O.amips_quality(self.V, self.T[pay])
        if kind in ("newton", "trial_energy"):
            v = pay if kind == "newton" else pay[0]
            ring = self.tet_of[self.off[v]:self.off[v + 1]]
            goff = np.array([0, len(ring)], dtype=np.uint64)
            if kind == "newton":
                return O.amips_ring_ejh(self.V, self.T, goff, np.array([v], dtype=np.int32), t_ids=ring)
            old = self.V[v].copy()
            self.V[v] = pay[1]                                   # move v, getNewEnergy, undo: VertexSmoother.cpp:505-541
            e = O.amips_ring_energy(self.V, self.T, goff, t_ids=ring)[0]
            self.V[v] = old
            return e
        if kind == "nearest":
            return self.S.nearest(pay[None])
        raise ValueError(kind)

    def call_by_call(self, limit=None):
        calls = self.s["calls"] if limit is None else self.s["calls"][:limit]
        t0 = time.perf_counter()
        res = [self.one(k, p) for k, p in calls]
        return time.perf_counter() - t0, res


def same(kind, a, b, tol=1e-9):
    """decisions identical; energies / tensors within tol (relative to their norm); nearest: same d2"""
    if kind in ("point_out", "boundary_point_out", "faces_out"):
        return bool(a) == bool(b)
    if kind == "quality":
        a, b = np.asarray(a), np.asarray(b)
        big = (a >= 1e49) | (b >= 1e49)
        return np.array_equal(a >= 1e49, b >= 1e49) and np.all(np.abs(a[~big] - b[~big]) <= tol * np.abs(b[~big]))
    if kind == "trial_energy":
        return abs(a - b) <= tol * abs(b) or (a >= 1e49 and b >= 1e49)
    if kind == "newton":
        return all(np.abs(np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64)).max() <= tol * max(1e-300, np.abs(np.asarray(y, dtype=np.float64)).max()) for x, y in zip(a[:3], b[:3])) and int(a[3][0]) == int(b[3][0])
    if kind == "nearest":
        return float(np.asarray(a[2]).ravel()[0]) == float(np.asarray(b[2]).ravel()[0])
    raise ValueError(kind)
