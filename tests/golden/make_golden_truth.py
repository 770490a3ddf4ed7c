"""tests/golden/make_golden_truth.py -- higher-precision / exact ground truths (round 2; VERDICT r01 "weak" #1, #2).

Independent evidence for the arithmetic that cannot be pinned bit for bit to a third-party source absent from
/root/reference, and for the AMIPS 1e-9 question on the literal config 3:

  amips_truth_golden.json   300 tets of the LITERAL C3 (translation U(-10,10)^3 NOT scaled): inputs, the
                            reference's own text (LocalOperations.cpp:28-291) evaluated in IEEE binary128
                            (oracle/ref_quad.cpp, rounded to double), the mathematical AMIPS in binary128, and the
                            reference's double results; + error statistics over the 50 k sample they were drawn from
  trisq_exact_golden.json   2 000 point-triangle squared distances in EXACT rational arithmetic (fractions.Fraction on the
                            double inputs), rounded to the nearest double: bounds geogram's
                            point_triangle_squared_distance restatement (oracle/envelope.c, csrc/tw_math.cuh)
  dihedral_mp_golden.json   min / max dihedral angles of 600 tets from the mathematical definition at 40 digits: bounds the
                            restatement of CGAL's plane / projection constructions behind calTetQuality_AD (oracle/amips.c)
  winding_mp_golden.json    generalized winding numbers of ~2 000 (surface, query) pairs as 40-digit mpmath sums of
                            Van Oosterom-Strackee solid angles: bounds the libigl restatement (oracle/winding.c) and the
                            device's complex-product accumulation (csrc/winding.cu)

usage:  python tests/golden/make_golden_truth.py     (needs oracle/_ref, i.e. /root/reference, for the first file only)
"""
import json
import os
import sys
from fractions import Fraction as Fr

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tetwild_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def normwise(a, b):
    n = len(b[0])
    return [np.abs(a[k].reshape(n, -1) - b[k].reshape(n, -1)).max(1) / np.abs(b[k].reshape(n, -1)).max(1) for k in range(3)]


# ------------------------------------------------------------------------------------------------- AMIPS truth
def amips_truth():
    assert O.ref_available() and O.quad_available(), "oracle/_ref not built (oracle/ref_build.sh needs /root/reference)"
    T = synth.random_tets(50000, seed=7, trans_scales=False)
    R = O.ref_amips_ejh_soa(T, threads=8)
    Q = O.refq_amips_ejh_soa(T, threads=8)
    X = O.exactq_amips_ejh_soa(T, threads=8)
    eR = normwise(R, Q)
    worst = np.argsort(-np.maximum.reduce(eR))[:100]
    rest = np.setdiff1d(np.arange(T.shape[1]), worst)
    pick = np.concatenate([worst, np.random.default_rng(1).choice(rest, 200, replace=False)])
    eQX = normwise(Q, X)
    stats = {"sample": int(T.shape[1]), "generator": "synth.random_tets(50000, seed=7, trans_scales=False)",
             "reference_double_vs_its_binary128_value": {"max_E_J_H": [float(e.max()) for e in eR], "p99_E_J_H": [float(np.quantile(e, .99)) for e in eR],
                                                        "tets_over_1e-9": int(((eR[0] > 1e-9) | (eR[1] > 1e-9) | (eR[2] > 1e-9)).sum())},
             "binary128_text_vs_binary128_exact_function": {"max_E_J_H": [float(e.max()) for e in eQX],
                                                           "note": "the text's 15-digit literals (0.577350269189626, 1.15470053837925, ...) make the written expression differ from the mathematical AMIPS"}}
    Ts = np.ascontiguousarray(T[:, pick])
    json.dump({"source": "reference LocalOperations.cpp:28-291 compiled in IEEE binary128 (oracle/ref_quad.cpp), rounded to double",
               "n": int(len(pick)), "T_rows_12xn": hx(Ts),
               "Eq": hx(Q[0][pick]), "Jq": hx(Q[1][pick]), "Hq": hx(Q[2][pick]),
               "Ex": hx(X[0][pick]), "Jx": hx(X[1][pick]), "Hx": hx(X[2][pick]),
               "Er": hx(R[0][pick]), "Jr": hx(R[1][pick]), "Hr": hx(R[2][pick]), "stats_50k": stats},
              open(os.path.join(HERE, "amips_truth_golden.json"), "w"))
    print(json.dumps(stats, indent=1))


# ------------------------------------------------------------------------------------------------- exact point-triangle
def fr3(v):
    return [Fr(float(x)) for x in v]


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def seg_d2(p, a, b):
    ab, ap = sub(b, a), sub(p, a)
    l2 = dot(ab, ab)
    if l2 == 0:
        return dot(ap, ap)
    t = dot(ap, ab) / l2
    t = min(max(t, Fr(0)), Fr(1))
    q = [a[k] + t * ab[k] for k in range(3)]
    d = sub(p, q)
    return dot(d, d)


def tri_d2_exact(p, a, b, c):
    """exact minimum of |a + s (b-a) + t (c-a) - p|^2 over the closed triangle: interior stationary point if feasible, else
    the best of the three edges (a convex quadratic attains its constrained minimum on the boundary otherwise)"""
    p, a, b, c = fr3(p), fr3(a), fr3(b), fr3(c)
    e0, e1, d = sub(b, a), sub(c, a), sub(a, p)
    a00, a01, a11, b0, b1 = dot(e0, e0), dot(e0, e1), dot(e1, e1), dot(d, e0), dot(d, e1)
    det = a00 * a11 - a01 * a01
    if det != 0:
        s = (a01 * b1 - a11 * b0) / det
        t = (a01 * b0 - a00 * b1) / det
        if s >= 0 and t >= 0 and s + t <= 1:
            q = [a[k] + s * e0[k] + t * e1[k] - p[k] for k in range(3)]
            return dot(q, q)
    return min(seg_d2(p, a, b), seg_d2(p, a, c), seg_d2(p, b, c))


def trisq_cases(n=2000, seed=77):
    rng = np.random.default_rng(seed)
    P, T = [], []
    for i in range(n):
        scale = 10.0 ** rng.uniform(-4, 1)
        tri = rng.normal(size=(3, 3)) * scale + rng.uniform(-1, 1, size=3)
        kind = i % 10
        if kind == 0:   # needle
            tri[2] = tri[0] + (tri[1] - tri[0]) * rng.uniform(0.2, 0.8) + rng.normal(size=3) * scale * 10.0 ** rng.uniform(-5, -2)
        elif kind == 1:  # exactly degenerate: repeated vertex (boundary mesh facets, Preprocess.cpp:192-197)
            tri[2] = tri[1]
        w = rng.dirichlet([1, 1, 1])
        inside = w @ tri
        nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        nn = np.linalg.norm(nrm)
        nrm = nrm / nn if nn > 0 else rng.normal(size=3)
        m = i % 7
        if m == 0:
            p = inside                                                     # on the facet (up to rounding)
        elif m == 1:
            p = inside + nrm * scale * 10.0 ** rng.uniform(-9, -1)         # just above the interior
        elif m == 2:
            p = tri[rng.integers(3)].copy()                                # exactly a vertex
        elif m == 3:
            a, b = tri[rng.integers(3)], tri[rng.integers(3)]
            p = a + (b - a) * rng.uniform(-0.5, 1.5) + nrm * scale * 10.0 ** rng.uniform(-12, -2)  # along / beyond an edge
        elif m == 4:
            p = tri.mean(0) + rng.normal(size=3) * scale * 3                # anywhere around
        elif m == 5:
            p = tri.mean(0) + rng.normal(size=3) * scale * 100              # far
        else:
            k = rng.integers(3)
            p = tri[k] + (tri[k] - tri.mean(0)) * rng.uniform(0, 2) + rng.normal(size=3) * scale * 1e-6  # vertex regions
        P.append(p)
        T.append(tri)
    return np.array(P), np.array(T)


def trisq_truth():
    P, T = trisq_cases()
    d2, cond = [], []
    for p, t in zip(P, T):
        d2.append(float(tri_d2_exact(p, t[0], t[1], t[2])))   # Fraction -> nearest double
        a, b, c = fr3(t[0]), fr3(t[1]), fr3(t[2])
        e0, e1 = sub(b, a), sub(c, a)
        a00, a01, a11 = dot(e0, e0), dot(e0, e1), dot(e1, e1)
        det = a00 * a11 - a01 * a01
        # condition number of the 2x2 solve behind the interior / edge parameters: 1 / sin^2 of the angle at V0 (1 for the
        # degenerate facets, which take the three-segment branch)
        cond.append(float(a00 * a11 / det) if det != 0 else 1.0)
    json.dump({"source": "exact rational arithmetic (fractions.Fraction) on the double inputs, rounded to the nearest double",
               "n": len(P), "P": hx(P), "T": hx(T), "d2": hx(d2), "cond": hx(cond)}, open(os.path.join(HERE, "trisq_exact_golden.json"), "w"))


# ------------------------------------------------------------------------------------------------- mpmath winding
def winding_cases():
    """(name, V, F, Q): closed, reversed, open, doubled, self-intersecting soup; queries inside / outside / a hair off the surface"""
    rng = np.random.default_rng(5)
    out = []
    V, F = synth.icosphere(1)
    V = V + rng.normal(0, 0.01, V.shape)

    def queries(V, F, n):
        lo, hi = V.min(0), V.max(0)
        box = 0.5 * (lo + hi) + 0.7 * (hi - lo) * rng.uniform(-1, 1, size=(n // 2, 3))
        tri = V[F[rng.integers(len(F), size=n - n // 2)]]
        w = rng.dirichlet([1, 1, 1], size=len(tri))
        on = (w[:, :, None] * tri).sum(1)
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        near = on + nrm * (10.0 ** rng.uniform(-7, -1, size=(len(on), 1))) * rng.choice([-1, 1], size=(len(on), 1))
        return np.concatenate([box, near])

    out.append(("icosphere80_noisy", V, F, queries(V, F, 400)))
    out.append(("icosphere80_reversed", V, F[:, [0, 2, 1]], queries(V, F, 200)))
    out.append(("icosphere80_open", V, F[: len(F) * 2 // 3], queries(V, F, 300)))
    out.append(("icosphere80_doubled", V, np.concatenate([F, F]), queries(V, F, 200)))
    Vk, Fk = synth.torus_knot(24, 6)
    out.append(("knot288", Vk, Fk, queries(Vk, Fk, 400)))
    V2 = np.concatenate([V, V * 0.8 + np.array([0.3, 0.1, 0.0])])
    F2 = np.concatenate([F, F + len(V)])
    out.append(("two_spheres_intersecting", V2, F2, queries(V2, F2, 300)))
    soup = rng.normal(size=(60, 3, 3)) * 0.3
    Vs, Fs = soup.reshape(-1, 3), np.arange(180).reshape(60, 3)
    out.append(("soup60", Vs, Fs, queries(Vs, Fs, 200)))
    return out


def winding_truth():
    import mpmath as mp
    mp.mp.dps = 40
    cases = []
    for name, V, F, Q in winding_cases():
        Vm = [[mp.mpf(float(x)) for x in v] for v in V]
        W = []
        for q in Q:
            qm = [mp.mpf(float(x)) for x in q]
            tot = mp.mpf(0)
            for f in F:
                a, b, c = ([Vm[i][k] - qm[k] for k in range(3)] for i in f)
                la, lb, lc = (mp.sqrt(u[0] ** 2 + u[1] ** 2 + u[2] ** 2) for u in (a, b, c))
                det = (a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]))
                den = la * lb * lc + (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) * lc + (b[0] * c[0] + b[1] * c[1] + b[2] * c[2]) * la + \
                    (c[0] * a[0] + c[1] * a[1] + c[2] * a[2]) * lb
                tot += 2 * mp.atan2(det, den)
            W.append(float(tot / (4 * mp.pi)))
        cases.append({"name": name, "V": hx(V), "F": [int(x) for x in np.asarray(F).ravel()], "Q": hx(Q), "W": hx(W)})
        print(name, len(F), "facets", len(Q), "queries")
    json.dump({"source": "mpmath, 40 significant digits: W = sum_f 2 atan2(det[a b c], |a||b||c| + (a.b)|c| + (b.c)|a| + (c.a)|b|) / (4 pi)",
               "cases": cases}, open(os.path.join(HERE, "winding_mp_golden.json"), "w"))


# ------------------------------------------------------------------------------------------------- mpmath dihedral angles
def dihedral_truth():
    """min / max dihedral angle of 600 tets from the mathematical definition -- the angle between the two faces that meet in an
    edge, acos(-N_a . N_b) with N_i the unit normal of the face opposite vertex i pointing towards vertex i -- at 40 digits.
    Independent of the CGAL plane / projection constructions calTetQuality_AD goes through (LocalOperations.cpp:783-860)."""
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(12)
    T = synth.random_tets(600, seed=33, scale_lo=1e-2, scale_hi=1e2)         # (12, n)
    X = T.T.reshape(-1, 4, 3)
    X[::7, 3] = X[::7, :3].mean(1) + 0.05 * (X[::7, 3] - X[::7, :3].mean(1))  # some flat ones (small and large angles)
    lo, hi = [], []
    for x in X:
        v = [[mp.mpf(float(c)) for c in p] for p in x]

        def sub(a, b):
            return [a[k] - b[k] for k in range(3)]

        def cross(a, b):
            return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

        def dot(a, b):
            return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
        N = []
        for i in range(4):
            o = [v[(i + 1) % 4], v[(i + 2) % 4], v[(i + 3) % 4]]
            n = cross(sub(o[1], o[0]), sub(o[2], o[0]))
            if dot(n, sub(v[i], o[0])) < 0:
                n = [-c for c in n]
            ln = mp.sqrt(dot(n, n))
            N.append([c / ln for c in n])
        ang = [mp.acos(max(mp.mpf(-1), min(mp.mpf(1), -dot(N[a], N[b])))) for a in range(4) for b in range(a + 1, 4)]
        lo.append(float(min(ang)))
        hi.append(float(max(ang)))
    json.dump({"source": "mpmath, 40 digits: dihedral angle at the edge shared by the faces opposite vertices a and b = acos(-N_a . N_b)",
               "n": len(X), "X": hx(X), "min_d_angle": hx(lo), "max_d_angle": hx(hi)}, open(os.path.join(HERE, "dihedral_mp_golden.json"), "w"))


if __name__ == "__main__":
    what = sys.argv[1:] or ["amips", "trisq", "winding", "dihedral"]
    O.build()
    if "amips" in what:
        amips_truth()
    if "trisq" in what:
        trisq_truth()
    if "winding" in what:
        winding_truth()
    if "dihedral" in what:
        dihedral_truth()
    for fn in ("amips_truth_golden.json", "trisq_exact_golden.json", "winding_mp_golden.json", "dihedral_mp_golden.json"):
        p = os.path.join(HERE, fn)
        if os.path.exists(p):
            print(fn, os.path.getsize(p), "bytes")
