"""CPU tier: the product's numeric core (tw_math.cuh / sampling.cuh, the source the kernels execute) compiled for the
host (tests/host_harness.cpp) against the oracle. Bit exact for distances, samples and predicates; AMIPS closed form
within 1e-9 of the tensor max-norm."""
import ctypes as C

import numpy as np

from conftest import amips_close
from tetwild_b200 import synth

_dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(_dp)


def normwise(a, b):
    a2, b2 = a.reshape(len(a), -1), b.reshape(len(b), -1)
    return (np.abs(a2 - b2).max(1) / np.maximum(np.abs(b2).max(1), 1e-300)).max()


def test_amips_closed_form(harness, oracle):
    T = np.ascontiguousarray(synth.random_tets(20000, seed=7))
    n = T.shape[1]
    E, J, H = np.empty(n), np.empty((n, 3)), np.empty((n, 9))
    harness.hh_amips_ejh_batch(P(T), C.c_uint64(n), P(E), P(J), P(H))
    Eo, Jo, Ho = oracle.amips_ejh_soa(T, threads=4)
    assert max(amips_close(T, (E, J, H), (Eo, Jo, Ho))) < 1e-9
    if oracle.ref_available():
        assert max(amips_close(T, (E, J, H), oracle.ref_amips_ejh_soa(T, threads=4))) < 1e-9


def test_point_triangle_distance_bit_exact(harness, oracle):
    rng = np.random.default_rng(3)
    for it in range(20000):
        v = rng.normal(size=(3, 3)) * rng.choice([1, 0.01])
        p = rng.normal(size=3) * rng.choice([1, 0.02, 3])
        if it % 7 == 0:
            p = v[0] * 0.3 + v[1] * 0.3 + v[2] * 0.4
        if it % 11 == 0:
            v[2] = v[0] + (v[1] - v[0]) * rng.uniform()
        if it % 13 == 0:
            v[2] = v[1].copy()
        near = np.empty(3)
        d = harness.hh_tri_sqdist(P(p), P(v[0]), P(v[1]), P(v[2]), P(near))
        d2, n2 = oracle.point_triangle_sqdist(p, v[0], v[1], v[2])
        assert d == d2 and np.array_equal(near, n2)


def test_point_triangle_distance_every_outcome_class(harness, oracle):
    """The product routine selects one of eight outcome classes by comparisons and then runs straight-line arithmetic
    (tw_math.cuh::tri_sqdist_rec); queries ON vertices / edges / the facet, just off an edge line beyond its ends, and far
    away reach every class and every region boundary. Bit-exact d2 and nearest point against the oracle."""
    rng = np.random.default_rng(99)
    for it in range(30000):
        v = rng.normal(size=(3, 3)) * rng.choice([1, 0.01, 100])
        k = it % 10
        if k == 0:
            p = v[rng.integers(3)].copy()
        elif k == 1:
            a, b = rng.choice(3, 2, replace=False)
            p = v[a] + (v[b] - v[a]) * rng.uniform()
        elif k == 2:
            a, b = rng.choice(3, 2, replace=False)
            p = v[a] + (v[b] - v[a]) * rng.uniform(-1, 2) + 1e-3 * rng.normal(size=3)
        elif k == 3:
            p = (v * rng.dirichlet([1, 1, 1])[:, None]).sum(0)
        elif k == 4:
            w = rng.uniform(-1, 2, 3)
            p = (v * w[:, None]).sum(0) + rng.normal(size=3) * rng.choice([0, 1e-6, 1])
        else:
            p = rng.normal(size=3) * rng.choice([1, 0.02, 3, 300])
        near = np.empty(3)
        d = harness.hh_tri_sqdist(P(p), P(v[0]), P(v[1]), P(v[2]), P(near))
        d2, n2 = oracle.point_triangle_sqdist(p, v[0], v[1], v[2])
        assert d == d2 and np.array_equal(near, n2)


def test_sampling_bit_exact(harness, oracle):
    rng = np.random.default_rng(4)
    for it in range(400):
        tri = rng.normal(size=(3, 3)) * rng.choice([0.003, 0.01, 0.05, 0.2])
        if it % 5 == 0:
            tri = np.round(tri * 100) / 100
        if it % 17 == 0:
            tri = np.array([[0, 0, 0], [0.05, 0, 0], [0, 0.05, 0]]) + rng.integers(-3, 3, size=3)
        sd = 1e-3 if it % 2 else 2.5e-3
        tri = np.ascontiguousarray(tri.reshape(9), dtype=np.float64)
        ref = oracle.sample_triangle(tri, sd)
        out = np.empty((len(ref) + 8, 3))
        n = harness.hh_sample_triangle(P(tri), C.c_double(sd), P(out), C.c_uint64(len(out)))
        assert n == len(ref) and np.array_equal(out[:n], ref)


def test_predicates(harness, oracle):
    rng = np.random.default_rng(2)
    for it in range(3000):
        a, b, c = rng.normal(size=(3, 3))
        if it % 3 == 0:
            w = rng.dirichlet([1, 1, 1])
            d = w[0] * a + w[1] * b + w[2] * c
        elif it % 3 == 1:
            a, b, c = np.round(a * 8) / 8, np.round(b * 8) / 8, np.round(c * 8) / 8
            w = np.round(rng.dirichlet([1, 1, 1]) * 4) / 4
            w[2] = 1 - w[0] - w[1]
            d = w[0] * a + w[1] * b + w[2] * c
        else:
            d = rng.normal(size=3)
        a, b, c, d = [np.ascontiguousarray(x) for x in (a, b, c, d)]
        s = oracle.orient3d_exact(a, b, c, d)
        assert harness.hh_orient3d(P(a), P(b), P(c), P(d)) == s
        assert harness.hh_orient3d_exact(P(a), P(b), P(c), P(d)) == s
        assert harness.hh_cgal_orientation(P(d), P(a), P(b), P(c)) == s
    z = [np.array(x, dtype=np.float64) for x in ([0, 0, 0], [1, 1, 1], [2, 2, 2], [2, 2, 2.0000000001])]
    assert harness.hh_triangle_is_degenerate(P(z[0]), P(z[1]), P(z[2])) == 1
    assert harness.hh_triangle_is_degenerate(P(z[0]), P(z[1]), P(z[3])) == 0


def test_oracle_dihedral_known_answers(oracle):
    """calTetQuality_AD (LocalOperations.cpp:783-860): regular tet -> acos(1/3) six times; corner tet -> three right
    angles and three acos(1/sqrt 3); a vertex on the opposite plane or a degenerate plane -> (0, pi) (:790-798)."""
    R = np.array([[0, 0, 0], [1, 0, 0], [.5, 3 ** .5 / 2, 0], [.5, 3 ** .5 / 6, 6 ** .5 / 3]])
    Cn = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.]])
    flat = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.]])
    dup = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 1.]])
    V = np.concatenate([R, Cn, flat, dup])
    lo, hi = oracle.tet_dihedral(V, np.arange(16, dtype=np.int32).reshape(4, 4))
    assert abs(lo[0] - np.arccos(1 / 3)) < 1e-14 and abs(hi[0] - np.arccos(1 / 3)) < 1e-14
    assert abs(lo[1] - np.arccos(1 / 3 ** .5)) < 1e-14 and abs(hi[1] - np.pi / 2) < 1e-14
    assert lo[2] == 0 and hi[2] == np.pi and lo[3] == 0 and hi[3] == np.pi
    # angles of a tet sum to more than 2 pi and less than 3 pi; invariant under uniform scale and translation
    from tetwild_b200 import synth
    Vg, Tg = synth.grid_tet_mesh(5, 4, 3)
    a, b = oracle.tet_dihedral(Vg, Tg)
    a2, b2 = oracle.tet_dihedral(Vg * 37.0 + 5.0, Tg)
    assert np.abs(a - a2).max() < 1e-9 and np.abs(b - b2).max() < 1e-9 and (a > 0).all() and (b < np.pi).all()


def _angle_sum(harness, x, y, skip=None, tile=32):
    x, y = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64)
    k = C.c_int(0)
    sk = np.ascontiguousarray(skip, dtype=np.uint8) if skip is not None else None
    v = harness.hh_angle_sum(P(x), P(y), sk.ctypes.data_as(C.POINTER(C.c_uint8)) if sk is not None else None, C.c_uint64(len(x)), C.c_int(tile), C.byref(k))
    return v, k.value


def test_winding_angle_accumulator(harness):
    """The winding kernel never evaluates atan2 per triangle: sum_k atan2(y_k, x_k) = arg(prod (x_k + i y_k)) + 2 pi K with K
    kept by branch-free sign-bit bookkeeping (winding_math.cuh::Angle). Same source, on the host: random factor streams of
    wildly different magnitudes, long one-directional runs (K large), factors at +-pi, zero factors and chain starts."""
    import math
    rng = np.random.default_rng(17)
    for it in range(300):
        n = int(rng.integers(1, 400))
        th = rng.uniform(-math.pi, math.pi, n) * rng.choice([1.0, 0.3, 0.02])
        if it % 5 == 0:
            th = np.abs(th)                                   # only counter-clockwise: K grows
        if it % 7 == 0:
            th = -np.abs(th)
        mag = 10.0 ** rng.uniform(-15, 15, n)
        x, y = mag * np.cos(th), mag * np.sin(th)
        skip = (rng.random(n) < 0.05).astype(np.uint8)
        if it % 3 == 0:
            zi = rng.integers(0, n, max(1, n // 20))
            x[zi] = 0.0; y[zi] = 0.0                          # atan2(0, 0) = 0 in the reference
        ref = math.fsum(math.atan2(b, a) for a, b, s in zip(x, y, skip) if not s)
        got, K = _angle_sum(harness, x, y, skip, tile=int(rng.choice([1, 7, 32])))
        assert abs(got - ref) <= 1e-13 * max(1.0, n), (it, got, ref, K)
    # factors exactly on the negative real axis: atan2(+0, -1) = +pi, atan2(-0, -1) = -pi
    for y0, want in ((0.0, math.pi), (-0.0, -math.pi)):
        got, K = _angle_sum(harness, [-1.0], [y0])
        assert got == want
    got, K = _angle_sum(harness, [-1.0] * 6, [0.0] * 6)          # six half turns = 6 pi
    assert abs(got - 6 * math.pi) < 1e-14
    got, K = _angle_sum(harness, [-2.0, -3.0], [1e-300, -1e-300])  # just short of +pi, then just short of -pi: total ~ 0
    assert abs(got) < 1e-290 or abs(got) < 1e-15


def test_winding_norm_and_closed_surface(harness):
    """|v| through one third-order correction of a 22-bit reciprocal-root seed is correctly rounded to ~1 ulp, and the factors
    of a closed surface add up to 4 pi (inside) / 0 (outside) through the accumulator."""
    import math
    rng = np.random.default_rng(5)
    v = rng.normal(size=(20000, 3)) * 10.0 ** rng.uniform(-100, 100, (20000, 1))
    got = np.array([harness.hh_norm3(C.c_double(a), C.c_double(b), C.c_double(c)) for a, b, c in v])
    ref = np.sqrt((v.astype(np.longdouble) ** 2).sum(1)).astype(np.float64)
    assert (np.abs(got - ref) <= 4e-16 * ref).all()
    assert harness.hh_norm3(C.c_double(0), C.c_double(0), C.c_double(0)) > 0       # a query ON a vertex: tiny positive length
    V, F = synth.icosphere(2)
    for q, want in (([0.05, -0.1, 0.02], 4 * math.pi), ([0.9, 0.2, 0.1], 0.0), ([0.0, 0.0, 0.49], 4 * math.pi)):
        xs, ys = [], []
        for f in F.astype(np.int64):
            xy = np.empty(2)
            harness.hh_solid_angle_factor(P(np.array(q, dtype=np.float64)), P(V[f[0]].copy()), P(V[f[1]].copy()), P(V[f[2]].copy()), P(xy))
            xs.append(xy[0]); ys.append(xy[1])
        got, K = _angle_sum(harness, xs, ys)
        assert abs(2.0 * got - want) < 1e-11            # Omega = 2 atan2(y, x) per triangle


def test_conservative_box_bound_is_rigorous(harness):
    """Every traversal prunes with tw_math.cuh::box_d2_lb: it must NEVER exceed the true squared distance from the (double)
    query to the (float) box, or a subtree holding a facet within eps could be skipped and a decision would flip. Checked
    against exact rational arithmetic, at scales from 1e-6 to 1e6 and for points inside, on and barely outside boxes."""
    from fractions import Fraction
    rng = np.random.default_rng(23)
    worst = 0.0
    for it in range(6000):
        scale = 10.0 ** rng.uniform(-6, 6)
        lo = (rng.normal(size=3) * scale).astype(np.float32)
        hi = (lo.astype(np.float64) + np.abs(rng.normal(size=3)) * scale * rng.choice([1.0, 1e-3, 0.0])).astype(np.float32)
        hi = np.maximum(hi, lo)
        k = it % 4
        if k == 0:
            p = rng.normal(size=3) * scale * 3
        elif k == 1:
            p = lo.astype(np.float64) + (hi.astype(np.float64) - lo.astype(np.float64)) * rng.uniform(0, 1, 3)      # inside
        elif k == 2:
            p = hi.astype(np.float64) * (1 + rng.choice([1e-16, 1e-12, 1e-8, 1e-4], 3))                            # barely outside
        else:
            p = lo.astype(np.float64) - np.abs(rng.normal(size=3)) * scale * rng.choice([1e-9, 1e-3, 1.0])
        box = np.concatenate([lo, hi]).astype(np.float32)
        lb = float(harness.hh_box_d2_lb(P(np.ascontiguousarray(p)), box.ctypes.data_as(C.POINTER(C.c_float))))
        exact = Fraction(0)
        for c in range(3):
            d = max(Fraction(float(lo[c])) - Fraction(float(p[c])), Fraction(float(p[c])) - Fraction(float(hi[c])), Fraction(0))
            exact += d * d
        assert Fraction(lb) <= exact, (it, lb, float(exact))
        coord = max(float(np.abs(p).max()), float(np.abs(box).max()))
        if float(exact) > (1e-2 * coord) ** 2:   # separations well above float resolution: the bound is also tight
            worst = max(worst, 1.0 - lb / float(exact))
    assert worst < 1e-3
