"""Synthetic workloads for the TetWild hot path (BASELINE.json configs).

numpy only; fixed seeds; every surface is normalised to a unit bounding-box diagonal so that `eps_rel` of the
reference (src/tetwild/State.cpp:24,36-41) can be applied literally.
"""
import math

import numpy as np


def normalise_unit_diag(V):
    lo, hi = V.min(0), V.max(0)
    diag = float(np.linalg.norm(hi - lo))
    return (V - 0.5 * (lo + hi)) / diag


def state_eps(eps_rel=1e-3, diag=1.0, stage=1, sub_stage=1):
    """State::State, src/tetwild/State.cpp:24,36-41 -> (sampling_dist, eps, eps_2)."""
    eps_input = diag * eps_rel
    sampling_dist = eps_input / stage
    eps = eps_input - sampling_dist / math.sqrt(3) * (stage + 1 - sub_stage)
    return sampling_dist, eps, eps * eps


def icosphere(subdiv=5, radius=0.5):
    """20 * 4**subdiv triangles (subdiv=5 -> 20480: BASELINE.json config 1)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    V = [np.array(v, dtype=np.float64) / math.sqrt(1 + t * t) for v in V]
    for _ in range(subdiv):
        cache, F2 = {}, []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = V[a] + V[b]
                V.append(m / np.linalg.norm(m))
                cache[key] = len(V) - 1
            return cache[key]

        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = F2
    return np.array(V) * radius, np.array(F, dtype=np.uint32)


def torus_knot(nu=1000, nv=100, p=2, q=3, tube=0.06):
    """(p,q) torus-knot tube, nu x nv quads -> 2*nu*nv triangles (1000 x 100 -> 200 000: config 2)."""
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)

    def curve(t):
        r = 1.0 + 0.4 * np.cos(q * t)
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), 0.4 * np.sin(q * t)], 1)

    c = curve(u)
    h = 1e-4
    tan = curve(u + h) - curve(u - h)
    tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    # rotation-minimising frame by parallel transport, closed up with a uniform twist
    n = np.zeros_like(c)
    ref = np.array([0.0, 0.0, 1.0])
    n0 = ref - tan[0] * (ref @ tan[0])
    n[0] = n0 / np.linalg.norm(n0)
    for i in range(1, nu):
        v = n[i - 1] - tan[i] * (n[i - 1] @ tan[i])
        n[i] = v / np.linalg.norm(v)
    b = np.cross(tan, n)
    v_last = n[-1] - tan[0] * (n[-1] @ tan[0])
    v_last /= np.linalg.norm(v_last)
    ang = math.atan2(float(v_last @ b[0]), float(v_last @ n[0]))
    tw = -ang * np.arange(nu) / nu
    n2 = n * np.cos(tw)[:, None] + b * np.sin(tw)[:, None]
    b2 = np.cross(tan, n2)
    v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    V = (c[:, None, :] + tube * (np.cos(v)[None, :, None] * n2[:, None, :] + np.sin(v)[None, :, None] * b2[:, None, :]))
    V = V.reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * nv + j).ravel()
    bb = (((i + 1) % nu) * nv + j).ravel()
    cc = (((i + 1) % nu) * nv + (j + 1) % nv).ravel()
    d = (i * nv + (j + 1) % nv).ravel()
    F = np.concatenate([np.stack([a, bb, cc], 1), np.stack([a, cc, d], 1)]).astype(np.uint32)
    return normalise_unit_diag(V), F


def uv_sphere(nu=708, nv=708, radius=0.5, noise=0.01, seed=11, center=(0.0, 0.0, 0.0), normalise=True):
    """Closed, outward-oriented UV sphere with radial noise: 2*nu*(nv-1) triangles (708 x 708 -> 1 001 112: config 4)."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0, np.pi, nv + 1)[1:-1]  # nv-1 rings
    ph = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    ring = np.stack([np.sin(th)[:, None] * np.cos(ph)[None, :], np.sin(th)[:, None] * np.sin(ph)[None, :],
                     np.repeat(np.cos(th)[:, None], nu, 1)], 2).reshape(-1, 3)
    V = np.concatenate([[[0, 0, 1.0]], ring, [[0, 0, -1.0]]])
    V = V * (radius * (1.0 + noise * rng.uniform(-1, 1, size=(len(V), 1))))
    nr = nv - 1
    south = 1 + nr * nu
    j = np.arange(nu)
    jn = (j + 1) % nu
    F = [np.stack([np.zeros(nu, dtype=np.int64), 1 + j, 1 + jn], 1)]
    for r in range(nr - 1):
        a, b = 1 + r * nu + j, 1 + r * nu + jn
        c, d = 1 + (r + 1) * nu + j, 1 + (r + 1) * nu + jn
        F.append(np.stack([a, c, d], 1))
        F.append(np.stack([a, d, b], 1))
    last = 1 + (nr - 1) * nu
    F.append(np.stack([last + j, np.full(nu, south), last + jn], 1))
    F = np.concatenate(F).astype(np.uint32)
    V = V + np.asarray(center)
    return (normalise_unit_diag(V) if normalise else V), F


def sphere_union(k=64, nu=128, nv=129, seed=5, noise=0.005):
    """Concatenation (no boolean) of k noisy overlapping UV spheres: config 5 geometry (k=64, 128x129 -> 2.1 M tris)."""
    rng = np.random.default_rng(seed)
    Vs, Fs, off = [], [], 0
    for i in range(k):
        c = rng.uniform(-1, 1, 3)
        r = rng.uniform(0.3, 0.6)
        V, F = uv_sphere(nu, nv, radius=r, noise=noise, seed=seed * 1000 + i, center=c, normalise=False)
        Vs.append(V)
        Fs.append(F + off)
        off += len(V)
    return normalise_unit_diag(np.concatenate(Vs)), np.concatenate(Fs).astype(np.uint32)


def envelope_points(V, F, n, eps, seed=20240501):
    """Config 2 query mix: 50 % surface sample + N(0, eps) normal offset, 25 % uniform in 1.1x bbox, 25 % on facets."""
    rng = np.random.default_rng(seed)
    n_near, n_box = n // 2, n // 4
    n_on = n - n_near - n_box
    tri = V[F.astype(np.int64)]
    area2 = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cdf = np.cumsum(area2)
    cdf /= cdf[-1]

    def on_surface(m):
        f = np.searchsorted(cdf, rng.uniform(size=m)).clip(0, len(F) - 1)
        r1, r2 = np.sqrt(rng.uniform(size=m)), rng.uniform(size=m)
        w = np.stack([1 - r1, r1 * (1 - r2), r1 * r2], 1)
        p = (tri[f] * w[:, :, None]).sum(1)
        nrm = np.cross(tri[f, 1] - tri[f, 0], tri[f, 2] - tri[f, 0])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
        return p, nrm

    p, nrm = on_surface(n_near)
    near = p + nrm * rng.normal(0.0, eps, size=(n_near, 1))
    lo, hi = V.min(0), V.max(0)
    c, h = 0.5 * (lo + hi), 0.55 * (hi - lo)
    box = c + h * rng.uniform(-1, 1, size=(n_box, 3))
    on, _ = on_surface(n_on)
    P = np.concatenate([near, box, on])
    return np.ascontiguousarray(P[rng.permutation(n)])


def random_tets(n, seed=7, scale_lo=1e-3, scale_hi=1e3, trans=10.0, trans_scales=True, sigma=0.15):
    """Config 3: regular unit tet + N(0, sigma), random rotation, log-uniform scale, translation.

    Returns T as a (12, n) SoA array (row 3*i+k = coordinate k of vertex i), all CGAL-POSITIVE, volume >= 1e-6 l^3.
    With trans_scales=True the translation is U(-trans, trans)*scale (coordinates stay commensurate with edge
    lengths, as in a real mesh); with False it is the literal U(-trans, trans) of SURVEY.md 8d, for which the
    reference's absolute-coordinate formula is itself ill-conditioned at small scales (DESIGN.md "AMIPS tolerance").
    """
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [1, 0, 0], [0.5, math.sqrt(3) / 2, 0], [0.5, math.sqrt(3) / 6, math.sqrt(6) / 3]])
    out = np.empty((12, n))
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.3) + 16
        X = base[None] + rng.normal(0, sigma, size=(m, 4, 3))
        q = rng.normal(size=(m, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        w, x, y, z = q.T
        R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                      np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                      np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
        X = np.einsum("mij,mvj->mvi", R, X)
        e = X[:, 1:] - X[:, :1]
        det = np.einsum("mi,mi->m", e[:, 0], np.cross(e[:, 1], e[:, 2]))
        ok = det > 6e-6 * 1.0
        X = X[ok]
        s = np.exp(rng.uniform(math.log(scale_lo), math.log(scale_hi), size=(len(X), 1, 1)))
        t = rng.uniform(-trans, trans, size=(len(X), 1, 3))
        X = X * s + (t * s if trans_scales else t)
        k = min(len(X), n - filled)
        out[:, filled:filled + k] = X[:k].reshape(k, 12).T
        filled += k
    return out


def ring_groups(n_groups, seed=7, kmin=12, kmax=36, **kw):
    """Config 3 'smoothing-candidate batch layout': groups of k~U{kmin..kmax} tets sharing a center vertex.

    Returns (V[nV,3], tets[nT,4] int32, group_off[nG+1] uint64, center[nG] int32). The center vertex sits at a
    random slot of each tet (NewtonsUpdate rotates it to slot 0, VertexSmoother.cpp:640-651) and every tet is
    CGAL-POSITIVE with the center in that slot order.
    """
    rng = np.random.default_rng(seed + 1)
    k = rng.integers(kmin, kmax + 1, size=n_groups)
    off = np.zeros(n_groups + 1, dtype=np.uint64)
    off[1:] = np.cumsum(k)
    nT = int(off[-1])
    T = random_tets(nT, seed=seed + 2, **kw).T.reshape(nT, 4, 3)
    g = np.repeat(np.arange(n_groups), k)
    first = off[:-1].astype(np.int64)
    # make all tets of a group share vertex 0 of the group's first tet (translate each tet onto it)
    T = T - T[:, :1] + T[first[g], :1]
    V = np.empty((n_groups + 3 * nT, 3))
    V[:n_groups] = T[first, 0]
    V[n_groups:] = T[:, 1:].reshape(-1, 3)
    tets = np.empty((nT, 4), dtype=np.int32)
    tets[:, 0] = g
    tets[:, 1:] = n_groups + np.arange(3 * nT).reshape(nT, 3)
    # cyclic rotations of 4 elements are odd permutations; an even rotation (by 2) keeps orientation, odd ones
    # are compensated by swapping the last two non-center vertices
    rot = rng.integers(0, 4, size=nT)
    out = np.empty_like(tets)
    for r in range(4):
        m = rot == r
        tt = tets[m].copy()
        if r % 2 == 1:
            tt[:, [2, 3]] = tt[:, [3, 2]]
        out[m] = np.roll(tt, r, axis=1)
    return V, out, off, np.arange(n_groups, dtype=np.int32)


def grid_tet_mesh(nx, ny, nz, jitter=0.25, seed=13, shuffle=True):
    """A connected tet mesh of the scheduler's shape (`tet_vertices[].posf`, `tets`): an (nx, ny, nz)-cell grid, every
    cell cut into 6 tets along its main diagonal (Kuhn), vertices jittered by `jitter` cell widths (< 0.29 keeps
    every tet positive). All tets are CGAL-POSITIVE; interior vertices have one-rings of 24 tets (the mean ring size
    the reference reports). Returns (V[nV,3] f64, tets[nT,4] int32); tets and vertex slots randomly permuted."""
    rng = np.random.default_rng(seed)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    V = np.stack([gx, gy, gz], -1).reshape(-1, 3).astype(np.float64)
    V += rng.uniform(-jitter, jitter, V.shape)
    V /= max(nx, ny, nz)

    def vid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    tets = []
    import itertools
    for perm in itertools.permutations(range(3)):
        d = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            d[s + 1] = d[s]
            d[s + 1, ax] += 1
        corners = [vid(ci + d[s, 0], cj + d[s, 1], ck + d[s, 2]) for s in range(4)]
        # orientation = sign of the permutation; swap two vertices of the odd ones
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b])
        if inv % 2 == 1:
            corners[2], corners[3] = corners[3], corners[2]
        tets.append(np.stack(corners, 1))
    tets = np.concatenate(tets).astype(np.int32)
    if shuffle:
        tets = tets[rng.permutation(len(tets))]
        pv = rng.permutation(len(V))          # new slot of old vertex i
        Vn = np.empty_like(V)
        Vn[pv] = V
        V, tets = Vn, pv[tets].astype(np.int32)
        # even permutations of the 4 slots keep orientation: rotate (1,2,3) at random
        r = rng.integers(0, 3, len(tets))
        for q in (1, 2):
            m = r == q
            tets[m, 1:] = np.roll(tets[m, 1:], q, axis=1)
    return V, np.ascontiguousarray(tets)


def cube_surface(k=41, half=0.5):
    """Closed, outward-oriented cube of half-size `half`, every face a k x k grid of quads split in two: 12 k^2 triangles
    (k = 41 -> 20 172, the size of the config-1 icosphere) with LARGE FLAT regions -- the surface on which a candidate face of
    edge ~ diag/20 can lie wholly inside the envelope, so that all ~1.4 k samples of it are tested."""
    V, F = [], []
    lin = np.linspace(-half, half, k + 1)
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            base = len(V)
            u, v = np.meshgrid(lin, lin, indexing="ij")
            pts = np.empty((k + 1, k + 1, 3))
            pts[..., axis] = sgn * half
            pts[..., (axis + 1) % 3] = u
            pts[..., (axis + 2) % 3] = v
            V.extend(pts.reshape(-1, 3))
            for i in range(k):
                for j in range(k):
                    a, b, c, d = base + i * (k + 1) + j, base + (i + 1) * (k + 1) + j, base + (i + 1) * (k + 1) + j + 1, base + i * (k + 1) + j + 1
                    if sgn > 0:
                        F += [(a, b, c), (a, c, d)]
                    else:
                        F += [(a, c, b), (a, d, c)]
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.uint32)


def winding_queries(V, n, seed=11, scale=1.2):
    rng = np.random.default_rng(seed)
    lo, hi = V.min(0), V.max(0)
    c, h = 0.5 * (lo + hi), 0.5 * scale * (hi - lo)
    return c + h * rng.uniform(-1, 1, size=(n, 3))


def face_queries(V, F, n, edge, eps, seed=3):
    """Config 1-shaped envelope calls: triangles of edge ~`edge` lying near the surface (offset ~N(0, eps/2))."""
    rng = np.random.default_rng(seed)
    tri = V[F.astype(np.int64)]
    f = rng.integers(0, len(F), size=n)
    w = rng.dirichlet([1, 1, 1], size=n)
    c = (tri[f] * w[:, :, None]).sum(1)
    nrm = np.cross(tri[f, 1] - tri[f, 0], tri[f, 2] - tri[f, 0])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    a = np.cross(nrm, rng.normal(size=(n, 3)))
    a /= np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-300)
    b = np.cross(nrm, a)
    ang = rng.uniform(0, 2 * np.pi, size=(n, 1)) + np.array([[0, 2 * np.pi / 3, 4 * np.pi / 3]])
    r = edge / math.sqrt(3) * rng.uniform(0.6, 1.2, size=(n, 3))
    T = c[:, None, :] + r[:, :, None] * (np.cos(ang)[:, :, None] * a[:, None, :] + np.sin(ang)[:, :, None] * b[:, None, :])
    T = T + nrm[:, None, :] * rng.normal(0, eps / 2, size=(n, 1, 1))
    return np.ascontiguousarray(T.reshape(n, 9))
