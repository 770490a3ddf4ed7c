"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes binding of the CPU oracle (oracle/liboracle.so, built by oracle/Makefile) and, when present, of the pieces of
the unmodified reference compiled into oracle/_ref/libtetwild_ref.so (oracle/ref_build.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package --
as the checker or as the timed CPU baseline. Nothing under tetwild_b200/ imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libtetwild_ref.so")
_REFQ = os.path.join(_HERE, "_ref", "libtetwild_ref_quad.so")

MAX_ENERGY = 1e50
NO_FACET = 0xFFFFFFFF

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile liboracle.so (and _ref when /root/reference exists). Building the checker is not using it."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in ("predicates.c", "amips.c", "envelope.c", "winding.c", "tw_oracle.h")
    ):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/tetwild") and (force or not os.path.exists(_REF) or not os.path.exists(_REFQ)
                                                        or os.path.getmtime(os.path.join(_HERE, "ref_quad.cpp")) > os.path.getmtime(_REFQ)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


_lib = None
_lib_native = None
_use_native = False


def _load(path):
    L = C.CDLL(path)
    _restypes(L)
    return L


class native:
    """`with oracle.native():` routes every oracle call through the same C sources compiled with -O3 -march=native ON THIS
    MACHINE (the upper-bound CPU mode; the default build is the reference's Release flags, -O2, no
    -march). Built on first use into /tmp, never shipped."""

    def __enter__(self):
        global _lib_native, _use_native
        if _lib_native is None:
            import hashlib
            import tempfile
            tag = hashlib.sha1(open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0].encode()).hexdigest()[:10] if os.path.exists("/proc/cpuinfo") else "x"
            out = os.path.join(tempfile.gettempdir(), "liboracle_native_%s.so" % tag)
            srcs = [os.path.join(_HERE, f) for f in ("predicates.c", "amips.c", "envelope.c", "winding.c")]
            if not os.path.exists(out) or any(os.path.getmtime(f) > os.path.getmtime(out) for f in srcs):
                cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
                subprocess.check_call([cc, "-O3", "-march=native", "-fPIC", "-fopenmp", "-ffp-contract=off", "-std=gnu11", "-shared", "-o", out] + srcs + ["-lm"])
            _lib_native = _load(out)
        _use_native = True
        return self

    def __exit__(self, *a):
        global _use_native
        _use_native = False


def _restypes(L):
    if True:
        if True:
            pass
        L.ora_amips_energy.restype = C.c_double
        L.ora_solid_angle_w.restype = C.c_double
        L.ora_point_triangle_sqdist.restype = C.c_double
        L.ora_surface_create.restype = C.c_void_p
        L.ora_wtree_create.restype = C.c_void_p
        L.ora_sample_triangle.restype = C.c_uint64
        L.ora_wtree_stats.restype = C.c_uint64
        L.ora_surface_num_facets.restype = C.c_uint32


def lib():
    global _lib
    if _use_native and _lib_native is not None:
        return _lib_native
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = _load(_LIB)
    return _lib


def max_threads():
    """Host threads available to this process. Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, and every
    oracle entry point passes its thread count explicitly (num_threads clause), which overrides that variable."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------- predicates
def cgal_orientation(p, q, r, s):
    a = [_f64(x) for x in (p, q, r, s)]
    return int(lib().ora_cgal_orientation(*[_p(x, _dp) for x in a]))


def orient3d_exact(a, b, c, d):
    v = [_f64(x) for x in (a, b, c, d)]
    return int(lib().ora_orient3d_exact(*[_p(x, _dp) for x in v]))


def triangle_is_degenerate(p, q, r):
    v = [_f64(x) for x in (p, q, r)]
    return bool(lib().ora_triangle_is_degenerate(*[_p(x, _dp) for x in v]))


# ---------------------------------------------------------------------------------------------- AMIPS
def _soa_ptrs(T):
    """T: (12, n) array (row k = coordinate k of every tet) -> ctypes array of 12 row pointers"""
    T = _f64(T)
    assert T.ndim == 2 and T.shape[0] == 12
    arr = (_dp * 12)(*[T[k].ctypes.data_as(_dp) for k in range(12)])
    return T, arr


def amips_energy(T12):
    t = _f64(T12)
    return float(lib().ora_amips_energy(_p(t, _dp)))


def amips_jacobian(T12):
    t = _f64(T12)
    J = np.empty(3)
    lib().ora_amips_jacobian(_p(t, _dp), _p(J, _dp))
    return J


def amips_hessian(T12):
    t = _f64(T12)
    H = np.empty(9)
    lib().ora_amips_hessian(_p(t, _dp), _p(H, _dp))
    return H.reshape(3, 3)


def amips_energy_soa(T, threads=1):
    T, ptrs = _soa_ptrs(T)
    n = T.shape[1]
    E = np.empty(n)
    lib().ora_amips_energy_soa(ptrs, _p(E, _dp), C.c_uint64(n), C.c_int(threads))
    return E


def amips_ejh_soa(T, threads=1):
    T, ptrs = _soa_ptrs(T)
    n = T.shape[1]
    E, J, H = np.empty(n), np.empty((n, 3)), np.empty((n, 9))
    lib().ora_amips_ejh_soa(ptrs, _p(E, _dp), _p(J, _dp), _p(H, _dp), C.c_uint64(n), C.c_int(threads))
    return E, J, H


def amips_quality(V, tets, threads=1):
    V, tets = _f64(V), _i32(tets)
    n = tets.shape[0]
    out = np.empty(n)
    lib().ora_amips_quality(_p(V, _dp), _p(tets, _i32p), C.c_uint64(n), _p(out, _dp), C.c_int(threads))
    return out


def amips_ring_ejh(V, tets, group_off, center, t_ids=None, threads=1):
    V, tets = _f64(V), _i32(tets)
    off = np.ascontiguousarray(group_off, dtype=np.uint64)
    center = _i32(center)
    tid = _i32(t_ids) if t_ids is not None else None
    g = center.shape[0]
    E, J, H, ok = np.empty(g), np.empty((g, 3)), np.empty((g, 9)), np.empty(g, dtype=np.uint8)
    lib().ora_amips_ring_ejh(_p(V, _dp), _p(tets, _i32p), _p(tid, _i32p), _p(off, _u64p), _p(center, _i32p),
                             C.c_uint64(g), _p(E, _dp), _p(J, _dp), _p(H, _dp), _p(ok, _u8p), C.c_int(threads))
    return E, J, H, ok


def amips_ring_energy(V, tets, group_off, t_ids=None, threads=1):
    V, tets = _f64(V), _i32(tets)
    off = np.ascontiguousarray(group_off, dtype=np.uint64)
    tid = _i32(t_ids) if t_ids is not None else None
    g = off.shape[0] - 1
    E = np.empty(g)
    lib().ora_amips_ring_energy(_p(V, _dp), _p(tets, _i32p), _p(tid, _i32p), _p(off, _u64p), C.c_uint64(g),
                                _p(E, _dp), C.c_int(threads))
    return E


def tet_dihedral(V, tets, threads=1):
    """calTetQuality_AD (LocalOperations.cpp:783-860): (min_d_angle, max_d_angle) per tet"""
    V, tets = _f64(V), _i32(tets)
    n = tets.shape[0]
    lo, hi = np.empty(n), np.empty(n)
    lib().ora_tet_dihedral(_p(V, _dp), _p(tets, _i32p), C.c_uint64(n), _p(lo, _dp), _p(hi, _dp), C.c_int(threads))
    return lo, hi


# ---------------------------------------------------------------------------------------------- envelope
def point_triangle_sqdist(p, v0, v1, v2):
    a = [_f64(x) for x in (p, v0, v1, v2)]
    near = np.empty(3)
    d = lib().ora_point_triangle_sqdist(*[_p(x, _dp) for x in a], _p(near, _dp))
    return float(d), near


class Surface:
    """MeshFacetsAABBWithEps restated (src/tetwild/geogram/mesh_AABB.cpp). Facet ids are in the caller's numbering."""

    def __init__(self, V, F, order=1):
        self.V, self.F = _f64(V), _u32(F)
        self.h = C.c_void_p(lib().ora_surface_create(_p(self.V, _dp), C.c_uint32(len(self.V)), _p(self.F, _u32p),
                                                     C.c_uint32(len(self.F)), C.c_int(order)))
        if not self.h:
            raise ValueError("empty surface")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_surface_destroy(self.h)
            self.h = None

    def order(self):
        out = np.empty(len(self.F), dtype=np.uint32)
        lib().ora_surface_get_order(self.h, _p(out, _u32p))
        return out

    def nearest(self, P, threads=1):
        P = _f64(P)
        n = len(P)
        f, q, d = np.empty(n, dtype=np.uint32), np.empty((n, 3)), np.empty(n)
        lib().ora_nearest(self.h, _p(P, _dp), C.c_uint64(n), _p(f, _u32p), _p(q, _dp), _p(d, _dp), C.c_int(threads))
        return f, q, d

    def sqdist_brute(self, P, threads=1):
        P = _f64(P)
        n = len(P)
        d, f = np.empty(n), np.empty(n, dtype=np.uint32)
        lib().ora_point_sqdist_brute(self.h, _p(P, _dp), C.c_uint64(n), _p(d, _dp), _p(f, _u32p), C.c_int(threads))
        return d, f

    def points_out(self, P, eps2, threads=1, brute=False):
        P = _f64(P)
        n = len(P)
        out = np.empty(n, dtype=np.uint8)
        fn = lib().ora_envelope_points_out_brute if brute else lib().ora_envelope_points_out
        fn(self.h, _p(P, _dp), C.c_uint64(n), C.c_double(eps2), _p(out, _u8p), C.c_int(threads))
        return out

    def faces_out(self, tris, sampling_dist, eps2, threads=1, degenerate_shortcut=True):
        """degenerate_shortcut=True: isFaceOutEnvelop_sampling (LocalOperations.cpp:1046-1109); False: the per-face body of
        Preprocess::isOutEnvelop (Preprocess.cpp:652-739)"""
        T = _f64(tris).reshape(-1, 9)
        n = len(T)
        out, ns = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint64)
        lib().ora_envelope_faces_out_ex(self.h, _p(T, _dp), C.c_uint64(n), C.c_double(sampling_dist), C.c_double(eps2),
                                        C.c_int(1 if degenerate_shortcut else 0), _p(out, _u8p), _p(ns, _u64p), C.c_int(threads))
        return out, ns


def sample_triangle(tri, sampling_dist):
    t = _f64(tri).reshape(9)
    n = lib().ora_sample_triangle(_p(t, _dp), C.c_double(sampling_dist), None, C.c_uint64(0))
    out = np.empty((n, 3))
    lib().ora_sample_triangle(_p(t, _dp), C.c_double(sampling_dist), _p(out, _dp), C.c_uint64(n))
    return out


# ---------------------------------------------------------------------------------------------- winding
def winding_direct(V, F, Q, threads=1):
    V, F, Q = _f64(V), _u32(F), _f64(Q)
    W = np.empty(len(Q))
    lib().ora_winding_direct(_p(V, _dp), C.c_uint32(len(V)), _p(F, _u32p), C.c_uint32(len(F)), _p(Q, _dp),
                             C.c_uint64(len(Q)), _p(W, _dp), C.c_int(threads))
    return W


class WindingTree:
    def __init__(self, V, F):
        self.V, self.F = _f64(V), _u32(F)
        self.h = C.c_void_p(lib().ora_wtree_create(_p(self.V, _dp), C.c_uint32(len(self.V)), _p(self.F, _u32p),
                                                   C.c_uint32(len(self.F))))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_wtree_destroy(self.h)
            self.h = None

    def eval(self, Q, threads=1):
        Q = _f64(Q)
        W = np.empty(len(Q))
        lib().ora_wtree_eval(self.h, _p(Q, _dp), C.c_uint64(len(Q)), _p(W, _dp), C.c_int(threads))
        return W

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().ora_wtree_stats(self.h, C.byref(a), C.byref(b))
        return {"nodes": a.value, "cap_faces": b.value}


def inout_filter(V, F, Q, hierarchical=False, threads=1):
    V, F, Q = _f64(V), _u32(F), _f64(Q)
    keep, W = np.empty(len(Q), dtype=np.uint8), np.empty(len(Q))
    r = lib().ora_inout_filter(_p(V, _dp), C.c_uint32(len(V)), _p(F, _u32p), C.c_uint32(len(F)), _p(Q, _dp),
                               C.c_uint64(len(Q)), _p(keep, _u8p), _p(W, _dp), C.c_int(int(hierarchical)),
                               C.c_int(threads))
    return keep, W, bool(r)


# ---------------------------------------------------------------------------------------------- the real reference
_ref = None


def ref_available():
    return os.path.exists(_REF)


def ref():
    """oracle/_ref/libtetwild_ref.so: pieces of the UNMODIFIED reference (see oracle/ref_wrap.cpp)."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF):
            raise RuntimeError("oracle/_ref/libtetwild_ref.so not built (needs /root/reference; run oracle/ref_build.sh)")
        R = C.CDLL(_REF)
        R.ref_amips_energy.restype = C.c_double
        R.ref_sample_triangle.restype = C.c_uint64
        R.ref_tree_create.restype = C.c_void_p
        _ref = R
    return _ref


def ref_amips_ejh_soa(T, threads=1, want=(True, True, True)):
    T, ptrs = _soa_ptrs(T)
    n = T.shape[1]
    E = np.empty(n) if want[0] else None
    J = np.empty((n, 3)) if want[1] else None
    H = np.empty((n, 9)) if want[2] else None
    ref().ref_amips_ejh_soa(ptrs, _p(E, _dp), _p(J, _dp), _p(H, _dp), C.c_uint64(n), C.c_int(threads))
    return E, J, H


_refq = None


def quad_available():
    return os.path.exists(_REFQ)


def _quad(fn, T, threads):
    """oracle/_ref/libtetwild_ref_quad.so (oracle/ref_quad.cpp): binary128 evaluations, results rounded to double"""
    global _refq
    if _refq is None:
        if not os.path.exists(_REFQ):
            raise RuntimeError("oracle/_ref/libtetwild_ref_quad.so not built (needs /root/reference; run oracle/ref_build.sh)")
        _refq = C.CDLL(_REFQ)
    T, ptrs = _soa_ptrs(T)
    n = T.shape[1]
    E, J, H = np.empty(n), np.empty((n, 3)), np.empty((n, 9))
    getattr(_refq, fn)(ptrs, _p(E, _dp), _p(J, _dp), _p(H, _dp), C.c_uint64(n), C.c_int(threads))
    return E, J, H


def refq_amips_ejh_soa(T, threads=1):
    """the reference's own text (LocalOperations.cpp:28-291) evaluated in IEEE binary128: the exact value of the
    reference's expression as written, for the same double inputs"""
    return _quad("refq_amips_ejh_soa", T, threads)


def exactq_amips_ejh_soa(T, threads=1):
    """the mathematical conformal AMIPS energy / gradient / Hessian (exact constants) in binary128"""
    return _quad("exactq_amips_ejh_soa", T, threads)


def ref_amips_ring_ejh(V, tets, group_off, center, t_ids=None, threads=1):
    """NewtonsUpdate over many one-rings through the reference's own E / J / H text (oracle/ref_wrap.cpp)"""
    V, tets = _f64(V), _i32(tets)
    off = np.ascontiguousarray(group_off, dtype=np.uint64)
    center = _i32(center)
    tid = _i32(t_ids) if t_ids is not None else None
    g = center.shape[0]
    E, J, H, ok = np.empty(g), np.empty((g, 3)), np.empty((g, 9)), np.empty(g, dtype=np.uint8)
    ref().ref_amips_ring_ejh(_p(V, _dp), _p(tets, _i32p), _p(tid, _i32p), _p(off, _u64p), _p(center, _i32p), C.c_uint64(g),
                             _p(E, _dp), _p(J, _dp), _p(H, _dp), _p(ok, _u8p), C.c_int(threads))
    return E, J, H, ok


def ref_sample_triangle(tri, sampling_dist):
    t = _f64(tri).reshape(9)
    n = ref().ref_sample_triangle(_p(t, _dp), C.c_double(sampling_dist), None, C.c_uint64(0))
    out = np.empty((n, 3))
    ref().ref_sample_triangle(_p(t, _dp), C.c_double(sampling_dist), _p(out, _dp), C.c_uint64(n))
    return out


class RefTree:
    """The reference's GEO::MeshFacetsAABBWithEps itself (mesh_AABB.cpp compiled unmodified), reorder=false."""

    def __init__(self, V, F_sorted):
        self.V, self.F = _f64(V), _u32(F_sorted)
        self.h = C.c_void_p(ref().ref_tree_create(_p(self.V, _dp), C.c_uint32(len(self.V)), _p(self.F, _u32p),
                                                  C.c_uint32(len(self.F))))

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_tree_destroy(self.h)
            self.h = None

    def nearest(self, P, threads=1):
        P = _f64(P)
        n = len(P)
        f, q, d = np.empty(n, dtype=np.uint32), np.empty((n, 3)), np.empty(n)
        ref().ref_tree_nearest(self.h, _p(P, _dp), C.c_uint64(n), _p(f, _u32p), _p(q, _dp), _p(d, _dp), C.c_int(threads))
        return f, q, d

    def faces_out(self, tris, sampling_dist, eps2, threads=1, degenerate_shortcut=True):
        """isFaceOutEnvelop_sampling around the reference's own sampleTriangle / DistanceQuery.h / tree (ref_wrap.cpp)"""
        T = _f64(tris).reshape(-1, 9)
        n = len(T)
        out, ns = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint64)
        ref().ref_tree_faces_out(self.h, _p(T, _dp), C.c_uint64(n), C.c_double(sampling_dist), C.c_double(eps2),
                                 C.c_int(1 if degenerate_shortcut else 0), _p(out, _u8p), _p(ns, _u64p), C.c_int(threads))
        return out, ns

    def points_out(self, P, eps2, threads=1):
        P = _f64(P)
        n = len(P)
        out, f, d = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint32), np.empty(n)
        ref().ref_tree_envelope_points_out(self.h, _p(P, _dp), C.c_uint64(n), C.c_double(eps2), _p(out, _u8p),
                                           _p(f, _u32p), _p(d, _dp), C.c_int(threads))
        return out, f, d
