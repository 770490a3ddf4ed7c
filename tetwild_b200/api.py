"""ctypes binding of include/tetwild_gpu.h.

Host-buffer methods take/return numpy arrays (copies are part of the call, like the C ABI); methods ending in `_dev`
take raw device pointers (ints, e.g. torch.Tensor.data_ptr()) and a CUDA stream handle and are asynchronous.
Names follow the reference: Surface ~ GEO::MeshFacetsAABBWithEps, Context.amips_* ~ LocalOperations /
VertexSmoother members, Winding ~ igl::winding_number.
"""
import ctypes as C
import os

import numpy as np

MAX_ENERGY = 1e50
NO_FACET = 0xFFFFFFFF

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtetwild_gpu.so")
_lib = None

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class TetWildGPUError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load_library():
    """Load libtetwild_gpu.so. Fails loudly when the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise TetWildGPUError(
            "libtetwild_gpu.so is missing at %s: build it with `python -m tetwild_b200.build` "
            "(there is no CPU fallback)" % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    L.twg_last_error.restype = C.c_char_p
    L.twg_version.restype = C.c_char_p
    L.twg_launch_count.restype = C.c_uint64
    L.twg_surface_num_facets.restype = C.c_uint32
    L.twg_last_error.argtypes = [_vp]
    L.twg_launch_count.argtypes = [_vp]
    L.twg_destroy.argtypes = [_vp]
    L.twg_surface_destroy.argtypes = [_vp]
    L.twg_winding_destroy.argtypes = [_vp]
    L.twg_destroy.restype = None
    L.twg_surface_destroy.restype = None
    L.twg_winding_destroy.restype = None
    L.twg_mesh_destroy.argtypes = [_vp]
    L.twg_mesh_destroy.restype = None
    L.twg_mesh_num_vertices.argtypes = [_vp]
    L.twg_mesh_num_vertices.restype = C.c_uint32
    L.twg_mesh_num_tets.argtypes = [_vp]
    L.twg_mesh_num_tets.restype = C.c_uint64
    L.twg_mesh_vertices_dev.argtypes = [_vp]
    L.twg_mesh_vertices_dev.restype = C.c_void_p
    L.twg_mesh_tets_dev.argtypes = [_vp]
    L.twg_mesh_tets_dev.restype = C.c_void_p
    L.twg_device_context.argtypes = [_vp, C.c_int]
    L.twg_device_context.restype = C.c_void_p
    L.twg_surface_replica.argtypes = [_vp, C.c_int]
    L.twg_surface_replica.restype = C.c_void_p
    L.twg_winding_replica.argtypes = [_vp, C.c_int]
    L.twg_winding_replica.restype = C.c_void_p
    L.twg_num_devices.argtypes = [_vp]
    L.twg_set_option.argtypes = [_vp, C.c_char_p, C.c_double]
    L.twg_get_option.argtypes = [_vp, C.c_char_p, C.POINTER(C.c_double)]
    L.twg_debug_counter.argtypes = [_vp, C.c_int, C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _dev(p):
    return C.c_void_p(int(p) if p else 0)


class Context:
    """twg_ctx: one device (device = int) or several devices of this process (device = sequence of ints, twg_create_multi:
    handles are replicated on every device and host-buffer batches are split by index range)."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self._L.twg_create_multi(C.byref(h), ids, C.c_int(len(device)))
        else:
            rc = self._L.twg_create(C.byref(h), C.c_int(device))
        if rc != 0:
            raise TetWildGPUError("twg_create(device=%r) failed with code %d (no sm_100-class GPU? there is no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device

    @property
    def num_devices(self):
        return int(self._L.twg_num_devices(self.h))

    def set_option(self, name, value):
        self._check(self._L.twg_set_option(self.h, name.encode(), C.c_double(value)))

    def get_option(self, name):
        v = C.c_double(0)
        self._check(self._L.twg_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def debug_counter(self, which=0):
        """0: envelope queries that overflowed the traversal stack and were re-decided by the exact descent"""
        v = C.c_uint64(0)
        self._check(self._L.twg_debug_counter(self.h, C.c_int(which), C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "h", None):
            self._L.twg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise TetWildGPUError("libtetwild_gpu error %d: %s" % (rc, self._L.twg_last_error(self.h).decode()))

    @property
    def launches(self):
        return int(self._L.twg_launch_count(self.h))

    def synchronize(self):
        self._check(self._L.twg_synchronize(self.h))

    def debug_sort_points(self, P, box=None, want_keys=True, want_sorted=True):
        """the library's own Morton radix sort (csrc/qsort.cu) -> (perm, sorted keys or None, sorted points or None)"""
        P = _f64(P)
        n = len(P)
        perm = np.empty(n, dtype=np.uint32)
        keys = np.empty(n, dtype=np.uint32) if want_keys else None
        srt = np.empty((n, 3)) if want_sorted else None
        b = _f64(box).reshape(6) if box is not None else None
        self._check(self._L.twg_debug_sort_points(self.h, _ptr(P), C.c_uint64(n), _ptr(b), _ptr(perm), _ptr(keys), _ptr(srt)))
        return perm, keys, srt

    # ---- roofline denominators (microbenchmarks, not on the hot path) ----
    def measure_fp64_tflops(self):
        v = C.c_double(0)
        self._check(self._L.twg_measure_fp64_tflops(self.h, C.byref(v)))
        return v.value

    def measure_fp64_tflops_distinct(self):
        v = C.c_double(0)
        self._check(self._L.twg_measure_fp64_tflops_distinct(self.h, C.byref(v)))
        return v.value

    def measure_copy_gbs(self, nbytes=1 << 30):
        v = C.c_double(0)
        self._check(self._L.twg_measure_copy_gbs(self.h, C.c_uint64(nbytes), C.byref(v)))
        return v.value

    # ---- AMIPS ----
    @staticmethod
    def _soa_ptrs(T):
        T = _f64(T)
        assert T.ndim == 2 and T.shape[0] == 12, "T must be (12, n): row 3*i+k = coordinate k of vertex i"
        return T, (C.c_void_p * 12)(*[T[k].ctypes.data for k in range(12)])

    def amips_energy_soa(self, T):
        """energy_ispc(V1_x..V4_z, E, count) (src/ispc/energy.ispc:7-21)"""
        T, ptrs = self._soa_ptrs(T)
        n = T.shape[1]
        E = np.empty(n)
        self._check(self._L.twg_amips_energy_soa(self.h, ptrs, _ptr(E), C.c_uint64(n)))
        return E

    def amips_ejh_soa(self, T, want=(True, True, True), out=None):
        """comformalAMIPS{Energy,Jacobian,Hessian}_new per tet; `out` = optional preallocated (E, J3, H9) (e.g. pinned)"""
        T, ptrs = self._soa_ptrs(T)
        n = T.shape[1]
        if out is not None:
            E, J, H = out
        else:
            E = np.empty(n) if want[0] else None
            J = np.empty((n, 3)) if want[1] else None
            H = np.empty((n, 9)) if want[2] else None
        self._check(self._L.twg_amips_ejh_soa(self.h, ptrs, _ptr(E), _ptr(J), _ptr(H), C.c_uint64(n)))
        return E, J, H

    def amips_ejh_soa_dev(self, dT12, dE, dJ3, dH9, n, stream=0):
        ptrs = (C.c_void_p * 12)(*[int(p) for p in dT12])
        self._check(self._L.twg_amips_ejh_soa_dev(self.h, ptrs, _dev(dE), _dev(dJ3), _dev(dH9), C.c_uint64(n), _dev(stream)))

    def amips_energy_soa_dev(self, dT12, dE, n, stream=0):
        ptrs = (C.c_void_p * 12)(*[int(p) for p in dT12])
        self._check(self._L.twg_amips_energy_soa_dev(self.h, ptrs, _dev(dE), C.c_uint64(n), _dev(stream)))

    def amips_quality(self, V, tets):
        """calTetQualities (LocalOperations.cpp:695-773): slim_energy per tet"""
        V = _f64(V)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        out = np.empty(len(tets))
        self._check(self._L.twg_amips_quality(self.h, _ptr(V), C.c_uint32(len(V)), _ptr(tets), C.c_uint64(len(tets)), _ptr(out)))
        return out

    def amips_quality_dev(self, dV, nV, dTets, nT, dSlim, stream=0):
        self._check(self._L.twg_amips_quality_dev(self.h, _dev(dV), C.c_uint32(nV), _dev(dTets), C.c_uint64(nT), _dev(dSlim), _dev(stream)))

    def amips_ring_ejh(self, V, tets, group_off, center, t_ids=None):
        """VertexSmoother::NewtonsUpdate for many one-rings (VertexSmoother.cpp:627-702)"""
        V = _f64(V)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        off = np.ascontiguousarray(group_off, dtype=np.uint64)
        center = np.ascontiguousarray(center, dtype=np.int32)
        tid = np.ascontiguousarray(t_ids, dtype=np.int32) if t_ids is not None else None
        g = len(center)
        E, J, H, ok = np.empty(g), np.empty((g, 3)), np.empty((g, 9)), np.empty(g, dtype=np.uint8)
        self._check(self._L.twg_amips_ring_ejh(self.h, _ptr(V), C.c_uint32(len(V)), _ptr(tets), C.c_uint64(len(tets)), _ptr(tid),
                                               _ptr(off), _ptr(center), C.c_uint64(g), _ptr(E), _ptr(J), _ptr(H), _ptr(ok)))
        return E, J, H, ok

    def amips_ring_ejh_dev(self, dV, nV, dTets, nT, dTids, dOff, dCenter, nG, dE, dJ3, dH9, dOk, stream=0):
        self._check(self._L.twg_amips_ring_ejh_dev(self.h, _dev(dV), C.c_uint32(nV), _dev(dTets), C.c_uint64(nT), _dev(dTids), _dev(dOff),
                                                   _dev(dCenter), C.c_uint64(nG), _dev(dE), _dev(dJ3), _dev(dH9), _dev(dOk), _dev(stream)))

    def amips_ring_energy(self, V, tets, group_off, t_ids=None):
        """VertexSmoother::getNewEnergy for many one-rings (VertexSmoother.cpp:544-625)"""
        V = _f64(V)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        off = np.ascontiguousarray(group_off, dtype=np.uint64)
        tid = np.ascontiguousarray(t_ids, dtype=np.int32) if t_ids is not None else None
        g = len(off) - 1
        E = np.empty(g)
        self._check(self._L.twg_amips_ring_energy(self.h, _ptr(V), C.c_uint32(len(V)), _ptr(tets), C.c_uint64(len(tets)), _ptr(tid),
                                                  _ptr(off), C.c_uint64(g), _ptr(E)))
        return E

    # ---- sampling (debug / parity) ----
    def sample_triangle(self, tri, sampling_dist):
        t = _f64(tri).reshape(9)
        cnt = C.c_uint64(0)
        self._check(self._L.twg_sample_triangle(self.h, _ptr(t), C.c_double(sampling_dist), C.c_void_p(0), C.c_uint64(0), C.byref(cnt)))
        out = np.empty((cnt.value, 3))
        if cnt.value:
            self._check(self._L.twg_sample_triangle(self.h, _ptr(t), C.c_double(sampling_dist), _ptr(out), C.c_uint64(cnt.value), C.byref(cnt)))
        return out

    # ---- winding one-shots ----
    def winding_number(self, V, F, Q, want_w=True, want_keep=True):
        """igl::winding_number(V,F,O,W) (InoutFiltering.cpp:45)"""
        V, Q = _f64(V), _f64(Q)
        F = np.ascontiguousarray(F, dtype=np.uint32)
        W = np.empty(len(Q)) if want_w else None
        keep = np.empty(len(Q), dtype=np.uint8) if want_keep else None
        self._check(self._L.twg_winding_number(self.h, _ptr(V), C.c_uint32(len(V)), _ptr(F), C.c_uint32(len(F)), _ptr(Q),
                                               C.c_uint64(len(Q)), _ptr(W), _ptr(keep)))
        return W, keep

    def inout_filter(self, V, F, Q):
        """InoutFiltering::filter decision incl. flip-and-retry (InoutFiltering.cpp:45-75) -> (keep, retried)"""
        V, Q = _f64(V), _f64(Q)
        F = np.ascontiguousarray(F, dtype=np.uint32)
        keep = np.empty(len(Q), dtype=np.uint8)
        r = C.c_int(0)
        self._check(self._L.twg_inout_filter(self.h, _ptr(V), C.c_uint32(len(V)), _ptr(F), C.c_uint32(len(F)), _ptr(Q),
                                             C.c_uint64(len(Q)), _ptr(keep), C.byref(r)))
        return keep, bool(r.value)


class Surface:
    """twg_surface ~ GEO::MeshFacetsAABBWithEps (src/tetwild/geogram/mesh_AABB.h:64-421), built on the device."""

    def __init__(self, ctx, V, F):
        self.ctx = ctx
        self._L = ctx._L
        V = _f64(V)
        F = np.ascontiguousarray(F, dtype=np.uint32)
        h = C.c_void_p()
        ctx._check(self._L.twg_surface_create(ctx.h, _ptr(V), C.c_uint32(len(V)), _ptr(F), C.c_uint32(len(F)), C.byref(h)))
        self.h = h
        self.num_facets = len(F)

    def close(self):
        if getattr(self, "h", None):
            self._L.twg_surface_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def points_out(self, P, eps2, out=None):
        """out[i] = isPointOutEnvelop(P[i]) (LocalOperations.cpp:1034-1044)"""
        P = _f64(P)
        if out is None:
            out = np.empty(len(P), dtype=np.uint8)
        self.ctx._check(self._L.twg_envelope_points_out(self.h, _ptr(P), C.c_uint64(len(P)), C.c_double(eps2), _ptr(out)))
        return out

    def points_out_dev(self, dP, n, eps2, dOut, stream=0):
        self.ctx._check(self._L.twg_envelope_points_out_dev(self.h, _dev(dP), C.c_uint64(n), C.c_double(eps2), _dev(dOut), _dev(stream)))

    def faces_out(self, tris, sampling_dist, eps2, degenerate_shortcut=True):
        """out[i] = isFaceOutEnvelop(tris[i]) (LocalOperations.cpp:967-976, :1046-1109); degenerate_shortcut=False: the
        per-face body of Preprocess::isOutEnvelop (Preprocess.cpp:643-747), which samples degenerate faces too"""
        T = _f64(tris).reshape(-1, 9)
        out = np.empty(len(T), dtype=np.uint8)
        self.ctx._check(self._L.twg_envelope_faces_out_ex(self.h, _ptr(T), C.c_uint64(len(T)), C.c_double(sampling_dist), C.c_double(eps2),
                                                          C.c_uint32(0 if degenerate_shortcut else 1), _ptr(out)))
        return out

    def faces_out_dev(self, dTris, n, sampling_dist, eps2, dOut, stream=0):
        self.ctx._check(self._L.twg_envelope_faces_out_dev(self.h, _dev(dTris), C.c_uint64(n), C.c_double(sampling_dist), C.c_double(eps2),
                                                           _dev(dOut), _dev(stream)))

    def nearest(self, P, out=None):
        """nearest_facet (mesh_AABB.h:130-141) -> (facet ids in the caller's numbering, nearest points, d2);
        `out` = optional preallocated (facet u32[n], nearest f64[n,3], d2 f64[n]) (e.g. pinned)"""
        P = _f64(P)
        n = len(P)
        f, q, d = out if out is not None else (np.empty(n, dtype=np.uint32), np.empty((n, 3)), np.empty(n))
        self.ctx._check(self._L.twg_nearest(self.h, _ptr(P), C.c_uint64(n), _ptr(f), _ptr(q), _ptr(d)))
        return f, q, d

    def nearest_dev(self, dP, n, dFacet, dNearest, dD2, stream=0):
        self.ctx._check(self._L.twg_nearest_dev(self.h, _dev(dP), C.c_uint64(n), _dev(dFacet), _dev(dNearest), _dev(dD2), _dev(stream)))

    def squared_distance(self, P):
        """squared_distance (mesh_AABB.h:221-226)"""
        P = _f64(P)
        d = np.empty(len(P))
        self.ctx._check(self._L.twg_nearest(self.h, _ptr(P), C.c_uint64(len(P)), C.c_void_p(0), C.c_void_p(0), _ptr(d)))
        return d


class TetMesh:
    """twg_mesh: device-resident mirror of the scheduler's tet mesh -- `tet_vertices[].posf`, `tets`, `t_is_removed`,
    `tet_vertices[].conn_tets` (src/tetwild/LocalOperations.h:35-45). A removed tet has a negative first index."""

    def __init__(self, ctx, V, tets):
        self.ctx = ctx
        self._L = ctx._L
        V = _f64(V)
        tets = np.ascontiguousarray(tets, dtype=np.int32).reshape(-1, 4)
        h = C.c_void_p()
        ctx._check(self._L.twg_mesh_create(ctx.h, _ptr(V), C.c_uint32(len(V)), _ptr(tets), C.c_uint64(len(tets)), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.twg_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_vertices(self):
        return int(self._L.twg_mesh_num_vertices(self.h))

    @property
    def num_tets(self):
        return int(self._L.twg_mesh_num_tets(self.h))

    @property
    def vertices_dev(self):
        return int(self._L.twg_mesh_vertices_dev(self.h) or 0)

    @property
    def tets_dev(self):
        return int(self._L.twg_mesh_tets_dev(self.h) or 0)

    def resize(self, nV, nT):
        self.ctx._check(self._L.twg_mesh_resize(self.h, C.c_uint32(nV), C.c_uint64(nT)))

    def set_vertices(self, v_ids, xyz):
        """posf[v_ids[i]] = xyz[i] (after an accepted smoothing / split / collapse)"""
        ids = np.ascontiguousarray(v_ids, dtype=np.int32)
        xyz = _f64(xyz).reshape(-1, 3)
        assert len(ids) == len(xyz)
        self.ctx._check(self._L.twg_mesh_set_vertices(self.h, _ptr(ids), _ptr(xyz), C.c_uint64(len(ids))))

    def set_tets(self, t_ids, tets):
        """tets[t_ids[i]] = tets[i]; a negative first index marks t_is_removed"""
        ids = np.ascontiguousarray(t_ids, dtype=np.int32)
        tets = np.ascontiguousarray(tets, dtype=np.int32).reshape(-1, 4)
        assert len(ids) == len(tets)
        self.ctx._check(self._L.twg_mesh_set_tets(self.h, _ptr(ids), _ptr(tets), C.c_uint64(len(ids))))

    def get_vertices(self):
        out = np.empty((self.num_vertices, 3))
        self.ctx._check(self._L.twg_mesh_get_vertices(self.h, _ptr(out)))
        return out

    def build_rings(self):
        self.ctx._check(self._L.twg_mesh_build_rings(self.h))

    def get_rings(self):
        """conn_tets as CSR: (off[nV+1], tets[off[nV]]), tets ascending within a vertex"""
        off = np.empty(self.num_vertices + 1, dtype=np.uint64)
        self.ctx._check(self._L.twg_mesh_get_rings(self.h, _ptr(off), C.c_void_p(0)))
        t = np.empty(int(off[-1]), dtype=np.int32)
        if len(t):
            self.ctx._check(self._L.twg_mesh_get_rings(self.h, _ptr(off), _ptr(t)))
        return off, t

    def _tids(self, t_ids):
        if t_ids is None:
            return None, self.num_tets
        ids = np.ascontiguousarray(t_ids, dtype=np.int32)
        return ids, len(ids)

    def quality(self, t_ids=None, out=None):
        """calTetQuality_AMIPS over the resident mesh (LocalOperations.cpp:862-884); t_ids None = every slot;
        `out` = optional preallocated result (e.g. pinned)"""
        ids, n = self._tids(t_ids)
        if out is None:
            out = np.empty(n)
        self.ctx._check(self._L.twg_mesh_quality(self.h, _ptr(ids), C.c_uint64(n), _ptr(out)))
        return out

    def quality_dev(self, dTids, n, dSlim, stream=0):
        self.ctx._check(self._L.twg_mesh_quality_dev(self.h, _dev(dTids), C.c_uint64(n), _dev(dSlim), _dev(stream)))

    def dihedral(self, t_ids=None):
        """calTetQuality_AD (LocalOperations.cpp:783-860) -> (min_d_angle, max_d_angle)"""
        ids, n = self._tids(t_ids)
        lo, hi = np.empty(n), np.empty(n)
        self.ctx._check(self._L.twg_mesh_dihedral(self.h, _ptr(ids), C.c_uint64(n), _ptr(lo), _ptr(hi)))
        return lo, hi

    def dihedral_dev(self, dTids, n, dMin, dMax, stream=0):
        self.ctx._check(self._L.twg_mesh_dihedral_dev(self.h, _dev(dTids), C.c_uint64(n), _dev(dMin), _dev(dMax), _dev(stream)))

    def vertex_ring_ejh(self, v_ids, out=None):
        """NewtonsUpdate (VertexSmoother.cpp:627-702) for the one-rings conn_tets[v], v in v_ids"""
        ids = np.ascontiguousarray(v_ids, dtype=np.int32)
        g = len(ids)
        if out is not None:
            E, J, H, ok = out
        else:
            E, J, H, ok = np.empty(g), np.empty((g, 3)), np.empty((g, 9)), np.empty(g, dtype=np.uint8)
        self.ctx._check(self._L.twg_mesh_vertex_ring_ejh(self.h, _ptr(ids), C.c_uint64(g), _ptr(E), _ptr(J), _ptr(H), _ptr(ok)))
        return E, J, H, ok

    def vertex_ring_ejh_dev(self, dVids, n, dE, dJ3, dH9, dOk, stream=0):
        self.ctx._check(self._L.twg_mesh_vertex_ring_ejh_dev(self.h, _dev(dVids), C.c_uint64(n), _dev(dE), _dev(dJ3), _dev(dH9), _dev(dOk),
                                                             _dev(stream)))

    def ring_ejh(self, t_ids, group_off, center):
        tid = np.ascontiguousarray(t_ids, dtype=np.int32)
        off = np.ascontiguousarray(group_off, dtype=np.uint64)
        center = np.ascontiguousarray(center, dtype=np.int32)
        g = len(center)
        E, J, H, ok = np.empty(g), np.empty((g, 3)), np.empty((g, 9)), np.empty(g, dtype=np.uint8)
        self.ctx._check(self._L.twg_mesh_ring_ejh(self.h, _ptr(tid), _ptr(off), _ptr(center), C.c_uint64(g), _ptr(E), _ptr(J), _ptr(H), _ptr(ok)))
        return E, J, H, ok

    def vertex_trial_energy(self, v_ids, xyz):
        """getNewEnergy of the one-ring of v_ids[i] with that vertex at xyz[i] (the smoother's line search, VertexSmoother.cpp:505-541);
        the resident mesh is not modified"""
        ids = np.ascontiguousarray(v_ids, dtype=np.int32)
        xyz = _f64(xyz).reshape(-1, 3)
        assert len(ids) == len(xyz)
        E = np.empty(len(ids))
        self.ctx._check(self._L.twg_mesh_vertex_trial_energy(self.h, _ptr(ids), _ptr(xyz), C.c_uint64(len(ids)), _ptr(E)))
        return E

    def ring_energy(self, t_ids, group_off):
        tid = np.ascontiguousarray(t_ids, dtype=np.int32)
        off = np.ascontiguousarray(group_off, dtype=np.uint64)
        g = len(off) - 1
        E = np.empty(g)
        self.ctx._check(self._L.twg_mesh_ring_energy(self.h, _ptr(tid), _ptr(off), C.c_uint64(g), _ptr(E)))
        return E


class Winding:
    """twg_winding: winding-number hierarchy over a surface (igl::winding_number, InoutFiltering.cpp:45)."""

    def __init__(self, ctx, V, F):
        self.ctx = ctx
        self._L = ctx._L
        V = _f64(V)
        F = np.ascontiguousarray(F, dtype=np.uint32)
        h = C.c_void_p()
        ctx._check(self._L.twg_winding_create(ctx.h, _ptr(V), C.c_uint32(len(V)), _ptr(F), C.c_uint32(len(F)), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self._L.twg_winding_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, Q, want_w=True, want_keep=True, out=None):
        Q = _f64(Q)
        if out is not None:
            W, keep = out
        else:
            W = np.empty(len(Q)) if want_w else None
            keep = np.empty(len(Q), dtype=np.uint8) if want_keep else None
        self.ctx._check(self._L.twg_winding_eval(self.h, _ptr(Q), C.c_uint64(len(Q)), _ptr(W), _ptr(keep)))
        return W, keep

    def eval_dev(self, dQ, n, dW, dKeep, stream=0):
        self.ctx._check(self._L.twg_winding_eval_dev(self.h, _dev(dQ), C.c_uint64(n), _dev(dW), _dev(dKeep), _dev(stream)))

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._L.twg_winding_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"nodes": a.value, "cap_segments": b.value, "triangles": c.value}

    def download(self):
        """test hook: the hierarchy as it lives on the device -> (nodes as raw 64-byte records, caps [n,4], tris [nF+2,9])"""
        st = self.stats()
        nodes = np.empty((st["nodes"], 64), dtype=np.uint8)
        caps = np.empty((st["cap_segments"], 4))
        tris = np.empty((st["triangles"] + 2, 9))
        self.ctx._check(self._L.twg_debug_winding_download(self.h, _ptr(nodes), _ptr(caps), _ptr(tris)))
        return nodes, caps, tris
