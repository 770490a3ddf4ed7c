"""tests/golden/make_golden.py -- regenerates the committed golden vectors from the REFERENCE ITSELF.

Runs the pieces of the unmodified reference compiled into oracle/_ref/libtetwild_ref.so (oracle/ref_build.sh; needs
/root/reference, so this script only runs in the build container, never on the GPU box) on small seeded inputs and
stores inputs + outputs as JSON (hex floats, bit exact):

  amips_golden.json   comformalAMIPS{Energy,Jacobian,Hessian}_new       src/tetwild/LocalOperations.cpp:28-291
  sample_golden.json  sampleTriangle                                    src/tetwild/Common.cpp:143-255
  tree_golden.json    MeshFacetsAABBWithEps nearest_facet /             src/tetwild/geogram/mesh_AABB.cpp (whole file,
                      facet_in_envelope_with_hint                       over the geogram API shim of oracle/shim)
  faces_golden.json   isFaceOutEnvelop_sampling decisions               src/tetwild/LocalOperations.cpp:1046-1109 composed from
                      (+ the Preprocess::isOutEnvelop per-face body)    the reference's sampleTriangle, DistanceQuery.h and tree

usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tetwild_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    O.build()
    assert O.ref_available(), "oracle/_ref not built"
    # ---- AMIPS: the two known-answer tets of SURVEY.md 8c + seeded random tets
    reg = [0, 0, 0, 1, 0, 0, 0.5, math.sqrt(3) / 2, 0, 0.5, math.sqrt(3) / 6, math.sqrt(6) / 3]
    corner = [0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1]
    T = np.concatenate([np.array([reg, corner]).T, synth.random_tets(62, seed=101)], axis=1)
    E, J, H = O.ref_amips_ejh_soa(T)
    json.dump({"source": "reference LocalOperations.cpp:28-291 via oracle/_ref", "T_rows_12xn": hx(T), "n": T.shape[1],
               "E": hx(E), "J": hx(J), "H": hx(H)}, open(os.path.join(HERE, "amips_golden.json"), "w"))
    # ---- sampleTriangle
    rng = np.random.default_rng(202)
    cases = []
    tris = [np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]]) * 0.05, np.array([[0, 0, 0], [1e-4, 0, 0], [0, 1e-4, 0.0]]),
            np.array([[0, 0, 0], [0.03, 0, 0], [0.015, 1e-4, 0.0]])]
    tris += [rng.normal(size=(3, 3)) * s for s in (0.004, 0.01, 0.02, 0.05, 0.05, 0.08) for _ in range(3)]
    for t in tris:
        for sd in (1e-3, 2.5e-3):
            ps = O.ref_sample_triangle(t, sd)
            cases.append({"tri": hx(t), "sd": float(sd).hex(), "count": int(len(ps)),
                          "sha256": hashlib.sha256(np.ascontiguousarray(ps).tobytes()).hexdigest(),
                          "first": hx(ps[0]), "last": hx(ps[-1])})
    json.dump({"source": "reference Common.cpp:143-255 via oracle/_ref", "cases": cases}, open(os.path.join(HERE, "sample_golden.json"), "w"))
    # ---- tree: small torus knot, facets pre-sorted with the oracle's Morton order, reference tree built with reorder=false
    V, F = synth.torus_knot(60, 12)
    S = O.Surface(V, F)
    order = S.order()
    RT = O.RefTree(V, F[order])
    sd, eps, eps2 = synth.state_eps(4e-3)
    P = synth.envelope_points(V, F, 600, eps, seed=303)
    f, q, d = RT.nearest(P)
    out, fe, de = RT.points_out(P, eps2)
    json.dump({"source": "reference mesh_AABB.cpp via oracle/_ref (geogram API shim; leaf distance = oracle restatement)",
               "surface": "synth.torus_knot(60, 12)", "eps2": float(eps2).hex(), "P": hx(P),
               "nearest_facet_original_ids": [int(x) for x in order[f]], "nearest_d2": hx(d), "nearest_pt": hx(q),
               "out": [int(x) for x in out]}, open(os.path.join(HERE, "tree_golden.json"), "w"))
    # ---- isFaceOutEnvelop_sampling (LocalOperations.cpp:1046-1109) and the Preprocess variant (Preprocess.cpp:652-739) composed
    # from the reference's own sampleTriangle, DistanceQuery.h and tree (oracle/ref_wrap.cpp::ref_tree_faces_out)
    T = synth.face_queries(V, F, 160, 0.01, eps, seed=404)                 # ~40 % out, ~64 samples each at sd / 4
    T[::20] = np.array([0, 0, 5, 1, 1, 6, 2, 2, 7.0])                      # exactly collinear
    sd = sd / 4
    o1, n1 = RT.faces_out(T, sd, eps2)
    o2, n2 = RT.faces_out(T, sd, eps2 * 0.64, degenerate_shortcut=False)
    json.dump({"source": "reference sampleTriangle + DistanceQuery.h + mesh_AABB.cpp via oracle/_ref, loop of LocalOperations.cpp:1046-1109",
               "surface": "synth.torus_knot(60, 12)", "sd": float(sd).hex(), "eps2": float(eps2).hex(), "tris": hx(T),
               "out": [int(x) for x in o1], "num_samples": [int(x) for x in n1],
               "preprocess_eps2": float(eps2 * 0.64).hex(), "preprocess_out": [int(x) for x in o2], "preprocess_num_samples": [int(x) for x in n2]},
              open(os.path.join(HERE, "faces_golden.json"), "w"))
    for fn in ("amips_golden.json", "sample_golden.json", "tree_golden.json", "faces_golden.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
