#!/usr/bin/env python
"""bench.py -- TetWild hot path on B200: envelope points/s (headline, BASELINE.json configs[1]), plus AMIPS
tet-evals/s (configs[2]) and winding queries/s (configs[3]) as `parts` of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--parts envelope,amips,winding] [--scale S]

A step = one pass of one part of the hot path over one synthetic batch:
  envelope  C2  10 M sampled points vs the 200 000-triangle torus knot, eps_rel = 1e-3 through State.cpp:36-41
  amips     C3  50 M random non-degenerate tets, flat SoA, E + J + H (FP64)
  winding   C4  100 M centroids vs the 1.0 M-triangle closed noisy sphere, keep = W > 0.5
`value` = units/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` = the same
metric through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H inside the timed region).
N > 1 (torchrun, one rank per GPU): the surface is replicated, every rank processes its own full-size batch
(weak scaling) and the 1-byte decisions are all-gathered over NCCL inside the timed region (double-buffered: the gather
of step k overlaps the kernels of step k+1; the region ends only after the last gather).
`--impl reference`: the reference's own CPU path (oracle/_ref where the reference compiles here, else the oracle
port) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FULL = {"envelope": 10_000_000, "amips": 50_000_000, "amips_literal": 50_000_000, "amips_ring": 50_000_000, "winding": 100_000_000, "winding_oneshot": 100_000_000,
        "envelope_faces": 400_000, "envelope_faces_c1": 100_000, "nearest": 10_000_000, "amips_quality": 50_000_000,
        "envelope_strong": 10_000_000, "winding_strong": 100_000_000, "pass_stream": 30_000}
UNIT = {"envelope": "points/s", "amips": "tets/s", "amips_literal": "tets/s", "amips_ring": "tets/s", "winding": "queries/s", "winding_oneshot": "queries/s",
        "envelope_faces": "faces/s", "envelope_faces_c1": "faces/s", "nearest": "points/s", "amips_quality": "tets/s",
        "envelope_strong": "points/s", "winding_strong": "queries/s", "pass_stream": "calls/s"}
METRIC = {"envelope": "envelope points/s", "amips": "AMIPS E+J+H tet-evals/s", "amips_literal": "AMIPS E+J+H tet-evals/s (literal C3)",
          "amips_ring": "AMIPS one-ring E+J+H tet-evals/s", "winding": "winding-number queries/s",
          "winding_oneshot": "winding-number queries/s, one shot (hierarchy build + evaluation + decision, InoutFiltering::filter)",
          "envelope_faces": "envelope faces/s (isFaceOutEnvelop)", "envelope_faces_c1": "envelope faces/s (isFaceOutEnvelop, faces of edge diag/20)",
          "nearest": "nearest-facet projections/s", "amips_quality": "AMIPS tet-quality evals/s (calTetQualities)",
          "pass_stream": "scheduler-shaped call stream, hot-path calls/s (one pass of MeshRefinement::doOperations)",
          "envelope_strong": "envelope points/s, fixed batch split by index over the ranks", "winding_strong": "winding-number queries/s, fixed batch split by index over the ranks"}
FACE_EDGE = 0.02     # C1-shaped candidate faces: small enough that the flat face stays within eps of the curved icosphere about half the time
FACE_EDGE_C1 = 0.05  # the edge SURVEY.md 8d states for C1 (diag/20): ~1.4 k samples per face
WORKLOAD = {
    "envelope": "C2: %d sampled points vs 200000-triangle (2,3) torus knot, eps_rel=1e-3 -> eps_2=(0.42265e-3)^2 (State.cpp:36-41)",
    "envelope_strong": "C2, strong scaling: ONE batch of %d sampled points vs 200000-triangle (2,3) torus knot, split by contiguous index range over the ranks (tetwild_b200/shard.py), decisions all-gathered with NCCL",
    "amips": "C3: %d random non-degenerate tets, flat SoA (12 arrays), E+J[3]+H[9] per tet, FP64",
    "amips_literal": "C3 as SURVEY.md 8d writes it (translation U(-10,10)^3 NOT scaled with the tet): %d random non-degenerate tets, flat SoA, E+J[3]+H[9] per tet, FP64",
    "amips_ring": "C3 smoothing-candidate layout: %d random non-degenerate tets in one-rings of k~U{12..36} around a centre vertex (indexed gather, centre rotated to slot 0), E+J[3]+H[9] per ring (NewtonsUpdate), FP64",
    "winding": "C4: %d centroids uniform in 1.2x bbox vs 1001112-triangle closed noisy UV sphere, keep = W > 0.5",
    "winding_strong": "C4 as BASELINE.json configs[3] states it, strong scaling: ONE batch of %d centroids vs the 1001112-triangle closed noisy UV sphere, sharded by contiguous index range over the ranks, decisions all-gathered with NCCL",
    "winding_oneshot": "C4 as the reference calls it (InoutFiltering.cpp:40-52): ONE call with host buffers -- build the hierarchy over the 1001112-triangle surface, evaluate %d centroids, keep = W > 0.5, flip-and-retry check",
    "amips_quality": "C3 indexed layout: calTetQualities over %d random non-degenerate tets (int4 tet -> 4 gathered vertices, exact orientation gate, energy, MAX_ENERGY rules of LocalOperations.cpp:862-884) on the resident tet mesh, FP64",
    "nearest": "C2 points, full nearest search: %d points vs 200000-triangle torus knot -> nearest facet id + nearest point + d2 (nearest_facet, mesh_AABB.h:130-176; the projection callers VertexSmoother.cpp:354-362, Preprocess.cpp:529)",
    "envelope_faces": "C1-shaped call stream: %d candidate faces (edge ~ diag/50, sampled on the device at sampling_dist = 1e-3 diag like Common.cpp:143-255) vs the 20480-triangle icosphere, eps through State.cpp:36-41",
    "pass_stream": "C1 proxy: a generated pass-shaped stream of ~%d hot-path calls on the 20480-triangle icosphere and a 16 k-tet mesh (tetwild_b200/callstream.py: the call mix of split / collapse / smooth candidates, MeshRefinement.cpp:120-183) -- no TetWild binary can be built here to record a real one",
    "envelope_faces_c1": "C1 at its stated size: %d candidate faces of edge ~ diag/20 (~1.4 k samples each at sampling_dist = 1e-3 diag) vs the 20480-triangle icosphere, eps through State.cpp:36-41",
}
L2NOTE = {
    "envelope": "inputs larger than L2: 24 B per point streamed per step (240 MB at full size); the 38 MB surface structure is meant to stay L2-resident",
    "envelope_strong": "inputs larger than L2 up to 4 ranks (240 MB of points over the ranks); the 38 MB surface structure is meant to stay L2-resident",
    "nearest": "inputs larger than L2: 24 B per point in, 36 B per point out per step",
    "envelope_faces": "L2 flushed by construction: every step re-reads 72 B per face (29 MB) between 15.7 G samples of traversal; the 4 MB surface structure stays L2-resident",
    "envelope_faces_c1": "72 B per face per step; the work is on-chip (samples are generated on the device), the 4 MB surface structure stays L2-resident",
    "amips": "inputs larger than L2: 96 B read + 104 B written per tet per step (10 GB at full size)",
    "amips_literal": "inputs larger than L2: 96 B read + 104 B written per tet per step (10 GB at full size)",
    "amips_quality": "inputs larger than L2: 16 B of indices + 73 B of vertices gathered + 8 B written per tet per step",
    "amips_ring": "inputs larger than L2: 16 B of indices + 73 B of vertices gathered per tet, 105 B written per ring per step",
    "winding": "inputs larger than L2: 24 B per query per step (2.4 GB at full size); the ~210 MB hierarchy is re-read from L2/HBM",
    "winding_strong": "inputs larger than L2: 24 B per query (2.4 GB over the ranks); the ~210 MB hierarchy is re-read from L2/HBM",
    "winding_oneshot": "one call per step: 2.4 GB of host queries + the surface in, 100 MB of decisions out",
    "pass_stream": "n/a: latency-bound host round trips (call by call) or a handful of small batches (re-batched)",
}
# algorithmic HBM bytes per unit. ring: 16 B indices + 72 B of unshared gathered vertices + (24 B centre + 105 B results) / 24;
# quality: 16 B tet + 72 B of unshared vertices + 24 B / 24 of the shared centre + 8 B out (what ncu measures: 97 B/tet)
ALG_BYTES = {"envelope": 25.0, "amips": 200.0, "amips_literal": 200.0, "amips_ring": 93.0, "winding": 25.0, "envelope_faces": 73.0, "envelope_faces_c1": 73.0,
             "nearest": 60.0, "amips_quality": 97.0, "envelope_strong": 25.0, "winding_strong": 25.0, "winding_oneshot": 25.0, "pass_stream": 0.0}
# what binds the dominant kernel of each part (from the committed ncu captures, profiles/): HBM bandwidth, L1 / LSU throughput of
# the divergent tree walks, or the FP64 pipe
BOUND = {"envelope": "l1", "envelope_strong": "l1", "envelope_faces": "l1", "envelope_faces_c1": "l1", "nearest": "l1", "amips": "hbm", "amips_literal": "hbm",
         "amips_quality": "hbm", "amips_ring": "hbm", "winding": "fp64", "winding_strong": "fp64", "winding_oneshot": "fp64", "pass_stream": "latency"}
SM_COUNT, L1_BYTES_PER_CLK = 148, 128   # B200: 148 SMs, 128 B/clk/SM of L1 (l1tex) bandwidth


def config_for(part, n, world):
    """the same dict in both arms (the driver compares them)"""
    par = "single GPU" if world == 1 else ("fixed batch split by contiguous index range over %d ranks, replicated surface, NCCL all_gather of the decisions" % world
                                           if part.endswith("_strong") else
                                           "replicated surface, every rank its own full-size batch (weak scaling), NCCL all_gather of decisions overlapped with the next step's kernels (double-buffered)")
    return {"workload": WORKLOAD[part] % n, "l2": L2NOTE[part], "parallelism": par}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        return [s for (t, s) in self.samples if t0 - 0.05 <= t <= t1 + 0.15]

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarise(lines):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for k, nm in enumerate(names):
                    if f[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- workloads
def knot_surface():
    from tetwild_b200 import synth
    return synth.torus_knot(1000, 100)


def sphere_surface():
    from tetwild_b200 import synth
    return synth.uv_sphere(708, 708)


def envelope_points_fast(V, F, n, eps, seed):
    """C2 query mix (50 % near-surface N(0,eps) normal offsets, 25 % uniform in 1.1x bbox, 25 % exactly on facets),
    float64, vectorised for 10 M points."""
    rng = np.random.default_rng(seed)
    tri = V[F.astype(np.int64)]
    n_near, n_box = n // 2, n // 4
    n_on = n - n_near - n_box
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cdf = np.cumsum(area)
    cdf /= cdf[-1]

    def on_surface(m):
        f = np.searchsorted(cdf, rng.random(m)).clip(0, len(F) - 1)
        r1, r2 = np.sqrt(rng.random(m)), rng.random(m)
        a, b, c = tri[f, 0], tri[f, 1], tri[f, 2]
        p = a * (1 - r1)[:, None] + b * (r1 * (1 - r2))[:, None] + c * (r1 * r2)[:, None]
        return p, f

    p, f = on_surface(n_near)
    nrm = np.cross(tri[f, 1] - tri[f, 0], tri[f, 2] - tri[f, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    near = p + nrm * rng.normal(0.0, eps, size=(n_near, 1))
    lo, hi = V.min(0), V.max(0)
    box = 0.5 * (lo + hi) + 0.55 * (hi - lo) * rng.uniform(-1, 1, size=(n_box, 3))
    on, _ = on_surface(n_on)
    P = np.concatenate([near, box, on])
    return np.ascontiguousarray(P[rng.permutation(n)])


def tets_on_device(n, seed, device, unit=False):
    """C3 generator on the GPU (same recipe as synth.random_tets: regular tet + N(0,0.15), random rotation, log-uniform
    scale 1e-3..1e3, translation U(-10,10)*scale; inverted draws are mirrored, (near-)flat draws replaced)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    base = torch.tensor([[0, 0, 0], [1, 0, 0], [0.5, math.sqrt(3) / 2, 0], [0.5, math.sqrt(3) / 6, math.sqrt(6) / 3]], device=device, dtype=torch.float64)
    out = torch.empty((12, n), device=device, dtype=torch.float64)
    step = 5_000_000
    for b in range(0, n, step):
        m = min(step, n - b)
        X = base[None] + 0.15 * torch.randn((m, 4, 3), generator=g, device=device, dtype=torch.float64)
        e = X[:, 1:] - X[:, :1]
        det = (e[:, 0] * torch.linalg.cross(e[:, 1], e[:, 2])).sum(1)
        bad = det.abs() < 6e-6
        X[bad] = base
        neg = (det < 0) & ~bad
        X[neg] = X[neg][:, [0, 2, 1, 3]]
        q = torch.randn((m, 4), generator=g, device=device, dtype=torch.float64)
        q = q / q.norm(dim=1, keepdim=True)
        w, x, y, z = q.unbind(1)
        R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                         torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                         torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
        X = torch.einsum("mij,mvj->mvi", R, X)
        if not unit:
            s = torch.exp(torch.empty((m, 1, 1), device=device, dtype=torch.float64).uniform_(math.log(1e-3), math.log(1e3), generator=g))
            t = torch.empty((m, 1, 3), device=device, dtype=torch.float64).uniform_(-10, 10, generator=g)
            X = X * s + t * s
        out[:, b:b + m] = X.reshape(m, 12).t()
    return out


def rings_on_device(n_tets, seed, device):
    """C3 ring layout on the GPU (same recipe as synth.ring_groups): returns V[nV,3] f64, tets[nT,4] i32, off[nG+1] u64 (as
    int64 storage), center[nG] i32. Every tet is CGAL-POSITIVE in its stored order; the centre sits at a random slot."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed + 1)
    n_groups = max(1, n_tets // 24)
    k = torch.randint(12, 37, (n_groups,), generator=g, device=device)
    # trim / pad the last groups so that the total is exactly n_tets
    off = torch.zeros(n_groups + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(k, 0)
    ng = int(torch.searchsorted(off, torch.tensor([n_tets], device=device), right=False).item())
    ng = max(1, min(ng, n_groups))
    off = off[:ng + 1].clone()
    off[ng] = n_tets
    if ng > 1 and off[ng] <= off[ng - 1]:
        ng -= 1
        off = off[:ng + 1].clone()
        off[ng] = n_tets
    n_groups = ng
    kk = off[1:] - off[:-1]
    # every ring has ONE scale and position (log-uniform 1e-3..1e3, centre U(-10,10)*scale), like a one-ring of a real mesh:
    # its member tets are perturbed regular tets of that size glued at the centre vertex
    X = tets_on_device(n_tets, seed + 2, device, unit=True).t().contiguous().view(n_tets, 4, 3)
    gi = torch.repeat_interleave(torch.arange(n_groups, device=device), kk)
    sg = torch.exp(torch.empty((n_groups, 1), device=device, dtype=torch.float64).uniform_(math.log(1e-3), math.log(1e3), generator=g))
    cen = torch.empty((n_groups, 3), device=device, dtype=torch.float64).uniform_(-10, 10, generator=g) * sg
    X = (X - X[:, :1]) * sg[gi][:, None, :] + cen[gi][:, None, :]
    V = torch.cat([cen, X[:, 1:].reshape(-1, 3)], 0).contiguous()
    base = n_groups + 3 * torch.arange(n_tets, device=device, dtype=torch.int64)
    tets = torch.stack([gi, base, base + 1, base + 2], 1)
    rot = torch.randint(0, 4, (n_tets,), generator=g, device=device)
    odd = (rot % 2) == 1
    t2 = tets.clone()
    t2[odd, 2], t2[odd, 3] = tets[odd, 3], tets[odd, 2]
    idx = (torch.arange(4, device=device)[None, :] - rot[:, None]) % 4  # out[j] = t2[(j - rot) % 4]  (np.roll by rot)
    out = torch.gather(t2, 1, idx).to(torch.int32).contiguous()
    del X
    return V, out, off.contiguous(), torch.arange(n_groups, device=device, dtype=torch.int32)


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_workload(part, n_full):
    """-> (run(m, threads) -> seconds for m units, probe size, kind, description). The reference's own code where it compiles
    here (oracle/_ref), else the oracle port."""
    import oracle as O
    from tetwild_b200 import synth
    have_ref = O.ref_available()
    base = part.replace("_strong", "")
    if base == "envelope":
        V, F = knot_surface()
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        RT = O.RefTree(V, F[S.order()]) if have_ref else None
        pts = {}

        def run(m, threads):
            if m not in pts:
                pts[m] = envelope_points_fast(V, F, m, eps, seed=20240501)
            t = time.perf_counter()
            (RT.points_out(pts[m], eps2, threads=threads) if have_ref and not O._use_native else S.points_out(pts[m], eps2, threads=threads))
            return time.perf_counter() - t
        return run, 200_000, ("reference" if have_ref else "port"), ("reference mesh_AABB.cpp facet_in_envelope_with_hint (compiled unmodified; geogram leaf distance restated)"
                                                                    if have_ref else "oracle port of mesh_AABB.cpp:482-548") + ", OpenMP over queries"
    if base in ("amips", "amips_literal"):
        lit = base == "amips_literal"
        data = {}

        def run(m, threads):
            if m not in data:
                data[m] = synth.random_tets(m, seed=7, trans_scales=not lit)
            t = time.perf_counter()
            (O.ref_amips_ejh_soa(data[m], threads=threads) if have_ref and not O._use_native else O.amips_ejh_soa(data[m], threads=threads))
            return time.perf_counter() - t
        return run, 200_000, ("reference" if have_ref else "port"), ("reference LocalOperations.cpp:28-291 text (compiled unmodified, -O2)" if have_ref
                                                                    else "oracle port (forward-mode AD)") + ", OpenMP over tets"
    if base in ("amips_quality", "amips_ring"):
        meshes = {}

        def mesh(m):
            if m not in meshes:
                meshes[m] = synth.ring_groups(max(1, m // 24), seed=7, scale_lo=0.1, scale_hi=10.0)
            return meshes[m]
        if base == "amips_quality":
            def run(m, threads):
                V, tets, off, cen = mesh(m)
                t = time.perf_counter()
                O.amips_quality(V, tets, threads=threads)
                return (time.perf_counter() - t) * m / len(tets)
            return run, 200_000, "port", "oracle port of calTetQuality_AMIPS (exact orientation predicate + the energy of LocalOperations.cpp:28-81), OpenMP over tets"

        def run(m, threads):
            V, tets, off, cen = mesh(m)
            fn = O.ref_amips_ring_ejh if have_ref and not O._use_native else O.amips_ring_ejh
            t = time.perf_counter()
            fn(V, tets, off, cen, threads=threads)
            return (time.perf_counter() - t) * m / int(off[-1])
        return run, 100_000, ("reference" if have_ref else "port"), ("NewtonsUpdate restated around the reference's own LocalOperations.cpp:28-291 E/J/H text (oracle/ref_wrap.cpp)"
                                                                    if have_ref else "oracle port of NewtonsUpdate") + ", OpenMP over rings"
    if base == "nearest":
        V, F = knot_surface()
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        RT = O.RefTree(V, F[S.order()]) if have_ref else None
        pts = {}

        def run(m, threads):
            if m not in pts:
                pts[m] = envelope_points_fast(V, F, m, eps, seed=20240501)
            t = time.perf_counter()
            (RT.nearest(pts[m], threads=threads) if have_ref and not O._use_native else S.nearest(pts[m], threads=threads))
            return time.perf_counter() - t
        return run, 50_000, ("reference" if have_ref else "port"), ("reference mesh_AABB.cpp nearest_facet (compiled unmodified; geogram leaf distance restated)"
                                                                   if have_ref else "oracle port of mesh_AABB.cpp:418-480") + ", OpenMP over queries"
    if base in ("envelope_faces", "envelope_faces_c1"):
        edge = FACE_EDGE if base == "envelope_faces" else FACE_EDGE_C1
        V, F = synth.icosphere(5)
        V = synth.normalise_unit_diag(V)
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        RT = O.RefTree(V, F[S.order()]) if have_ref else None
        faces = {}

        def run(m, threads):
            if m not in faces:
                faces[m] = synth.face_queries(V, F, m, edge, eps, seed=3)
            t = time.perf_counter()
            (RT.faces_out(faces[m], sd, eps2, threads=threads) if have_ref and not O._use_native else S.faces_out(faces[m], sd, eps2, threads=threads))
            return time.perf_counter() - t
        return run, 2000, ("reference" if have_ref else "port"), (
            "reference sampleTriangle (Common.cpp:143-255) + DistanceQuery.h + mesh_AABB.cpp compiled unmodified, composed by the loop of LocalOperations.cpp:1046-1109 (oracle/ref_wrap.cpp)"
            if have_ref else "oracle port of isFaceOutEnvelop_sampling (LocalOperations.cpp:1046-1109)") + ", first OUT sample stops the face, OpenMP over faces"
    if base == "pass_stream":
        from tetwild_b200 import callstream as cs
        cache = {}

        def run(m, threads):
            if "s" not in cache:
                cache["s"] = cs.make_pass_stream(*stream_sizes(n_full), seed=1)
                cache["r"] = cs.CpuReplayer(O, cache["s"])
            m = min(m, len(cache["s"]["calls"]))
            return cache["r"].call_by_call(limit=m)[0]
        return run, 2000, "port", "oracle, ONE call at a time on ONE core (how the reference's sequential scheduler runs; Python ctypes call overhead included, as in the GPU arm)"
    if base in ("winding", "winding_oneshot"):
        V, F = sphere_surface()
        state = {}

        def run(m, threads):
            t0 = time.perf_counter()
            if "wt" not in state:
                state["wt"] = O.WindingTree(V, F)
                state["build_s"] = time.perf_counter() - t0
            Q = synth.winding_queries(V, m, seed=11)
            t = time.perf_counter()
            state["wt"].eval(Q, threads=threads)
            dt = time.perf_counter() - t
            state["eval_rate"] = m / dt
            return dt
        what = "oracle port of libigl's exact winding-number hierarchy (libigl not vendored), OpenMP over queries"
        if base == "winding_oneshot":
            what += "; a one-shot call = hierarchy build (single-threaded, like libigl's) + evaluation of ALL queries: extrapolated from the measured build time and the sampled evaluation rate"
        else:
            what += " (hierarchy build excluded)"
        run.state = state
        return run, 20_000, "port", what
    raise ValueError(part)


def stream_sizes(n_calls):
    """(collapse, smooth, split) candidates that give about n_calls calls (~3.5 calls per candidate)"""
    k = max(30, int(n_calls / 3.5 / 8))
    return 3 * k, 3 * k, 2 * k


def cpu_rate(part, n_full, threads, budget_s=8.0, modes=False):
    """Times the reference CPU path of one part on a bounded sample. -> dict for the JSON line's cpu_baseline."""
    import oracle as O
    O.build()
    run, probe, kind, what = cpu_workload(part, n_full)
    if part == "pass_stream":
        threads, modes = 1, False      # sequential by construction
    probe = min(probe, n_full)
    r0 = probe / run(probe, threads)
    m = int(min(n_full, 20_000_000, max(probe, r0 * budget_s)))
    dt = run(m, threads)
    rate = m / dt
    out = {"value": rate, "unit": UNIT[part], "cores": threads, "kind": kind, "sample": "%d of %d units, %s" % (m, n_full, what)}
    if part == "winding_oneshot":   # whole call at full size: build + n_full / eval rate
        st = run.state
        total = st["build_s"] + n_full / st["eval_rate"]
        out["value"] = n_full / total
        out["hierarchy_build_s"] = st["build_s"]
        out["evaluation_rate_sampled"] = st["eval_rate"]
    if modes:
        # (i) single thread = how the reference really runs envelope and AMIPS; (ii) -O3 -march=native
        # upper bound (oracle port rebuilt on this machine; the reference-compiled pieces keep their -O2 build)
        m1 = int(max(probe // 4, min(m, rate / max(1, threads) * 2.0)))
        out["single_thread"] = {"value": m1 / run(m1, 1), "unit": UNIT[part], "cores": 1, "sample": "%d units" % m1}
        try:
            with O.native():
                run_n, _, _, what_n = cpu_workload(part, n_full)
                run_n(probe, threads)
                mn = int(min(m, max(probe, r0 * 3.0)))
                out["o3_march_native"] = {"value": mn / run_n(mn, threads), "unit": UNIT[part], "cores": threads, "kind": "port",
                                          "sample": "%d units, oracle port built with gcc -O3 -march=native on this machine" % mn}
        except Exception as ex:  # noqa: BLE001
            out["o3_march_native"] = {"unavailable": repr(ex)[:200]}
    return out


def run_reference(args, parts):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle as O
    threads = O.max_threads()
    lines = {}
    for part in parts:
        n_full = max(1000, int(FULL[part] * args.scale))
        rates = []
        for _ in range(max(1, min(args.steps, 3)) if part == parts[0] else 1):   # the headline part: median of up to three samples
            cb = cpu_rate(part, n_full, threads, budget_s=6.0)
            rates.append(cb["value"])
        v = float(np.median(rates))
        cb["value"] = v
        lines[part] = {"metric": METRIC[part], "value": v, "unit": UNIT[part], "ms_per_step": None, "cpu_baseline": cb,
                       "e2e": {"value": v, "unit": UNIT[part], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                       "config": config_for(part, n_full, world)}
    head = parts[0]
    out = {"impl": "reference", "metric": METRIC[head], "value": lines[head]["value"], "unit": UNIT[head], "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": lines[head]["config"],
           "cpu_baseline": lines[head]["cpu_baseline"], "e2e": lines[head]["e2e"], "gpu_launches": 0,
           "parts": {p: lines[p] for p in parts if p != head}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------- GPU arm
def tets_on_device_literal(n, seed, device):
    """C3 exactly as SURVEY.md 8d writes it: translation U(-10,10)^3 NOT multiplied by the scale (synth.random_tets(trans_scales=False))"""
    import torch
    T = tets_on_device(n, seed, device, unit=True)
    g = torch.Generator(device=device).manual_seed(seed + 77)
    step = 5_000_000
    for b in range(0, n, step):
        m = min(step, n - b)
        sc = torch.exp(torch.empty((1, m), device=device, dtype=torch.float64).uniform_(math.log(1e-3), math.log(1e3), generator=g))
        t = torch.empty((3, m), device=device, dtype=torch.float64).uniform_(-10, 10, generator=g)
        T[:, b:b + m] = T[:, b:b + m] * sc + t.repeat(4, 1)
    return T


def run_gpu(args, parts):
    import torch
    import torch.distributed as dist
    import tetwild_b200 as tw
    from tetwild_b200 import shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    unpinned = []

    def pin(t):
        """page-lock a host tensor; a box that cannot lock that much memory (8 ranks x ~10 GB) still runs, with pageable buffers"""
        try:
            return t.pin_memory()
        except RuntimeError:
            unpinned.append(int(t.numel() * t.element_size()))
            return t

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = tw.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream
    hbm_peak, peak_src = peaks()
    # roofline denominators measured in this very run on this device (peaks.cu): FP64 DFMA rate and the library's own copy kernel
    fp64_peak = ctx.measure_fp64_tflops()
    fp64_peak_distinct = ctx.measure_fp64_tflops_distinct()
    copy_gbs = ctx.measure_copy_gbs(1 << 30)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    head = parts[0]
    prof = load_ncu_traffic()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Pipe:
        """Double-buffered 1-byte decisions of one rank + their NCCL all_gather. The gather of step k is issued asynchronously
        (NCCL's own stream, ordered after step k's kernels) and overlaps the kernels of step k+1, which write the other
        buffer; a buffer is handed out again only after its previous gather has completed. drain() makes the launching
        stream wait for every outstanding gather, so the timed region contains all of them."""

        def __init__(self, n):
            self.buf = [torch.empty(n, device=dev, dtype=torch.uint8) for _ in range(2)]
            self.gath = [torch.empty(n * world, device=dev, dtype=torch.uint8) for _ in range(2)] if world > 1 else None
            self.work = [None, None]
            self.k = 0

        def out(self):
            p = self.k & 1
            if self.work[p] is not None:
                self.work[p].wait()
                self.work[p] = None
            return self.buf[p]

        def gather(self):
            p = self.k & 1
            if world > 1:
                self.work[p] = dist.all_gather_into_tensor(self.gath[p], self.buf[p], async_op=True)
            self.k += 1

        def drain(self):
            for p in range(2):
                if self.work[p] is not None:
                    self.work[p].wait()
                    self.work[p] = None

        def last(self):
            return self.buf[(self.k - 1) & 1]

        def last_gathered(self):
            return self.gath[(self.k - 1) & 1] if world > 1 else self.buf[(self.k - 1) & 1]

    def timed(part, step_fn, gather_fn=None, drain_fn=None):
        """W warm-up steps, then K steps bracketed by barrier+sync; device time (CUDA events on the launching stream),
        max over ranks. Returns (ms_per_step, kernel_ms_avg, launches, clock window, steps). The headline part runs exactly
        --steps; the other parts at most 5 (stated per part) so that the default run stays within minutes."""
        K = args.steps if part == head else min(args.steps, 5)
        Wm = args.warmup if part == head else min(args.warmup, 3)
        for _ in range(Wm):
            step_fn()
            if gather_fn:
                gather_fn()
        if drain_fn:
            drain_fn()
        barrier()
        l0 = ctx.launches
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 2)]
        t0 = time.perf_counter()
        ev[0].record(stream)
        for k in range(K):
            ev[2 + 2 * k].record(stream)
            step_fn()
            ev[3 + 2 * k].record(stream)
            if gather_fn:
                gather_fn()
        if drain_fn:
            drain_fn()
        ev[1].record(stream)
        barrier()
        t1 = time.perf_counter()
        total = max_over_ranks(ev[0].elapsed_time(ev[1]))
        kern = float(np.mean([ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) for k in range(K)]))
        return total / K, kern, ctx.launches - l0, (t0, t1), K, Wm

    def e2e_timed(part, call):
        K = args.steps if part == head else min(args.steps, 3)
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return max_over_ranks(dt) / K

    def roofline(part, n, kms, sm_mhz, extra=None):
        """achieved = algorithmic units of the bound resource per launch / measured kernel time; what binds each part is read off
        the committed ncu captures (profiles/ncu_traffic.json): HBM bytes (SURVEY.md 8d figures), L1 bytes per unit as ncu counted
        them (l1tex__t_bytes), or FP64-pipe instructions per unit (sm__inst_executed_pipe_fp64, one DFMA slot each)."""
        t = prof.get(part.replace("_strong", "").replace("_oneshot", "").replace("amips_literal", "amips").replace("envelope_faces_c1", "envelope_faces"), {})
        bound = BOUND[part]
        ach_hbm = ALG_BYTES[part] * n / (kms * 1e-3) / 1e9
        r = {"bound": bound, "kernel_ms": kms, "hbm_algorithmic_bytes_per_unit": ALG_BYTES[part], "hbm_achieved_gbs": ach_hbm, "hbm_peak_gbs": hbm_peak,
             "hbm_frac": ach_hbm / hbm_peak, "peak_source": peak_src, "fp64_dfma_peak_tflops_measured_in_run": fp64_peak,
             "fp64_distinct_operand_peak_tflops_measured_in_run": fp64_peak_distinct, "copy_gbs_measured_in_run": copy_gbs,
             "traffic": (t["dram_bytes"] * n / t["units"]) if "dram_bytes" in t else None,
             "kernel_ms_scope": "CUDA events around ALL launches of one step on the launching stream (for the point / nearest / winding parts that includes the Morton ordering of the batch), so `achieved` is a lower bound for the traversal kernel alone"}
        if "file" in t:
            r["traffic_source"] = "%s (ncu --set full, per launch of %d units, scaled per unit)" % (t["file"], t["units"])
        if bound == "hbm":
            r.update({"achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_hbm / hbm_peak})
        elif bound == "l1":
            clk = (sm_mhz or 1965.0) * 1e6
            peak = SM_COUNT * L1_BYTES_PER_CLK * clk / 1e9
            if "l1_bytes" in t:
                ach = t["l1_bytes"] / t["units"] * n / (kms * 1e-3) / 1e9
                r.update({"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                          "l1_bytes_per_unit_ncu": t["l1_bytes"] / t["units"], "l1_peak_source": "148 SMs x 128 B/clk x SM clock under load"})
                if "lts_bytes" in t:
                    r["l2_achieved_gbs"] = t["lts_bytes"] / t["units"] * n / (kms * 1e-3) / 1e9
                    r["l2_bytes_per_unit_ncu"] = t["lts_bytes"] / t["units"]
            else:
                r.update({"achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_hbm / hbm_peak, "note": "no L1 byte count committed for this kernel: HBM convention"})
        elif bound == "latency":
            r.update({"achieved": None, "peak": None, "unit": None, "frac": None, "note": "host round trips: no device roofline applies"})
        elif bound == "fp64":
            if "fp64_inst" in t:
                flops = t["fp64_inst"] / t["units"] * 64.0 * n / (kms * 1e-3) / 1e12   # one FP64-pipe warp instruction = 32 lanes x (1 DFMA = 2 flops)
                r.update({"achieved": flops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops / fp64_peak,
                          "frac_of_distinct_operand_peak": flops / fp64_peak_distinct,
                          "fp64_pipe_instructions_per_unit_ncu": t["fp64_inst"] / t["units"],
                          "convention": "DFMA-equivalents: every FP64-pipe instruction counted as one fused multiply-add (2 flops)"})
            else:
                r.update({"achieved": ach_hbm, "peak": hbm_peak, "unit": "GB/s", "frac": ach_hbm / hbm_peak, "note": "no FP64 instruction count committed: HBM convention"})
        if "fp64_pipe_pct" in t:
            r["fp64_pipe_active_pct_ncu"] = t["fp64_pipe_pct"]
        # what else the committed ncu capture of this kernel says about its limiter (percent of ncu's own peaks)
        for k in ("l1tex_throughput_pct", "lts_throughput_pct", "dram_throughput_pct", "issue_active_pct", "lanes_per_inst", "warps_active_pct"):
            if k in t:
                r[k + "_ncu"] = t[k]
        if part == "amips_ring":
            r["note"] = ("HBM convention (the part's algorithmic traffic), but the kernel is bound by instruction issue and gather latency: see issue_active_pct_ncu, "
                         "l1tex_throughput_pct_ncu, dram_throughput_pct_ncu and DESIGN.md 3.1")
        if extra:
            r.update(extra)
        return r

    results = {}
    import oracle as O  # cpu_baseline leg + post-timing parity samples only (rank 0)

    for part in parts:
        n = max(1000, int(FULL[part] * args.scale))
        res = {"metric": METRIC[part], "unit": UNIT[part], "config": config_for(part, n, world), "scaling": "strong" if part.endswith("_strong") else "weak"}
        roof_extra = None
        units_per_step = n * world
        if part in ("envelope", "envelope_strong"):
            strong = part == "envelope_strong"
            V, F = knot_surface()
            sd, eps, eps2 = synth.state_eps(1e-3)
            S = tw.Surface(ctx, V, F)
            if strong:   # ONE batch (same seed on every rank), this rank's contiguous index range
                Pall = envelope_points_fast(V, F, n, eps, seed=20240501)
                b0, e0 = shard.shard_bounds(n, world, rank)
                P = np.ascontiguousarray(Pall[b0:e0])
                units_per_step = n
            else:
                P = envelope_points_fast(V, F, n, eps, seed=20240501 + rank)
            m = len(P)
            hP = pin(torch.from_numpy(P))
            dP = hP.to(dev, non_blocking=True)
            pipe = Pipe(max(shard.shard_sizes(n, world)) if strong else n)
            step = lambda: S.points_out_dev(dP.data_ptr(), m, eps2, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step, pipe.gather, pipe.drain)
            dO = pipe.last()
            if strong:
                # e2e of the sharded batch: this rank's slice through the host-buffer C-ABI call (pinned host memory in, chunked H2D
                # overlapped with the kernels, decisions back to the host), then the 1-byte decisions of all ranks gathered over
                # NVLink (NCCL all_gather of device copies) and the whole result on the host of rank 0
                hLoc = pin(torch.empty(m, dtype=torch.uint8))
                hO = pin(torch.empty(n if rank == 0 else 1, dtype=torch.uint8))
                dLoc = torch.empty(max(shard.shard_sizes(n, world)), device=dev, dtype=torch.uint8)

                def call():
                    S.points_out(hP.numpy(), eps2, out=hLoc.numpy())
                    if world > 1:
                        dLoc[:m].copy_(hLoc, non_blocking=True)
                        full = shard.all_gather_ragged(dLoc[:m], n)
                        if rank == 0:
                            hO.copy_(full, non_blocking=True)
                    torch.cuda.synchronize()
                e2e_s = e2e_timed(part, call)
                h2d, d2h = m * 24 + (m if world > 1 else 0), m + (n if (rank == 0 and world > 1) else 0)
            else:
                hO = pin(torch.empty(n, dtype=torch.uint8))
                e2e_s = e2e_timed(part, lambda: S.points_out(hP.numpy(), eps2, out=hO.numpy()))
                h2d, d2h = n * 24, n
            out_frac = float(dO[:m].float().mean().item())
            # parity inside the bench: a 100k sample of this very batch against the oracle (decisions must be identical)
            idx = np.random.default_rng(5).choice(m, min(m, 100_000), replace=False)
            mism = None
            if rank == 0:
                OS = O.Surface(V, F)
                mism = int((OS.points_out(P[idx], eps2, threads=O.max_threads()) != dO.cpu().numpy()[idx]).sum())
                if strong and world > 1:   # the gathered array holds every rank's slice in index order
                    full = pipe.last_gathered().cpu().numpy()
                    sz = max(shard.shard_sizes(n, world))
                    other = world - 1
                    bo, eo = shard.shard_bounds(n, world, other)
                    j = np.random.default_rng(6).choice(eo - bo, min(eo - bo, 20_000), replace=False)
                    mism += int((OS.points_out(Pall[bo:eo][j], eps2, threads=O.max_threads()) != full[other * sz:other * sz + (eo - bo)][j]).sum())
            res.update({"h2d": h2d, "d2h": d2h, "extra": {"out_of_envelope_fraction": out_frac, "decision_mismatches_vs_oracle_100k_sample": mism,
                                                           "surface_triangles": int(len(F)), "env_stack_overflow_fallbacks": ctx.debug_counter(0)}})
            if strong:
                res["extra"]["e2e_pcie_gbs_per_rank"] = m * 24 / e2e_s / 1e9
                res["extra"]["e2e_host_to_device_gbs_all_ranks"] = n * 24 / e2e_s / 1e9
            del dP, dO, S, pipe
        elif part == "nearest":
            V, F = knot_surface()
            sd, eps, eps2 = synth.state_eps(1e-3)
            S = tw.Surface(ctx, V, F)
            P = envelope_points_fast(V, F, n, eps, seed=20240501 + rank)
            hP = pin(torch.from_numpy(P))
            dP = hP.to(dev, non_blocking=True)
            dF = torch.empty(n, device=dev, dtype=torch.int32)
            dN = torch.empty((n, 3), device=dev, dtype=torch.float64)
            dD = torch.empty(n, device=dev, dtype=torch.float64)
            step = lambda: S.nearest_dev(dP.data_ptr(), n, dF.data_ptr(), dN.data_ptr(), dD.data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step)
            hF, hN, hD = pin(torch.empty(n, dtype=torch.int32)), pin(torch.empty((n, 3), dtype=torch.float64)), pin(torch.empty(n, dtype=torch.float64))
            outs = (hF.numpy().view(np.uint32), hN.numpy(), hD.numpy())
            e2e_s = e2e_timed(part, lambda: S.nearest(hP.numpy(), out=outs))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 20_000), replace=False)
                dref = O.Surface(V, F).sqdist_brute(P[idx], threads=O.max_threads())[0]
                got = dD.cpu().numpy()[idx]
                pt = dN.cpu().numpy()[idx]
                mism = {"d2_mismatches_vs_brute_force": int((got != dref).sum()), "sample": int(len(idx)),
                        "nearest_point_realises_d2_max_err_rel_to_d2_plus_ulp": float((np.abs(((P[idx] - pt) ** 2).sum(1) - dref) / (dref + 1e-15 * np.sqrt(dref) + 1e-30)).max())}
            # the projection callers (VertexSmoother.cpp:354-362) only ever ask for points ON or next to the surface; a quarter of the
            # C2 mix is uniform in the bounding box, and an exact nearest search for a point far from a curved surface has
            # hundreds of candidate facets whatever the hierarchy. The same kernels on the near-surface part of the batch:
            sub = torch.nonzero(dD <= (4.0 * eps) ** 2).flatten()
            ms_near, m_near = None, int(sub.numel())
            if m_near > 0:
                dPs = dP[sub].contiguous()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for it in range(4):
                    if it == 1:
                        e0.record(stream)
                    S.nearest_dev(dPs.data_ptr(), m_near, dF.data_ptr(), dN.data_ptr(), dD.data_ptr(), sh)
                e1.record(stream)
                torch.cuda.synchronize()
                ms_near = e0.elapsed_time(e1) / 3.0
                del dPs
            res.update({"h2d": n * 24, "d2h": n * 36, "extra": {"parity_vs_brute_force": mism, "surface_triangles": int(len(F)),
                                                                 "kernel": "nearest_packet_kernel" if ctx.get_option("nearest_mode") == 1 else "nearest_kernel",
                                                                 "near_surface_subset": {"what": "the points of the same batch within 4 eps of the surface (the projection callers' case), same call, 3 timed steps",
                                                                                         "points": m_near, "ms_per_step": ms_near,
                                                                                         "points_per_s": (m_near / (ms_near * 1e-3)) if ms_near else None}}})
            del dP, dF, dN, dD, S
        elif part in ("envelope_faces", "envelope_faces_c1"):
            edge = FACE_EDGE if part == "envelope_faces" else FACE_EDGE_C1
            V, F = synth.icosphere(5)
            V = synth.normalise_unit_diag(V)
            sd, eps, eps2 = synth.state_eps(1e-3)
            S = tw.Surface(ctx, V, F)
            T = synth.face_queries(V, F, n, edge, eps, seed=3 + rank)
            hT = pin(torch.from_numpy(T))
            dTr = hT.to(dev, non_blocking=True)
            pipe = Pipe(n)
            step = lambda: S.faces_out_dev(dTr.data_ptr(), n, sd, eps2, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step, pipe.gather, pipe.drain)
            dO = pipe.last()
            e2e_s = e2e_timed(part, lambda: S.faces_out(hT.numpy(), sd, eps2))
            mism = nsamp = None
            extra = {}
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 3000 if part == "envelope_faces_c1" else 5000), replace=False)
                ref, cnt = O.Surface(V, F).faces_out(T[idx], sd, eps2, threads=O.max_threads())
                mism = int((ref != dO.cpu().numpy()[idx]).sum())
                nsamp = float(np.mean(cnt))
            if part == "envelope_faces_c1":
                # the same faces on a surface with large flat regions (a 20 172-triangle cube): a face of edge diag/20 can lie wholly
                # inside the envelope there, so ALL of its ~1.4 k samples are tested -- the cost SURVEY.md 8d has in mind. On the
                # icosphere a flat face of that size leaves the envelope of the curved surface (sagitta 1.5e-3 > eps) and stops early.
                Vc, Fc = synth.cube_surface(41)
                Vc = synth.normalise_unit_diag(Vc)
                Sc = tw.Surface(ctx, Vc, Fc)
                nf = min(n, 100_000)
                Tc = synth.face_queries(Vc, Fc, nf, edge, eps, seed=9 + rank)
                dTc = torch.from_numpy(Tc).to(dev)
                dOc = torch.empty(nf, device=dev, dtype=torch.uint8)
                stepc = lambda: Sc.faces_out_dev(dTc.data_ptr(), nf, sd, eps2, dOc.data_ptr(), sh)  # noqa: E731
                msc, kmsc, _, _, _, _ = timed(part, stepc)
                flat = {"surface": "cube, 20172 triangles, unit diagonal", "faces": nf, "faces_per_s": nf * world / (msc * 1e-3), "out_of_envelope_fraction": float(dOc.float().mean().item())}
                if rank == 0:
                    j = np.random.default_rng(5).choice(nf, min(nf, 1500), replace=False)
                    refc, cntc = O.Surface(Vc, Fc).faces_out(Tc[j], sd, eps2, threads=O.max_threads())
                    flat["decision_mismatches_vs_oracle_1500_sample"] = int((refc != dOc.cpu().numpy()[j]).sum())
                    flat["mean_samples_per_face_sampleTriangle"] = float(np.mean(cntc))
                    flat["samples_per_s"] = flat["faces_per_s"] * float(np.mean(cntc[refc == 0])) * (1 - flat["out_of_envelope_fraction"])
                extra["flat_surface"] = flat
                Sc.close()
                del dTc, dOc
            res.update({"h2d": n * 72, "d2h": n, "extra": dict(extra, out_of_envelope_fraction=float(dO.float().mean().item()),
                                                               decision_mismatches_vs_oracle_sample=mism,
                                                               mean_samples_per_face_sampleTriangle=nsamp, surface_triangles=int(len(F)))})
            del dTr, dO, S, pipe
        elif part in ("amips", "amips_literal"):
            lit = part == "amips_literal"
            dT = tets_on_device_literal(n, 7 + rank, dev) if lit else tets_on_device(n, 7 + rank, dev)
            dE = torch.empty(n, device=dev, dtype=torch.float64)
            dJ = torch.empty((n, 3), device=dev, dtype=torch.float64)
            dH = torch.empty((n, 9), device=dev, dtype=torch.float64)
            ptrs = [dT[k].data_ptr() for k in range(12)]
            step = lambda: ctx.amips_ejh_soa_dev(ptrs, dE.data_ptr(), dJ.data_ptr(), dH.data_ptr(), n, sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step)
            hT = pin(torch.empty((12, n), dtype=torch.float64))
            hT.copy_(dT)
            hE, hJ, hH = (pin(torch.empty(s, dtype=torch.float64)) for s in ((n,), (n, 3), (n, 9)))
            e2e_s = e2e_timed(part, lambda: ctx.amips_ejh_soa(hT.numpy(), out=(hE.numpy(), hJ.numpy(), hH.numpy())))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 50_000), replace=False)
                Ts = np.ascontiguousarray(hT.numpy()[:, idx])
                got = (dE.cpu().numpy()[idx], dJ.cpu().numpy()[idx], dH.cpu().numpy()[idx])

                def nw(a, b2):
                    k = len(b2[0])
                    return [np.abs(a[c].reshape(k, -1) - b2[c].reshape(k, -1)).max(1) / np.abs(b2[c].reshape(k, -1)).max(1) for c in range(3)]
                ref = O.ref_amips_ejh_soa(Ts, threads=O.max_threads()) if O.ref_available() else O.amips_ejh_soa(Ts, threads=O.max_threads())
                e = nw(got, ref)
                mism = {"criterion": "per tet and tensor: max-norm error / max-norm of the tensor (DESIGN.md 3.1)",
                        "vs_reference_text_double": {"max_E_J_H": [float(x.max()) for x in e], "over_1e-9": int(((e[0] > 1e-9) | (e[1] > 1e-9) | (e[2] > 1e-9)).sum())},
                        "sample": int(len(idx))}
                if O.quad_available():   # the reference's text in IEEE binary128 = the exact value of its expression
                    Qt = O.refq_amips_ejh_soa(Ts, threads=O.max_threads())
                    eg, er = nw(got, Qt), nw(ref, Qt)
                    mism["gpu_vs_truth"] = {"max_E_J_H": [float(x.max()) for x in eg], "p99_E_J_H": [float(np.quantile(x, .99)) for x in eg],
                                            "over_1e-9": int(((eg[0] > 1e-9) | (eg[1] > 1e-9) | (eg[2] > 1e-9)).sum())}
                    mism["ref_vs_truth"] = {"max_E_J_H": [float(x.max()) for x in er], "p99_E_J_H": [float(np.quantile(x, .99)) for x in er],
                                            "over_1e-9": int(((er[0] > 1e-9) | (er[1] > 1e-9) | (er[2] > 1e-9)).sum())}
                    mism["truth"] = "reference LocalOperations.cpp:28-291 compiled in IEEE binary128 (oracle/ref_quad.cpp)"
            res.update({"h2d": n * 96, "d2h": n * 104, "extra": {("parity_literal_c3" if lit else "parity_vs_reference_text"): mism}})
            del dT, dE, dJ, dH, hT, hE, hJ, hH
        elif part == "amips_quality":
            dV, dT4, dOff, dCen = rings_on_device(n, 7 + rank, dev)
            hV, hT4 = pin(dV.cpu()), pin(dT4.cpu())
            nV = int(dV.shape[0])
            del dV, dT4, dOff, dCen
            torch.cuda.empty_cache()
            M = tw.TetMesh(ctx, hV.numpy(), hT4.numpy())
            dQ = torch.empty(n, device=dev, dtype=torch.float64)
            step = lambda: M.quality_dev(0, n, dQ.data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step)
            hQ = pin(torch.empty(n, dtype=torch.float64))
            e2e_s = e2e_timed(part, lambda: M.quality(out=hQ.numpy()))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 50_000), replace=False)
                sub = hT4.numpy()[idx]
                uniq, inv = np.unique(sub.ravel(), return_inverse=True)
                ref = O.amips_quality(hV.numpy()[uniq], inv.reshape(-1, 4).astype(np.int32), threads=O.max_threads())
                got = dQ.cpu().numpy()[idx]
                gate = int(((got == tw.MAX_ENERGY) != (ref == O.MAX_ENERGY)).sum())
                ok = ref != O.MAX_ENERGY
                mism = {"gate_mismatches": gate, "max_rel_err": float((np.abs(got[ok] - ref[ok]) / ref[ok]).max()), "sample": int(len(idx))}
            res.update({"h2d": 0, "d2h": n * 8, "extra": {"vertices": nV, "parity_vs_oracle": mism,
                                                           "e2e_path": "twg_mesh_quality over every slot of the resident mesh (nothing in, 8 B per tet out)"}})
            M.close()
            del dQ, hV, hT4
        elif part == "amips_ring":
            dV, dT4, dOff, dCen = rings_on_device(n, 7 + rank, dev)
            nG, nV = int(dCen.numel()), int(dV.shape[0])
            dE = torch.empty(nG, device=dev, dtype=torch.float64)
            dJ = torch.empty((nG, 3), device=dev, dtype=torch.float64)
            dH = torch.empty((nG, 9), device=dev, dtype=torch.float64)
            dOk = torch.empty(nG, device=dev, dtype=torch.uint8)
            step = lambda: ctx.amips_ring_ejh_dev(dV.data_ptr(), nV, dT4.data_ptr(), n, 0, dOff.data_ptr(), dCen.data_ptr(), nG, dE.data_ptr(),  # noqa: E731
                                                  dJ.data_ptr(), dH.data_ptr(), dOk.data_ptr(), sh)
            ms, kms, launches, win, K, Wm = timed(part, step)
            hV, hT4, hOff, hCen = pin(dV.cpu()), pin(dT4.cpu()), pin(dOff.cpu()), pin(dCen.cpu())
            ship_s = e2e_timed(part, lambda: ctx.amips_ring_ejh(hV.numpy(), hT4.numpy(), hOff.numpy().view(np.uint64), hCen.numpy()))
            # the integration the scheduler uses: the tet mesh is RESIDENT on the device (uploaded once,
            # kept in step by scatter updates), a Newton batch ships 4 B of vertex id per ring in and 105 B per ring out
            t0 = time.perf_counter()
            M = tw.TetMesh(ctx, hV.numpy(), hT4.numpy())
            M.build_rings()
            mesh_build_s = time.perf_counter() - t0
            hE, hJ, hH = (pin(torch.empty(sz, dtype=torch.float64)) for sz in ((nG,), (nG, 3), (nG, 9)))
            hOk = pin(torch.empty(nG, dtype=torch.uint8))
            outs = (hE.numpy(), hJ.numpy(), hH.numpy(), hOk.numpy())
            e2e_s = e2e_timed(part, lambda: M.vertex_ring_ejh(hCen.numpy(), out=outs))
            same = bool(np.array_equal(hE.numpy(), dE.cpu().numpy()) and np.array_equal(hH.numpy(), dH.cpu().numpy()))
            M.close()
            mism = None
            if rank == 0:
                gsel = np.random.default_rng(5).choice(nG, min(nG, 4000), replace=False)
                offs = hOff.numpy()
                mem = np.concatenate([np.arange(offs[a], offs[a + 1]) for a in gsel])
                soff = np.zeros(len(gsel) + 1, dtype=np.uint64)
                soff[1:] = np.cumsum(offs[gsel + 1] - offs[gsel])
                ring_ref = O.ref_amips_ring_ejh if O.ref_available() else O.amips_ring_ejh
                Eo, Jo, Ho, oko = ring_ref(hV.numpy(), hT4.numpy(), soff, hCen.numpy()[gsel], t_ids=mem.astype(np.int32), threads=O.max_threads())
                got = (dE.cpu().numpy()[gsel], dJ.cpu().numpy()[gsel], dH.cpu().numpy()[gsel])
                eE = np.abs(got[0] - Eo) / np.abs(Eo)
                eJ = np.abs(got[1] - Jo).max(1) / np.abs(Jo).max(1)
                eH = np.abs(got[2] - Ho).max(1) / np.abs(Ho).max(1)
                mism = {"max_rel_err_E": float(eE.max()), "max_rel_err_J_normwise": float(eJ.max()), "max_rel_err_H_normwise": float(eH.max()),
                        "ok_flag_mismatches": int((dOk.cpu().numpy()[gsel] != oko).sum()), "rings_checked": int(len(gsel))}
            res.update({"h2d": nG * 4, "d2h": nG * 105,
                        "extra": {"rings": nG, "vertices": nV, "parity_vs_oracle": mism,
                                  "e2e_path": "twg_mesh_vertex_ring_ejh on the resident tet mesh (host ids in, host E/J/H/ok out)",
                                  "resident_results_identical_to_device_batch": same, "resident_mesh_upload_and_ring_build_s": mesh_build_s,
                                  "e2e_ship_everything": {"value": n * world / ship_s, "unit": "tets/s", "path": "twg_amips_ring_ejh (vertices + tets + CSR shipped with every call)",
                                                          "h2d_bytes_per_step": nV * 24 + n * 16 + (nG + 1) * 8 + nG * 4, "d2h_bytes_per_step": nG * 105}}})
            del dV, dT4, dOff, dCen, dE, dJ, dH, dOk, hV, hT4, hOff, hCen
        elif part in ("winding", "winding_strong"):
            strong = part == "winding_strong"
            V, F = sphere_surface()
            t0 = time.perf_counter()
            Wt = tw.Winding(ctx, V, F)
            build_s = time.perf_counter() - t0
            lo, hi = torch.tensor(V.min(0), device=dev), torch.tensor(V.max(0), device=dev)
            if strong:   # ONE batch: the same generator state on every rank, this rank keeps its contiguous index range
                b0, e0 = shard.shard_bounds(n, world, rank)
                m = e0 - b0
                g = torch.Generator(device=dev).manual_seed(11)
                dQ = torch.empty((m, 3), device=dev, dtype=torch.float64)
                blk = 10_000_000
                for bb in range(0, n, blk):   # every rank draws the whole stream block by block and keeps its rows
                    r_ = torch.rand((min(blk, n - bb), 3), generator=g, device=dev, dtype=torch.float64)
                    lo_, hi_ = max(bb, b0), min(bb + len(r_), e0)
                    if lo_ < hi_:
                        dQ[lo_ - b0:hi_ - b0] = 0.5 * (lo + hi) + 0.6 * (hi - lo) * (2 * r_[lo_ - bb:hi_ - bb] - 1)
                    del r_
                units_per_step = n
            else:
                m = n
                g = torch.Generator(device=dev).manual_seed(11 + rank)
                dQ = (0.5 * (lo + hi) + 0.6 * (hi - lo) * (2 * torch.rand((n, 3), generator=g, device=dev, dtype=torch.float64) - 1)).contiguous()
            pipe = Pipe(max(shard.shard_sizes(n, world)) if strong else n)
            p0 = ctx.debug_counter(2)
            step = lambda: Wt.eval_dev(dQ.data_ptr(), m, 0, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win, K, Wm = timed(part, step, pipe.gather, pipe.drain)
            pairs_per_query = (ctx.debug_counter(2) - p0) / float((K + Wm) * m)
            dK = pipe.last()
            hQ = pin(torch.empty((m, 3), dtype=torch.float64))
            hQ.copy_(dQ)
            if strong:
                hLoc = pin(torch.empty(m, dtype=torch.uint8))
                hK = pin(torch.empty(n if rank == 0 else 1, dtype=torch.uint8))
                dLoc = torch.empty(max(shard.shard_sizes(n, world)), device=dev, dtype=torch.uint8)

                def call():   # host-buffer call on this rank's slice, then the decisions of all ranks gathered over NVLink, whole result on rank 0's host
                    Wt.eval(hQ.numpy(), want_w=False, out=(None, hLoc.numpy()))
                    if world > 1:
                        dLoc[:m].copy_(hLoc, non_blocking=True)
                        full = shard.all_gather_ragged(dLoc[:m], n)
                        if rank == 0:
                            hK.copy_(full, non_blocking=True)
                    torch.cuda.synchronize()
                e2e_s = e2e_timed(part, call)
                h2d, d2h = m * 24 + (m if world > 1 else 0), m + (n if (rank == 0 and world > 1) else 0)
                del dLoc
            else:
                hK = pin(torch.empty(n, dtype=torch.uint8))
                e2e_s = e2e_timed(part, lambda: Wt.eval(hQ.numpy(), want_w=False, out=(None, hK.numpy())))
                h2d, d2h = n * 24, n
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(m, min(m, 20_000), replace=False)
                Wo = O.WindingTree(V, F).eval(hQ.numpy()[idx], threads=O.max_threads())
                mism = int(((Wo > 0.5).astype(np.uint8) != dK.cpu().numpy()[idx]).sum())
            res.update({"h2d": h2d, "d2h": d2h, "extra": {"inside_fraction": float(dK[:m].float().mean().item()), "hierarchy_build_s": build_s,
                                                           "decision_mismatches_vs_oracle_20k_sample": mism, **Wt.stats()}})
            roof_extra = {"pairs_per_query": pairs_per_query,
                          "pairs_note": "(query, cap point or leaf facet) evaluations per query, counted by the kernel itself (twg_debug_counter 2)"}
            del dQ, dK, Wt, pipe
        elif part == "pass_stream":
            from tetwild_b200 import callstream as cs
            strm = cs.make_pass_stream(*stream_sizes(n), seed=1)
            ncalls = len(strm["calls"])
            counts, units = cs.mix(strm)
            G = cs.GpuReplayer(ctx, strm)
            G.batched(); G.call_by_call()                      # warm-up (scratch, pinned slabs, ring build)
            barrier()
            l0 = ctx.launches
            t0 = time.perf_counter()
            K, Wm = min(args.steps, 3), 1
            tb = [G.batched() for _ in range(K)]
            t1 = time.perf_counter()
            ta, ra = G.call_by_call()
            launches = ctx.launches - l0
            win = (t0, t1)
            rb = tb[-1][1]
            bsec = float(np.median([x[0] for x in tb]))
            mism = None
            if rank == 0:
                tc, rc = cs.CpuReplayer(O, strm).call_by_call()
                mism = {"call_by_call_vs_batched": int(sum(not cs.same(k, ra[i], rb[i], tol=1e-12) for i, (k, _) in enumerate(strm["calls"]))),
                        "gpu_vs_oracle": int(sum(not cs.same(k, ra[i], rc[i]) for i, (k, _) in enumerate(strm["calls"])))}
            G.close()
            units_per_step = ncalls * world
            ms = kms = bsec * 1e3
            e2e_s = bsec
            res.update({"h2d": 0, "d2h": 0, "extra": {"calls": ncalls, "call_mix": counts, "units": units, "mismatches": mism,
                                                     "rebatched_by_kind_calls_per_s": ncalls / bsec,
                                                     "call_by_call_calls_per_s": ncalls / ta, "call_by_call_us_per_call": ta / ncalls * 1e6,
                                                     "value_is": "the stream re-batched by kind through the host-buffer C ABI (7 calls per pass); call by call is reported beside it; both include the Python ctypes call overhead"}})
        elif part == "winding_oneshot":
            # InoutFiltering::filter as the reference calls it (InoutFiltering.cpp:40-52): surface + all centroids in, decisions out,
            # ONE call: hierarchy build + H2D + evaluation + D2H + the all-removed check. The device-resident `value` of this part is
            # the same call with the host buffers page-locked; there is no device-pointer form of a one-shot call.
            V, F = sphere_surface()
            g = np.random.default_rng(11 + rank)
            lo, hi = V.min(0), V.max(0)
            hQ = pin(torch.from_numpy(0.5 * (lo + hi) + 0.6 * (hi - lo) * (2 * g.random((n, 3)) - 1)))
            hK = pin(torch.empty(n, dtype=torch.uint8))
            L = tw.load_library()
            import ctypes as C
            Vc, Fc = np.ascontiguousarray(V, dtype=np.float64), np.ascontiguousarray(F, dtype=np.uint32)
            retried = C.c_int(0)

            def call():
                rc = L.twg_inout_filter(ctx.h, C.c_void_p(Vc.ctypes.data), C.c_uint32(len(Vc)), C.c_void_p(Fc.ctypes.data), C.c_uint32(len(Fc)),
                                        C.c_void_p(hQ.data_ptr()), C.c_uint64(n), C.c_void_p(hK.data_ptr()), C.byref(retried))
                if rc != 0:
                    raise RuntimeError("twg_inout_filter failed: %d" % rc)
            for _ in range(2):
                call()
            barrier()
            K = min(args.steps, 3)
            Wm = 2
            l0 = ctx.launches
            t0 = time.perf_counter()
            for _ in range(K):
                call()
            t1 = time.perf_counter()
            e2e_s = max_over_ranks(t1 - t0) / K
            ms = kms = e2e_s * 1e3
            launches = ctx.launches - l0
            win = (t0, t1)
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 10_000), replace=False)
                Wo = O.WindingTree(V, F).eval(hQ.numpy()[idx], threads=O.max_threads())
                mism = int(((Wo > 0.5).astype(np.uint8) != hK.numpy()[idx]).sum())
            res.update({"h2d": n * 24 + Vc.nbytes + Fc.nbytes, "d2h": n,
                        "extra": {"decision_mismatches_vs_oracle_10k_sample": mism, "retried": bool(retried.value), "kept_fraction": float(hK.numpy().mean()),
                                  "value_is": "the one-shot host call itself (build + copies + evaluation): identical to e2e"}})
            del hQ, hK
        torch.cuda.empty_cache()
        res["value"] = units_per_step / (ms * 1e-3)
        res["ms_per_step"] = ms
        res["steps"], res["warmup"] = K, Wm
        res["gpu_launches"] = launches
        res["e2e"] = {"value": units_per_step / e2e_s, "unit": UNIT[part], "h2d_bytes_per_step": res.pop("h2d"), "d2h_bytes_per_step": res.pop("d2h")}
        res["clocks"] = Clocks.summarise(clocks.window(*win)) if rank == 0 else None
        n_kernel = (units_per_step // world) if part.endswith("_strong") else n
        res["roofline"] = roofline(part, n_kernel, kms, (res["clocks"] or {}).get("sm_mhz"), roof_extra)
        if rank == 0 and world == 1 and not args.no_cpu:
            # the single-thread and -O3 -march=native modes for the four parts the north star names; the
            # other parts keep the all-cores figure only, so that the default run stays well inside the driver's time limit
            res["cpu_baseline"] = cpu_rate(part, n, O.max_threads(), budget_s=args.cpu_budget, modes=part in ("envelope", "amips", "amips_ring", "winding"))
        results[part] = res
    clocks.stop()
    if rank == 0:
        h = results[head]
        out = {"metric": h["metric"], "value": h["value"], "unit": h["unit"], "n_gpus": world, "steps": h["steps"], "warmup": h["warmup"],
               "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": h["scaling"], "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": h["config"],
               "roofline": h["roofline"], "e2e": h["e2e"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"],
               "cpu_baseline": h.get("cpu_baseline"), "extra": dict(h.get("extra") or {}, host_buffers_not_page_locked_bytes=sum(unpinned)),
               "parts": {p: results[p] for p in parts if p != head}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def load_ncu_traffic():
    """dram bytes per launch of the dominant kernels from the committed ncu summary (profiles/ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--parts", default="auto", help="comma list; auto = every part (+ the strong-scaling parts when launched on more than one rank)")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full BASELINE.json batch sizes (1.0 = as named)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=6.0, help="seconds of CPU work per part for the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.parts == "auto":
        args.parts = "envelope,envelope_faces,envelope_faces_c1,nearest,amips,amips_literal,amips_quality,amips_ring,winding,winding_oneshot,pass_stream"
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            args.parts = args.parts.replace(",winding_oneshot,pass_stream", "") + ",envelope_strong,winding_strong"
    parts = [p for p in args.parts.split(",") if p in FULL]
    if args.impl == "reference":
        run_reference(args, parts)
    else:
        run_gpu(args, parts)


if __name__ == "__main__":
    main()
