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

FULL = {"envelope": 10_000_000, "amips": 50_000_000, "amips_ring": 50_000_000, "winding": 100_000_000, "envelope_faces": 400_000, "nearest": 10_000_000, "amips_quality": 50_000_000}
UNIT = {"envelope": "points/s", "amips": "tets/s", "amips_ring": "tets/s", "winding": "queries/s", "envelope_faces": "faces/s", "nearest": "points/s", "amips_quality": "tets/s"}
METRIC = {"envelope": "envelope points/s", "amips": "AMIPS E+J+H tet-evals/s", "amips_ring": "AMIPS one-ring E+J+H tet-evals/s",
          "winding": "winding-number queries/s", "envelope_faces": "envelope faces/s (isFaceOutEnvelop)", "nearest": "nearest-facet projections/s", "amips_quality": "AMIPS tet-quality evals/s (calTetQualities)"}
FACE_EDGE = 0.02  # C1-shaped candidate faces: small enough that the flat face stays within eps of the curved icosphere about half the time
WORKLOAD = {
    "envelope": "C2: %d sampled points vs 200000-triangle (2,3) torus knot, eps_rel=1e-3 -> eps_2=(0.42265e-3)^2 (State.cpp:36-41)",
    "amips": "C3: %d random non-degenerate tets, flat SoA (12 arrays), E+J[3]+H[9] per tet, FP64",
    "amips_ring": "C3 smoothing-candidate layout: %d random non-degenerate tets in one-rings of k~U{12..36} around a centre vertex (indexed gather, centre rotated to slot 0), E+J[3]+H[9] per ring (NewtonsUpdate), FP64",
    "winding": "C4: %d centroids uniform in 1.2x bbox vs 1001112-triangle closed noisy UV sphere, keep = W > 0.5",
    "amips_quality": "C3 indexed layout: calTetQualities over %d random non-degenerate tets (int4 tet -> 4 gathered vertices, exact orientation gate, energy, MAX_ENERGY rules of LocalOperations.cpp:862-884) on the resident tet mesh, FP64",
    "nearest": "C2 points, full nearest search: %d points vs 200000-triangle torus knot -> nearest facet id + nearest point + d2 (nearest_facet, mesh_AABB.h:130-176; the projection callers VertexSmoother.cpp:354-362, Preprocess.cpp:529)",
    "envelope_faces": "C1-shaped call stream: %d candidate faces (edge ~ diag/50, sampled on the device at sampling_dist = 1e-3 diag like Common.cpp:143-255) vs the 20480-triangle icosphere, eps through State.cpp:36-41",
}
# algorithmic HBM bytes per unit (ring: 16 B indices + 72 B gathered vertices + 128 B of per-ring data / 24;
# quality: 16 B tet + 96 B gathered vertices + 8 B out)
ALG_BYTES = {"envelope": 25.0, "amips": 200.0, "amips_ring": 93.0, "winding": 25.0, "envelope_faces": 73.0, "nearest": 60.0, "amips_quality": 120.0}


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
def cpu_rate(part, n_full, threads, budget_s=8.0):
    """Times the reference CPU path of one part on a bounded sample; returns (units/s, kind, sample description)."""
    import oracle as O
    from tetwild_b200 import synth
    O.build()
    have_ref = O.ref_available()
    if part == "envelope":
        V, F = knot_surface()
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        if have_ref:
            RT = O.RefTree(V, F[S.order()])
            fn = lambda P: RT.points_out(P, eps2, threads=threads)[0]  # noqa: E731
            kind, what = "reference", "reference mesh_AABB.cpp facet_in_envelope_with_hint (compiled unmodified; geogram leaf distance restated)"
        else:
            fn = lambda P: S.points_out(P, eps2, threads=threads)  # noqa: E731
            kind, what = "port", "oracle port of mesh_AABB.cpp:482-548"
        P = envelope_points_fast(V, F, 200_000, eps, seed=99)
        t = time.perf_counter(); fn(P); r0 = len(P) / (time.perf_counter() - t)
        m = int(min(n_full, max(200_000, r0 * budget_s)))
        P = envelope_points_fast(V, F, m, eps, seed=20240501)
        t = time.perf_counter(); fn(P); dt = time.perf_counter() - t
        return m / dt, kind, "%d of %d points, %s, OpenMP over queries" % (m, n_full, what)
    if part == "amips":
        fn = (lambda T: O.ref_amips_ejh_soa(T, threads=threads)) if have_ref else (lambda T: O.amips_ejh_soa(T, threads=threads))
        kind = "reference" if have_ref else "port"
        what = "reference LocalOperations.cpp:28-291 text (compiled unmodified, -O2)" if have_ref else "oracle port (forward-mode AD)"
        T = synth.random_tets(200_000, seed=1)
        t = time.perf_counter(); fn(T); r0 = T.shape[1] / (time.perf_counter() - t)
        m = int(min(n_full, 20_000_000, max(200_000, r0 * budget_s)))
        T = synth.random_tets(m, seed=7)
        t = time.perf_counter(); fn(T); dt = time.perf_counter() - t
        return m / dt, kind, "%d of %d tets, %s, OpenMP over tets" % (m, n_full, what)
    if part == "amips_quality":
        m = int(min(n_full, 4_000_000))
        V, tets, off, cen = synth.ring_groups(max(1, m // 24), seed=7, scale_lo=0.1, scale_hi=10.0)
        O.amips_quality(V, tets[:20000], threads=threads)
        t = time.perf_counter(); O.amips_quality(V, tets, threads=threads); dt = time.perf_counter() - t
        return len(tets) / dt, "port", ("%d of %d tets, oracle port of calTetQuality_AMIPS (exact orientation predicate + the energy of "
                                        "LocalOperations.cpp:28-81), OpenMP over tets" % (len(tets), n_full))
    if part == "amips_ring":
        # NewtonsUpdate over one-rings (VertexSmoother.cpp:627-702): per member tet the reference's own E, J, H text
        m = int(min(n_full, 4_000_000))
        V, tets, off, cen = synth.ring_groups(max(1, m // 24), seed=7, scale_lo=0.1, scale_hi=10.0)
        nt = int(off[-1])
        fn = O.ref_amips_ring_ejh if have_ref else O.amips_ring_ejh
        kind = "reference" if have_ref else "port"
        fn(V, tets, off[:1001], cen[:1000], threads=threads)
        t = time.perf_counter(); fn(V, tets, off, cen, threads=threads); dt = time.perf_counter() - t
        what = "NewtonsUpdate restated around the reference's own LocalOperations.cpp:28-291 E/J/H text (oracle/ref_wrap.cpp)" if have_ref else "oracle port of NewtonsUpdate"
        return nt / dt, kind, "%d of %d tets in %d one-rings, %s, OpenMP over rings" % (nt, n_full, len(cen), what)
    if part == "nearest":
        V, F = knot_surface()
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        if have_ref:
            RT = O.RefTree(V, F[S.order()])
            fn = lambda P: RT.nearest(P, threads=threads)  # noqa: E731
            kind, what = "reference", "reference mesh_AABB.cpp nearest_facet (compiled unmodified; geogram leaf distance restated)"
        else:
            fn = lambda P: S.nearest(P, threads=threads)  # noqa: E731
            kind, what = "port", "oracle port of mesh_AABB.cpp:418-480"
        P = envelope_points_fast(V, F, 50_000, eps, seed=99)
        t = time.perf_counter(); fn(P); r0 = len(P) / (time.perf_counter() - t)
        m = int(min(n_full, max(50_000, r0 * budget_s)))
        P = envelope_points_fast(V, F, m, eps, seed=20240501)
        t = time.perf_counter(); fn(P); dt = time.perf_counter() - t
        return m / dt, kind, "%d of %d points, %s, OpenMP over queries" % (m, n_full, what)
    if part == "envelope_faces":
        V, F = synth.icosphere(5)
        V = synth.normalise_unit_diag(V)
        sd, eps, eps2 = synth.state_eps(1e-3)
        S = O.Surface(V, F)
        if have_ref:  # the loop of LocalOperations.cpp:1046-1109 around the reference's own sampleTriangle, DistanceQuery.h and tree
            RT = O.RefTree(V, F[S.order()])
            fn = lambda T: RT.faces_out(T, sd, eps2, threads=threads)  # noqa: E731
            kind, what = "reference", "reference sampleTriangle (Common.cpp:143-255) + DistanceQuery.h + mesh_AABB.cpp compiled unmodified, composed by the loop of LocalOperations.cpp:1046-1109 (oracle/ref_wrap.cpp)"
        else:
            fn = lambda T: S.faces_out(T, sd, eps2, threads=threads)  # noqa: E731
            kind, what = "port", "oracle port of isFaceOutEnvelop_sampling (LocalOperations.cpp:1046-1109)"
        T = synth.face_queries(V, F, 2000, FACE_EDGE, eps, seed=99)
        t = time.perf_counter(); fn(T); r0 = len(T) / (time.perf_counter() - t)
        m = int(min(n_full, max(2000, r0 * budget_s)))
        T = synth.face_queries(V, F, m, FACE_EDGE, eps, seed=3)
        t = time.perf_counter(); fn(T); dt = time.perf_counter() - t
        return m / dt, kind, "%d of %d faces, %s, first OUT sample stops the face, OpenMP over faces" % (m, n_full, what)
    if part == "winding":
        V, F = sphere_surface()
        WT = O.WindingTree(V, F)
        Q = synth.winding_queries(V, 20_000, seed=3)
        t = time.perf_counter(); WT.eval(Q, threads=threads); r0 = len(Q) / (time.perf_counter() - t)
        m = int(min(n_full, max(20_000, r0 * budget_s)))
        Q = synth.winding_queries(V, m, seed=11)
        t = time.perf_counter(); WT.eval(Q, threads=threads); dt = time.perf_counter() - t
        return m / dt, "port", "%d of %d queries, oracle port of libigl's exact winding-number hierarchy (libigl not vendored), OpenMP over queries (hierarchy build excluded)" % (m, n_full)
    raise ValueError(part)


def run_reference(args, parts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    threads = O.max_threads()
    lines = {}
    for part in parts:
        n_full = max(1000, int(FULL[part] * args.scale))
        rates = []
        for _ in range(max(1, min(args.steps, 3))):
            r, kind, sample = cpu_rate(part, n_full, threads, budget_s=6.0)
            rates.append(r)
        v = float(np.median(rates))
        lines[part] = {"metric": METRIC[part], "value": v, "unit": UNIT[part], "ms_per_step": None,
                       "cpu_baseline": {"value": v, "unit": UNIT[part], "cores": threads, "kind": kind, "sample": sample},
                       "e2e": {"value": v, "unit": UNIT[part], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                       "config": {"workload": WORKLOAD[part] % n_full}}
    head = parts[0]
    out = {"impl": "reference", "metric": METRIC[head], "value": lines[head]["value"], "unit": UNIT[head], "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": lines[head]["config"],
           "cpu_baseline": lines[head]["cpu_baseline"], "e2e": lines[head]["e2e"], "gpu_launches": 0,
           "parts": {p: lines[p] for p in parts if p != head}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, parts):
    import torch
    import torch.distributed as dist
    import tetwild_b200 as tw
    from tetwild_b200 import synth

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
    copy_gbs = ctx.measure_copy_gbs(1 << 30)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    K, Wm = args.steps, args.warmup

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
            self.gath = [[torch.empty(n, device=dev, dtype=torch.uint8) for _ in range(world)] for _ in range(2)] if world > 1 else None
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
                self.work[p] = dist.all_gather(self.gath[p], self.buf[p], async_op=True)
            self.k += 1

        def drain(self):
            for p in range(2):
                if self.work[p] is not None:
                    self.work[p].wait()
                    self.work[p] = None

        def last(self):
            return self.buf[(self.k - 1) & 1]

    def timed(step_fn, gather_fn=None, drain_fn=None):
        """W warm-up steps, then K steps bracketed by barrier+sync; device time (CUDA events on the launching stream),
        max over ranks. Returns (ms_per_step, kernel_ms_avg, launches, clock window)."""
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
        return total / K, kern, ctx.launches - l0, (t0, t1)

    def e2e_timed(call):
        for _ in range(min(Wm, 2)):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return max_over_ranks(dt) / K

    results = {}
    import oracle as O  # cpu_baseline leg only (rank 0, N = 1)
    for part in parts:
        n = max(1000, int(FULL[part] * args.scale))
        res = {"metric": METRIC[part], "unit": UNIT[part], "config": {"workload": WORKLOAD[part] % n}}
        if part == "envelope":
            V, F = knot_surface()
            sd, eps, eps2 = synth.state_eps(1e-3)
            S = tw.Surface(ctx, V, F)
            P = envelope_points_fast(V, F, n, eps, seed=20240501 + rank)
            hP = pin(torch.from_numpy(P))
            dP = hP.to(dev, non_blocking=True)
            pipe = Pipe(n)
            step = lambda: S.points_out_dev(dP.data_ptr(), n, eps2, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win = timed(step, pipe.gather, pipe.drain)
            dO = pipe.last()
            hO = pin(torch.empty(n, dtype=torch.uint8))
            e2e_s = e2e_timed(lambda: S.points_out(hP.numpy(), eps2, out=hO.numpy()))
            out_frac = float(dO.float().mean().item())
            # parity inside the bench: a 100k sample of this very batch against the oracle (decisions must be identical)
            idx = np.random.default_rng(5).choice(n, min(n, 100_000), replace=False)
            mism = None
            if rank == 0:
                OS = O.Surface(V, F)
                mism = int((OS.points_out(P[idx], eps2, threads=O.max_threads()) != dO.cpu().numpy()[idx]).sum())
            res.update({"h2d": n * 24, "d2h": n, "extra": {"out_of_envelope_fraction": out_frac, "decision_mismatches_vs_oracle_100k_sample": mism,
                                                           "surface_triangles": int(len(F))}})
            res["config"]["l2"] = "inputs larger than L2: 240 MB of points streamed per step; the 38 MB surface structure is meant to stay L2-resident"
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
            ms, kms, launches, win = timed(step)
            hF, hN, hD = pin(torch.empty(n, dtype=torch.int32)), pin(torch.empty((n, 3), dtype=torch.float64)), pin(torch.empty(n, dtype=torch.float64))
            outs = (hF.numpy().view(np.uint32), hN.numpy(), hD.numpy())
            e2e_s = e2e_timed(lambda: S.nearest(hP.numpy(), out=outs))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 20_000), replace=False)
                dref = O.Surface(V, F).sqdist_brute(P[idx], threads=O.max_threads())[0]
                mism = int((dD.cpu().numpy()[idx] != dref).sum())
            res.update({"h2d": n * 24, "d2h": n * 36, "extra": {"d2_mismatches_vs_brute_force_20k_sample": mism, "surface_triangles": int(len(F))}})
            res["config"]["l2"] = "inputs larger than L2: 240 MB of points in, 360 MB of results out per step"
            del dP, dF, dN, dD, S
        elif part == "envelope_faces":
            V, F = synth.icosphere(5)
            V = synth.normalise_unit_diag(V)
            sd, eps, eps2 = synth.state_eps(1e-3)
            S = tw.Surface(ctx, V, F)
            T = synth.face_queries(V, F, n, FACE_EDGE, eps, seed=3 + rank)
            hT = pin(torch.from_numpy(T))
            dTr = hT.to(dev, non_blocking=True)
            pipe = Pipe(n)
            step = lambda: S.faces_out_dev(dTr.data_ptr(), n, sd, eps2, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win = timed(step, pipe.gather, pipe.drain)
            dO = pipe.last()
            e2e_s = e2e_timed(lambda: S.faces_out(hT.numpy(), sd, eps2))
            mism = nsamp = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 5000), replace=False)
                ref, cnt = O.Surface(V, F).faces_out(T[idx], sd, eps2, threads=O.max_threads())
                mism = int((ref != dO.cpu().numpy()[idx]).sum())
                nsamp = float(np.mean(cnt))
            res.update({"h2d": n * 72, "d2h": n, "extra": {"out_of_envelope_fraction": float(dO.float().mean().item()),
                                                           "decision_mismatches_vs_oracle_5k_sample": mism,
                                                           "mean_samples_per_face_sampleTriangle": nsamp, "surface_triangles": int(len(F))}})
            res["config"]["l2"] = "L2 flushed by construction: every step re-reads %.0f MB of faces; the 4 MB surface structure stays L2-resident" % (n * 72 / 1e6)
            del dTr, dO, S, pipe
        elif part == "amips":
            dT = tets_on_device(n, 7 + rank, dev)
            dE = torch.empty(n, device=dev, dtype=torch.float64)
            dJ = torch.empty((n, 3), device=dev, dtype=torch.float64)
            dH = torch.empty((n, 9), device=dev, dtype=torch.float64)
            ptrs = [dT[k].data_ptr() for k in range(12)]
            step = lambda: ctx.amips_ejh_soa_dev(ptrs, dE.data_ptr(), dJ.data_ptr(), dH.data_ptr(), n, sh)  # noqa: E731
            ms, kms, launches, win = timed(step)
            hT = pin(torch.empty((12, n), dtype=torch.float64))
            hT.copy_(dT)
            hE, hJ, hH = (pin(torch.empty(s, dtype=torch.float64)) for s in ((n,), (n, 3), (n, 9)))
            e2e_s = e2e_timed(lambda: ctx.amips_ejh_soa(hT.numpy(), out=(hE.numpy(), hJ.numpy(), hH.numpy())))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 50_000), replace=False)
                Ts = np.ascontiguousarray(hT.numpy()[:, idx])
                ref = O.ref_amips_ejh_soa(Ts, threads=O.max_threads()) if O.ref_available() else O.amips_ejh_soa(Ts, threads=O.max_threads())
                got = (dE.cpu().numpy()[idx], dJ.cpu().numpy()[idx], dH.cpu().numpy()[idx])
                X = Ts.T.reshape(-1, 4, 3)
                l2 = sum(((X[:, a] - X[:, b]) ** 2).sum(1) for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))) / 6.0
                eE = np.abs(got[0] - ref[0]) / np.abs(ref[0])
                eJ = np.abs(got[1] - ref[1]).max(1) / np.maximum(np.abs(ref[1]).max(1), np.abs(ref[0]) / np.sqrt(l2))
                eH = np.abs(got[2] - ref[2]).max(1) / np.maximum(np.abs(ref[2]).max(1), np.abs(ref[0]) / l2)
                mism = {"max_rel_err_E": float(eE.max()), "max_rel_err_J": float(eJ.max()), "max_rel_err_H": float(eH.max()),
                        "over_1e-9": int(((eE > 1e-9) | (eJ > 1e-9) | (eH > 1e-9)).sum()), "sample": int(len(idx))}
            res.update({"h2d": n * 96, "d2h": n * 104, "extra": {"parity_vs_reference_text": mism}})
            res["config"]["l2"] = "inputs larger than L2: 4.8 GB read + 5.2 GB written per step"
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
            ms, kms, launches, win = timed(step)
            hQ = pin(torch.empty(n, dtype=torch.float64))
            e2e_s = e2e_timed(lambda: M.quality(out=hQ.numpy()))
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
            res["config"]["l2"] = "inputs larger than L2: %.1f GB of vertices + %.1f GB of tets gathered per step" % (nV * 24 / 1e9, n * 16 / 1e9)
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
            ms, kms, launches, win = timed(step)
            hV, hT4, hOff, hCen = pin(dV.cpu()), pin(dT4.cpu()), pin(dOff.cpu()), pin(dCen.cpu())
            ship_s = e2e_timed(lambda: ctx.amips_ring_ejh(hV.numpy(), hT4.numpy(), hOff.numpy().view(np.uint64), hCen.numpy()))
            # the integration the scheduler uses: the tet mesh is RESIDENT on the device (uploaded once,
            # kept in step by scatter updates), a Newton batch ships 4 B of vertex id per ring in and 105 B per ring out
            t0 = time.perf_counter()
            M = tw.TetMesh(ctx, hV.numpy(), hT4.numpy())
            M.build_rings()
            mesh_build_s = time.perf_counter() - t0
            hE, hJ, hH = (pin(torch.empty(sz, dtype=torch.float64)) for sz in ((nG,), (nG, 3), (nG, 9)))
            hOk = pin(torch.empty(nG, dtype=torch.uint8))
            outs = (hE.numpy(), hJ.numpy(), hH.numpy(), hOk.numpy())
            e2e_s = e2e_timed(lambda: M.vertex_ring_ejh(hCen.numpy(), out=outs))
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
            res["config"]["l2"] = "inputs larger than L2: %.1f GB of vertices + %.1f GB of indices gathered per step" % (nV * 24 / 1e9, n * 16 / 1e9)
            del dV, dT4, dOff, dCen, dE, dJ, dH, dOk, hV, hT4, hOff, hCen
        elif part == "winding":
            V, F = sphere_surface()
            t0 = time.perf_counter()
            Wt = tw.Winding(ctx, V, F)
            build_s = time.perf_counter() - t0
            g = torch.Generator(device=dev).manual_seed(11 + rank)
            lo, hi = torch.tensor(V.min(0), device=dev), torch.tensor(V.max(0), device=dev)
            dQ = (0.5 * (lo + hi) + 0.6 * (hi - lo) * (2 * torch.rand((n, 3), generator=g, device=dev, dtype=torch.float64) - 1)).contiguous()
            pipe = Pipe(n)
            step = lambda: Wt.eval_dev(dQ.data_ptr(), n, 0, pipe.out().data_ptr(), sh)  # noqa: E731
            ms, kms, launches, win = timed(step, pipe.gather, pipe.drain)
            dK = pipe.last()
            hQ = pin(torch.empty((n, 3), dtype=torch.float64))
            hQ.copy_(dQ)
            hK = pin(torch.empty(n, dtype=torch.uint8))
            e2e_s = e2e_timed(lambda: Wt.eval(hQ.numpy(), want_w=False, out=(None, hK.numpy())))
            mism = None
            if rank == 0:
                idx = np.random.default_rng(5).choice(n, min(n, 20_000), replace=False)
                Wo = O.WindingTree(V, F).eval(hQ.numpy()[idx], threads=O.max_threads())
                mism = int(((Wo > 0.5).astype(np.uint8) != dK.cpu().numpy()[idx]).sum())
            res.update({"h2d": n * 24, "d2h": n, "extra": {"inside_fraction": float(dK.float().mean().item()), "hierarchy_build_s": build_s,
                                                           "decision_mismatches_vs_oracle_20k_sample": mism, **Wt.stats()}})
            res["config"]["l2"] = "inputs larger than L2: 2.4 GB of queries per step; the ~210 MB hierarchy is re-read from L2/HBM"
            del dQ, dK, Wt, pipe
        torch.cuda.empty_cache()
        total_units = n * world
        res["value"] = total_units / (ms * 1e-3)
        res["ms_per_step"] = ms
        res["gpu_launches"] = launches
        res["e2e"] = {"value": total_units / e2e_s, "unit": UNIT[part], "h2d_bytes_per_step": res.pop("h2d"), "d2h_bytes_per_step": res.pop("d2h")}
        ach = ALG_BYTES[part] * n / (kms * 1e-3) / 1e9
        res["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                           "peak_source": peak_src, "kernel_ms": kms, "algorithmic_bytes_per_unit": ALG_BYTES[part],
                           "fp64_dfma_peak_tflops_measured_in_run": fp64_peak, "copy_gbs_measured_in_run": copy_gbs,
                           "kernel_ms_scope": "CUDA events around ALL launches of one step on the launching stream (for the point / nearest / winding parts that includes the Morton sort of the batch), so `achieved` is a lower bound for the traversal kernel alone"}
        res["clocks"] = Clocks.summarise(clocks.window(*win)) if rank == 0 else None
        if rank == 0 and world == 1 and not args.no_cpu:
            v, kind, sample = cpu_rate(part, n, O.max_threads(), budget_s=args.cpu_budget)
            res["cpu_baseline"] = {"value": v, "unit": UNIT[part], "cores": O.max_threads(), "kind": kind, "sample": sample}
        results[part] = res
    clocks.stop()
    if rank == 0:
        head = parts[0]
        h = results[head]
        traffic = load_ncu_traffic()
        for p in parts:
            t = traffic.get(p)
            if t:  # dram bytes per launch measured by ncu at t["units"] units, scaled to this launch's size
                n_p = max(1000, int(FULL[p] * args.scale))
                results[p]["roofline"]["traffic"] = t["dram_bytes"] * n_p / t["units"]
                results[p]["roofline"]["traffic_source"] = "%s (ncu --set full, dram__bytes_read+write, n=%d, scaled per unit)" % (t["file"], t["units"])
                if "fp64_pipe_pct" in t:
                    results[p]["roofline"]["fp64_pipe_active_pct"] = t["fp64_pipe_pct"]
        out = {"metric": h["metric"], "value": h["value"], "unit": h["unit"], "n_gpus": world, "steps": K, "warmup": Wm,
               "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": dict(h["config"], parallelism="replicated surface, batches split by rank, NCCL all_gather of decisions overlapped with the next step's kernels (double-buffered)" if world > 1 else "single GPU"),
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
    ap.add_argument("--parts", default="envelope,envelope_faces,nearest,amips,amips_quality,amips_ring,winding")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full BASELINE.json batch sizes (1.0 = as named)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=8.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    parts = [p for p in args.parts.split(",") if p in FULL]
    if args.impl == "reference":
        run_reference(args, parts)
    else:
        run_gpu(args, parts)


if __name__ == "__main__":
    main()
