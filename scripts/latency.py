"""Latency of small calls through the C ABI (tools/latency.cpp) next to the oracle's CPU time for the same calls.
    python scripts/latency.py   -> one JSON line (also appended to gpurun_out/latency.jsonl)"""
import json, os, struct, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tetwild_b200 import synth, build
import oracle

def main():
    build.build(); oracle.build()
    V, F = synth.icosphere(5); V = synth.normalise_unit_diag(V)
    sd, eps, eps2 = synth.state_eps(1e-3)
    T = synth.face_queries(V, F, 4096, 0.02, eps, seed=3)
    P = synth.envelope_points(V, F, 16384, eps)
    MV, MT = synth.grid_tet_mesh(40, 40, 40, seed=3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "latency_input.bin")
    with open(path, "wb") as f:
        for a in (V.astype(np.float64), F.astype(np.uint32), T.astype(np.float64), P.astype(np.float64), MV.astype(np.float64), MT.astype(np.int32),
                  np.array([sd, eps2])):
            a = np.ascontiguousarray(a)
            f.write(struct.pack("<Q", a.size)); f.write(a.tobytes())
    exe = os.path.join(ROOT, "tests", "_build", "latency")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "latency.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "tetwild_b200"), "-ltetwild_gpu", "-Wl,-rpath," + os.path.join(ROOT, "tetwild_b200")])
    res = json.loads(subprocess.run([exe, path], capture_output=True, text=True, check=True).stdout)
    # the same calls on ONE host core through the oracle (how the reference's sequential scheduler runs them)
    OS = oracle.Surface(V, F)
    cpu = {}
    for n in (1, 16, 256):
        t = time.perf_counter(); reps = max(3, 2000 // n)
        for _ in range(reps): OS.faces_out(T[:n], sd, eps2, threads=1)
        cpu[str(n)] = (time.perf_counter() - t) / reps * 1e6
    res["cpu_one_core_faces_out_us"] = cpu
    off = np.zeros(MV.shape[0] + 1, dtype=np.int64); np.add.at(off, MT.ravel() + 1, 1); off = np.cumsum(off)
    order = np.argsort(MT.ravel(), kind="stable"); adj = (order // 4).astype(np.int32)
    ids = ((np.arange(1024) * 7919) % len(MV)).astype(np.int32)
    goff = np.concatenate([[0], np.cumsum(off[ids + 1] - off[ids])]).astype(np.uint64)
    tids = np.concatenate([adj[off[v]:off[v + 1]] for v in ids]).astype(np.int32)
    fn = oracle.ref_amips_ring_ejh if oracle.ref_available() else oracle.amips_ring_ejh
    cpu = {}
    for n in (1, 16, 1024):
        g = goff[:n + 1]; t = time.perf_counter(); reps = max(3, 4000 // n)
        for _ in range(reps): fn(MV, MT, g, ids[:n], t_ids=tids[:int(g[-1])], threads=1)
        cpu[str(n)] = (time.perf_counter() - t) / reps * 1e6
    res["cpu_one_core_ring_ejh_us"] = cpu
    res["note"] = "host wall time per call incl. H2D, kernel(s), D2H and the final synchronize; faces at ~208 samples each; one-rings of ~24 tets"
    print(json.dumps(res))
    with open(os.path.join(ROOT, "gpurun_out", "latency.jsonl"), "a") as f: f.write(json.dumps(res) + "\n")

if __name__ == "__main__":
    main()
