"""BASELINE config 5 surface (self-intersecting union of 64 noisy spheres, ~2.1 M triangles): the full TetWild run cannot be
built here, so this drives the three hot-path primitives against THAT surface at full size and checks decisions on samples.
    python scripts/c5_stress.py  -> one JSON line (also appended to gpurun_out/c5_stress.jsonl)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tetwild_b200 as tw
from tetwild_b200 import synth
import bench, oracle

def dev_time(fn, iters=3):
    st = torch.cuda.current_stream()
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

def main():
    oracle.build()
    ctx = tw.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream()); s = torch.cuda.current_stream().cuda_stream
    V, F = synth.sphere_union(64, 128, 129)
    V = synth.normalise_unit_diag(V)
    res = {"surface_triangles": int(len(F)), "surface_vertices": int(len(V))}
    sd, eps, eps2 = synth.state_eps(5e-4)          # -e 1/2000
    t = time.perf_counter(); S = tw.Surface(ctx, V, F); ctx.synchronize(); res["surface_build_s"] = time.perf_counter() - t
    n = 5_000_000
    P = bench.envelope_points_fast(V, F, n, eps, seed=5)
    dP = torch.from_numpy(P).cuda(); dO = torch.empty(n, device="cuda", dtype=torch.uint8)
    ms = dev_time(lambda: S.points_out_dev(dP.data_ptr(), n, eps2, dO.data_ptr(), s))
    idx = np.random.default_rng(1).choice(n, 50_000, replace=False)
    OS = oracle.Surface(V, F)
    ref = OS.points_out(P[idx], eps2, threads=oracle.max_threads())
    res["envelope"] = {"points": n, "ms": ms, "gpts_s": n / ms / 1e6, "out_fraction": float(dO.float().mean()),
                       "decision_mismatches_50k": int((dO.cpu().numpy()[idx] != ref).sum())}
    T = synth.face_queries(V, F, 200_000, 0.0125, eps, seed=7)    # -l 1/40: faces of edge ~ diag/80 .. diag/40
    dT = torch.from_numpy(T).cuda(); dOf = torch.empty(len(T), device="cuda", dtype=torch.uint8)
    ms = dev_time(lambda: S.faces_out_dev(dT.data_ptr(), len(T), sd, eps2, dOf.data_ptr(), s))
    fidx = np.random.default_rng(2).choice(len(T), 3000, replace=False)
    fref, ns = OS.faces_out(T[fidx], sd, eps2, threads=oracle.max_threads())
    res["faces"] = {"faces": int(len(T)), "ms": ms, "mfaces_s": len(T) / ms / 1e3, "mean_samples": float(ns.mean()), "out_fraction": float(dOf.float().mean()),
                    "decision_mismatches_3k": int((dOf.cpu().numpy()[fidx] != fref).sum())}
    del S, dP, dO, dT, dOf
    t = time.perf_counter(); W = tw.Winding(ctx, V, F); res["winding_build_s"] = time.perf_counter() - t
    res["winding_hierarchy"] = W.stats()
    nq = 10_000_000
    Q = synth.winding_queries(V, nq, seed=11)
    dQ = torch.from_numpy(Q).cuda(); dK = torch.empty(nq, device="cuda", dtype=torch.uint8); dW = torch.empty(nq, device="cuda", dtype=torch.float64)
    ms = dev_time(lambda: W.eval_dev(dQ.data_ptr(), nq, dW.data_ptr(), dK.data_ptr(), s))
    widx = np.random.default_rng(3).choice(nq, 600, replace=False)
    Wd = oracle.winding_direct(V, F, Q[widx], threads=oracle.max_threads())
    Wg = dW.cpu().numpy()
    res["winding"] = {"queries": nq, "ms": ms, "mqueries_s": nq / ms / 1e3, "inside_fraction": float(dK.float().mean()), "max_W": float(Wg.max()),
                      "max_abs_err_vs_direct_sum_600": float(np.abs(Wg[widx] - Wd).max()),
                      "decision_mismatches_600": int(((Wd > 0.5).astype(np.uint8) != dK.cpu().numpy()[widx]).sum())}
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "c5_stress.jsonl"), "a") as f: f.write(json.dumps(res) + "\n")

if __name__ == "__main__":
    main()
