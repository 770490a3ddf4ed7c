"""Envelope / nearest A-B timing on the C2 workload (development aid; bench.py is the contract).
    TWG_ENV_HINT=0|1 python scripts/env_ab.py [n]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tetwild_b200 as tw
from tetwild_b200 import synth
import bench, oracle

def timeit(fn, iters=5, warm=2):
    st = torch.cuda.current_stream()
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    ctx = tw.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    s = torch.cuda.current_stream().cuda_stream
    V, F = synth.torus_knot(1000, 100)
    S = tw.Surface(ctx, V, F)
    sd, eps, eps2 = synth.state_eps(1e-3)
    Ph = bench.envelope_points_fast(V, F, n, eps, seed=20240501)
    P = torch.from_numpy(Ph).cuda()
    O = torch.empty(n, device="cuda", dtype=torch.uint8)
    res = {"policy": os.environ.get("TWG_ENV_POLICY", "1"), "group": os.environ.get("TWG_ENV_GROUP", "64"), "n": n}
    for name, e2 in (("env_state_eps", eps2), ("env_1e-3", 1e-6)):
        t = timeit(lambda: S.points_out_dev(P.data_ptr(), n, e2, O.data_ptr(), s))
        res[name] = dict(ms=t[0], med=t[1], gpts_s=n / t[0] / 1e6, out_frac=float(O.float().mean()))
        m = min(n, 100000)
        idx = np.random.default_rng(1).choice(n, m, replace=False)
        ref = oracle.Surface(V, F).points_out(Ph[idx], e2, threads=16)
        res[name]["mismatch_100k"] = int((O.cpu().numpy()[idx] != ref).sum())
    if os.environ.get("ENV_AB_SKIP_NEAREST"):
        print(json.dumps(res))
        return
    D = torch.empty(n, device="cuda", dtype=torch.float64)
    t = timeit(lambda: S.nearest_dev(P.data_ptr(), n, 0, 0, D.data_ptr(), s), iters=3, warm=1)
    res["nearest"] = dict(ms=t[0], mpts_s=n / t[0] / 1e3)
    m = min(n, 5000)
    idx = np.random.default_rng(2).choice(n, m, replace=False)
    dref = oracle.Surface(V, F).sqdist_brute(Ph[idx], threads=16)[0]
    res["nearest"]["mismatch_5k_vs_brute"] = int((D.cpu().numpy()[idx] != dref).sum())
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/env_ab.jsonl", "a") as f: f.write(json.dumps(res) + "\n")

if __name__ == "__main__":
    main()
