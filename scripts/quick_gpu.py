"""Quick device-side timing of the three parts (development aid; bench.py is the contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tetwild_b200 as tw
from tetwild_b200 import synth

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
    torch.cuda.init()
    ctx = tw.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())   # a real (non-NULL) stream handle: NULL means "the context's stream"
    s = torch.cuda.current_stream().cuda_stream
    res = {}
    # ---- AMIPS flat SoA, 16M tets
    n = int(os.environ.get("QN_AMIPS", 16_000_000))
    g = torch.Generator(device="cuda").manual_seed(7)
    base = torch.tensor([[0,0,0],[1,0,0],[0.5,3**0.5/2,0],[0.5,3**0.5/6,6**0.5/3]], device="cuda", dtype=torch.float64)
    X = base[None] + 0.15*torch.randn(n,4,3, generator=g, device="cuda", dtype=torch.float64)
    T = X.reshape(n,12).t().contiguous()
    E = torch.empty(n, device="cuda", dtype=torch.float64); J = torch.empty(n,3, device="cuda", dtype=torch.float64); H = torch.empty(n,9, device="cuda", dtype=torch.float64)
    ptrs = [T[k].data_ptr() for k in range(12)]
    t = timeit(lambda: ctx.amips_ejh_soa_dev(ptrs, E.data_ptr(), J.data_ptr(), H.data_ptr(), n, s))
    res["amips_ejh"] = dict(ms=t[0], gtets_s=n/t[0]/1e6, gbs=n*200/t[0]/1e6)
    t = timeit(lambda: ctx.amips_energy_soa_dev(ptrs, E.data_ptr(), n, s))
    res["amips_e"] = dict(ms=t[0], gtets_s=n/t[0]/1e6, gbs=n*104/t[0]/1e6)
    del X, T, E, J, H
    # ---- envelope 10M pts vs 200k tris
    V, F = synth.torus_knot(1000, 100)
    t0 = time.time(); S = tw.Surface(ctx, V, F); res["surface_build_s"] = time.time()-t0
    sd, eps, eps2 = synth.state_eps(1e-3)
    n = int(os.environ.get("QN_ENV", 10_000_000))
    P = synth.envelope_points(V, F, n, eps)
    dP = torch.from_numpy(P).cuda(); dO = torch.empty(n, device="cuda", dtype=torch.uint8)
    for name, e2 in (("env_state_eps", eps2), ("env_1e-3", 1e-6)):
        t = timeit(lambda: S.points_out_dev(dP.data_ptr(), n, e2, dO.data_ptr(), s))
        res[name] = dict(ms=t[0], gpts_s=n/t[0]/1e6, out_frac=float(dO.float().mean()))
    os.environ["TWG_ENVELOPE_SORT"] = "0"
    S2 = tw.Surface(ctx, V, F)
    t = timeit(lambda: S2.points_out_dev(dP.data_ptr(), n, eps2, dO.data_ptr(), s))
    res["env_state_eps_nosort"] = dict(ms=t[0], gpts_s=n/t[0]/1e6)
    del os.environ["TWG_ENVELOPE_SORT"]
    # sorted points
    t = timeit(lambda: S.nearest_dev(dP.data_ptr(), n, 0, 0, torch.empty(n, device="cuda", dtype=torch.float64).data_ptr(), s), iters=3, warm=1)
    res["nearest"] = dict(ms=t[0], gpts_s=n/t[0]/1e6)
    # faces
    nf = 20000
    Tq = synth.face_queries(V, F, nf, 0.05, eps)
    dT = torch.from_numpy(Tq).cuda(); dOf = torch.empty(nf, device="cuda", dtype=torch.uint8)
    t = timeit(lambda: S.faces_out_dev(dT.data_ptr(), nf, sd, eps2, dOf.data_ptr(), s), iters=3, warm=1)
    res["faces"] = dict(ms=t[0], faces_s=nf/t[0]*1e3, out_frac=float(dOf.float().mean()))
    del dP, dO
    # ---- winding 1M tris
    V, F = synth.uv_sphere(708, 708)
    n = int(os.environ.get("QN_WIND", 4_000_000))
    Q = synth.winding_queries(V, n)
    dQ = torch.from_numpy(Q).cuda(); dK = torch.empty(n, device="cuda", dtype=torch.uint8)
    for leaf in os.environ.get("QN_LEAVES", "32,64,128").split(","):
        os.environ["TWG_WINDING_LEAF"] = leaf
        t0 = time.time(); W = tw.Winding(ctx, V, F); bt = time.time()-t0
        t = timeit(lambda: W.eval_dev(dQ.data_ptr(), n, 0, dK.data_ptr(), s), iters=2, warm=1)
        res["winding_leaf%s" % leaf] = dict(ms=t[0], mq_s=n/t[0]/1e3, build_s=bt, inside=float(dK.float().mean()), **W.stats())
        W.close()
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/quick_gpu.json", "w"), indent=1)

if __name__ == "__main__":
    main()
