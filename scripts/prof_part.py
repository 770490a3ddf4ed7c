"""One part of the hot path at a profiling-friendly size (for ncu; see profiles/README.md).
    python scripts/prof_part.py envelope|amips|ring|winding|faces [n] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tetwild_b200 as tw
from tetwild_b200 import synth

def main():
    part = sys.argv[1]
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ctx = tw.Context(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    s = torch.cuda.current_stream().cuda_stream
    if part in ("envelope", "faces", "nearest"):
        V, F = synth.torus_knot(1000, 100)
        S = tw.Surface(ctx, V, F)
        sd, eps, eps2 = synth.state_eps(1e-3)
        if part == "faces":
            n = n or 20000
            T = torch.from_numpy(synth.face_queries(V, F, n, 0.05, eps)).cuda()
            O = torch.empty(n, device="cuda", dtype=torch.uint8)
            for _ in range(iters): S.faces_out_dev(T.data_ptr(), n, sd, eps2, O.data_ptr(), s)
        else:
            n = n or 10_000_000
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import bench
            P = torch.from_numpy(bench.envelope_points_fast(V, F, n, eps, seed=20240501)).cuda()
            O = torch.empty(n, device="cuda", dtype=torch.uint8)
            if part == "envelope":
                for _ in range(iters): S.points_out_dev(P.data_ptr(), n, eps2, O.data_ptr(), s)
            else:
                D = torch.empty(n, device="cuda", dtype=torch.float64)
                for _ in range(iters): S.nearest_dev(P.data_ptr(), n, 0, 0, D.data_ptr(), s)
    elif part == "amips":
        import bench
        n = n or 16_000_000
        T = bench.tets_on_device(n, 7, torch.device("cuda", 0))
        E = torch.empty(n, device="cuda", dtype=torch.float64); J = torch.empty(n, 3, device="cuda", dtype=torch.float64); H = torch.empty(n, 9, device="cuda", dtype=torch.float64)
        ptrs = [T[k].data_ptr() for k in range(12)]
        for _ in range(iters): ctx.amips_ejh_soa_dev(ptrs, E.data_ptr(), J.data_ptr(), H.data_ptr(), n, s)
    elif part == "winding":
        n = n or 2_000_000
        V, F = synth.uv_sphere(708, 708)
        W = tw.Winding(ctx, V, F)
        Q = torch.from_numpy(synth.winding_queries(V, n, seed=11)).cuda()
        K = torch.empty(n, device="cuda", dtype=torch.uint8)
        for _ in range(iters): W.eval_dev(Q.data_ptr(), n, 0, K.data_ptr(), s)
    torch.cuda.synchronize()
    print("done", part, n)

if __name__ == "__main__":
    main()
