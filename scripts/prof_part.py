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
        if part == "faces":   # the bench's envelope_faces workload: icosphere, faces of edge ~ diag/50
            V, F = synth.icosphere(5)
            V = synth.normalise_unit_diag(V)
        else:
            V, F = synth.torus_knot(1000, 100)
        S = tw.Surface(ctx, V, F)
        sd, eps, eps2 = synth.state_eps(1e-3)
        if part == "faces":
            n = n or 100000
            T = torch.from_numpy(synth.face_queries(V, F, n, 0.02, eps, seed=3)).cuda()
            O = torch.empty(n, device="cuda", dtype=torch.uint8)
            fn = lambda: S.faces_out_dev(T.data_ptr(), n, sd, eps2, O.data_ptr(), s)
        else:
            n = n or 10_000_000
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import bench
            P = torch.from_numpy(bench.envelope_points_fast(V, F, n, eps, seed=20240501)).cuda()
            O = torch.empty(n, device="cuda", dtype=torch.uint8)
            if part == "envelope":
                fn = lambda: S.points_out_dev(P.data_ptr(), n, eps2, O.data_ptr(), s)
            else:
                D = torch.empty(n, device="cuda", dtype=torch.float64)
                fn = lambda: S.nearest_dev(P.data_ptr(), n, 0, 0, D.data_ptr(), s)
    elif part == "amips":
        import bench
        n = n or 16_000_000
        T = bench.tets_on_device(n, 7, torch.device("cuda", 0))
        E = torch.empty(n, device="cuda", dtype=torch.float64); J = torch.empty(n, 3, device="cuda", dtype=torch.float64); H = torch.empty(n, 9, device="cuda", dtype=torch.float64)
        ptrs = [T[k].data_ptr() for k in range(12)]
        fn = lambda: ctx.amips_ejh_soa_dev(ptrs, E.data_ptr(), J.data_ptr(), H.data_ptr(), n, s)
    elif part == "peaks":
        n = 1
        fn = lambda: print("fp64 TFLOP/s", ctx.measure_fp64_tflops(), "distinct-operand", ctx.measure_fp64_tflops_distinct(), "copy GB/s", ctx.measure_copy_gbs(1 << 30))
    elif part == "ring":
        import bench
        n = n or 16_000_000
        dV, dT4, dOff, dCen = bench.rings_on_device(n, 7, torch.device("cuda", 0))
        nG = dCen.numel()
        E = torch.empty(nG, device="cuda", dtype=torch.float64); J = torch.empty(nG, 3, device="cuda", dtype=torch.float64); H = torch.empty(nG, 9, device="cuda", dtype=torch.float64)
        K = torch.empty(nG, device="cuda", dtype=torch.uint8)
        fn = lambda: ctx.amips_ring_ejh_dev(dV.data_ptr(), dV.shape[0], dT4.data_ptr(), n, 0, dOff.data_ptr(), dCen.data_ptr(), nG, E.data_ptr(), J.data_ptr(), H.data_ptr(), K.data_ptr(), s)
    elif part == "quality":   # the bench's amips_quality workload: C3 indexed layout on the resident mesh
        import bench
        n = n or 15_000_000
        dV, dT4, dOff, dCen = bench.rings_on_device(n, 7, torch.device("cuda", 0))
        M = tw.TetMesh(ctx, dV.cpu().numpy(), dT4.cpu().numpy())
        del dV, dT4
        q = torch.empty(n, device="cuda", dtype=torch.float64)
        fn = lambda: M.quality_dev(0, n, q.data_ptr(), s)
    elif part == "mesh":
        # resident tet mesh: whole-mesh quality + dihedral passes and one-ring Newton terms for every vertex
        g = n or 150
        V, T = synth.grid_tet_mesh(g, g, g, seed=3)
        M = tw.TetMesh(ctx, V, T)
        nT, nV = len(T), len(V)
        q = torch.empty(nT, device="cuda", dtype=torch.float64); a = torch.empty(nT, device="cuda", dtype=torch.float64); b = torch.empty(nT, device="cuda", dtype=torch.float64)
        ids = torch.arange(nV, device="cuda", dtype=torch.int32)
        E = torch.empty(nV, device="cuda", dtype=torch.float64); J = torch.empty(nV, 3, device="cuda", dtype=torch.float64); H = torch.empty(nV, 9, device="cuda", dtype=torch.float64)
        K = torch.empty(nV, device="cuda", dtype=torch.uint8)
        def fn():
            M.quality_dev(0, nT, q.data_ptr(), s)
            M.dihedral_dev(0, nT, a.data_ptr(), b.data_ptr(), s)
            M.vertex_ring_ejh_dev(ids.data_ptr(), nV, E.data_ptr(), J.data_ptr(), H.data_ptr(), K.data_ptr(), s)
        n = nT
    elif part == "winding":
        n = n or 2_000_000
        V, F = synth.uv_sphere(708, 708)
        W = tw.Winding(ctx, V, F)
        Q = torch.from_numpy(synth.winding_queries(V, n, seed=11)).cuda()
        K = torch.empty(n, device="cuda", dtype=torch.uint8)
        fn = lambda: W.eval_dev(Q.data_ptr(), n, 0, K.data_ptr(), s)
    st = torch.cuda.current_stream()
    ts = []
    for _ in range(iters):
        a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(st); fn(); b0.record(st); torch.cuda.synchronize()
        ts.append(a0.elapsed_time(b0))
    print("done %s n=%d iters=%d ms_min=%.4f ms_med=%.4f units_per_s=%.4g (CUDA events; meaningless under ncu)" % (part, n, iters, min(ts), float(np.median(ts)), n / min(ts) * 1e3))

if __name__ == "__main__":
    main()
