"""nearest-facet work counters (option trace): 8-wide steps, oriented bounds, exact tests and warp rounds per query"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import tetwild_b200 as tw
from tetwild_b200 import synth
import importlib.util
spec = importlib.util.spec_from_file_location("bench", "bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
V, F = b.knot_surface()
sd, eps, eps2 = synth.state_eps(1e-3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
P = b.envelope_points_fast(V, F, n, eps, seed=1)
for what, sel in (("all", slice(None)), ("near", None), ("far", None)):
    c = tw.Context(0)
    c.set_option("trace", 1)
    c.set_option("nearest_mode", 2)   # the round-scheduled kernel carries the counters
    S = tw.Surface(c, V, F)
    if sel is None:
        d = S.squared_distance(P)
        Q = P[d <= 1e-4] if what == "near" else P[d > 1e-4]
    else:
        Q = P
    dev = torch.device("cuda", 0)
    dQ = torch.from_numpy(np.ascontiguousarray(Q)).to(dev)
    dD = torch.empty(len(Q), device=dev, dtype=torch.float64)
    base = [c.debug_counter(k) for k in range(3, 7)]
    torch.cuda.synchronize(); t = time.perf_counter()
    S.nearest_dev(dQ.data_ptr(), len(Q), 0, 0, dD.data_ptr(), 0)
    c.synchronize(); dt = time.perf_counter() - t
    cnt = [c.debug_counter(k) - b0 for k, b0 in zip(range(3, 7), base)]
    print("%-5s n=%d  %.2f ms  per query: steps %.1f bounds %.1f exact %.1f ; warp rounds per 32 queries %.1f" % (what, len(Q), dt * 1e3, cnt[0] / len(Q), cnt[1] / len(Q), cnt[2] / len(Q), cnt[3] * 32.0 / len(Q)))
    S.close(); c.close()
