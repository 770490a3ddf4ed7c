#!/bin/bash
# round 2, session 12: where does the one-shot winding call spend its time; then the ncu captures of the final kernels
TAG=r2s12
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/${TAG}_oneshot_phases.log 2>&1 <<'PY'
import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch, tetwild_b200 as tw
from tetwild_b200 import synth
V, F = synth.uv_sphere(708, 708)
n = 100_000_000
g = np.random.default_rng(11)
lo, hi = V.min(0), V.max(0)
hQ = torch.from_numpy(0.5 * (lo + hi) + 0.6 * (hi - lo) * (2 * g.random((n, 3)) - 1)).pin_memory()
hK = torch.empty(n, dtype=torch.uint8).pin_memory()
c = tw.Context(0)
for rep in range(3):
    t0 = time.perf_counter(); W = tw.Winding(c, V, F); t1 = time.perf_counter()
    W.eval(hQ.numpy(), want_w=False, out=(None, hK.numpy())); t2 = time.perf_counter()
    s = int(hK.numpy().sum()); t3 = time.perf_counter()
    W.close(); t4 = time.perf_counter()
    keep, retried = None, None
    t5 = time.perf_counter()
    L = tw.load_library(); r = C.c_int(0)
    Vc, Fc = np.ascontiguousarray(V), np.ascontiguousarray(F, dtype=np.uint32)
    rc = L.twg_inout_filter(c.h, C.c_void_p(Vc.ctypes.data), C.c_uint32(len(Vc)), C.c_void_p(Fc.ctypes.data), C.c_uint32(len(Fc)), C.c_void_p(hQ.data_ptr()), C.c_uint64(n), C.c_void_p(hK.data_ptr()), C.byref(r))
    t6 = time.perf_counter()
    print("rep %d: create %.1f ms, eval %.1f ms, numpy sum %.1f ms, close %.1f ms | twg_inout_filter %.1f ms rc=%d" % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t6-t5)*1e3, rc))
PY
cat gpurun_out/${TAG}_oneshot_phases.log
bash scripts/r2_profile.sh r02
ls gpurun_out/r02_* | head -40
