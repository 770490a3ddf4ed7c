"""Wall time of twg_winding_create (device build vs host build) at three surface sizes; `--once` builds the 1.0 M-facet
hierarchy twice (for an ncu launch list of the build's kernels)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, tetwild_b200 as tw
from tetwild_b200 import synth
if "--once" in sys.argv:
    V, F = synth.uv_sphere(708, 708)
    c = tw.Context(0)
    for _ in range(2):
        W = tw.Winding(c, V, F); c.synchronize(); W.close()
    c.close()
    sys.exit(0)
for name, (V, F) in (("sphere 1.0M", synth.uv_sphere(708, 708)), ("sphere 100k", synth.uv_sphere(224, 224)), ("icosphere 20k", synth.icosphere(5))):
    for dev in (1, 0):
        c = tw.Context(0); c.set_option("winding_device_build", dev)
        ts = []
        for _ in range(4):
            t = time.perf_counter(); W = tw.Winding(c, V, F); c.synchronize(); ts.append(time.perf_counter() - t); st = W.stats(); W.close()
        print("%-14s device_build=%d: %.1f ms (first %.1f ms) %s" % (name, dev, min(ts[1:]) * 1e3, ts[0] * 1e3, st))
        c.close()
