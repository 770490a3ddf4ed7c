"""CPU tier: the reference arm of bench.py (the reference's CPU path timed on host cores) runs without a GPU at a tiny
scale and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.0005", "--steps", "1", "--warmup", "0",
                        "--parts", "envelope,envelope_faces,amips,amips_ring,amips_quality"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "envelope points/s" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert set(d["parts"]) == {"envelope_faces", "amips", "amips_ring", "amips_quality"}
    for p in d["parts"].values():
        assert p["value"] > 0 and p["cpu_baseline"]["value"] == p["value"]


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--parts", "amips", "--scale", "0.001"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
