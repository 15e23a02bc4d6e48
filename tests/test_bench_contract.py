"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm
(`--impl reference`: the oracle port of the reference's CPU path on the host cores) prints ONE JSON line with the
agreed keys, under torchrun only rank 0 prints, and the measured arm refuses to run without a B200."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-images", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("img/sec encode (CLIP+entropy+rANS)")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in line, key
    assert line["vs_baseline"] is None and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 3000 < line["config"]["bits_per_img"] < 6000  # 512 symbols of the b005 bottleneck + 32-bit record header


def test_reference_arm_other_ranks_stay_silent():
    r = _run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_measured_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = _run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
