"""The bench line's contract, checked on the one arm that runs without a GPU (`--impl reference`: the reference's algorithm,
oracle port, on the host cores) and, statically, on the keys the GPU arm emits."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--examples", "20000", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout          # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "examples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "examples/sec FFM training" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("c3:") and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "examples per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_emits_every_contract_key():
    """Static check of run_ours(): the keys of the line and of its roofline / e2e / cpu_baseline / clocks objects."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert re.search(r'"%s"\s*:' % key, src), key
    for key in ("bound", "achieved", "peak", "frac", "traffic", "h2d_bytes_per_step", "d2h_bytes_per_step", "cores", "kind", "sample", "sm_mhz", "sm_max_mhz", "reasons"):
        assert re.search(r'"%s"\s*:' % key, src), key
    assert "no CUDA device; the product has no CPU path" in src   # the GPU arm refuses to run without a GPU
