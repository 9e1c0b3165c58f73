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


def test_both_arms_build_the_same_config_and_traffic_is_hash_checked():
    """The reference arm and the GPU arm describe the workload with ONE function (the driver compares the two lines' configs),
    and profiles/traffic_<workload>.json is only believed while it carries the hash of the kernel sources the bench runs on."""
    sys.path.insert(0, ROOT)
    import importlib

    bench = importlib.import_module("bench")
    from fwumious_wabbit_b200 import synth

    w = synth.workload("c3")
    a = bench.make_config(w, bench.STEP_EXAMPLES["c3"], 1, False, None)
    b = bench.make_config(synth.workload("c3"), bench.STEP_EXAMPLES["c3"], 1, False, object())
    assert a == b and a["examples_per_step_per_gpu"] == 2_000_000 and a["workload"].startswith("c3:")
    assert "one model" in bench.make_config(w, 1, 8, True, None)["parallelism"] and "replicas x8" in bench.make_config(w, 1, 8, False, None)["parallelism"]
    sha = bench.kernel_source_sha()
    assert re.fullmatch(r"[0-9a-f]{16}", sha) and sha == bench.kernel_source_sha()
    for name in ("c2", "c3", "c4x1", "c4x1_uniform"):
        t = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{name}.json")))
        assert set(t) >= {"dram_bytes_per_example", "kernel_source_sha", "kernel", "examples_in_launch"}
        assert t["dram_bytes_per_example"] > 0 and re.fullmatch(r"[0-9a-f]{16}", t["kernel_source_sha"])


def test_nvlink_counter_parser(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    class R:
        stdout = "GPU 0: NVIDIA B200\n\t Link 0: Data Tx: 100 KiB\n\t Link 0: Data Rx: 40 KiB\n\t Link 1: Data Tx: 28 KiB\n\t Link 1: Data Rx: 2 KiB\n"

    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: R())
    assert bench.nvlink_bytes(0) == (128 * 1024, 42 * 1024)
    R.stdout = "GPU 0: NVIDIA B200\n\t Link 0: Data Tx: N/A\n"
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no nvidia-smi")))
    assert bench.nvlink_bytes(0) is None
