"""The driver's bench contract, checked on CPU: the reference arm prints exactly one JSON line with the agreed keys,
and the GPU arm's line (built in bench.run_ours) carries every key the contract names."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample", "400"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fastq_to_uq_encode_reads_per_s" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_line_carries_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours"):src.index("def _ref_worker")]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "decode"):
        assert re.search(r'"%s"\s*:' % key, body), key
    for key in ("bound", "achieved", "peak", "frac", "traffic"):
        assert re.search(r'"%s"\s*:' % key, body), "roofline." + key
    for key in ("h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in body
    # the product path never touches the oracle except for the cpu_baseline leg
    assert body.count("from oracle import") == 1
    at = body.index("from oracle import")
    assert "CPU baseline" in body[at - 400:at] and "not args.no_cpu" in body[at - 400:at]
