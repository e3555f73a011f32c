"""bench.py's reference arm runs without a GPU: check the JSON contract of its line (keys the driver reads)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + list(extra),
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    line = _run("--config", "cunet2", "--steps", "1", "--cpu-sample", "1")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] / 1000.0 - 1.0) < 1e-6      # 1-image sample
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "CU-Net-2" in line["config"]["workload"]


def test_reference_arm_other_ranks_and_quantized_config():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""          # only rank 0 runs and prints
    line = _run("--config", "cunet8bin")
    assert line["impl"] == "reference" and "unavailable" in line


def test_reference_arm_stdout_is_one_line():
    """stdout carries exactly the record (bench.py sends everything else, e.g. NCCL's version banner, to stderr)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cunet2",
                          "--steps", "1", "--cpu-sample", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1 and lines[0].startswith("{")


def test_traffic_key_matches_the_kernel_source():
    """bench.py reports `roofline.traffic` only while profiles/traffic.json is keyed to the sha1 of the dominant kernel's
    source; a kernel edit without a fresh `ncu --set full` capture must show up here, not as a silent null in the line."""
    import hashlib
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["conv_bwd1x1"]
    src = os.path.join(ROOT, "cu-net_b200", "csrc", tj["kernel_file"])
    assert hashlib.sha1(open(src, "rb").read()).hexdigest() == tj["kernel_sha1"]
    assert 0.9 * 138412032 < tj["bytes"] < 1.3 * 138412032      # close to the op's algorithmic bytes
