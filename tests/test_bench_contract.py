"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the reference's
own CPU engine, oracle/_ref) prints exactly one JSON line with the agreed keys, the main arm refuses to
run without a CUDA device (there is no CPU fallback), and the algorithmic flops / bytes the roofline
fractions are computed from are the ones of SURVEY.md section 8(d) / Appendix A."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line(oracle_libs):
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"] == "YOLOv2-tiny 416x416 images/sec" and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["global_batch"] == 4096 and "workload" in d["config"]


def test_reference_arm_under_torchrun_env_only_rank0_prints(oracle_libs):
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--ref-sample", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_main_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_algorithmic_work_matches_the_survey():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    assert abs(sum(bench.GFLOP_PER_IMAGE) - 6.9710) < 1e-3
    for i in range(9):     # 2 * OH * OW * K * K * IC * OC
        fl = 2 * bench.HW_IN[i] ** 2 * bench.KS[i] ** 2 * bench.CH[i] * bench.CH[i + 1] / 1e9
        assert abs(fl - bench.GFLOP_PER_IMAGE[i]) < 5e-4, (i, fl)
    # 79.82 MB per image at batch 1 in fp32 (63.4 MB of it weights), 16.60 MB per image at batch 256
    b1 = sum(bench.algorithmic_bytes(i, 1, 4) for i in range(9))
    b256 = sum(bench.algorithmic_bytes(i, 256, 4) for i in range(9)) / 256
    assert abs(b1 / 1e6 - 79.82) < 0.05, b1
    assert abs(b256 / 1e6 - 16.60) < 0.05, b256
    # bf16 activations: layer 0 still reads the fp32 image and layer 8 still writes fp32
    assert bench.algorithmic_bytes(0, 1, 2) == 416 * 416 * 3 * 4 + 208 * 208 * 16 * 2 + 27 * 16 * 2 + 2 * 16 * 4
