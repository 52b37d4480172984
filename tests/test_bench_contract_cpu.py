"""bench.py's reference arm on CPU (the arm the driver launches as `bench.py --impl reference ...`): one JSON line with
the contract's keys, `impl` = reference, a cpu_baseline describing the run, zero-byte e2e; ranks other than 0 print
nothing and exit 0. The native arm must refuse to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    r = run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--ref-batch", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"].startswith("captions/sec") and d["unit"] == "captions/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "cfg2_vgg_normal_b256" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "2 images" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present")
def test_native_arm_refuses_to_run_on_cpu():
    r = run(["--gpus", "1", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
