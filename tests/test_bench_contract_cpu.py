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


def test_algorithmic_work_matches_the_survey():
    """The FLOP / byte counts behind `roofline.achieved` are SURVEY 8(d)'s per-unit figures times the units of a cfg 2 step."""
    sys.path.insert(0, ROOT)
    import bench as B
    w = B.WORKLOADS[B.DEFAULT_WORKLOAD]
    p = B.params_for(w)
    work = B.family_work(p, w["B"])
    n_img, n_cap = w["B"], w["B"] * B.C
    assert abs(work["conv"][1] / n_img - 30.69e9) < 0.05e9              # 13 convolutions of VGG16, forward, per image
    assert abs((work["conv"][1] + work["fc"][1]) / n_img - 30.932e9) < 0.01e9
    assert abs(work["logits_fwd"][1] / n_cap - 231.69e6) < 0.01e6       # 11.5845 MFLOP x 20 tokens
    assert abs(work["z_rnn"][1] / n_cap - 7.68e6) < 1e3
    assert abs(work["lstm_fwd_step"][1] / n_cap - 135.3e6) < 0.1e6      # 3.1457 MFLOP x (21 encoder + 22 decoder steps)
    assert work["lstm_fwd_seq"] == work["lstm_fwd_step"] and work["lstm_bwd_seq"] == work["lstm_bwd_step"]
    assert abs(work["adam"][1] / 28 - 19.79e6) < 0.01e6                 # 19.79 M non-CNN parameters, 28 B each
    assert abs(work["ce"][1] / (n_cap * B.T) - 2 * 2 * B.V) < 1         # bf16 logits read once, dlogits written once
    # whole step, Normal prior, frozen VGG16 forward: 1.131 GFLOP per caption + 30.932 GFLOP per image
    assert abs(B.total_flops(p, w["B"], True) - (n_cap * 1.131e9 + n_img * 30.932e9)) < 0.005 * n_img * 30.932e9
