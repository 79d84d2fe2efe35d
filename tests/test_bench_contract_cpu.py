"""bench.py contract checks that need no GPU: the reference arm's JSON line (the driver runs `bench.py --impl reference` beside ours
and computes the ratio itself), rank > 0 staying silent under a multi-rank launch, and the product arm refusing to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("audio-seconds/sec") and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["value"] > 0 and line["steps"] == 1 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    # the UNCHANGED reference module when a reference tree is reachable (/root/reference here, oracle/_ref on the GPU box), else the port
    from oracle import refshim

    assert cb["kind"] == ("reference" if refshim.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["config"]["global_batch"] == 32   # the whole batch per step: same configuration as the product arm
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_fails_loudly_without_cuda():
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
