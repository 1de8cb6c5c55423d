"""bench.py's contract as far as it can be checked without a GPU: the reference arm (`--impl reference`: the CPU port of
the reference's loop on the host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses
to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "img/s"
    assert "DCGAN-64" in d["config"]["workload"] and d["config"]["global_batch"] == 1024 and "bs1024" in base["metric"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert d["vs_baseline"] is None                                  # BASELINE.json publishes no number for this metric
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank_zero_works_under_torchrun_env():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
