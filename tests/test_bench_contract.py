"""bench.py's contract as far as it can be checked without a GPU: the reference arm (`--impl reference`: the CPU port of
the reference's loop on the host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses
to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT, env=e)


@pytest.mark.parametrize("cfg,word", [("cfg2", "DCGAN-64"), ("cfg3", "SN-DCGAN"), ("cfg4", "SNGAN projection"), ("cfg5", "ACGAN")])
def test_reference_arm_prints_one_contract_line(cfg, word):
    """--global-batch 32 keeps the CPU suite short; without it the arm times the configuration's full batch (the step the
    product arm runs), which takes ~10 s per step on 16 host cores."""
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "models", "dcgan.py"))
    if cfg == "cfg3" and not have_ref:
        pytest.skip("cfg3's reference arm needs baseline/_ref (oracle/install_ref.py)")
    r = _run("--impl", "reference", "--config", cfg, "--global-batch", "32", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "img/s"
    assert word in d["config"]["workload"] and d["config"]["global_batch"] == 32 and d["config"]["name"] == cfg
    assert "bs1024" in base["metric"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert abs(d["value"] - 32 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]     # the printed step time is the measured one
    assert d["vs_baseline"] is None                                  # BASELINE.json publishes no number for this metric
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "32-image" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_default_batches_follow_baseline_json():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.CONFIGS["cfg2"]["batch"] == 1024 and mod.CONFIGS["cfg2"]["scaling"] == "strong"
    assert mod.CONFIGS["cfg5"]["batch"] == 512 and mod.CONFIGS["cfg5"]["scaling"] == "weak"      # 512 per GPU x 8
    assert abs(mod.CONFIGS["cfg2"]["flops_img"] - 9.7994e9) < 1e6


def test_reference_arm_only_rank_zero_works_under_torchrun_env():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
