"""The drop-in boundary executed (north_star: "the main_dcgan.py and main_sngan.py scripts run unchanged"):
the UNMODIFIED reference training scripts — baseline/_ref/main_dcgan.py and main_sngan.py, byte-for-byte copies made by
oracle/install_ref.py — run for one short epoch in a child process with THIS repository's `models` / `utils` packages
first on sys.path, so `from models import dcgan_blur` / `sngan_projection` and `from utils.criterion import GANLoss`
resolve to the B200-native mirrors. Only the dataset is replaced (torchvision's CelebA / MNIST download -> FakeData of the
same image type). The script's own loop, torch.optim.Adam, `.item()` reads, the `netG(fixed_noise)` + save_image
sampling path and its save_model checkpoint all run as written (main_dcgan.py:33-123, main_sngan.py:33-128)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")

_DRIVER = r"""
import os, runpy, sys
ROOT, SCRIPT, KIND = sys.argv[1], sys.argv[2], sys.argv[3]
ARGS = sys.argv[4:]
sys.path.insert(0, ROOT)                       # this repo's `models` / `utils` shims, not the reference's
import torch, torchvision
from torchvision import transforms

class _Fake(torchvision.datasets.FakeData):   # stands in for the dataset download only
    def __init__(self, root=None, split=None, train=True, transform=None, download=False, **kw):
        shape = (3, 72, 72) if KIND == "celeba" else (1, 28, 28)
        super().__init__(size=48, image_size=shape, num_classes=10, transform=transform)

torchvision.datasets.CelebA = _Fake
torchvision.datasets.MNIST = _Fake

class _Loader(torch.utils.data.DataLoader):    # a fixed shuffle: the default one draws its seed from the global CPU generator,
    def __init__(self, *a, **kw):              # which the script advances differently on cuda and cpu (torch.randn(device=...))
        kw["generator"] = torch.Generator().manual_seed(7)
        super().__init__(*a, **kw)

torch.utils.data.DataLoader = _Loader
try:
    from gan_playground_b200 import _lib
    count = _lib.launch_count
except ImportError:                            # the control run: the reference's own models on the CPU
    count = lambda: 0
sys.argv = [SCRIPT] + ARGS
torch.manual_seed(0)                           # same-seed construction == the reference's weights (tests/test_host_logic.py)
runpy.run_path(SCRIPT, run_name="__main__")    # a script FILE: runpy leaves sys.path alone
import models
print("MODELS_FROM", os.path.dirname(os.path.abspath(models.__file__)))
print("NATIVE_LAUNCHES", count())
"""


def _run_script(tmp_path, script, kind, extra, control=False):
    """control=True: the SAME unmodified script on the reference's OWN models, on the CPU (no CUDA device visible) — with
    the same seed it builds the same weights and draws the same first batch, so the first progress line's D(x) (mean
    discriminator logit of that batch, main_dcgan.py:71) is the reference's answer for what the mirrors computed."""
    if not os.path.exists(os.path.join(REF, script)):
        pytest.skip("baseline/_ref/%s absent: run oracle/install_ref.py where /root/reference exists" % script)
    sub = "ctl" if control else "run"
    args = ["--n_epochs", "1", "--batch_size", "16", "--n_workers", "0", "--data_root", str(tmp_path / "data"),
            "--checkpoint_path", str(tmp_path / sub / "ckpt"), "--result_path", str(tmp_path / sub / "res")] + extra
    env = dict(os.environ)
    if control:
        env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-c", _DRIVER, REF if control else ROOT, os.path.join(REF, script), kind] + args,
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, "script failed:\n%s\n%s" % (r.stdout[-1500:], r.stderr[-3000:])
    return r.stdout


def _same_first_iteration(ours, ctl, rel=1e-2):
    """Progress line of iteration 0: Loss_D, Loss_G, D(x), D(G(z))_1, D(G(z))_2. D(x) comes from identical weights and an
    identical first batch: north_star's activation bar. The other four involve noise drawn on the device (the CUDA and CPU
    generators differ), so they agree only statistically: same sign conventions / magnitude."""
    a, b = _check_progress_line(ours), _check_progress_line(ctl)
    assert abs(a[2] - b[2]) <= rel * max(abs(b[2]), 0.05), "D(x): ours %.4f vs the reference's %.4f" % (a[2], b[2])
    assert abs(a[0] - b[0]) <= 0.35 * abs(b[0]) + 0.1, "Loss_D: ours %.4f vs the reference's %.4f" % (a[0], b[0])
    return a, b


def _check_progress_line(out):
    line = [l for l in out.splitlines() if l.startswith("[1/1][0/3]")]
    assert line, out[-800:]
    nums = [float(t) for t in line[0].replace("/", " ").split() if t.replace(".", "").replace("-", "").isdigit() and "." in t]
    assert len(nums) == 5 and all(n == n and abs(n) < 1e3 for n in nums), line      # Loss_D, Loss_G, D(x), D(G(z)) x2
    return nums


def test_unmodified_main_dcgan_runs_on_the_mirrors(tmp_path):
    out = _run_script(tmp_path, "main_dcgan.py", "celeba", ["--ngf", "16", "--ndf", "16", "--save_name", "t"])
    assert "MODELS_FROM %s" % os.path.join(ROOT, "models") in out
    assert int(out.split("NATIVE_LAUNCHES")[1].split()[0]) > 500          # the native kernels did the work
    ours, ref = _same_first_iteration(out, _run_script(tmp_path, "main_dcgan.py", "celeba",
                                                       ["--ngf", "16", "--ndf", "16", "--save_name", "t"], control=True))
    print("main_dcgan.py iteration 0: ours %s | reference models on CPU %s" % (ours, ref))
    assert os.path.exists(tmp_path / "run" / "res" / "t" / "fake_epoch001_0001.jpg")       # netG(fixed_noise) + save_image
    ck = torch.load(tmp_path / "run" / "ckpt" / "t" / "checkpoint_001.pth", map_location="cpu", weights_only=False)
    assert sorted(ck) == ["epoch", "optimizer", "state_dict"] and sorted(ck["state_dict"]) == ["discriminator", "generator"]
    sd = ck["state_dict"]["generator"]
    assert "blocks.0.1.weight" in sd and "blocks.0.2.filt" in sd and sd["blocks.0.3.num_batches_tracked"] > 0
    assert ck["optimizer"]["generator"]["state"][0]["exp_avg"].abs().sum() > 0    # torch.optim.Adam stepped on our grads


def test_unmodified_main_sngan_runs_on_the_mirrors(tmp_path):
    out = _run_script(tmp_path, "main_sngan.py", "mnist",
                      ["--ngf", "16", "--ndf", "16", "--n_disc_update", "1", "--save_name", "t"])
    assert "MODELS_FROM %s" % os.path.join(ROOT, "models") in out
    assert int(out.split("NATIVE_LAUNCHES")[1].split()[0]) > 500
    ours, ref = _same_first_iteration(out, _run_script(tmp_path, "main_sngan.py", "mnist",
                                                       ["--ngf", "16", "--ndf", "16", "--n_disc_update", "1", "--save_name", "t"],
                                                       control=True))
    print("main_sngan.py iteration 0: ours %s | reference models on CPU %s" % (ours, ref))
    assert os.path.exists(tmp_path / "run" / "res" / "t" / "fake_epoch001_0001.jpg")
    ck = torch.load(tmp_path / "run" / "ckpt" / "t" / "checkpoint_001.pth", map_location="cpu", weights_only=False)
    sd = ck["state_dict"]["discriminator"]
    assert "block1.c1.weight_orig" in sd and "block1.c1.weight_u" in sd and "l_y.weight_orig" in sd
