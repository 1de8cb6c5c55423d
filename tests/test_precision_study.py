"""tools/precision_study.py is the CPU emulation of the kernels' roundings that the per-pass precision policy was designed
with. Its value rests on its calibration: at the configuration of tests/test_gpu_precision.py (DCGAN-64, width 64, batch
32, same seeds) the emulated gradient cosines must stay where the B200 measured them
(profiles/r01_pytest_gpu_fp16_mode.txt):

    mode      D-real     D-fake     G-step        (measured on the GPU)
    bf16      0.999913   0.993763   0.971050
    fp16      0.999994   0.999404   0.996810
    bf16x3    1.000000   0.999976   0.999873
"""
import contextlib
import io
import os
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))

MEASURED = {"bf16": (0.999913, 0.993763, 0.971050), "fp16": (0.999994, 0.999404, 0.996810),
            "bf16x3": (1.000000, 0.999976, 0.999873)}


@pytest.fixture(scope="module")
def study():
    import precision_study as S
    from gan_playground_b200.models import dcgan
    from oracle import gan_oracle as O

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        netG, netD = dcgan.Generator(), dcgan.Discriminator()
    sd_g = {k: v.clone() for k, v in netG.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netD.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(32, 3, 64, 64, generator=gen) * 2 - 1
    z1, z2 = torch.randn(32, 100, generator=gen), torch.randn(32, 100, generator=gen)
    ref = O.dcgan_step_grads(sd_g, sd_d, x, z1, z2)
    return S, (sd_g, sd_d, x, z1, z2, ref)


@pytest.mark.parametrize("mode", ["bf16", "fp16", "bf16x3"])
def test_emulation_reproduces_the_measured_cosines(study, mode):
    S, args = study
    P = {"bf16": S.policy(S.BF16), "bf16x3": S.policy(S.X3),
         "fp16": S.policy(S.BF16, all=dict(x="h", w="h", y="f32", a="h", col="f32"))}[mode]
    r = S.run(P, *args)
    got = (r["cos D-real"], r["cos D-fake"], r["cos G-step"])
    for g, m in zip(got, MEASURED[mode]):
        # compare the DEFECTS 1 - cos: within a factor of 2 (+ the 2e-4 floor where the backward's own rounding and the
        # run-to-run spread of the GPU's atomics dominate)
        assert abs((1 - g) - (1 - m)) <= 0.5 * (1 - m) + 2e-4, (mode, got, MEASURED[mode])
