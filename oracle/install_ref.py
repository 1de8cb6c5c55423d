"""Recipe for the reference arm: copy the UNMODIFIED reference (flat script collection — no setup.py / pyproject.toml, so
`pip install /root/reference` is impossible) from /root/reference into the git-ignored `baseline/_ref/`,
which travels to the GPU box with the gpurun snapshot (it is NOT in .gpurunignore). Only the files the hot path's
configurations need are taken: `models/`, `utils/` and the three training scripts. Nothing here enters the repository
history, and nothing under baseline/_ref is imported by the product path — only by

  * `bench.py --impl reference` / bench.py's cpu_baseline leg (the reference's own modules timed on the host cores), and
  * tests/test_gpu_reference_scripts.py (the unmodified main_dcgan.py / main_sngan.py executed on top of THIS repo's
    `models` / `utils` packages — the drop-in check of SURVEY.md §8b).

    python oracle/install_ref.py            # no-op when /root/reference is absent (GPU box: uses the shipped copy)
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
WANT = ["models", "utils", "main_dcgan.py", "main_sngan.py", "main_acgan.py"]


def install(verbose=False):
    """Returns the path of the copy, or None when neither the reference nor an earlier copy exists."""
    if not os.path.isdir(SRC):
        return DST if os.path.exists(os.path.join(DST, "main_dcgan.py")) else None
    os.makedirs(DST, exist_ok=True)
    for name in WANT:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copy2(s, d)
        if verbose:
            print("copied", s, "->", d)
    return DST


if __name__ == "__main__":
    print(install(verbose="-v" in sys.argv))
