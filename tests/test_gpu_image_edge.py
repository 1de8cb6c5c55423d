"""Fused image-edge kernels (csrc/image_edge.cu; SURVEY.md §8 f3: D's first Conv2d and G's last ConvTranspose2d + Tanh of
models/dcgan.py:41-44,106-109 without a column buffer) against torch in true fp32, through the C-ABI."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("check_image_edge", os.path.join(ROOT, "tools", "check_image_edge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.gpu
def test_image_conv_forward_and_weight_gradient_match_torch():
    m = _tool()
    m.case(3, 64, 128, 1)       # DCGAN-64's D block 0 (3 -> 128 at 32x32), odd image count
    m.case(2, 32, 128, 2)       # 32x32 images: eight output rows per tile
    m.case(2, 128, 64, 4)       # 128x128 images: two output rows per tile
    m.case(160, 64, 128, 6)     # more tiles than resident CTAs: persistent loops, store-ring reuse, stage wrap-around
    assert not m.FAILED, m.FAILED


@pytest.mark.gpu
def test_image_conv_transpose_matches_torch():
    m = _tool()
    m.case_t(4, 64, 64, 10)
    m.case_t(3, 32, 128, 11)    # 16-wide rows: two image rows per warp
    m.case_t(160, 64, 128, 14)  # G's last layer of DCGAN-64 (128 -> 3), more tiles than CTAs
    assert not m.FAILED, m.FAILED


@pytest.mark.gpu
def test_unsupported_shapes_are_refused_and_routed_to_the_column_buffer_path():
    m = _tool()
    m.unsupported()
    assert not m.FAILED, m.FAILED
