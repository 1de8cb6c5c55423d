"""`utils.criterion` as imported by the reference's scripts (main_dcgan.py:12) -> fused-kernel GANLoss."""
from gan_playground_b200.criterion import GANLoss  # noqa: F401
