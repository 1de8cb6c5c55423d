"""Drop-in nn.Module mirrors of the reference's DCGAN-family networks (same constructors, forward signatures,
sub-module names and state_dict layout), computing on hand-written sm_100a kernels."""
from . import dcgan  # noqa: F401
from . import acgan, dcgan_blur, dcgan_specnorm, dcgan_specnorm_up, sngan_projection  # noqa: F401
