"""`models.dcgan_specnorm_up` of the reference -> B200-native mirror (gan_playground_b200.models.dcgan_specnorm_up)."""
from gan_playground_b200.models.dcgan_specnorm_up import *  # noqa: F401,F403
from gan_playground_b200.models.dcgan_specnorm_up import __dict__ as _d  # noqa: F401

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
