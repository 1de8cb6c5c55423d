"""`models.dcgan_blur` as imported by the reference's main_dcgan.py:11 -> B200-native mirror."""
from gan_playground_b200.models.dcgan_blur import *  # noqa: F401,F403
from gan_playground_b200.models.dcgan_blur import __dict__ as _d  # noqa: F401

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
