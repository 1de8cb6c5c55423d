"""`models.ops` as imported by the reference's models/dcgan_blur.py:5 -> B200-native BlurPool2d."""
from gan_playground_b200.models.ops import *  # noqa: F401,F403
from gan_playground_b200.models.ops import __dict__ as _d  # noqa: F401

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
