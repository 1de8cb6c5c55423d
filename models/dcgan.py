"""`models.dcgan` as imported by the reference's scripts -> B200-native mirror (gan_playground_b200.models.dcgan)."""
from gan_playground_b200.models.dcgan import *  # noqa: F401,F403
from gan_playground_b200.models.dcgan import __dict__ as _d  # noqa: F401

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
