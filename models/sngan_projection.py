"""`models.sngan_projection` as imported by the reference's scripts -> B200-native mirror (gan_playground_b200.models.sngan_projection)."""
from gan_playground_b200.models.sngan_projection import *  # noqa: F401,F403
from gan_playground_b200.models.sngan_projection import __dict__ as _d  # noqa: F401

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
