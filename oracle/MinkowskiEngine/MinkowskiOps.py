from . import MinkowskiLinear, cat  # noqa: F401
