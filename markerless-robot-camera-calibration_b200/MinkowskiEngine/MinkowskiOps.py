"""ME.MinkowskiOps alias module (model/robotnet_segmentation.py:44 uses ME.MinkowskiOps.MinkowskiLinear)."""
from .nn import MinkowskiLinear  # noqa: F401
from .core import cat  # noqa: F401
