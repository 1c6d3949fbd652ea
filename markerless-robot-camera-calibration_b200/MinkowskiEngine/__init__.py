"""MinkowskiEngine-compatible subset on B200: hand-written sm_100a CUDA behind a C ABI (libb2me.so).

Drop-in for `import MinkowskiEngine as ME` in model/robotnet_segmentation.py, model/robotnet_vote.py,
model/robotnet_encode.py, model/backbone/minkunet.py and model/backbone/resnet.py of
bcsefercik/markerless-robot-camera-calibration (SURVEY.md §8b lists the symbols). Inference only, CUDA
only: there is no CPU or PyTorch fallback."""
__version__ = "0.5.4+b200.0"

from .core import (SparseTensor, TensorField, SparseTensorQuantizationMode, MinkowskiAlgorithm,  # noqa: F401
                   SparseTensorOperationMode, CoordinateManager, CoordinateMapKey, cat, set_compute_dtype,
                   get_compute_dtype, get_compute_mode, set_tc_operand_path, get_tc_operand_path,
                   set_tc_rot128, set_tc_prefetch, set_fuse_head, reset_launch_count, launch_count, set_profile, set_mask_sort, set_mask_sort_block,
                   set_mask_sort_two_level, set_k3_block_min_rows, set_mask_sort_morton, mask_sorted_perm, tile_masks)
from .nn import *  # noqa: F401,F403
from .nn import MinkowskiModuleBase  # noqa: F401
from . import utils, modules, MinkowskiOps, ops  # noqa: F401
from ._lib import lib as _C, B2MEError  # noqa: F401
