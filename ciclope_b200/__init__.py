"""ciclope_b200 -- B200-native (sm_100a) drop-in for ciclope's CT -> CalculiX .INP hot path.

    from ciclope_b200.preprocess import remove_unconnected
    from ciclope_b200.voxelFE import vol2ugrid, mesh2voxelfe, matpropdictionary

mirror ``ciclope.utils.preprocess.{remove_unconnected, remove_largest, fill_voids, segment, add_cap, resample}`` and
``ciclope.core.voxelFE.{vol2ugrid, mesh2voxelfe, matpropdictionary, vol2h5ParOSol}``.  The compute is hand-written CUDA behind the C ABI declared in
``include/ciclope_b200.h`` (``libciclope_b200.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import preprocess, voxelFE  # noqa: E402,F401
from .preprocess import (add_cap, embed, fill_voids, gaussian, mask_image, periosteummask,  # noqa: E402,F401
                         remove_largest, remove_unconnected, resample, segment)
from .voxelFE import matpropdictionary, mesh2voxelfe, vol2h5ParOSol, vol2ugrid  # noqa: E402,F401
