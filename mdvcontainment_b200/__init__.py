"""mdvcontainment_b200 -- B200-native voxel hot path of MDVContainment behind the reference's API.

    from mdvcontainment_b200 import Containment, VoxelContainment

``Containment(selection, resolution, closing=...)`` and ``VoxelContainment(grid)`` have the signatures
and results of the reference (BartBruininks/mdvcontainment v1.1.1); the array work runs in
hand-written sm_100a CUDA kernels behind the C-ABI of ``include/mdvc.h``.  There is no CPU fallback.
"""
from .containment_main import VoxelContainment  # noqa: F401
from .mda_containment import Containment  # noqa: F401

__version__ = "0.1.0"
