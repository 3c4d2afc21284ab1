"""spvd_b200 -- sm_100a (B200) backend for the SPVD sparse point-voxel denoiser hot path.

Importing the package loads (and if necessary builds) libspvd_b200.so; there is no CPU fallback.
"""
from . import _C, ops, functional  # noqa: F401
from .geometry import Geometry, KernelMap  # noqa: F401
from .sparse import PointTensor, SparseTensor  # noqa: F401
from .denoiser import PRESETS, CondSPVUnet, SPVUnet, build_model  # noqa: F401

__version__ = "0.1.0"
