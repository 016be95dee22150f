"""welding_robot_b200 — B200-native (sm_100a) ACS hot path of mhsitu/welding_robot.

STL -> bit-packed voxel grid -> rank-based 3-D ant-colony path search -> ant-colony seam
ordering, in hand-written CUDA behind a C ABI (include/wr_gpu.h, lib/libwrgpu.so).
There is no CPU fallback: importing works anywhere, computing needs the CUDA library and a GPU.
"""
from ._lib import AcsParams, WrError, UPDATE_ATOMIC, UPDATE_FUSED, UPDATE_SPLIT  # noqa: F401
from .api import ACS_GTSP, ACS_Rank, Agent, BS_Basic, GridMap, Point3f, STLReader, Vertex3  # noqa: F401

__version__ = "0.1.0"
