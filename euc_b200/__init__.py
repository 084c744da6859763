"""euc_b200 — B200-native raster back end for euc's `Pipeline::render` hot path.

Host-side mirror of the reference's surface (core.py, pipelines.py) over a C-ABI CUDA library
(csrc/libeuc_b200.so, include/euc_b200.h).  There is no CPU fallback: constructing a Context without the
compiled library or without a CUDA device raises."""
from . import abi, io, scenes, vek
from ._lib import EucError, LIB_PATH, load
from .core import (AaMode, Buffer2d, Context, CoordinateMode, CullMode, DepthMode, Empty, Geometry, IndexedVertices,
                   LineList, LineTriangleList, Pipeline, PixelMode, Sampler, TriangleList, default_context)
from .pipelines import (VERTEX_P4C4, VERTEX_P4UV, VERTEX_PN, VERTEX_VOXEL, BlendTris, Cube, Teapot, TeapotShadow,
                        UserPipeline, VertexColor, VoxelIcon, Wireframe)
