"""The benchmarked `impl Pipeline` types (host side).  Each mirrors a struct of the reference's benches/examples:
its fields are the uniforms, its getters fix the modes, and the shader stages run as CUDA device functions
(euc_b200/csrc/shaders.cuh) selected by `pipeline_id`."""
import numpy as np

from . import abi
from .core import AaMode, CullMode, DepthMode, LineList, LineTriangleList, Pipeline, PixelMode, TriangleList

# vertex layouts (include/euc_b200.h)
VERTEX_PN = np.dtype([("pos", np.float32, 3), ("normal", np.float32, 3)])                                   # 24 B
VERTEX_P4UV = np.dtype([("pos", np.float32, 4), ("uv", np.float32, 2), ("_pad", np.float32, 2)])            # 32 B
VERTEX_P4C4 = np.dtype([("pos", np.float32, 4), ("rgba", np.float32, 4)])                                   # 32 B
VERTEX_VOXEL = np.dtype([("pos", np.float32, 3), ("normal", np.float32, 3), ("rgba", np.uint8, 4), ("_pad", np.uint32)])  # 32 B


def _mat(m):
    """column-major 16 floats, like vek::Mat4 (m is a (4,4) row-indexed numpy matrix: m[row, col])."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T).tobytes()


def _vec4(v):
    out = np.zeros(4, dtype=np.float32)
    out[: len(v)] = np.asarray(v, dtype=np.float32)
    return out.tobytes()


class _ModeMixin(Pipeline):
    """Lets tests/benches override the trait getters per instance (aa=..., depth=..., cull=..., coords=..., pixel=...)."""

    def __init__(self, aa=None, depth=None, cull=None, coords=None, pixel=None, primitives=None):
        self._aa, self._depth, self._cull, self._coords, self._pixel = aa, depth, cull, coords, pixel
        if primitives is not None:
            self.Primitives = primitives  # type Primitives = TriangleList | LineList | LineTriangleList

    def aa_mode(self):
        return self._aa if self._aa is not None else super().aa_mode()

    def depth_mode(self):
        return self._depth if self._depth is not None else self._default_depth()

    def rasterizer_config(self):
        return self._cull if self._cull is not None else self._default_cull()

    def coordinate_mode(self):
        return self._coords if self._coords is not None else super().coordinate_mode()

    def pixel_mode(self):
        return self._pixel if self._pixel is not None else self._default_pixel()

    def _default_depth(self):
        return DepthMode.NONE

    def _default_cull(self):
        return CullMode.Back

    def _default_pixel(self):
        return PixelMode.WRITE


class TeapotShadow(_ModeMixin):
    """benches/teapot.rs:10-51: PixelMode::PASS, DepthMode::LESS_WRITE, CullMode::None."""
    pipeline_id = abi.PIPE_TEAPOT_SHADOW
    vertex_dtype = VERTEX_PN

    def __init__(self, mvp, **kw):
        super().__init__(**kw)
        self.mvp = np.asarray(mvp, dtype=np.float32)

    def _default_depth(self):
        return DepthMode.LESS_WRITE

    def _default_cull(self):
        return CullMode.NONE

    def _default_pixel(self):
        return PixelMode.PASS

    def uniform_block(self):
        return _mat(self.mvp)


class Teapot(_ModeMixin):
    """benches/teapot.rs:53-142: DepthMode::LESS_WRITE, defaults otherwise; shadow = Clamped<Linear<&Buffer2d<f32>>>."""
    pipeline_id = abi.PIPE_TEAPOT_PHONG
    vertex_dtype = VERTEX_PN

    def __init__(self, m, v, p, light_pos, shadow, light_vp, cam_pos, **kw):
        super().__init__(**kw)
        self.m, self.v, self.p, self.light_vp = (np.asarray(x, dtype=np.float32) for x in (m, v, p, light_vp))
        self.light_pos, self.cam_pos, self.shadow = light_pos, cam_pos, shadow

    def _default_depth(self):
        return DepthMode.LESS_WRITE

    def uniform_block(self):
        return _mat(self.m) + _mat(self.v) + _mat(self.p) + _mat(self.light_vp) + _vec4(self.light_pos) + _vec4(self.cam_pos)

    def samplers(self):
        return [self.shadow]


class Cube(_ModeMixin):
    """examples/texture_mapping.rs:5-35: all-default modes (DepthMode::NONE, CullMode::Back, VULKAN)."""
    pipeline_id = abi.PIPE_TEX_CUBE
    vertex_dtype = VERTEX_P4UV

    def __init__(self, mvp, sampler, **kw):
        super().__init__(**kw)
        self.mvp, self.sampler = np.asarray(mvp, dtype=np.float32), sampler

    def uniform_block(self):
        return _mat(self.mvp)

    def samplers(self):
        return [self.sampler]


class BlendTris(_ModeMixin):
    """BASELINE config 4: pre-transformed rgba triangles, LESS_WRITE, CullMode::None, src-over blend."""
    pipeline_id = abi.PIPE_BLEND_TRIS
    vertex_dtype = VERTEX_P4C4

    def _default_depth(self):
        return DepthMode.LESS_WRITE

    def _default_cull(self):
        return CullMode.NONE


class VoxelIcon(_ModeMixin):
    """BASELINE config 5: lit voxel mesh, LESS_WRITE, CullMode::Back, src-over blend."""
    pipeline_id = abi.PIPE_VOXEL_ICON
    vertex_dtype = VERTEX_VOXEL

    def __init__(self, mvp, light_dir, **kw):
        super().__init__(**kw)
        self.mvp, self.light_dir = np.asarray(mvp, dtype=np.float32), light_dir

    def _default_depth(self):
        return DepthMode.LESS_WRITE

    def uniform_block(self):
        return _mat(self.mvp) + _vec4(self.light_dir)


class VertexColor(_ModeMixin):
    """examples/triangle.rs:7-25 (mvp = identity, no depth) and examples/spinning_cube.rs:5-29."""
    pipeline_id = abi.PIPE_VERTEX_COLOR
    vertex_dtype = VERTEX_P4C4

    def __init__(self, mvp=None, **kw):
        super().__init__(**kw)
        self.mvp = np.eye(4, dtype=np.float32) if mvp is None else np.asarray(mvp, dtype=np.float32)

    def uniform_block(self):
        return _mat(self.mvp)


class Wireframe(_ModeMixin):
    """examples/wireframes.rs:5-37: LineTriangleList, constant red fragment, no depth, BGRA pack."""
    pipeline_id = abi.PIPE_WIREFRAME
    vertex_dtype = VERTEX_PN
    Primitives = LineTriangleList

    def __init__(self, m, v, p, **kw):
        super().__init__(**kw)
        self.m, self.v, self.p = (np.asarray(x, dtype=np.float32) for x in (m, v, p))

    def uniform_block(self):
        return _mat(self.m) + _mat(self.v) + _mat(self.p)


class UserPipeline(_ModeMixin):
    """A pipeline whose shader stages were registered at run time (Context.register_pipeline).  `uniforms` is the raw
    uniform block (bytes, laid out like the source's `Uniforms` struct); `samplers` binds up to two textures."""

    def __init__(self, pipeline_id, vertex_dtype, uniforms=b"", samplers=(), **kw):
        super().__init__(**kw)
        self.pipeline_id, self.vertex_dtype = pipeline_id, vertex_dtype
        self._uniforms, self._samplers = bytes(uniforms), list(samplers)

    def uniform_block(self):
        return self._uniforms

    def samplers(self):
        return self._samplers
