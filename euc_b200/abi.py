"""ctypes mirror of include/euc_b200.h (POD structs and enums only — no library loading here)."""
import ctypes as C

ABI_VERSION = 2
MAX_SAMPLERS = 2
PIPE_USER_BASE = 1000
IPC_HANDLE_BYTES = 64
MAX_MIRRORS = 7
MAX_GROUP = 8
GATHER_NONE, GATHER_ROOT, GATHER_ALL = range(3)

# enum euc_status
OK, E_INVALID, E_SIZE_MISMATCH, E_UNSUPPORTED, E_CUDA, E_OOM, E_OUT_OF_BOUNDS = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "EUC_OK", -1: "EUC_E_INVALID", -2: "EUC_E_SIZE_MISMATCH", -3: "EUC_E_UNSUPPORTED", -4: "EUC_E_CUDA",
                -5: "EUC_E_OOM", -6: "EUC_E_OUT_OF_BOUNDS"}

# enum euc_pipeline_id
PIPE_TEAPOT_SHADOW, PIPE_TEAPOT_PHONG, PIPE_TEX_CUBE, PIPE_BLEND_TRIS, PIPE_VOXEL_ICON, PIPE_VERTEX_COLOR, PIPE_WIREFRAME = range(7)
# enum euc_primitive_kind
PRIM_TRIANGLE_LIST, PRIM_LINE_LIST, PRIM_LINE_TRIANGLE_LIST = range(3)
# enum euc_cull_mode
CULL_NONE, CULL_BACK, CULL_FRONT = range(3)
# enum euc_depth_test
DEPTH_NONE, DEPTH_LESS, DEPTH_EQUAL, DEPTH_GREATER = range(4)
HAND_LEFT, HAND_RIGHT = range(2)
FILTER_NEAREST, FILTER_LINEAR = range(2)
WRAP_NONE, WRAP_CLAMP, WRAP_TILE, WRAP_MIRROR = range(4)
TEXEL_F32, TEXEL_RGBA8_TO_F32 = range(2)
STAGE_NAMES = ("setup", "alloc", "fill", "resolve", "raster", "classify")


class SamplerDesc(C.Structure):
    _fields_ = [("buf", C.c_uint64), ("format", C.c_int32), ("filter", C.c_int32), ("wrap", C.c_int32), ("_pad", C.c_int32)]


class PipelineDesc(C.Structure):
    _fields_ = [
        ("pipeline_id", C.c_int32), ("primitive_kind", C.c_int32), ("cull_mode", C.c_int32), ("depth_test", C.c_int32),
        ("depth_write", C.c_int32), ("pixel_write", C.c_int32), ("y_axis_up", C.c_int32), ("handedness", C.c_int32),
        ("z_clip_enabled", C.c_int32), ("z_clip_min", C.c_float), ("z_clip_max", C.c_float), ("msaa_level", C.c_int32),
        ("uniforms", C.c_void_p), ("uniform_bytes", C.c_uint32), ("_pad", C.c_uint32),
        ("samplers", SamplerDesc * MAX_SAMPLERS),
    ]


class BatchDraw(C.Structure):
    _fields_ = [("first", C.c_uint32), ("count", C.c_uint32), ("base_vertex", C.c_int32), ("layer", C.c_uint32)]


class RenderStats(C.Structure):
    _fields_ = [("primitives", C.c_uint64), ("binned_pairs", C.c_uint64), ("fragments", C.c_uint64)]


# every symbol include/euc_b200.h declares: name -> (restype, argtypes)
_ctx_p = C.c_void_p
SYMBOLS = {
    "euc_abi_version": (C.c_int, []),
    "euc_init": (C.c_int, [C.c_int, C.POINTER(_ctx_p)]),
    "euc_shutdown": (C.c_int, [_ctx_p]),
    "euc_last_error": (C.c_char_p, [_ctx_p]),
    "euc_set_stream": (C.c_int, [_ctx_p, C.c_void_p]),
    "euc_sync": (C.c_int, [_ctx_p]),
    "euc_set_async": (C.c_int, [_ctx_p, C.c_int]),
    "euc_blocking_waits": (C.c_uint64, [_ctx_p]),
    "euc_set_stats": (C.c_int, [_ctx_p, C.c_int]),
    "euc_get_stats": (C.c_int, [_ctx_p, C.POINTER(RenderStats)]),
    "euc_buf_create": (C.c_int, [_ctx_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_buf_destroy": (C.c_int, [_ctx_p, C.c_uint64]),
    "euc_buf_clear": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p]),
    "euc_buf_clear_rows": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32]),
    "euc_render_clear": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p]),
    "euc_buf_upload": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_size_t]),
    "euc_buf_download": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_size_t]),
    "euc_host_alloc": (C.c_int, [_ctx_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "euc_host_free": (C.c_int, [_ctx_p, C.c_void_p]),
    "euc_buf_download_async": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "euc_ticket_wait": (C.c_int, [_ctx_p, C.c_uint64]),
    "euc_buf_download_rows_async": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_buf_device_ptr": (C.c_int, [_ctx_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "euc_buf_size": (C.c_int, [_ctx_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "euc_buf_wrap": (C.c_int, [_ctx_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_set_profiling": (C.c_int, [_ctx_p, C.c_int]),
    "euc_get_profile": (C.c_int, [_ctx_p, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.c_int]),
    "euc_launch_count": (C.c_uint64, [_ctx_p]),
    "euc_geom_create": (C.c_int, [_ctx_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_geom_wrap": (C.c_int, [_ctx_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_geom_update": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "euc_geom_update_range": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]),
    "euc_geom_destroy": (C.c_int, [_ctx_p, C.c_uint64]),
    "euc_render": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]),
    "euc_render_geom": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_uint64, C.c_uint64, C.c_uint64]),
    "euc_render_geom_rows": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]),
    "euc_buf_ipc_export": (C.c_int, [_ctx_p, C.c_uint64, C.c_void_p]),
    "euc_buf_ipc_import": (C.c_int, [_ctx_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "euc_render_geom_rows_mirrored": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                                C.POINTER(C.c_uint64), C.c_uint32]),
    "euc_group_create": (C.c_int, [_ctx_p, C.c_char_p, C.c_uint32, C.c_uint32]),
    "euc_group_destroy": (C.c_int, [_ctx_p]),
    "euc_group_share_buf": (C.c_int, [_ctx_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "euc_group_barrier": (C.c_int, [_ctx_p]),
    "euc_group_rows": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "euc_group_frames": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "euc_group_allgather_geom": (C.c_int, [_ctx_p, C.c_uint64]),
    "euc_group_render": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64, C.c_int]),
    "euc_pipeline_register": (C.c_int, [_ctx_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]),
    "euc_pipeline_log": (C.c_char_p, [_ctx_p]),
    "euc_render_batch": (C.c_int, [_ctx_p, C.POINTER(PipelineDesc), C.c_uint64, C.POINTER(BatchDraw), C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64]),
}
