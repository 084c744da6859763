//! Raw declarations of include/euc_b200.h (ABI version 1).  Every function returns 0 (EUC_OK) or a negative EUC_E_*.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct euc_ctx { _private: [u8; 0] }
pub type euc_buf = u64; // 0 == euc::Empty
pub type euc_geom = u64;

pub const EUC_OK: c_int = 0;
pub const EUC_E_INVALID: c_int = -1;
pub const EUC_E_SIZE_MISMATCH: c_int = -2;
pub const EUC_E_UNSUPPORTED: c_int = -3;
pub const EUC_E_CUDA: c_int = -4;
pub const EUC_E_OOM: c_int = -5;
pub const EUC_E_OUT_OF_BOUNDS: c_int = -6;

pub const EUC_PIPE_TEAPOT_SHADOW: i32 = 0;
pub const EUC_PIPE_TEAPOT_PHONG: i32 = 1;
pub const EUC_PIPE_TEX_CUBE: i32 = 2;
pub const EUC_PIPE_BLEND_TRIS: i32 = 3;
pub const EUC_PIPE_VOXEL_ICON: i32 = 4;
pub const EUC_PIPE_VERTEX_COLOR: i32 = 5;
pub const EUC_PIPE_WIREFRAME: i32 = 6;

pub const EUC_PRIM_TRIANGLE_LIST: i32 = 0;
pub const EUC_PRIM_LINE_LIST: i32 = 1;
pub const EUC_PRIM_LINE_TRIANGLE_LIST: i32 = 2;

pub const EUC_TEXEL_F32: i32 = 0;
pub const EUC_TEXEL_RGBA8_TO_F32: i32 = 1;
pub const EUC_FILTER_NEAREST: i32 = 0;
pub const EUC_FILTER_LINEAR: i32 = 1;
pub const EUC_WRAP_NONE: i32 = 0;
pub const EUC_WRAP_CLAMP: i32 = 1;
pub const EUC_WRAP_TILE: i32 = 2;
pub const EUC_WRAP_MIRROR: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct euc_sampler_desc { pub buf: euc_buf, pub format: i32, pub filter: i32, pub wrap: i32, pub _pad: i32 }

#[repr(C)]
pub struct euc_pipeline_desc {
    pub pipeline_id: i32, pub primitive_kind: i32, pub cull_mode: i32, pub depth_test: i32,
    pub depth_write: i32, pub pixel_write: i32, pub y_axis_up: i32, pub handedness: i32,
    pub z_clip_enabled: i32, pub z_clip_min: f32, pub z_clip_max: f32, pub msaa_level: i32,
    pub uniforms: *const c_void, pub uniform_bytes: u32, pub _pad: u32,
    pub samplers: [euc_sampler_desc; 2],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct euc_batch_draw { pub first: u32, pub count: u32, pub base_vertex: i32, pub layer: u32 }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct euc_render_stats { pub primitives: u64, pub binned_pairs: u64, pub fragments: u64 }

extern "C" {
    pub fn euc_abi_version() -> c_int;
    pub fn euc_init(device_ordinal: c_int, out_ctx: *mut *mut euc_ctx) -> c_int;
    pub fn euc_shutdown(ctx: *mut euc_ctx) -> c_int;
    pub fn euc_last_error(ctx: *mut euc_ctx) -> *const c_char;
    pub fn euc_set_stream(ctx: *mut euc_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn euc_sync(ctx: *mut euc_ctx) -> c_int;
    pub fn euc_set_stats(ctx: *mut euc_ctx, enabled: c_int) -> c_int;
    pub fn euc_get_stats(ctx: *mut euc_ctx, out: *mut euc_render_stats) -> c_int;
    pub fn euc_buf_create(ctx: *mut euc_ctx, w: u32, h: u32, layers: u32, texel_bytes: u32, out: *mut euc_buf) -> c_int;
    pub fn euc_buf_destroy(ctx: *mut euc_ctx, buf: euc_buf) -> c_int;
    pub fn euc_buf_clear(ctx: *mut euc_ctx, buf: euc_buf, texel: *const c_void) -> c_int;
    pub fn euc_buf_clear_rows(ctx: *mut euc_ctx, buf: euc_buf, texel: *const c_void, row_begin: u32, row_end: u32) -> c_int;
    /// The next render first clears its targets (null = leave alone), fused into its kernels.
    pub fn euc_render_clear(ctx: *mut euc_ctx, pixel_texel: *const c_void, depth_texel: *const c_void) -> c_int;
    pub fn euc_buf_upload(ctx: *mut euc_ctx, buf: euc_buf, host: *const c_void, bytes: usize) -> c_int;
    pub fn euc_buf_download(ctx: *mut euc_ctx, buf: euc_buf, host: *mut c_void, bytes: usize) -> c_int;
    pub fn euc_geom_create(ctx: *mut euc_ctx, vertices: *const c_void, stride: u32, n_vertices: u32,
                           indices: *const u32, n_indices: u32, out: *mut euc_geom) -> c_int;
    pub fn euc_geom_update(ctx: *mut euc_ctx, geom: euc_geom, vertices: *const c_void, indices: *const u32) -> c_int;
    pub fn euc_geom_destroy(ctx: *mut euc_ctx, geom: euc_geom) -> c_int;
    pub fn euc_render(ctx: *mut euc_ctx, desc: *const euc_pipeline_desc, vertices: *const c_void, stride: u32,
                      n_vertices: u32, indices: *const u32, n_indices: u32, pixel: euc_buf, depth: euc_buf) -> c_int;
    pub fn euc_render_geom(ctx: *mut euc_ctx, desc: *const euc_pipeline_desc, geom: euc_geom, pixel: euc_buf, depth: euc_buf) -> c_int;
    pub fn euc_render_geom_rows(ctx: *mut euc_ctx, desc: *const euc_pipeline_desc, geom: euc_geom, pixel: euc_buf,
                                depth: euc_buf, row_begin: u32, row_end: u32) -> c_int;
    pub fn euc_render_batch(ctx: *mut euc_ctx, desc: *const euc_pipeline_desc, geom: euc_geom, draws: *const euc_batch_draw,
                            n_draws: u32, uniforms: *const c_void, pixel: euc_buf, depth: euc_buf) -> c_int;
}
