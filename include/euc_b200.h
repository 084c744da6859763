/*
 * euc_b200.h — C ABI of the B200-native raster back end for euc's `Pipeline::render` hot path.
 *
 * Every entry point returns `int`: 0 = EUC_OK, < 0 = EUC_E_*.  A human-readable message for the last
 * failure on a context is available through euc_last_error().  Nothing here unwinds, aborts or prints.
 * Host pointers are borrowed for the duration of the call only.  Device buffers are owned by the
 * context and named by opaque 64-bit handles (0 is never a valid handle: it stands for euc's
 * `Empty` target, reference src/texture.rs:285-319).
 *
 * What each entry point replaces in the reference (paths relative to the euc crate root):
 *
 *   euc_buf_create / euc_buf_destroy   Buffer2d::fill / drop              src/buffer.rs:60-83
 *   euc_buf_clear                      Target::clear                      src/buffer.rs:213-218
 *   euc_buf_upload / euc_buf_download  Buffer::raw_mut / Buffer::raw      src/buffer.rs:104-114
 *   euc_geom_create / euc_geom_destroy the vertex slice + IndexedVertices src/index.rs:4-55
 *   euc_render / euc_render_geom       Pipeline::render                   src/pipeline.rs:248-300
 *                                      (+ render_par :304-366, render_inner :396-614,
 *                                         Triangles::rasterize src/rasterizer/triangles.rs:15-306,
 *                                         Lines::rasterize src/rasterizer/lines.rs:12-120)
 *   euc_render_batch                   a loop of Pipeline::render calls over independent targets
 *   euc_pipeline_desc                  the trait getters pixel_mode/depth_mode/coordinate_mode/aa_mode/
 *                                      rasterizer_config                  src/pipeline.rs:178-209
 *   euc_sampler_desc                   Texture::linear()/nearest() + Sampler::clamped()/tiled()/mirrored()
 *                                                                         src/texture.rs:51-95, src/sampler/mod.rs:44-70
 *
 * The reference is generic over user closures (vertex/fragment/blend).  A device back end cannot call
 * host closures per fragment, so the shader stages of the benchmarked pipelines exist as CUDA device
 * functions selected by `pipeline_id`; their uniform blocks are the POD structs below.
 */
#ifndef EUC_B200_H
#define EUC_B200_H

#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#else /* NVRTC (run-time compiled pipelines, euc_pipeline_register): no system headers */
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long long uintptr_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EUC_B200_ABI_VERSION 2

typedef struct euc_ctx euc_ctx;
typedef uint64_t euc_buf;  /* device Buffer2d handle; 0 == euc `Empty` target */
typedef uint64_t euc_geom; /* device-resident vertex (+ optional index) buffer */

enum euc_status {
    EUC_OK = 0,
    EUC_E_INVALID = -1,       /* bad argument / unknown handle / malformed desc */
    EUC_E_SIZE_MISMATCH = -2, /* reference: assert_eq!(pixel.size(), depth.size()) src/pipeline.rs:262-266 */
    EUC_E_UNSUPPORTED = -3,   /* e.g. width > 20000 (reference divides by zero, src/pipeline.rs:329-330) */
    EUC_E_CUDA = -4,          /* CUDA runtime failure, message is sticky */
    EUC_E_OOM = -5,
    EUC_E_OUT_OF_BOUNDS = -6  /* index >= n_vertices (reference: slice index panic, src/index.rs:53) */
};

/* Shader-stage sets.  Each one is the CUDA restatement of a concrete `impl Pipeline`. */
enum euc_pipeline_id {
    EUC_PIPE_TEAPOT_SHADOW = 0, /* benches/teapot.rs:10-51   depth-only shadow pass            */
    EUC_PIPE_TEAPOT_PHONG = 1,  /* benches/teapot.rs:53-142  Phong + shadow-map lookup         */
    EUC_PIPE_TEX_CUBE = 2,      /* examples/texture_mapping.rs:5-35  sampler lookup            */
    EUC_PIPE_BLEND_TRIS = 3,    /* BASELINE config 4: pre-transformed rgba triangles, src-over */
    EUC_PIPE_VOXEL_ICON = 4,    /* BASELINE config 5: lit voxel meshes, src-over               */
    EUC_PIPE_VERTEX_COLOR = 5,  /* examples/triangle.rs:7-25, examples/spinning_cube.rs:5-29   */
    EUC_PIPE_WIREFRAME = 6,     /* examples/wireframes.rs:5-37  constant-colour LineTriangleList */
    EUC_PIPE_COUNT = 7
};

enum euc_primitive_kind { /* src/primitives.rs:21, :82, :49 */
    EUC_PRIM_TRIANGLE_LIST = 0,
    EUC_PRIM_LINE_LIST = 1,
    EUC_PRIM_LINE_TRIANGLE_LIST = 2
};

enum euc_cull_mode { EUC_CULL_NONE = 0, EUC_CULL_BACK = 1, EUC_CULL_FRONT = 2 }; /* src/rasterizer/mod.rs:10-18 */

/* DepthMode.test : Option<Ordering>  (src/pipeline.rs:14-19) */
enum euc_depth_test { EUC_DEPTH_NONE = 0, EUC_DEPTH_LESS = 1, EUC_DEPTH_EQUAL = 2, EUC_DEPTH_GREATER = 3 };

enum euc_handedness { EUC_HAND_LEFT = 0, EUC_HAND_RIGHT = 1 }; /* stored, never read by the rasterisers */

enum euc_filter { EUC_FILTER_NEAREST = 0, EUC_FILTER_LINEAR = 1 };                      /* src/sampler/{nearest,linear}.rs */
enum euc_wrap { EUC_WRAP_NONE = 0, EUC_WRAP_CLAMP = 1, EUC_WRAP_TILE = 2, EUC_WRAP_MIRROR = 3 }; /* src/sampler/mod.rs:100-179 */

enum euc_texel_format {
    EUC_TEXEL_F32 = 0,         /* Buffer2d<f32> sampled as f32 (shadow map)                               */
    EUC_TEXEL_RGBA8_TO_F32 = 1 /* Buffer2d<[u8;4]>.map(|p| p as f32): 0..255, not normalised (texture.rs:140-176) */
};

typedef struct euc_sampler_desc {
    euc_buf buf; /* device buffer bound as texture; 0 = unbound */
    int32_t format; /* euc_texel_format */
    int32_t filter; /* euc_filter */
    int32_t wrap;   /* euc_wrap */
    int32_t _pad;
} euc_sampler_desc;

#define EUC_MAX_SAMPLERS 2

/* POD mirror of the `Pipeline` trait getters (src/pipeline.rs:178-209) plus the pipeline's uniform state. */
typedef struct euc_pipeline_desc {
    int32_t pipeline_id;    /* euc_pipeline_id */
    int32_t primitive_kind; /* euc_primitive_kind (type Primitives) */
    int32_t cull_mode;      /* rasterizer_config(): euc_cull_mode */
    int32_t depth_test;     /* depth_mode().test: euc_depth_test */
    int32_t depth_write;    /* depth_mode().write */
    int32_t pixel_write;    /* pixel_mode().write */
    int32_t y_axis_up;      /* coordinate_mode().y_axis_direction == Up */
    int32_t handedness;     /* coordinate_mode().handedness */
    int32_t z_clip_enabled; /* coordinate_mode().z_clip_range.is_some() */
    float z_clip_min;       /* inclusive (src/pipeline.rs:151-156) */
    float z_clip_max;       /* inclusive */
    int32_t msaa_level;     /* aa_mode(): 0 = AaMode::None, n = Msaa{level:n}; clamped to 0..6 (src/pipeline.rs:291-294) */
    const void* uniforms;   /* host pointer to the pipeline's uniform block (structs below) */
    uint32_t uniform_bytes;
    uint32_t _pad;
    euc_sampler_desc samplers[EUC_MAX_SAMPLERS];
} euc_pipeline_desc;

/* ---- uniform blocks and vertex layouts (all matrices column-major, like vek::Mat4) ------------------ */

/* EUC_PIPE_TEAPOT_SHADOW: vertex = euc_vertex_pn. */
typedef struct euc_uniforms_teapot_shadow { float mvp[16]; } euc_uniforms_teapot_shadow;

/* EUC_PIPE_TEAPOT_PHONG: vertex = euc_vertex_pn; sampler 0 = shadow map (F32, LINEAR, CLAMP in the bench). */
typedef struct euc_uniforms_teapot_phong {
    float m[16], v[16], p[16], light_vp[16];
    float light_pos[4]; /* xyz used */
    float cam_pos[4];   /* xyz used */
} euc_uniforms_teapot_phong;

/* EUC_PIPE_TEX_CUBE: vertex = euc_vertex_p4uv; sampler 0 = colour texture. */
typedef struct euc_uniforms_tex_cube { float mvp[16]; } euc_uniforms_tex_cube;

/* EUC_PIPE_BLEND_TRIS: vertex = euc_vertex_p4c4 with `pos` already in clip space; no uniforms. */

/* EUC_PIPE_VOXEL_ICON: vertex = euc_vertex_voxel. */
typedef struct euc_uniforms_voxel_icon {
    float mvp[16];
    float light_dir[4]; /* xyz used, unit length */
} euc_uniforms_voxel_icon;

/* EUC_PIPE_VERTEX_COLOR: vertex = euc_vertex_p4c4. */
typedef struct euc_uniforms_vertex_color { float mvp[16]; } euc_uniforms_vertex_color;

/* EUC_PIPE_WIREFRAME: vertex = euc_vertex_pn. */
typedef struct euc_uniforms_wireframe { float m[16], v[16], p[16]; } euc_uniforms_wireframe;

typedef struct euc_vertex_pn { float pos[3]; float normal[3]; } euc_vertex_pn;                     /* 24 B */
typedef struct euc_vertex_p4uv { float pos[4]; float uv[2]; float _pad[2]; } euc_vertex_p4uv;     /* 32 B */
typedef struct euc_vertex_p4c4 { float pos[4]; float rgba[4]; } euc_vertex_p4c4;                  /* 32 B */
typedef struct euc_vertex_voxel { float pos[3]; float normal[3]; uint8_t rgba[4]; uint32_t _pad; } euc_vertex_voxel; /* 32 B */

/* One draw of a batch (euc_render_batch): an index/vertex range of a geom, its own uniforms and targets. */
typedef struct euc_batch_draw {
    uint32_t first;       /* first index (indexed geom) or first vertex (non-indexed) */
    uint32_t count;       /* number of stream vertices; a trailing partial primitive is dropped (pipeline.rs:283) */
    int32_t base_vertex;  /* added to every index */
    uint32_t layer;       /* target layer inside the pixel/depth buffer arrays */
} euc_batch_draw;

/* Per-render statistics (device counters read back on request). */
typedef struct euc_render_stats {
    uint64_t primitives;      /* assembled primitives */
    uint64_t binned_pairs;    /* (tile, primitive) pairs produced by the binner */
    uint64_t fragments;       /* emit_fragment calls (depth-test passes), as the reference counts them */
} euc_render_stats;

/* ---- context ---------------------------------------------------------------------------------------- */
int euc_abi_version(void);
int euc_init(int device_ordinal, euc_ctx** out_ctx);
int euc_shutdown(euc_ctx* ctx);
const char* euc_last_error(euc_ctx* ctx);
/* Run all subsequent work of this context on `cuda_stream` (a cudaStream_t / CUstream); NULL = the context's own stream. */
int euc_set_stream(euc_ctx* ctx, void* cuda_stream);
int euc_sync(euc_ctx* ctx);
/* Asynchronous renders (default: enabled).  A render call then never waits for the device: (tile, primitive) pairs that do
 * not fit the bins sized from earlier renders are handled on the device, and what the host needs to size later renders
 * reaches it through mapped pinned memory.  Consequences for errors that only the device can detect: a vertex index out
 * of range in geometry whose index bounds the host has not seen (euc_geom_wrap; euc_geom_update / euc_render with more
 * than 65536 indices) makes that render draw nothing and is reported as EUC_E_OUT_OF_BOUNDS by the NEXT call on the
 * context (render, euc_sync, download, euc_get_stats).  A scene far denser than the ones its target has seen cannot fail:
 * tiles whose pairs fit neither their bin nor the overflow buffer are rendered by testing every primitive of the render
 * against the tile (slow for that one render; the buffers are sized for it afterwards).  Geometry created with euc_geom_create, small index streams and non-indexed
 * streams are validated on the host and fail in the render call itself, like the reference's slice-index panic
 * (src/index.rs:53).  The first render of a target shape, and every render after euc_set_async(ctx, 0), is checked:
 * the call waits for the set-up kernel's flags (not for the raster work).  A whole frame of asynchronous single-draw
 * renders can be captured into a CUDA graph on the stream given to euc_set_stream. */
int euc_set_async(euc_ctx* ctx, int enabled);
/* Number of render calls so far that had to wait for the device (diagnostics: constant in steady state). */
uint64_t euc_blocking_waits(euc_ctx* ctx);
/* Enable (1) / disable (0) fragment counting; counting costs one atomic per warp per tile. */
int euc_set_stats(euc_ctx* ctx, int enabled);
int euc_get_stats(euc_ctx* ctx, euc_render_stats* out); /* blocking; stats of the last render call */
/* Per-stage device timing (CUDA events on the context's stream around every kernel of a render call). */
enum euc_stage { EUC_STAGE_SETUP = 0, EUC_STAGE_ALLOC = 1, EUC_STAGE_FILL = 2, EUC_STAGE_RESOLVE = 3, EUC_STAGE_RASTER = 4, EUC_STAGE_CLASSIFY = 5, EUC_STAGE_COUNT = 6 };
int euc_set_profiling(euc_ctx* ctx, int enabled);
/* Blocking. ms[EUC_STAGE_COUNT] = accumulated milliseconds per stage since the last reset; calls[] = launches per stage. */
int euc_get_profile(euc_ctx* ctx, float* ms, uint64_t* calls, int reset);
/* Number of kernels this context has launched so far (clears included). */
uint64_t euc_launch_count(euc_ctx* ctx);

/* ---- Buffer2d --------------------------------------------------------------------------------------- */
/* `layers` > 1 creates an array of equally sized targets stored back to back (batch rendering). texel_bytes: 4. */
int euc_buf_create(euc_ctx* ctx, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes, euc_buf* out);
int euc_buf_destroy(euc_ctx* ctx, euc_buf buf);
int euc_buf_clear(euc_ctx* ctx, euc_buf buf, const void* texel);                     /* all layers */
/* Clear rows [row_begin, row_end) of every layer only (row-band rendering: the other rows belong to other ranks). */
int euc_buf_clear_rows(euc_ctx* ctx, euc_buf buf, const void* texel, uint32_t row_begin, uint32_t row_end);
/* Fused clear: the NEXT render call of this context behaves as if its pixel target had first been cleared to *pixel_texel
 * and its depth target to *depth_texel (rows it renders, every layer; NULL = leave that target alone) -- the sequence
 * `color.clear(..); depth.clear(..); pipe.render(..)` of benches/teapot.rs:183-204 in one call.  The tile kernels start
 * from the constants instead of loading the targets and write every tile of the rendered rows, which saves one full
 * write and one full read of both targets per frame.  A target the render does not use (pixel target of a depth-only
 * pass, depth target under DepthMode::NONE) is filled the ordinary way before the render.  The request is consumed by the
 * next euc_render* call, also when that call draws nothing; it is dropped if that call fails validation. */
int euc_render_clear(euc_ctx* ctx, const void* pixel_texel, const void* depth_texel);
int euc_buf_upload(euc_ctx* ctx, euc_buf buf, const void* host, size_t bytes);       /* row-major, x + w*y, layer-major */
int euc_buf_download(euc_ctx* ctx, euc_buf buf, void* host, size_t bytes);           /* blocking */
/* Row N4 (host I/O around the path): pinned host memory and asynchronous read-back, so that a consumer in the style of
 * `win.update_with_buffer(color.raw(), ..)` (examples/teapot.rs:224) can keep several frames in flight.
 * euc_buf_download_async returns at once; *out_ticket is complete when the bytes are in `host` (which should come from
 * euc_host_alloc for a truly asynchronous copy).  euc_ticket_wait blocks on one ticket only, not on the whole stream. */
int euc_host_alloc(euc_ctx* ctx, size_t bytes, void** out_ptr);
int euc_host_free(euc_ctx* ctx, void* ptr);
int euc_buf_download_async(euc_ctx* ctx, euc_buf buf, void* host, size_t bytes, uint64_t* out_ticket);
int euc_ticket_wait(euc_ctx* ctx, uint64_t ticket);
/* Rows [row_begin, row_end) of a single-layer buffer into `host` (tightly packed), asynchronously: a rank of a group
 * reads back the band it rendered. */
int euc_buf_download_rows_async(euc_ctx* ctx, euc_buf buf, void* host, uint32_t row_begin, uint32_t row_end, uint64_t* out_ticket);
int euc_buf_device_ptr(euc_ctx* ctx, euc_buf buf, void** out_ptr, size_t* out_bytes);
int euc_buf_size(euc_ctx* ctx, euc_buf buf, uint32_t* w, uint32_t* h, uint32_t* layers);
/* Wrap caller-owned device memory (e.g. a slice of a collective's receive buffer) as a Buffer2d. Not freed by destroy. */
int euc_buf_wrap(euc_ctx* ctx, void* device_ptr, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes, euc_buf* out);

/* ---- geometry --------------------------------------------------------------------------------------- */
/* indices may be NULL (non-indexed stream). Indices are u32 on the device (the reference uses usize). */
int euc_geom_create(euc_ctx* ctx, const void* vertices, uint32_t vertex_stride, uint32_t n_vertices,
                    const uint32_t* indices, uint32_t n_indices, euc_geom* out);
/* Wrap caller-owned device memory as a geom (e.g. buffers that a collective fills); not freed by destroy. */
int euc_geom_wrap(euc_ctx* ctx, void* device_vertices, uint32_t vertex_stride, uint32_t n_vertices, void* device_indices,
                  uint32_t n_indices, euc_geom* out);
int euc_geom_destroy(euc_ctx* ctx, euc_geom geom);
/* Re-upload vertices (and indices) into an existing geom of the same shape; asynchronous when the host memory is pinned. */
int euc_geom_update(euc_ctx* ctx, euc_geom geom, const void* vertices, const uint32_t* indices);
/* Re-upload a part: `vertices` holds n_vertices vertices that become vertices [first_vertex, ..), `indices` likewise
 * (either pointer may be NULL).  With euc_group_allgather_geom: every rank uploads its own slice only. */
int euc_geom_update_range(euc_ctx* ctx, euc_geom geom, const void* vertices, uint32_t first_vertex, uint32_t n_vertices,
                          const uint32_t* indices, uint32_t first_index, uint32_t n_indices);

/* ---- render ----------------------------------------------------------------------------------------- */
/* Pipeline::render with host geometry: uploads, renders, returns without waiting for the device. */
int euc_render(euc_ctx* ctx, const euc_pipeline_desc* desc, const void* vertices, uint32_t vertex_stride,
               uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices, euc_buf pixel, euc_buf depth);
/* Pipeline::render with device-resident geometry. */
int euc_render_geom(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth);
/* Row-restricted render: only target rows [row_begin, row_end) are produced (screen-space bands for
 * multi-GPU partitioning). The rows written are bit-identical to the same rows of a full render because
 * euc's row bands are independent (src/pipeline.rs:348-350). row_begin must be a multiple of 16. */
int euc_render_geom_rows(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth,
                         uint32_t row_begin, uint32_t row_end);
/* Multi-GPU: export a buffer so that another process on the same node can map it (CUDA IPC), and map one. */
#define EUC_IPC_HANDLE_BYTES 64
int euc_buf_ipc_export(euc_ctx* ctx, euc_buf buf, void* handle_out /* EUC_IPC_HANDLE_BYTES */);
int euc_buf_ipc_import(euc_ctx* ctx, const void* handle, uint32_t width, uint32_t height, uint32_t layers, uint32_t texel_bytes,
                       euc_buf* out);
/* Row-restricted render whose colour rows are ALSO stored, by the raster/resolve kernels themselves, into every
 * `mirrors[i]` (same size and layout as `pixel`; typically peer GPUs' framebuffers mapped with euc_buf_ipc_import):
 * the framebuffer gather is fused into the tile write-back as direct peer stores over NVLink instead of a separate
 * collective.  Rows [row_begin,row_end) of every mirror are fully written (tiles without primitives forward the local
 * colour).  The caller orders consumers after all ranks' renders (any stream-ordered barrier). n_mirrors <= 7. */
#define EUC_MAX_MIRRORS 7
int euc_render_geom_rows_mirrored(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, euc_buf pixel, euc_buf depth,
                                  uint32_t row_begin, uint32_t row_end, const euc_buf* mirrors, uint32_t n_mirrors);
/* ---- multi-GPU: one process (or host thread) per GPU of one node, one context each ------------------------------------
 * The reference spreads the row bands of a frame over a thread pool (render_par, src/pipeline.rs:304-366; bands are
 * independent, :340-362).  Here a band goes to a GPU.  Nothing below needs MPI, NCCL or torch: the ranks meet in a POSIX
 * shared-memory segment named after `name`, exchange CUDA IPC handles there, and synchronise on the device through flags
 * in each other's memory (NVLink / NVSwitch peer stores). */
#define EUC_MAX_GROUP 8
/* Collective over `world` ranks (blocks until all have joined; 120 s time-out).  A context belongs to at most one group. */
int euc_group_create(euc_ctx* ctx, const char* name, uint32_t rank, uint32_t world);
int euc_group_destroy(euc_ctx* ctx); /* collective */
/* Collective: every rank passes one buffer of the same size (created by euc_buf_create, >= 2 MiB); peers_out[r] is rank
 * r's buffer as mapped into this context (peers_out[own rank] == buf).  Stores into a peer's buffer travel over NVLink. */
int euc_group_share_buf(euc_ctx* ctx, euc_buf buf, euc_buf* peers_out /* world entries */);
/* Stream-ordered barrier on the device, no host wait: work queued on this context's stream after the call starts only when
 * every rank's work queued before ITS call has finished and is visible (also peer stores).  Every rank must call it the
 * same number of times.  A rank that never arrives makes the others give up after 10 s (EUC_E_CUDA from a later call). */
int euc_group_barrier(euc_ctx* ctx);
/* The two partitions of a job (pure arithmetic, no context): tile-aligned row bands of one large frame (ranks beyond the
 * last tile row get the empty band [0,0)), and contiguous ranges of independent frames (sizes differ by at most one). */
int euc_group_rows(uint32_t height, uint32_t rank, uint32_t world, uint32_t* row_begin, uint32_t* row_end);
int euc_group_frames(uint32_t n_frames, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end);
/* This rank's row band of one frame (euc_render_geom_rows over euc_group_rows), the band's colour rows being stored by the
 * raster / resolve kernels into the root's framebuffer as well (EUC_GATHER_ROOT: rank 0 ends up with the whole frame, each
 * row crosses NVLink once), into every peer's (EUC_GATHER_ALL), or nowhere else (EUC_GATHER_NONE), followed by
 * euc_group_barrier (not for EUC_GATHER_NONE).  pixel_peers: what euc_group_share_buf returned for the colour target;
 * depth: this rank's own depth target.  A pending euc_render_clear applies to the band.  Collective. */
enum euc_gather { EUC_GATHER_NONE = 0, EUC_GATHER_ROOT = 1, EUC_GATHER_ALL = 2 };
int euc_group_render(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, const euc_buf* pixel_peers, euc_buf depth, int gather);
/* Collective: every rank holds a geometry of the same shape (euc_geom_create; arrays >= 2 MiB) and has uploaded its own
 * 1/world slice of the vertices and of the indices (euc_geom_update_range over euc_group_frames of the two counts); each
 * rank's slice is copied into every peer's geometry over NVLink, then euc_group_barrier.  A frame loop that re-uploads its
 * geometry moves it over PCIe once in total instead of once per GPU. */
int euc_group_allgather_geom(euc_ctx* ctx, euc_geom geom);

/* Row N3 of SURVEY 8(f): pipelines whose shader stages are CUDA source compiled at run time (NVRTC), the device analogue
 * of writing `impl Pipeline for MyShader` in the reference (src/pipeline.rs:171-244).  `source` is placed inside
 * `namespace eucb` after the kernels of this library and must define `struct <struct_name>` with the static interface
 * documented in euc_b200/csrc/shaders.cuh (V, HAS_FRAGMENT, BLEND_IGNORES_OLD, Uniforms, VERTEX_BYTES, vertex(),
 * fragment(), blend()).  It is compiled with --fmad=false for sm_100a like the built-in pipelines.  On success
 * *out_pipeline_id (>= EUC_PIPE_USER_BASE) can be used as euc_pipeline_desc.pipeline_id with this context
 * (TriangleList only).  The compile log of the last call is available through euc_pipeline_log(). */
#define EUC_PIPE_USER_BASE 1000
int euc_pipeline_register(euc_ctx* ctx, const char* source, const char* struct_name, int32_t* out_pipeline_id);
const char* euc_pipeline_log(euc_ctx* ctx);

/* n_draws independent renders in one launch sequence. `uniforms` holds n_draws blocks of
 * desc->uniform_bytes each (desc->uniforms is ignored). Draw i renders into layer draws[i].layer. */
int euc_render_batch(euc_ctx* ctx, const euc_pipeline_desc* desc, euc_geom geom, const euc_batch_draw* draws,
                     uint32_t n_draws, const void* uniforms, euc_buf pixel, euc_buf depth);

#ifdef __cplusplus
}
#endif
#endif /* EUC_B200_H */
