//! Safe wrappers with the call shape of `euc::Pipeline::render` (euc src/pipeline.rs:248-255) for the pipelines of
//! euc's own bench (benches/teapot.rs).  NOT COMPILED in this repository's environment (no rustc).
use euc_b200_sys as sys;
use std::{cmp::Ordering, ffi::CStr, marker::PhantomData};

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }

/// One CUDA device. `Send`, not `Sync`: one host thread at a time (calls are asynchronous on the context's stream).
pub struct Context { raw: *mut sys::euc_ctx }
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::euc_init(device, &mut raw) };
        if rc != sys::EUC_OK { return Err(Error { code: rc, message: "euc_init failed: no CUDA device (there is no CPU fallback)".into() }); }
        Ok(Self { raw })
    }
    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == sys::EUC_OK { return Ok(()); }
        let message = unsafe { CStr::from_ptr(sys::euc_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
    pub fn sync(&self) -> Result<(), Error> { self.check(unsafe { sys::euc_sync(self.raw) }) }
    /// `color.clear(a); depth.clear(b);` of benches/teapot.rs:183-185 attached to the next `render`: the clears are fused into
    /// its kernels (`None` leaves that target alone).
    pub fn render_clear(&self, pixel: Option<u32>, depth: Option<f32>) -> Result<(), Error> {
        let p = pixel.as_ref().map_or(std::ptr::null(), |v| v as *const u32 as *const _);
        let d = depth.as_ref().map_or(std::ptr::null(), |v| v as *const f32 as *const _);
        self.check(unsafe { sys::euc_render_clear(self.raw, p, d) })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::euc_shutdown(self.raw); } } }

/// Device-resident `Buffer2d<T>` with 4-byte texels (euc src/buffer.rs).
pub struct Buffer2d<'c, T> { ctx: &'c Context, handle: sys::euc_buf, size: [usize; 2], _t: PhantomData<T> }

impl<'c, T: Copy> Buffer2d<'c, T> {
    /// `Buffer2d::fill(size, item)` (buffer.rs:60-67)
    pub fn fill(ctx: &'c Context, size: [usize; 2], item: T) -> Result<Self, Error> {
        assert_eq!(std::mem::size_of::<T>(), 4, "4-byte texels only");
        let mut handle = 0;
        ctx.check(unsafe { sys::euc_buf_create(ctx.raw, size[0] as u32, size[1] as u32, 1, 4, &mut handle) })?;
        let mut b = Self { ctx, handle, size, _t: PhantomData };
        b.clear(item)?;
        Ok(b)
    }
    /// `Target::clear` (buffer.rs:213-218)
    pub fn clear(&mut self, item: T) -> Result<(), Error> {
        self.ctx.check(unsafe { sys::euc_buf_clear(self.ctx.raw, self.handle, &item as *const T as *const _) })
    }
    pub fn size(&self) -> [usize; 2] { self.size }
    /// `Buffer::raw()` (buffer.rs:104-107): downloads (blocking)
    pub fn raw(&self) -> Result<Vec<T>, Error> {
        let n = self.size[0] * self.size[1];
        let mut v = Vec::<T>::with_capacity(n);
        self.ctx.check(unsafe { sys::euc_buf_download(self.ctx.raw, self.handle, v.as_mut_ptr() as *mut _, n * 4) })?;
        unsafe { v.set_len(n) };
        Ok(v)
    }
}
impl<T> Drop for Buffer2d<'_, T> { fn drop(&mut self) { unsafe { sys::euc_buf_destroy(self.ctx.raw, self.handle); } } }

/// `wavefront::Vertex` flattened: position + normal (include/euc_b200.h: euc_vertex_pn)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct VertexPn { pub pos: [f32; 3], pub normal: [f32; 3] }

/// POD mirror of the trait getters (pipeline.rs:178-209) for a pipeline of the reference.
fn desc_of<'r, P: euc::Pipeline<'r>>(p: &P, id: i32, uniforms: &[f32]) -> sys::euc_pipeline_desc
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    let dm = p.depth_mode();
    let cm = p.coordinate_mode();
    let cull: euc::CullMode = p.rasterizer_config().into();
    sys::euc_pipeline_desc {
        pipeline_id: id, primitive_kind: sys::EUC_PRIM_TRIANGLE_LIST,
        cull_mode: match cull { euc::CullMode::None => 0, euc::CullMode::Back => 1, euc::CullMode::Front => 2 },
        depth_test: match dm.test { None => 0, Some(Ordering::Less) => 1, Some(Ordering::Equal) => 2, Some(Ordering::Greater) => 3 },
        depth_write: dm.write as i32, pixel_write: p.pixel_mode().write as i32,
        y_axis_up: matches!(cm.y_axis_direction, euc::YAxisDirection::Up) as i32,
        handedness: matches!(cm.handedness, euc::Handedness::Right) as i32,
        z_clip_enabled: cm.z_clip_range.is_some() as i32,
        z_clip_min: cm.z_clip_range.as_ref().map_or(0.0, |r| r.start),
        z_clip_max: cm.z_clip_range.as_ref().map_or(0.0, |r| r.end),
        msaa_level: match p.aa_mode() { euc::AaMode::Msaa { level } => level as i32, _ => 0 },
        uniforms: uniforms.as_ptr() as *const _, uniform_bytes: (uniforms.len() * 4) as u32, _pad: 0,
        samplers: Default::default(),
    }
}

/// `TeapotShadow { mvp }.render(model.vertices(), &mut Empty::default(), &mut shadow)`  (benches/teapot.rs:188-192)
pub fn render_teapot_shadow<'r, P: euc::Pipeline<'r>>(ctx: &Context, pipe: &P, mvp: vek::Mat4<f32>, vertices: &[VertexPn],
                                                     shadow: &mut Buffer2d<f32>) -> Result<(), Error>
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    let u = mvp.into_col_array();
    let d = desc_of(pipe, sys::EUC_PIPE_TEAPOT_SHADOW, &u);
    ctx.check(unsafe { sys::euc_render(ctx.raw, &d, vertices.as_ptr() as *const _, 24, vertices.len() as u32,
                                       std::ptr::null(), 0, 0 /* euc::Empty */, shadow.handle) })
}

/// `Teapot { m, v, p, light_pos, shadow: (&shadow).linear().clamped(), light_vp, cam_pos }.render(.., &mut color, &mut depth)`
/// (benches/teapot.rs:195-204).  `uniforms` = m, v, p, light_vp (column-major), light_pos.xyz_, cam_pos.xyz_ (72 floats).
pub fn render_teapot<'r, P: euc::Pipeline<'r>>(ctx: &Context, pipe: &P, uniforms: &[f32; 72], shadow: &Buffer2d<f32>,
                                              vertices: &[VertexPn], color: &mut Buffer2d<u32>, depth: &mut Buffer2d<f32>) -> Result<(), Error>
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    let mut d = desc_of(pipe, sys::EUC_PIPE_TEAPOT_PHONG, uniforms);
    d.samplers[0] = sys::euc_sampler_desc { buf: shadow.handle, format: sys::EUC_TEXEL_F32, filter: sys::EUC_FILTER_LINEAR, wrap: sys::EUC_WRAP_CLAMP, _pad: 0 };
    ctx.check(unsafe { sys::euc_render(ctx.raw, &d, vertices.as_ptr() as *const _, 24, vertices.len() as u32,
                                       std::ptr::null(), 0, color.handle, depth.handle) })
}

/// `IndexedVertices::new(indices, verts)` for vertices of 32 bytes (euc_vertex_p4uv / euc_vertex_p4c4 / euc_vertex_voxel)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct Vertex32 { pub words: [f32; 8] }

fn indexed_render<'r, P: euc::Pipeline<'r>>(ctx: &Context, d: &sys::euc_pipeline_desc, indices: &[u32], vertices: &[Vertex32],
                                             color: sys::euc_buf, depth: sys::euc_buf) -> Result<(), Error> {
    ctx.check(unsafe { sys::euc_render(ctx.raw, d, vertices.as_ptr() as *const _, 32, vertices.len() as u32, indices.as_ptr(), indices.len() as u32, color, depth) })
}

/// `Cube { mvp, sampler }.render(IndexedVertices::new(INDICES, VERTICES), &mut color, &mut Empty::default())`
/// (examples/texture_mapping.rs:5-35, :146-151).  `texture`: RGBA8 texels uploaded as a `Buffer2d<u32>`; `wrap`: EUC_WRAP_*.
pub fn render_cube<'r, P: euc::Pipeline<'r>>(ctx: &Context, pipe: &P, mvp: vek::Mat4<f32>, texture: &Buffer2d<u32>, filter: i32, wrap: i32,
                                            indices: &[u32], vertices: &[Vertex32], color: &mut Buffer2d<u32>) -> Result<(), Error>
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    let u = mvp.into_col_array();
    let mut d = desc_of(pipe, sys::EUC_PIPE_TEX_CUBE, &u);
    d.samplers[0] = sys::euc_sampler_desc { buf: texture.handle, format: sys::EUC_TEXEL_RGBA8_TO_F32, filter, wrap, _pad: 0 };
    indexed_render::<P>(ctx, &d, indices, vertices, color.handle, 0 /* euc::Empty */)
}

/// BASELINE config 4: pre-transformed rgba triangles, depth test + src-over blend (pipeline id EUC_PIPE_BLEND_TRIS).
pub fn render_blend_tris<'r, P: euc::Pipeline<'r>>(ctx: &Context, pipe: &P, indices: &[u32], vertices: &[Vertex32],
                                                  color: &mut Buffer2d<u32>, depth: &mut Buffer2d<f32>) -> Result<(), Error>
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    let d = desc_of(pipe, sys::EUC_PIPE_BLEND_TRIS, &[]);
    indexed_render::<P>(ctx, &d, indices, vertices, color.handle, depth.handle)
}

/// Device-resident geometry (`euc_geom_create`): the vertex slice + `IndexedVertices` of a frame loop that renders it many times.
pub struct Geometry<'c> { ctx: &'c Context, handle: sys::euc_geom }
impl<'c> Geometry<'c> {
    pub fn new(ctx: &'c Context, vertices: &[Vertex32], indices: &[u32]) -> Result<Self, Error> {
        let mut handle = 0;
        ctx.check(unsafe { sys::euc_geom_create(ctx.raw, vertices.as_ptr() as *const _, 32, vertices.len() as u32, indices.as_ptr(), indices.len() as u32, &mut handle) })?;
        Ok(Self { ctx, handle })
    }
}
impl Drop for Geometry<'_> { fn drop(&mut self) { unsafe { sys::euc_geom_destroy(self.ctx.raw, self.handle); } } }

/// BASELINE config 5: `draws.len()` independent `VoxelIcon { mvp, light_dir }.render(..)` calls, icon i into layer `draws[i].layer`
/// of the layered targets, in one launch sequence.  `uniforms`: 20 floats per icon (mvp column-major, light_dir.xyz_).
pub fn render_voxel_batch<'r, P: euc::Pipeline<'r>>(ctx: &Context, pipe: &P, geom: &Geometry, draws: &[sys::euc_batch_draw], uniforms: &[[f32; 20]],
                                                   color: sys::euc_buf, depth: sys::euc_buf) -> Result<(), Error>
where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
    assert_eq!(draws.len(), uniforms.len());
    let mut d = desc_of(pipe, sys::EUC_PIPE_VOXEL_ICON, &[]);
    d.uniform_bytes = 80;
    ctx.check(unsafe { sys::euc_render_batch(ctx.raw, &d, geom.handle, draws.as_ptr(), draws.len() as u32, uniforms.as_ptr() as *const _, color, depth) })
}

/// One rank of a multi-GPU job on one node (one process or thread per GPU): what `render_par`'s thread pool over row bands
/// (src/pipeline.rs:304-366) becomes with a GPU per band.  Collective calls must be made by every rank in the same order.
pub struct Group<'c> { ctx: &'c Context, pub rank: u32, pub world: u32 }
impl<'c> Group<'c> {
    pub fn join(ctx: &'c Context, name: &str, rank: u32, world: u32) -> Result<Self, Error> {
        let c = std::ffi::CString::new(name).unwrap();
        ctx.check(unsafe { sys::euc_group_create(ctx.raw, c.as_ptr(), rank, world) })?;
        Ok(Self { ctx, rank, world })
    }
    /// Every rank passes its colour target; returns all ranks' targets as mapped into this context (NVLink peer memory).
    pub fn share(&self, color: &Buffer2d<u32>) -> Result<Vec<sys::euc_buf>, Error> {
        let mut peers = vec![0u64; self.world as usize];
        self.ctx.check(unsafe { sys::euc_group_share_buf(self.ctx.raw, color.handle, peers.as_mut_ptr()) })?;
        Ok(peers)
    }
    /// This rank's row band of the frame; the band's colour rows also land in rank 0's target (`gather_all`: in every rank's).
    pub fn render_blend_tris<'r, P: euc::Pipeline<'r>>(&self, pipe: &P, geom: &Geometry, peers: &[sys::euc_buf], depth: &mut Buffer2d<f32>, gather_all: bool) -> Result<(), Error>
    where <<P::Primitives as euc::primitives::PrimitiveKind<P::VertexData>>::Rasterizer as euc::rasterizer::Rasterizer>::Config: Into<euc::CullMode> {
        let d = desc_of(pipe, sys::EUC_PIPE_BLEND_TRIS, &[]);
        let mode = if gather_all { sys::EUC_GATHER_ALL } else { sys::EUC_GATHER_ROOT };
        self.ctx.check(unsafe { sys::euc_group_render(self.ctx.raw, &d, geom.handle, peers.as_ptr(), depth.handle, mode) })
    }
    pub fn barrier(&self) -> Result<(), Error> { self.ctx.check(unsafe { sys::euc_group_barrier(self.ctx.raw) }) }
}
impl Drop for Group<'_> { fn drop(&mut self) { unsafe { sys::euc_group_destroy(self.ctx.raw); } } }
