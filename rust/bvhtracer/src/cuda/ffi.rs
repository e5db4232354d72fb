//! Raw bindings of `include/bvht.h` (ABI version 4).  Layouts are `#[repr(C)]` mirrors of the C structs.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const BVHT_OK: c_int = 0;
pub const BVHT_FLAG_STRICT: u32 = 0x0;
pub const BVHT_FLAG_FAST: u32 = 0x1;
pub const BVHT_FLAG_LEAF_ACCEL: u32 = 0x2;
pub const BVHT_FLAG_STAMP_INSTANCE: u32 = 0x4;

pub const BVHT_SHADE_DEPTH: u32 = 1;
pub const BVHT_SHADE_INTERSECTION: u32 = 2;
pub const BVHT_SHADE_UV: u32 = 3;
pub const BVHT_SHADE_NORMAL: u32 = 4;
pub const BVHT_SHADE_TEXTURE: u32 = 5;

pub enum BvhtCtx {}

// bvht_bvh_node / bvht_tlas_node: defined next to the private node types they flatten (reference-accessors.patch)
pub use crate::model::BvhtBvhNode;
pub use crate::scene::BvhtTlasNode;
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct BvhtInstance { pub transform_inv: [f32; 16], pub blas_id: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct BvhtCamera { pub top_left_eye: [f32; 3], pub top_right_eye: [f32; 3], pub bottom_left_eye: [f32; 3], pub view_matrix_inv: [f32; 16] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct BvhtRay { pub origin: [f32; 3], pub direction: [f32; 3], pub t: f32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct BvhtHit { pub t: f32, pub u: f32, pub v: f32, pub id: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct BvhtRect { pub x0: u32, pub y0: u32, pub x1: u32, pub y1: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct BvhtShade {
    pub kind: u32, pub depth_scale: f32, pub depth_offset: f32,
    pub hit_rgba: [u8; 4], pub miss_rgba: [u8; 4], pub object0_transform: [f32; 16],
}

impl BvhtShade {
    fn base(kind: u32) -> Self {
        Self { kind, depth_scale: 0.0, depth_offset: 0.0, hit_rgba: [0; 4], miss_rgba: [0; 4], object0_transform: [0.0; 16] }
    }
    /// `DepthAccumulator::new()` + `DepthMappingShader::new(scale, offset)`
    pub fn depth(scale: f32, offset: f32) -> Self { Self { depth_scale: scale, depth_offset: offset, ..Self::base(BVHT_SHADE_DEPTH) } }
    /// `IntersectionAccumulator` + `IntersectionShader::new(hit, miss)`
    pub fn intersection(hit: [u8; 4], miss: [u8; 4]) -> Self { Self { hit_rgba: hit, miss_rgba: miss, ..Self::base(BVHT_SHADE_INTERSECTION) } }
    /// `UvMappingAccumulator` + `RadianceToRgbShader`
    pub fn uv() -> Self { Self::base(BVHT_SHADE_UV) }
    /// `NormalMappingAccumulator` + `RadianceToRgbShader` (object0_transform is filled per frame)
    pub fn normal() -> Self { Self::base(BVHT_SHADE_NORMAL) }
    /// `TextureMaterialAccumulator` + `RadianceToRgbShader` (needs tex coords + texture of scene object 0's model)
    pub fn texture() -> Self { Self::base(BVHT_SHADE_TEXTURE) }
}

extern "C" {
    pub fn bvht_abi_version() -> c_int;
    pub fn bvht_create(device: c_int, flags: u32, out: *mut *mut BvhtCtx) -> c_int;
    pub fn bvht_destroy(ctx: *mut BvhtCtx);
    pub fn bvht_last_error(ctx: *const BvhtCtx) -> *const c_char;
    pub fn bvht_status_string(status: c_int) -> *const c_char;
    /// scheduling overrides (1 = coverage raster, 2 = K0, 3 = bands); never changes a result; value < 0 = the library's rule
    pub fn bvht_set_option(ctx: *mut BvhtCtx, option: u32, value: i32) -> c_int;
    pub fn bvht_blas_create(ctx: *mut BvhtCtx, tris: *const f32, n_tris: u32, nodes: *const BvhtBvhNode, nodes_used: u32, out_id: *mut u32) -> c_int;
    pub fn bvht_blas_build(ctx: *mut BvhtCtx, tris: *const f32, n_tris: u32, out_id: *mut u32) -> c_int;
    pub fn bvht_blas_rebuild(ctx: *mut BvhtCtx, id: u32) -> c_int;
    pub fn bvht_blas_info(ctx: *mut BvhtCtx, id: u32, n_tris: *mut u32, nodes_used: *mut u32) -> c_int;
    pub fn bvht_blas_read_triangles(ctx: *mut BvhtCtx, id: u32, out: *mut f32, n_tris: u32) -> c_int;
    pub fn bvht_blas_read_permutation(ctx: *mut BvhtCtx, id: u32, out: *mut u32, n_tris: u32) -> c_int;
    pub fn bvht_blas_set_normals(ctx: *mut BvhtCtx, id: u32, normals: *const f32, n_tris: u32) -> c_int;
    pub fn bvht_blas_set_tex_coords(ctx: *mut BvhtCtx, id: u32, tex_coords: *const f32, n_tris: u32) -> c_int;
    pub fn bvht_blas_set_texture(ctx: *mut BvhtCtx, id: u32, rgb: *const u8, width: u32, height: u32) -> c_int;
    pub fn bvht_blas_update_vertices(ctx: *mut BvhtCtx, id: u32, tris: *const f32, n_tris: u32) -> c_int;
    pub fn bvht_blas_refit(ctx: *mut BvhtCtx, id: u32) -> c_int;
    pub fn bvht_blas_read_nodes(ctx: *mut BvhtCtx, id: u32, out: *mut BvhtBvhNode, max_nodes: u32) -> c_int;
    pub fn bvht_tlas_set(ctx: *mut BvhtCtx, nodes: *const BvhtTlasNode, nodes_used: u32, inst: *const BvhtInstance, n_inst: u32) -> c_int;
    /// `for o in objects { o.set_transform(..) }; tlas.rebuild(objects)` computed on the device (column-major `Transform3` matrices).
    pub fn bvht_scene_set_transforms(ctx: *mut BvhtCtx, transforms: *const f32, blas_ids: *const u32, n_inst: u32) -> c_int;
    /// The TLAS / cached inverses / world bounds the device computed, in the reference's layouts (any output may be null).
    pub fn bvht_tlas_read(ctx: *mut BvhtCtx, nodes_out: *mut BvhtTlasNode, max_nodes: u32, nodes_used_out: *mut u32,
                          inst_out: *mut BvhtInstance, bounds_out: *mut f32, max_inst: u32, n_inst_out: *mut u32) -> c_int;
    pub fn bvht_render_frame(ctx: *mut BvhtCtx, cam: *const BvhtCamera, width: u32, height: u32, tile: u32, region: BvhtRect,
                             shade: *const BvhtShade, frame_out: *mut u32, hits_out: *mut BvhtHit) -> c_int;
    /// bvht_render_frame with up to two frames in flight: `begin` queues the frame and returns, `end` waits for the oldest one.
    /// The host buffers must be page-locked (bvht_host_alloc) and stay untouched until the matching `end`.
    pub fn bvht_render_frame_begin(ctx: *mut BvhtCtx, cam: *const BvhtCamera, width: u32, height: u32, tile: u32, region: BvhtRect,
                                   shade: *const BvhtShade, frame_out: *mut u32, hits_out: *mut BvhtHit) -> c_int;
    pub fn bvht_render_frame_end(ctx: *mut BvhtCtx) -> c_int;
    pub fn bvht_trace_rays(ctx: *mut BvhtCtx, rays: *const BvhtRay, n: u64, out: *mut BvhtHit) -> c_int;
    pub fn bvht_host_alloc(ctx: *mut BvhtCtx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn bvht_host_free(ctx: *mut BvhtCtx, p: *mut c_void) -> c_int;
}
