//! `CudaPathTracer`: a second `Integrator` next to `PathTracer` (renderer.rs:337-385) that renders through
//! libbvht_cuda.so.  NOT COMPILED in the bvht-b200 build image (no Rust toolchain there); kept deliberately thin
//! and mechanical.  The C++ mirror of this file is bvhtracer_b200/host/bvhtracer.hpp (class CudaPathTracer).
mod ffi;
pub use ffi::{BvhtShade, BVHT_FLAG_FAST, BVHT_FLAG_LEAF_ACCEL, BVHT_FLAG_STAMP_INSTANCE, BVHT_FLAG_STRICT};

use crate::geometry::Aabb;
use crate::model::Model;
use crate::query::{InstancePrimitiveIndex, Intersection, Ray, SurfaceInteraction};
use crate::renderer::{Integrator, RendererState};
use crate::scene::Scene;
use crate::transform::Transform3;
use cglinalg::{Matrix4x4, Vector3};
use ffi::*;
use std::cell::RefCell;
use std::ffi::CStr;
use std::rc::Rc;

fn cols(m: &Matrix4x4<f32>) -> [f32; 16] {
    let mut out = [0_f32; 16];
    for c in 0..4 { for r in 0..4 { out[c * 4 + r] = m[c][r]; } }     // cglinalg is column-major: m[c][r]
    out
}
fn v3(v: Vector3<f32>) -> [f32; 3] { [v.x, v.y, v.z] }

struct Uploaded { model: *const RefCell<Model>, blas_id: u32, vertex_version: u64 }

pub struct CudaPathTracer {
    ctx: *mut BvhtCtx,
    shade: BvhtShade,
    tile: u32,
    uploaded: Vec<Uploaded>,
    frame_state_resident: bool,         // update_transforms left this frame's TLAS / instances on the device
}

impl CudaPathTracer {
    /// `flags`: BVHT_FLAG_STRICT (bit-identical to PathTracer) | BVHT_FLAG_LEAF_ACCEL, or BVHT_FLAG_FAST.
    pub fn new(flags: u32, shade: BvhtShade) -> Self {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { bvht_create(0, flags, &mut ctx) };
        if rc != BVHT_OK {
            // no CPU fallback by design: mirror the reference's unwrap() style
            panic!("bvht_create failed: {}", unsafe { CStr::from_ptr(bvht_status_string(rc)) }.to_string_lossy());
        }
        Self { ctx, shade, tile: 8, uploaded: vec![], frame_state_resident: false }
    }

    fn check(&self, rc: i32) {
        if rc != BVHT_OK {
            panic!("bvht: {}", unsafe { CStr::from_ptr(bvht_last_error(self.ctx)) }.to_string_lossy());
        }
    }

    /// Upload a model once per `Rc<RefCell<Model>>` (model.rs:16-18); returns its BLAS id.
    fn blas_id_for(&mut self, handle: &Rc<RefCell<Model>>) -> u32 {
        let key = Rc::as_ptr(handle);
        if let Some(u) = self.uploaded.iter().find(|u| u.model == key) { return u.blas_id; }
        let model = handle.borrow();
        let tris = model.primitives();                         // &[Triangle<f32>], 36 B each (mesh.rs:126-134)
        let nodes = model.bvh().export_nodes();                // first nodes_used nodes, field by field
        let mut id = 0_u32;
        self.check(unsafe { bvht_blas_create(self.ctx, tris.as_ptr() as *const f32, tris.len() as u32, nodes.as_ptr(), nodes.len() as u32, &mut id) });
        let normals = model.normals();                         // &[Normals<f32, 3>], 36 B each, never reordered
        if normals.len() == tris.len() {
            self.check(unsafe { bvht_blas_set_normals(self.ctx, id, normals.as_ptr() as *const f32, normals.len() as u32) });
        }
        let tex_coords = model.mesh().tex_coords();            // &[TextureCoordinates<f32, 3>], 24 B each, never reordered
        if tex_coords.len() == tris.len() {
            self.check(unsafe { bvht_blas_set_tex_coords(self.ctx, id, tex_coords.as_ptr() as *const f32, tex_coords.len() as u32) });
        }
        let texture = model.texture().texture();               // TextureBuffer2D<Rgb<u8>, Vec<u8>> (reference-accessors.patch)
        if texture.width() > 0 && texture.height() > 0 {
            self.check(unsafe { bvht_blas_set_texture(self.ctx, id, texture.as_ptr() as *const u8, texture.width() as u32, texture.height() as u32) });
        }
        self.uploaded.push(Uploaded { model: key, blas_id: id, vertex_version: 0 });
        id
    }

    /// What `ModelInstance::refit` (model.rs:31-33) becomes: new vertices to the device, refit there, boxes back.
    pub fn refit(&mut self, handle: &Rc<RefCell<Model>>) {
        let id = self.blas_id_for(handle);
        let mut model = handle.borrow_mut();
        let tris = model.primitives();
        self.check(unsafe { bvht_blas_update_vertices(self.ctx, id, tris.as_ptr() as *const f32, tris.len() as u32) });
        self.check(unsafe { bvht_blas_refit(self.ctx, id) });
        let mut nodes = model.bvh().export_nodes();
        self.check(unsafe { bvht_blas_read_nodes(self.ctx, id, nodes.as_mut_ptr(), nodes.len() as u32) });
        model.bvh_mut().import_bounds(&nodes);
    }

    /// What the per-frame loop `for (o, t) in objects.zip(transforms) { o.set_transform(t) }; scene.rebuild()` of the animated
    /// examples (sixteen_armadillos.rs:132-163) becomes when the scene is large: inverses, world boxes and the agglomerative
    /// `Tlas::rebuild` (tlas.rs:179-250) run on the device, bit-identical to the host code, and the scene adopts the results
    /// (`Scene::adopt_device_state`: cached inverse + bounds per object, `Tlas.nodes`).  Worth it from about 50 objects on.
    pub fn update_transforms(&mut self, scene: &mut Scene, transforms: &[Transform3<f32>]) {
        let n = scene.objects().len();
        assert_eq!(n, transforms.len());
        let mut mats = Vec::with_capacity(n * 16);
        let mut ids = Vec::with_capacity(n);
        for (object, t) in scene.objects().iter().zip(transforms.iter()) {
            mats.extend_from_slice(&cols(&t.compute_matrix()));
            ids.push(self.blas_id_for(&object.model().model()));
        }
        self.check(unsafe { bvht_scene_set_transforms(self.ctx, mats.as_ptr(), ids.as_ptr(), n as u32) });
        let mut nodes = vec![BvhtTlasNode::default(); 2 * n.max(1)];
        let mut inst = vec![BvhtInstance::default(); n];
        let mut bounds = vec![0.0f32; 6 * n];
        let (mut used, mut n_out) = (0u32, 0u32);
        self.check(unsafe { bvht_tlas_read(self.ctx, nodes.as_mut_ptr(), nodes.len() as u32, &mut used, inst.as_mut_ptr(),
                                           bounds.as_mut_ptr(), n as u32, &mut n_out) });
        nodes.truncate(used as usize);
        let inverses: Vec<Transform3<f32>> = inst.iter().map(|i| {
            let m = &i.transform_inv;                           // column-major, like cglinalg
            Transform3::from_matrix(Matrix4x4::new(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]))
        }).collect();
        let boxes: Vec<Aabb<f32>> = bounds.chunks(6).map(|b| Aabb::new(Vector3::new(b[0], b[1], b[2]), Vector3::new(b[3], b[4], b[5]))).collect();
        scene.adopt_device_state(transforms, &inverses, &boxes, &nodes);
        self.frame_state_resident = true;                       // evaluate() need not call bvht_tlas_set for this frame
    }

    fn upload_frame_state(&mut self, scene: &Scene) {
        if std::mem::take(&mut self.frame_state_resident) { return; }
        let mut inst = Vec::with_capacity(scene.objects().len());
        for object in scene.objects().iter() {
            let id = self.blas_id_for(&object.model().model());
            inst.push(BvhtInstance { transform_inv: cols(&object.get_transform_inv().compute_matrix()), blas_id: id });
        }
        let tlas = scene.tlas().export_nodes();                // Tlas.nodes[..nodes_used] after Tlas::rebuild (tlas.rs:204-250)
        self.check(unsafe { bvht_tlas_set(self.ctx, tlas.as_ptr(), tlas.len() as u32, inst.as_ptr(), inst.len() as u32) });
    }

    fn camera(scene: &Scene) -> BvhtCamera {
        let c = scene.active_camera();
        BvhtCamera { top_left_eye: v3(c.top_left_eye()), top_right_eye: v3(c.top_right_eye()), bottom_left_eye: v3(c.bottom_left_eye()),
                     view_matrix_inv: cols(c.view_matrix_inv()) }
    }

    /// `Scene::intersect(&Ray)` (scene.rs:32-34) on the device.
    pub fn intersect(&mut self, scene: &Scene, ray: &Ray<f32>) -> Option<Intersection<f32>> {
        self.upload_frame_state(scene);
        let r = BvhtRay { origin: v3(ray.origin), direction: v3(ray.direction), t: ray.t };
        let mut hit = BvhtHit::default();
        self.check(unsafe { bvht_trace_rays(self.ctx, &r, 1, &mut hit) });
        if hit.id == u32::MAX { return None; }
        let mut closest = *ray;                                // NOTE: the reference returns the MODEL-space ray here
        closest.t = hit.t;
        Some(Intersection::new(closest, SurfaceInteraction::new(hit.t, hit.u, hit.v), InstancePrimitiveIndex::from_primitive(hit.id & 0x000F_FFFF)))
    }
}

impl Integrator for CudaPathTracer {
    fn evaluate(&mut self, renderer_state: &mut RendererState, scene: &Scene) -> usize {
        self.upload_frame_state(scene);
        let cam = Self::camera(scene);
        let (w, h) = (renderer_state.frame_buffer().width() as u32, renderer_state.frame_buffer().height() as u32);
        let mut shade = self.shade;
        if shade.kind == ffi::BVHT_SHADE_NORMAL {
            // scene.get_unchecked(0).get_transform(): the reference's instance index is always 0 (renderer.rs:258-278)
            shade.object0_transform = cols(&scene.get_unchecked(0).get_transform().compute_matrix());
        }
        let pixels = renderer_state.frame_pixels_mut();       // Rgba<u8> storage, 4 B/pixel, row-major, not flipped
        debug_assert_eq!(pixels.len(), (w * h * 4) as usize);
        // ONE call replaces the 409,600 Accumulator::evaluate + PixelShader::evaluate calls of renderer.rs:353-381
        self.check(unsafe { bvht_render_frame(self.ctx, &cam, w, h, self.tile, BvhtRect { x0: 0, y0: 0, x1: w, y1: h }, &shade,
                                              pixels.as_mut_ptr() as *mut u32, std::ptr::null_mut()) });
        (w * h) as usize                                       // rays traced, like PathTracer (renderer.rs:383)
    }
}

impl CudaPathTracer {
    /// `evaluate` without the wait: the frame is queued into `pixels` (page-locked, width * height Rgba<u8>, e.g. from
    /// bvht_host_alloc; NOT the `Vec` of a `FrameBuffer`, which is pageable and may move) and the call returns.  At most two
    /// frames may be in flight; alternate two buffers and present frame n after `evaluate_end()` while frame n + 1 is traced.
    ///
    /// # Safety
    /// `pixels` must stay valid and untouched until the matching `evaluate_end`.
    pub unsafe fn evaluate_begin(&mut self, pixels: *mut u32, width: u32, height: u32, scene: &Scene) -> usize {
        self.upload_frame_state(scene);
        let cam = Self::camera(scene);
        let mut shade = self.shade;
        if shade.kind == ffi::BVHT_SHADE_NORMAL {
            shade.object0_transform = cols(&scene.get_unchecked(0).get_transform().compute_matrix());
        }
        self.check(bvht_render_frame_begin(self.ctx, &cam, width, height, self.tile, BvhtRect { x0: 0, y0: 0, x1: width, y1: height },
                                           &shade, pixels, std::ptr::null_mut()));
        (width * height) as usize
    }

    /// Wait for the oldest frame begun with `evaluate_begin`; its pixels are then in the buffer that was passed.
    pub fn evaluate_end(&mut self) {
        self.check(unsafe { bvht_render_frame_end(self.ctx) });
    }
}

impl Drop for CudaPathTracer {
    fn drop(&mut self) { unsafe { bvht_destroy(self.ctx) } }
}
