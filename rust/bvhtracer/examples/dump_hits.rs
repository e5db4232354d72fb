//! dump_hits -- what the REFERENCE ITSELF answers for every primary ray of the benchmark scenes, as raw 16-byte records.
//!
//! Uses nothing but the reference crate's public API (`Scene::intersect`, scene.rs:32-34, on rays from
//! `Camera::get_ray_world`, camera.rs:1004-1010, with u = x / W, v = y / H as in renderer.rs:358-361): no patch, no CUDA, no
//! window.  Copy this file to `bvhtracer/examples/dump_hits.rs` of the reference workspace and run
//!
//!     cargo run --release -p bvhtracer --example dump_hits -- <out_dir> [case ...]
//!
//! One file per case, `<out_dir>/<scene>_f<frame>_<W>x<H>.hits`: W * H little-endian records {t: f32, u: f32, v: f32, id: u32} in
//! row-major pixel order (v = 0 is the top row); a miss is {f32::MAX, 0, 0, 0xFFFFFFFF}; id = instance_index << 20 | primitive_index
//! (`InstancePrimitiveIndex`, intersection.rs:33-67) -- byte for byte the `bvht_hit` records of bvht-b200, whose
//! `tools/compare_dump.py <out_dir>` traces the same cases on the GPU (or with its CPU oracle) and diffs them.
//!
//! A case is `scene:frame:WxH`.  Scenes and their frames (bvht-b200 SURVEY.md 8d):
//!   cube                frame 0 = the scene as constructed (cube.rs:25-80)
//!   quad                frame 0 = tests/test_scene_quad.rs:52-131
//!   two_armadillos      frame 0 = both instances at identity (as constructed); frame 1 = translations (-1.3, 0, 0) / (+1.3, 0, 0)
//!   sixteen_armadillos  frame k = after k calls of AppState::update(1/60) (sixteen_armadillos.rs:132-163)
//!   trippy_teapots      frame k = likewise (trippy_teapots.rs)
//!   big_ben_clock       frame k = after k calls of animate() + ModelInstance::refit (big_ben_clock.rs:67-103)
extern crate bvhtracer;
extern crate cglinalg;

use bvhtracer::*;
use cglinalg::{Degrees, Radians, Rotation3, Vector3};
use std::fs::File;
use std::io::{self, BufWriter, Write};
use std::path::Path;

const DEFAULT_CASES: &[&str] = &[
    "cube:0:640x640", "quad:0:640x640",
    "two_armadillos:0:480x270", "two_armadillos:1:480x270",
    "sixteen_armadillos:0:480x270", "sixteen_armadillos:1:480x270", "sixteen_armadillos:37:480x270",
    "trippy_teapots:10:480x270", "big_ben_clock:3:480x270",
];

fn asset(name: &str) -> io::BufReader<File> {
    // the examples crate's assets, relative to this crate (bvhtracer/)
    let path = Path::new(env!("CARGO_MANIFEST_DIR")).join("../examples/assets").join(name);
    io::BufReader::new(File::open(&path).unwrap_or_else(|e| panic!("{}: {}", path.display(), e)))
}

fn tri_model(name: &str) -> ModelInstance {
    let mesh = TriMeshDecoder::new(asset(name)).read_mesh().unwrap();
    ModelBuilder::new().with_mesh(mesh).build()
}

fn obj_model(name: &str) -> ModelInstance {
    // the examples attach bricks_rgb.png through SimpleModelDecoder; the texture plays no part in Scene::intersect
    let mesh = ObjMeshDecoder::new(asset(name)).read_mesh().unwrap();
    ModelBuilder::new().with_mesh(mesh).build()
}

fn fov_camera(position: Vector3<f32>, forward: Vector3<f32>, right: Vector3<f32>, up: Vector3<f32>, near: f32, far: f32) -> Camera<f32, PerspectiveProjection<f32>> {
    let projection = SymmetricFovSpec::new(Degrees(90_f32), 1_f32, near, far);
    let attitude = CameraAttitudeSpec::new(position, forward, right, up, -forward);
    Camera::new(&projection, &attitude)
}

fn object(world: &mut World<f32>, model: ModelInstance, transform: Option<&Transform3<f32>>) -> SceneObject {
    let body = world.register_body(RigidBody::default());
    let builder = SceneObjectBuilder::new(model, body);
    match transform {
        Some(t) => builder.with_transform(t).build(),
        None => builder.build(),
    }
}

/// The closed-form `Physics` of sixteen_armadillos.rs:21-44 / trippy_teapots.rs, one per grid cell.
struct Cell { angle: f32, position_init: Vector3<f32>, height: f32, speed: f32, angular_velocity: f32, acceleration: f32, direction: f32 }

fn grid_scene(model: ModelInstance, camera: Camera<f32, PerspectiveProjection<f32>>, frame: usize) -> Scene {
    let height_init = [5_f32, 4_f32, 3_f32, 2_f32, 1_f32, 5_f32, 4_f32, 3_f32, 5_f32, 4_f32, 3_f32, 2_f32, 1_f32, 5_f32, 4_f32, 3_f32];
    let mut world = World::new();
    let mut objects = vec![];
    let mut cells = vec![];
    let mut i = 0;
    for x in 0..4 { for y in 0..4 {
        let even = ((x + y) & 1) == 0;
        // `& 7 + 2` parses as `& 9`, as in the example
        let angular_velocity = if even { (((i * 13) & 7 + 2) as f32) * 0.10 } else { 0_f32 };
        let horizontal = Vector3::new((x as f32 - 1.5) * 2.5, 0_f32, (y as f32 - 1.5) * 2.5);
        let vertical = if even { Vector3::zero() } else { Vector3::new(0_f32, height_init[i / 2], 0_f32) };
        let translation = horizontal + vertical;
        let rotation = Rotation3::from_angle_x(Radians(0_f32)) * Rotation3::from_angle_z(Radians(0_f32));
        let transform = Transform3::new(&Vector3::from_fill(0.75), &translation, rotation);
        objects.push(object(&mut world, model.clone(), Some(&transform)));
        cells.push(Cell { angle: 0_f32, position_init: translation, height: vertical.y, speed: 0_f32, angular_velocity,
                          acceleration: if even { 0_f32 } else { 9.8 }, direction: -1_f32 });
        i += 1;
    }}
    let mut scene = SceneBuilder::new(camera).with_objects(objects).build();
    let elapsed = 1.0_f64 / 60.0_f64;
    for _ in 0..frame {
        for (i, c) in cells.iter().enumerate() {
            let translation = c.position_init + Vector3::new(0_f32, c.height, 0_f32);
            let rotation = Rotation3::from_angle_x(Radians(c.angle)) * Rotation3::from_angle_z(Radians(c.angle));
            scene.get_mut_unchecked(i).set_transform(&Transform3::new(&Vector3::from_fill(0.75_f32), &translation, rotation));
        }
        for c in cells.iter_mut() {
            c.angle = c.angle + c.angular_velocity * (elapsed as f32);
            c.speed += c.acceleration * (elapsed as f32);
            c.height += c.direction * c.speed * (elapsed as f32);
            if c.height < -3_f32 {
                c.height = (-3_f32) + 0.01;
                c.direction = -c.direction;
                c.speed = 0.2;
            } else if c.height > c.position_init.y {
                c.height = c.position_init.y - 0.01;
                c.direction = -c.direction;
            }
        }
        scene.rebuild();
    }
    scene
}

fn big_ben_scene(frame: usize) -> Scene {
    let camera = fov_camera(Vector3::new(0_f32, 2.75_f32, -2.5_f32), Vector3::unit_z(), Vector3::unit_x(), Vector3::unit_y(), 2_f32, 10000_f32);
    let mut world = World::new();
    let scene_object = object(&mut world, tri_model("bigben.tri"), None);
    let mut scene = SceneBuilder::new(camera).with_object(scene_object).build();
    let originals: Vec<Triangle<f32>> = scene.get_unchecked(0).model().model().borrow().primitives().iter().map(|p| *p).collect();
    let mut r = 0_f32;
    for _ in 0..frame {
        r += 0.05;
        if r > std::f32::consts::FRAC_2_PI { r -= std::f32::consts::FRAC_2_PI; }
        let a = f32::sin(r) * 0.5;
        let handle = scene.get_unchecked(0).model().model();
        for (i, original) in originals.iter().enumerate() {
            let mut twisted = [Vector3::zero(); 3];
            for k in 0..3 {
                let o = original.vertices[k];
                let s = a * (o.y - 0.2) * 0.2;
                twisted[k] = Vector3::new(o.x * f32::cos(s) - o.y * f32::sin(s), o.x * f32::sin(s) + o.y * f32::cos(s), o.z);
            }
            handle.borrow_mut().primitives_mut()[i] = Triangle::new(twisted[0], twisted[1], twisted[2]);
        }
        scene.get_mut_unchecked(0).model().refit();
    }
    scene
}

fn build_scene(name: &str, frame: usize) -> Scene {
    match name {
        "cube" => {
            let position = Vector3::new(0_f32, 4_f32, 0_f32);
            let forward = (Vector3::zero() - position).normalize();
            let camera = fov_camera(position, forward, -Vector3::unit_x(), Vector3::unit_z(), 1_f32, 100_f32);
            let mut world = World::new();
            let transform = Transform3::from_scale_translation(&Vector3::from_fill(2_f32), &Vector3::new(-1_f32, -1_f32, -1_f32));
            let cube = object(&mut world, obj_model("cube.obj"), Some(&transform));
            SceneBuilder::new(camera).with_physics(world).with_object(cube).build()
        }
        "quad" => {
            let projection = BoxSpec::new(-1_f32, 1_f32, -1_f32, 1_f32, 1_f32, 100_f32);
            let attitude = CameraAttitudeSpec::new(Vector3::new(0_f32, 0_f32, 2_f32), -Vector3::unit_z(), Vector3::unit_x(), Vector3::unit_y(), -Vector3::unit_z());
            let camera = Camera::new(&projection, &attitude);
            let mesh = MeshBuilder::new()
                .with_primitive(
                    Triangle::new(Vector3::new(-1.0, -1.0, 0.0), Vector3::new(1.0, 1.0, 0.0), Vector3::new(-1.0, 1.0, 0.0)),
                    TextureCoordinates::from([cglinalg::Vector2::new(0.0, 0.0), cglinalg::Vector2::new(1.0, 1.0), cglinalg::Vector2::new(0.0, 1.0)]),
                    Normals::from([Vector3::new(0.0, 0.0, 1.0), Vector3::new(0.0, 0.0, 1.0), Vector3::new(0.0, 0.0, 1.0)]))
                .with_primitive(
                    Triangle::new(Vector3::new(-1.0, -1.0, 0.0), Vector3::new(1.0, -1.0, 0.0), Vector3::new(1.0, 1.0, 0.0)),
                    TextureCoordinates::from([cglinalg::Vector2::new(0.0, 0.0), cglinalg::Vector2::new(1.0, 0.0), cglinalg::Vector2::new(1.0, 1.0)]),
                    Normals::from([Vector3::new(0.0, 0.0, 1.0), Vector3::new(0.0, 0.0, 1.0), Vector3::new(0.0, 0.0, 1.0)]))
                .build();
            let mut world = World::new();
            let quad = object(&mut world, ModelBuilder::new().with_mesh(mesh).build(), Some(&Transform3::identity()));
            SceneBuilder::new(camera).with_physics(world).with_object(quad).build()
        }
        "two_armadillos" => {
            let camera = fov_camera(Vector3::new(0_f32, 1_f32, -2.5_f32), Vector3::unit_z(), Vector3::unit_x(), Vector3::unit_y(), 2_f32, 10000_f32);
            let model = tri_model("armadillo.tri");
            let mut world = World::new();
            let objects = vec![
                object(&mut world, model.clone(), Some(&Transform3::identity())),
                object(&mut world, model.clone(), Some(&Transform3::identity())),
            ];
            let mut scene = SceneBuilder::new(camera).with_physics(world).with_objects(objects).build();
            if frame >= 1 {
                // the pose the rigid bodies hold (two_armadillos.rs:55-82) with the physics' negligible rotation left out
                scene.get_mut_unchecked(0).set_transform(&Transform3::from_translation(&Vector3::new(-1.3_f32, 0_f32, 0_f32)));
                scene.get_mut_unchecked(1).set_transform(&Transform3::from_translation(&Vector3::new(1.3_f32, 0_f32, 0_f32)));
                scene.rebuild();
            }
            scene
        }
        "sixteen_armadillos" => {
            let camera = fov_camera(Vector3::new(0_f32, 1_f32, -5.5_f32), Vector3::unit_z(), Vector3::unit_x(), Vector3::unit_y(), 2_f32, 10000_f32);
            grid_scene(tri_model("armadillo.tri"), camera, frame)
        }
        "trippy_teapots" => {
            let camera = fov_camera(Vector3::new(0_f32, 1.5_f32, -5.5_f32), Vector3::unit_z(), -Vector3::unit_x(), Vector3::unit_y(), 2_f32, 10000_f32);
            grid_scene(obj_model("teapot.obj"), camera, frame)
        }
        "big_ben_clock" => big_ben_scene(frame),
        other => panic!("unknown scene {}", other),
    }
}

fn dump(out_dir: &Path, case: &str) -> io::Result<()> {
    let parts: Vec<&str> = case.split(':').collect();
    let name = parts[0];
    let frame: usize = parts.get(1).map(|s| s.parse().unwrap()).unwrap_or(0);
    let (width, height) = parts.get(2).map(|s| {
        let wh: Vec<usize> = s.split('x').map(|n| n.parse().unwrap()).collect();
        (wh[0], wh[1])
    }).unwrap_or((640, 640));
    let scene = build_scene(name, frame);
    let path = out_dir.join(format!("{}_f{}_{}x{}.hits", name, frame, width, height));
    let mut out = BufWriter::new(File::create(&path)?);
    let mut hits = 0_usize;
    for y in 0..height {
        for x in 0..width {
            let u = x as f32 / width as f32;
            let v = y as f32 / height as f32;
            let ray = scene.active_camera().get_ray_world(u, v);
            let (t, bu, bv, id) = match scene.intersect(&ray) {
                Some(hit) => {
                    hits += 1;
                    let ip = hit.instance_primitive;
                    (hit.interaction.t, hit.interaction.u, hit.interaction.v, (ip.instance_index() << 20) | ip.primitive_index())
                }
                None => (f32::MAX, 0_f32, 0_f32, u32::MAX),
            };
            out.write_all(&t.to_le_bytes())?;
            out.write_all(&bu.to_le_bytes())?;
            out.write_all(&bv.to_le_bytes())?;
            out.write_all(&id.to_le_bytes())?;
        }
    }
    out.flush()?;
    println!("{}: {} x {} rays, {} hits", path.display(), width, height, hits);
    Ok(())
}

fn main() -> io::Result<()> {
    let args: Vec<String> = std::env::args().skip(1).collect();
    if args.is_empty() {
        eprintln!("usage: dump_hits <out_dir> [scene:frame:WxH ...]");
        std::process::exit(2);
    }
    let out_dir = Path::new(&args[0]);
    std::fs::create_dir_all(out_dir)?;
    if args.len() == 1 {
        for case in DEFAULT_CASES { dump(out_dir, case)?; }
    } else {
        for case in &args[1..] { dump(out_dir, case)?; }
    }
    Ok(())
}
