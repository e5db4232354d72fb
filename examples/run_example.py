#!/usr/bin/env python3
"""Run one of the reference's examples (cube, quad, two_armadillos, sixteen_armadillos, trippy_teapots, big_ben_clock)
through the B200 backend and save the frame, as the reference's App would display it.

    python examples/run_example.py sixteen_armadillos --frames 30 --size 640 640 --out frame.png

Each example keeps its own scene, camera, accumulator and pixel shader (examples/*.rs `main`); only the integrator is
`CudaPathTracer` instead of `PathTracer`.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from bvhtracer_b200 import FLAG_FAST, FLAG_LEAF_ACCEL, FLAG_STRICT, examples, host

PIPELINES = {
    "cube": host.normal_pipeline,                                       # cube.rs:101-102
    "two_armadillos": lambda: host.depth_pipeline(80.0, 3.0),           # two_armadillos.rs:123-124
    "sixteen_armadillos": lambda: host.depth_pipeline(80.0, 3.0),       # sixteen_armadillos.rs:182-183
    "trippy_teapots": host.normal_pipeline,                             # trippy_teapots.rs:184-185
    "big_ben_clock": lambda: host.intersection_pipeline((255, 255, 255, 255), (0, 0, 0, 255)),   # big_ben_clock.rs:123-130
    "quad": host.texture_pipeline,                                      # quad.rs:137-138
}


def build_models(spec, renderer, on_device):
    """The example's ModelBuilder calls: on the host (C++ mirror of BvhBuilder) or on the device (CudaPathTracer::build_model)."""
    models = []
    for asset in spec.meshes:
        if asset == "<quad>":                                           # quad.rs:45-92: mesh + texture coordinates + texture
            mesh = host.Mesh.from_triangles(examples.QUAD_TRIS, examples.QUAD_NORMALS).set_tex_coords(examples.QUAD_TEX_COORDS)
            models.append(host.ModelBuilder().with_mesh(mesh).with_texture(examples.brick_texture()).build())
            continue
        mesh = host.load_asset_mesh(asset)
        models.append(renderer.build_model(mesh) if on_device else host.ModelBuilder().with_mesh(mesh).build())
    return models


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", choices=sorted(PIPELINES))
    ap.add_argument("--frames", type=int, default=1, help="number of update() calls before the saved frame")
    ap.add_argument("--size", type=int, nargs=2, default=None, help="W H (default: the example's 640 640)")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-build", action="store_true", help="run BvhBuilder::build_for on the device (bvht_blas_build)")
    args = ap.parse_args()

    spec = examples.quad_example(0) if args.name == "quad" else examples.CONFIGS[args.name]()
    w, h = args.size if args.size else spec.default_size
    flags = (FLAG_FAST if args.mode == "fast" else FLAG_STRICT) | FLAG_LEAF_ACCEL
    renderer = host.Renderer(flags=flags)
    t_build = time.perf_counter()
    scene, models = host.build_scene(spec, models=build_models(spec, renderer, args.device_build))
    t_build = time.perf_counter() - t_build
    state = host.RendererState(PIPELINES[args.name](), w, h)
    anim = examples.GridAnimation() if args.name in ("sixteen_armadillos", "trippy_teapots") else None
    bb = examples.BigBenAnimation(models[0].primitives()) if args.name == "big_ben_clock" else None
    t0 = time.perf_counter()
    rays = renderer.render(state, scene)
    for frame_no in range(1, args.frames + 1):
        if args.name == "quad":                                          # stand-in for the rigid-body spin (examples.quad_example)
            scene.set_transform(0, host.object_transform(examples.quad_example(frame_no).objects[0]))
            scene.rebuild()
        if anim is not None:                                             # AppState::update (sixteen_armadillos.rs:132-163)
            anim.update()
            for i, o in enumerate(anim.objects()):
                scene.set_transform(i, host.object_transform(o))
            scene.rebuild()
        if bb is not None:                                               # big_ben_clock.rs:67-103
            models[0].set_primitives(bb.animate())
            models[0].refit()
        rays += renderer.render(state, scene)
    dt = time.perf_counter() - t0
    # row 0 is the TOP of the image (v = 0, renderer.rs:358-361); bvhtracer_demos flips it only because GL textures start at
    # the bottom (lib.rs:114) -- an image file wants it as is
    frame = state.frame_buffer().reshape(h, w).copy()
    print(f"{args.name}: {args.frames + 1} frames of {w}x{h}, {rays} rays in {dt * 1e3:.1f} ms "
          f"({rays / dt / 1e6:.0f} Mrays/s incl. host updates), last trace {renderer.stats()['last_trace_ms']:.3f} ms; "
          f"scene built in {t_build * 1e3:.1f} ms ({'device' if args.device_build else 'host'} BVH build)")
    if args.out:
        rgba = frame.view(np.uint8).reshape(h, w, 4)
        if args.out.endswith(".ppm"):
            with open(args.out, "wb") as f:
                f.write(f"P6\n{w} {h}\n255\n".encode())
                f.write(rgba[:, :, :3].tobytes())
        else:
            from PIL import Image
            Image.fromarray(rgba[:, :, :3]).save(args.out)
        print("wrote", args.out)


if __name__ == "__main__":
    main()
