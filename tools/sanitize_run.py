#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel for compute-sanitizer (GPU box):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhtracer_b200 import examples, host

for flags in (0, 2, 3):
    # animated TLAS scene, depth shading, hits
    anim = examples.GridAnimation()
    scene, models = host.build_scene(examples.sixteen_armadillos(0))
    r = host.Renderer(flags=flags)
    st = host.RendererState(host.depth_pipeline(), 160, 96, keep_hits=True)
    for f in range(2):
        anim.update()
        for i, o in enumerate(anim.objects()):
            scene.set_transform(i, host.object_transform(o))
        scene.rebuild()
        r.render(st, scene)
    rays = np.array([[0, 1, -5.5, 0, 0, 1, 3.0e38], [0, 1, -5.5, 0.1, 0, 1, 3.0e38]], np.float32)
    r.intersect(scene, rays)
    # vertex animation + refit + normal shading
    scene2, models2 = host.build_scene(examples.big_ben_clock())
    a2 = examples.BigBenAnimation(models2[0].primitives())
    st2 = host.RendererState(host.normal_pipeline(), 128, 72, keep_hits=False)
    for f in range(2):
        models2[0].set_primitives(a2.animate())
        models2[0].refit()
        r.render(st2, scene2)
    print("flags", flags, "ok", int(st.frame_buffer().sum() % 1000), int(st2.frame_buffer().sum() % 1000))

# device-side construction (K5) + textured quad + a far ray batch (re-bake path)
r = host.Renderer(flags=2)
m = r.build_model(host.load_asset_mesh("teapot.obj"))
bb = r.build_model(host.load_asset_mesh("unity.tri"))
bb.set_primitives(examples.BigBenAnimation(bb.primitives()).animate())
r.rebuild_model(bb)
mesh = host.Mesh.from_triangles(examples.QUAD_TRIS, examples.QUAD_NORMALS).set_tex_coords(examples.QUAD_TEX_COORDS)
quad = host.ModelBuilder().with_mesh(mesh).with_texture(examples.brick_texture(32, 16)).build()
scene3, _ = host.build_scene(examples.quad_example(4), models=[quad])
st3 = host.RendererState(host.texture_pipeline(), 96, 96, keep_hits=False)
r.render(st3, scene3)
scene4, _ = host.build_scene(examples.trippy_teapots(2), models=[m])
st4 = host.RendererState(host.normal_pipeline(), 128, 72, keep_hits=True)
r.render(st4, scene4)
far = np.array([[300, 200, -500, -0.3, -0.2, 0.5, 3.0e38], [0, 1, -5.5, 0.1, 0, 40.0, 3.0e38]], np.float32)
r.intersect(scene4, far)
print("device build ok", int(st3.frame_buffer().sum() % 1000), int(st4.frame_buffer().sum() % 1000))

# K0 (classify + fill) forced on, the chain-skipping TLAS walk, and K6 (device set_transform + Tlas::rebuild)
anim = examples.GridAnimation()
scene5, _ = host.build_scene(examples.sixteen_armadillos(0))
r5 = host.Renderer(flags=2)
r5.engine().set_option(2, 1)      # BVHT_OPT_K0 forced on
st5 = host.RendererState(host.depth_pipeline(), 200, 120, keep_hits=True)
for f in range(3):
    anim.update()
    r5.update_transforms(scene5, [host.object_transform(o) for o in anim.objects()])
    r5.render(st5, scene5)
print("k0 + k6 ok", int(st5.frame_buffer().sum() % 1000))

# the pipelined host frame: several bands, every pull order, two copy streams, coverage raster forced on, a tile-row shard
r6 = host.Renderer(flags=2)
e6 = r6.engine()
st6 = host.RendererState(host.depth_pipeline(), 320, 184, keep_hits=True)
e6.set_option(1, 1)               # BVHT_OPT_COVER
for bands, order in ((4, 1), (7, 2), (23, 3), (2, 0)):
    e6.set_option(3, bands)       # BVHT_OPT_BANDS
    e6.set_option(4, order)       # BVHT_OPT_BAND_ORDER
    r6.render(st6, scene5)
e6.set_shard(1, 3)
r6.render(st6, scene5)
print("pipelined frame ok", int(st6.frame_buffer().sum() % 1000))

# two frames in flight: band copies of one frame under the kernels of the next, scene updates queued between them
r7 = host.Renderer(flags=2)
pair = [host.RendererState(host.depth_pipeline(), 320, 184, keep_hits=(i == 0)) for i in range(2)]
r7.engine().set_option(3, 5)      # BVHT_OPT_BANDS
for f in range(5):
    anim.update()
    for i, o in enumerate(anim.objects()):
        scene5.set_transform(i, host.object_transform(o))
    scene5.rebuild()
    r7.render_begin(pair[f & 1], scene5)
    if f:
        r7.render_end()
r7.render_end()
print("frames in flight ok", int(pair[0].frame_buffer().sum() % 1000), int(pair[1].frame_buffer().sum() % 1000))
