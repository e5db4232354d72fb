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
