"""Renderer::render (device-side band flags, cuStreamWaitValue32) against render_begin + render_end back to back (band flags in host
memory, copies issued by the host) on the same animation frames; per-step CUDA-event intervals, L2 flushed outside them."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from bvhtracer_b200 import host

steps = 20
for name in [a for a in sys.argv[1:] if a != "-v"] or ["sixteen_armadillos"]:
    for mode in ("render", "begin+end", "render", "begin+end"):
        wl = bench.GpuWorkload(name, bench.MODES["strict-accel"], 0)
        stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
        wl.renderer.set_stream(stream.cuda_stream)
        w, h = bench.frame_size(name, 1, "strong")
        state = host.RendererState(wl.pipeline, w, h, keep_hits=False)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        r, scene = wl.renderer, wl.scene
        ev = []
        for i in range(5 + steps):
            wl.advance(); flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            if mode == "render":
                r.render(state, scene)
            else:
                r.render_begin(state, scene); r.render_end()
            b.record(stream)
            if i >= 5: ev.append((a, b))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        print(f"{name:20s} {mode:10s}: {np.mean(ms):.4f} ms per frame (min {min(ms):.4f})  checksum {int(np.bitwise_xor.reduce(state.frame_buffer()))}")
        if "-v" in sys.argv:
            print("    ", " ".join(f"{m:.2f}" for m in ms))
        del wl, state, r, scene
