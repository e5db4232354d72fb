"""Where does a loop with two frames in flight spend its time?  Host timestamps around every call + device events."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "sixteen_armadillos"
steps = 12
wl = bench.GpuWorkload(name, bench.MODES["strict-accel"], 0)
from bvhtracer_b200 import host
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
wl.renderer.set_stream(stream.cuda_stream)
wl.eng.set_option(6, 1)            # BVHT_OPT_TIMELINE
w, h = bench.frame_size(name, 1, "strong")
states = [host.RendererState(wl.pipeline, w, h, keep_hits=False) for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
r, scene = wl.renderer, wl.scene
start = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for _ in range(start):
    wl.advance()
r.render(states[0], scene)
# the synchronous call on the same frames
for rep in range(2):
    ev = []
    for i in range(steps):
        wl.advance(); flush.zero_()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream); r.render(states[0], scene); eb.record(stream)
        ev.append((ea, eb))
    torch.cuda.synchronize()
    tl = wl.eng.debug_frame_timeline()
    print(f"== {name}: Renderer::render from frame {wl.frame - steps + 1}: {sum(a.elapsed_time(b) for a, b in ev) / steps:.3f} ms per frame; K1 of the last {r.stats()['last_k1_ms']:.3f} ms, kernels done {tl['bands'][0][0]:.3f} ms; re-bakes so far {r.stats()['rebakes']}")
for variant in ("animated scene + flush", "animated scene + flush", "animated scene, no flush"):
    animated, do_flush = variant.startswith("animated"), variant.endswith("+ flush")
    for i in range(4):
        r.render_begin(states[i & 1], scene)
        if i: r.render_end()
    r.render_end()
    torch.cuda.synchronize()
    wl.precompute(steps)
    rows = []
    evs = []
    t_start = time.perf_counter()
    p0 = torch.cuda.Event(enable_timing=True); p0.record(stream)
    for i in range(steps):
        t0 = time.perf_counter()
        if animated: wl.advance()
        if do_flush: flush.zero_()
        t1 = time.perf_counter()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        r.render_begin(states[i & 1], scene)
        eb.record(stream)
        t2 = time.perf_counter()
        if i: r.render_end()
        t3 = time.perf_counter()
        rows.append((t0 - t_start, t1 - t0, t2 - t1, t3 - t2)); evs.append((ea, eb))
    r.render_end()
    p1 = torch.cuda.Event(enable_timing=True); p1.record(stream)
    torch.cuda.synchronize()
    st = r.stats()
    print(f"   last frame in flight: K1 {st['last_k1_ms']:.3f} ms, first device op .. last kernel {st['last_trace_ms']:.3f} ms")
    print(f"== {name} {w}x{h}: {variant}, ending at frame {wl.frame}: {p0.elapsed_time(p1) / steps:.3f} ms per frame; re-bakes so far {r.stats()['rebakes']}")
    for i, ((ts, ta, tb, te), (ea, eb)) in list(enumerate(zip(rows, evs)))[:6]:
        print(f"  frame {i:2d}: host t={ts * 1e3:7.3f} ms  update+flush {ta * 1e3:6.3f}  begin {tb * 1e3:6.3f}  end {te * 1e3:6.3f} | device: kernels start {p0.elapsed_time(ea):7.3f}  end {p0.elapsed_time(eb):7.3f}  ({ea.elapsed_time(eb):.3f})")
