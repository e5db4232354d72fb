#!/usr/bin/env python3
"""bench.py -- primary closest-hit throughput (Mrays/s) of the B200 engine on the reference's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--mode strict-accel|...]

Workloads = BASELINE.json `configs`; the default is the one the metric's target is quoted on:
  cube               C1   640 x  640   1 instance of cube.obj (12 triangles), normal-mapping shader
  two_armadillos     C2  1920 x 1080   2 instances of armadillo.tri (30,001 triangles), depth shader
  sixteen_armadillos C3  3840 x 2160   16 animated instances of one armadillo BLAS, TLAS rebuilt per frame (DEFAULT)
  trippy_teapots     C4  3840 x 2160   16 animated, rotating instances of teapot.obj (1,024 triangles), normal-mapping shader
  big_ben_clock      C5  7680 x 4320   bigben.tri (20,945 triangles), every vertex animated per frame + BLAS refit
One STEP = one frame: the host-side scene update of the example (outside every timed interval, as in the reference's
AppState::update), then every primary ray of the frame traced to its closest hit.
  value  : rays / device time of the frame with the scene resident in HBM (CUDA events on the launching stream); the 16-byte
           hit records stay in HBM.  C5: the interval also holds the vertex upload, the sub-BVH refit + re-bake and K2 (Bvh::refit).
  e2e    : the same frame through the reference-facing plugin call `Renderer::render` of the C++ host mirror
           (CudaPathTracer::evaluate: model / TLAS / instance uploads from host memory + refit + trace + on-device accumulator and
           pixel shader + Rgba<u8> frame buffer copied back to page-locked host memory), CUDA events around the call.
           e2e.two_frames_in_flight (N = 1): the same frames through Renderer::render_begin / render_end over two frame buffers
           (frame n copied back under frame n+1's kernels), ONE interval around the whole loop, L2 flushes inside it.
Multi-GPU (torchrun, one rank per GPU), STRONG scaling by default: the frame is fixed, the scene replicated, tile rows
interleaved over the ranks.  value: every rank stores its shaded pixels straight into rank 0's frame buffer over NVLink P2P (CUDA
IPC mapping), no collective in the data path; hit records stay in the rank's own HBM.  e2e: every rank's `Renderer::render`
fills its tile rows of ONE page-locked frame in POSIX shared memory over its own PCIe link.  Both are sums of per-step CUDA-event
intervals, max over ranks -- the same clock as N = 1.  `--scaling weak` grows the frame to W x (H * N) instead.
`--impl reference` times the CPU oracle (the C restatement of the reference's Rust path; the reference itself cannot be compiled
here: no cargo/rustc) with all host threads on a bounded sample of the same frames.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "Mrays/s closest-hit (primary)"
MODES = {"strict-brute": 0x0, "strict-accel": 0x2, "fast-brute": 0x1, "fast-accel": 0x3}
# name -> (BASELINE config tag, shading pipeline of the example, scene update per frame)
WORKLOADS = {
    "cube": ("C1", "normal", "static"),
    "two_armadillos": ("C2", "depth", "static"),
    "sixteen_armadillos": ("C3", "depth", "grid"),
    "trippy_teapots": ("C4", "normal", "grid"),
    "big_ben_clock": ("C5", "intersection", "bigben"),
}
DEFAULT_WORKLOAD = "sixteen_armadillos"
TILE = 8


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 2 ms; nvidia-smi as fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, uuid=None, index=0, period_s=0.002):
        self.uuid, self.index, self.period_s = uuid, index, period_s
        self.sm, self.mask, self.sm_max = [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.error = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, self.uuid.encode()):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    pass
                time.sleep(self.period_s)
        except Exception as e:                                   # pragma: no cover
            self.error = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            return self._smi_once()
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "samples": len(self.sm),
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "source": f"nvml, {self.period_s * 1e3:g} ms period, whole timed region"}

    def _smi_once(self):
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1,
                    "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")],
                    "source": f"nvidia-smi after the timed region (nvml failed: {self.error})"}
        except Exception as e:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [f"unavailable: {e!r}"]}


def frame_size(workload, world, scaling):
    from bvhtracer_b200 import examples
    w, h = examples.CONFIGS[workload]().bench_size if workload not in ("sixteen_armadillos", "trippy_teapots") \
        else examples.CONFIGS[workload](0).bench_size
    return w, h * (world if scaling == "weak" else 1)


# ------------------------------------------------------------------------------------------ CPU oracle legs
class OracleWorkload:
    """The workload's frames built with the CPU oracle (tests/oracle_lib.py: the reference's path restated in C).  Test
    infrastructure used as the MEASURED CPU baseline only -- never on the GPU arm's data path."""

    def __init__(self, workload):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        import scene_build as SB
        from bvhtracer_b200 import examples
        self.O, self.SB, self.examples, self.workload = O, SB, examples, workload
        self.kind = WORKLOADS[workload][2]
        self._bigben = None

    def frame(self, index):
        """-> (oracle scene, oracle camera) of animation frame `index` (static workloads: the one frame)"""
        O, SB, ex = self.O, self.SB, self.examples
        if self.kind == "grid":
            return SB.oracle_scene(ex.CONFIGS[self.workload](index))
        if self.kind == "bigben":
            if self._bigben is None:
                blas = O.Blas(O.load_asset("bigben.tri"))             # private copy: its vertices get animated
                scene = O.Scene([blas], [(0, O.mat4_identity())], with_transform=False)
                _, cam = SB.oracle_scene(ex.big_ben_clock())
                self._bigben = [blas, scene, cam, ex.BigBenAnimation(blas.tris), 0]
            blas, scene, cam, anim, at = self._bigben
            if index < at:
                self._bigben = None
                return self.frame(index)
            while at < index:                                          # big_ben_clock.rs:67-103: animate + ModelInstance::refit
                blas.tris[:] = anim.animate()
                at += 1
            blas.refit()
            scene.refresh_blas()
            self._bigben[4] = at
            return scene, cam
        return SB.oracle_scene(ex.CONFIGS[self.workload]())


def oracle_sample_rows(height, tile, fraction):
    """Every k-th tile row of the frame: a spatially uniform sample (top/sky and model rows in proportion)."""
    n_rows = (height + tile - 1) // tile
    k = max(1, int(round(1.0 / fraction)))
    return [r for r in range(k // 2, n_rows, k)]


def oracle_time_frame(ow, frame_index, width, height, rows, threads, tile=TILE):
    """Render the sampled tile rows of one frame with the CPU oracle; -> (rays, seconds, counters dict)."""
    O = ow.O
    scene, cam = ow.frame(frame_index)
    hits = np.zeros(width * height, O.HIT)
    total = O.Counters()
    rays, t0 = 0, time.perf_counter()
    for r in rows:
        c = O.Counters()
        y0, y1 = r * tile, min(height, (r + 1) * tile)
        scene.render(cam, width, height, tile=tile, region=(0, y0, width, y1), threads=threads, counters=c, out=hits)
        rays += (y1 - y0) * width
        for n, _ in O.Counters._fields_:
            if n.startswith("max_"):
                setattr(total, n, max(getattr(total, n), getattr(c, n)))
            else:
                setattr(total, n, getattr(total, n) + getattr(c, n))
    return rays, time.perf_counter() - t0, total.as_dict()


def algorithmic_bytes(counters):
    """B = 32*N_blas_nodes + 36*N_tris + 32*N_tlas_nodes + 64*N_inst + 16 per ray (reference objects)."""
    tris = counters["tri_area"] + counters["tri_u"] + counters["tri_v"] + counters["tri_t"]
    return 32 * counters["blas_nodes"] + 36 * tris + 32 * counters["tlas_nodes"] + 64 * counters["inst"] + 16 * counters["rays"]


def algorithmic_flops(counters):
    """MT 20/30/46/53 by exit stage, slab 22, instance entry 59, ray generation 94."""
    return (20 * counters["tri_area"] + 30 * counters["tri_u"] + 46 * counters["tri_v"] + 53 * counters["tri_t"]
            + 22 * counters["box_tests"] + 59 * counters["inst"] + 94 * counters["rays"])


# a fraction of tile rows per CPU step that keeps the reference arm within minutes on 16 host cores
CPU_FRACTION = {"cube": 1.0, "two_armadillos": 0.1, "sixteen_armadillos": 0.05, "trippy_teapots": 0.25, "big_ben_clock": 0.02}


def run_reference(args, rank, world):
    if rank != 0:
        return
    ow = OracleWorkload(args.workload)
    width, height = frame_size(args.workload, world, args.scaling)
    threads = ow.O.max_threads()
    frac = args.cpu_fraction if args.cpu_fraction > 0 else CPU_FRACTION[args.workload]
    rows = oracle_sample_rows(height, TILE, frac)
    for w in range(args.warmup):
        oracle_time_frame(ow, w + 1, width, height, rows[:max(1, len(rows) // 8)], threads)
    rays = secs = 0
    t_begin = time.perf_counter()
    for k in range(args.steps):
        r, s, _ = oracle_time_frame(ow, args.warmup + 1 + k, width, height, rows, threads)
        rays += r; secs += s
    wall = time.perf_counter() - t_begin
    value = rays / secs / 1e6
    sample = f"every {max(1, int(round(1.0 / frac)))}th 8-pixel tile row of each frame ({len(rows)} rows, {rays // args.steps} rays/step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "reference assets, example camera/animation",
        "config": {"workload": args.workload, "baseline_config": WORKLOADS[args.workload][0], "width": width, "height": height, "tile": TILE,
                   "frames": f"{args.warmup + 1}..{args.warmup + args.steps}",
                   "note": "CPU oracle = C restatement of the reference's Rust path (no cargo/rustc here); bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class GpuWorkload:
    """One BASELINE config driven through the C++ host mirror (the reference-facing API): scene, per-frame host update."""

    def __init__(self, name, flags, device):
        from bvhtracer_b200 import examples, host
        self.name, self.host, self.examples = name, host, examples
        self.tag, self.shading, self.kind = WORKLOADS[name]
        spec = examples.CONFIGS[name](0) if self.kind == "grid" else examples.CONFIGS[name]()
        self.spec = spec
        self.scene, self.models = host.build_scene(spec)
        self.renderer = host.Renderer(flags=flags, device=device, tile=TILE)
        self.eng = self.renderer.engine()
        self.cam = self.scene.camera()
        self.anim = examples.GridAnimation() if self.kind == "grid" else None
        self.bigben = examples.BigBenAnimation(self.models[0].primitives()) if self.kind == "bigben" else None
        self.frame = 0
        self.transforms = [host.object_transform(o) for o in spec.objects]
        self.queue = collections.deque()
        # static scenes whose objects all carry a transform re-set them in one call per frame
        self.static_m = (np.stack([t.matrix for t in self.transforms])
                         if self.kind == "static" and all(o.with_transform for o in spec.objects) else None)
        self.pipeline = {"depth": host.depth_pipeline(80.0, 3.0), "normal": host.normal_pipeline(),
                         "intersection": host.intersection_pipeline()}[self.shading]

    def shade_params(self):
        e = self.eng
        if self.shading == "depth":
            return e.shade_depth(80.0, 3.0)                  # DepthAccumulator + DepthMappingShader::new(80, 3)
        if self.shading == "intersection":
            return e.shade_intersection()
        return e.shade_normal(self.transforms[0].matrix)     # scene.get_unchecked(0).get_transform(), renderer.rs:275-278

    def precompute(self, n):
        """The next n frames' animation results (the example's own arithmetic: transforms of the grid, vertices of big_ben_clock),
        so that a loop with frames in flight is not bound by this script's Python: advance() then only applies them to the scene."""
        host = self.host
        for _ in range(n):
            if self.kind == "grid":
                self.anim.update()
                tr = [host.object_transform(o) for o in self.anim.objects()]
                self.queue.append((tr, np.stack([t.matrix for t in tr])))
            elif self.kind == "bigben":
                self.queue.append(self.bigben.animate().copy())

    def advance(self):
        """AppState::update of the example, on the host (never inside a timed interval)."""
        host = self.host
        self.frame += 1
        if self.kind == "grid":                              # sixteen_armadillos.rs:132-163: 16 x set_transform + Scene::rebuild
            if self.queue:
                self.transforms, m = self.queue.popleft()
                self.scene.set_transforms(m)
            else:
                self.anim.update()
                self.transforms = [host.object_transform(o) for o in self.anim.objects()]
                for i, t in enumerate(self.transforms):
                    self.scene.set_transform(i, t)
            self.scene.rebuild()
        elif self.kind == "bigben":                          # big_ben_clock.rs:67-103: animate() + ModelInstance::refit
            self.models[0].set_primitives(self.queue.popleft() if self.queue else self.bigben.animate())
            self.models[0].refit()
        else:                                                # cube / two_armadillos: Scene::run re-sets the transforms and rebuilds
            if self.static_m is not None:                    # the TLAS every frame (scene.rs:40-50); fixed pose here
                self.scene.set_transforms(self.static_m)
            else:
                for i, t in enumerate(self.transforms):
                    if self.spec.objects[i].with_transform:
                        self.scene.set_transform(i, t)
            self.scene.rebuild()

    def goto(self, frame):
        """A fresh animation advanced to `frame` (for the work-counter pass after the timed loops)."""
        ex = self.examples
        if self.kind == "grid":
            self.anim = ex.GridAnimation()
        elif self.kind == "bigben":
            self.bigben.r = np.float32(0)
        self.frame = 0
        for _ in range(frame):
            self.advance()


def lane_op_weights():
    p = os.path.join(ROOT, "profiles", "lane_op_weights.json")
    try:
        return json.load(open(p))["weights_thread_instructions_per_event"], os.path.relpath(p, ROOT)
    except Exception:
        return None, None


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from bvhtracer_b200 import host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    flags = MODES[args.mode]
    width, height = frame_size(args.workload, world, args.scaling)
    npix = width * height
    tile = TILE

    wl = GpuWorkload(args.workload, flags, local_rank)
    renderer, eng, scene, cam = wl.renderer, wl.eng, wl.scene, wl.cam
    stream = torch.cuda.Stream()                         # a real (non-NULL) stream: the library launches on it, so
    torch.cuda.set_stream(stream)                        # torch.cuda.Event sees exactly the kernels we time
    renderer.set_stream(stream.cuda_stream)

    # ---- resident output.  N = 1: the frame of 16-byte hit records in HBM.  N > 1, --gather frame (default): the FRAME BUFFER
    # (Rgba<u8>, the example's accumulator + pixel shader fused into the trace kernel) lives in rank 0's HBM, peers map it and
    # store into it over NVLink P2P; the hit records of a rank's tile rows stay in that rank's HBM.  --gather hits: the 16-byte
    # records themselves are gathered on rank 0.
    gather_frame = world > 1 and args.gather == "frame"
    d_frame = None
    if gather_frame:
        d_hits = eng.device_alloc(npix * 16)             # local: only this rank's tile rows are ever written
        shared = eng.device_alloc(npix * 4) if rank == 0 else None
    else:
        shared = eng.device_alloc(npix * 16) if rank == 0 else None
    if world > 1:
        hb = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            hb.copy_(torch.frombuffer(bytearray(eng.ipc_export(shared)), dtype=torch.uint8))
        dist.broadcast(hb, 0)
        if rank != 0:
            shared = eng.ipc_open(bytes(hb.cpu().numpy().tobytes()))
        eng.set_shard(rank, world)
    if gather_frame:
        d_frame = shared
    else:
        d_hits = shared
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def resident_step():
        """-> list of (event, event) pairs whose intervals add up to the frame's device time"""
        pairs = []
        if wl.kind == "bigben":
            # the frame's device work starts with the new vertices: upload (754 KB) + sub-BVH refit + re-bake + K2 (Bvh::refit)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            renderer.sync_scene(scene)
            e1.record(stream)
            pairs.append((e0, e1))
        else:
            renderer.sync_scene(scene)                   # this frame's TLAS/instances -> HBM (outside the timed interval)
            flush.zero_()                                # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.render_frame_device(cam, width, height, wl.shade_params() if gather_frame else None, tile, None, d_frame, d_hits)
        e1.record(stream)
        pairs.append((e0, e1))
        return pairs

    # ---- warm-up
    for _ in range(args.warmup):
        wl.advance()
        resident_step()
    barrier()

    # ---- timed: resident (value)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(gpu_uuid, local_rank, args.clock_period_ms * 1e-3)
    sampler.start()
    launches0 = renderer.stats()["kernel_launches"]
    barrier()
    wall0 = time.perf_counter()
    events = []
    for _ in range(args.steps):
        wl.advance()
        events.append(resident_step())
    barrier()
    wall_resident = time.perf_counter() - wall0
    launches = renderer.stats()["kernel_launches"] - launches0
    ms = [sum(a.elapsed_time(b) for a, b in pairs) for pairs in events]
    trace_ms = [pairs[-1][0].elapsed_time(pairs[-1][1]) for pairs in events]
    t_resident = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_resident, op=dist.ReduceOp.MAX)
    total_ms = float(t_resident.item())
    kernel_ms_local = float(np.mean(trace_ms))
    first_timed_frame, last_timed_frame = wl.frame - args.steps + 1, wl.frame

    # the frame assembled in rank 0's HBM by all ranks (last timed frame) must equal the same frame rendered by rank 0 alone
    gathered_ok = None
    if gather_frame:
        barrier()
        if rank == 0:
            got = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_frame)
            eng.set_shard(0, 1)
            d_single = eng.device_alloc(npix * 4)
            eng.render_frame_device(cam, width, height, wl.shade_params(), tile, None, d_single, None)
            eng.sync()
            single = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_single)
            eng.device_free(d_single)
            eng.set_shard(rank, world)
            gathered_ok = bool(np.array_equal(single, got))
            del got, single
        barrier()

    # ---- timed: end to end through Renderer::render.  N > 1: every rank renders its tile rows (shaded frame) and copies them,
    # over ITS OWN PCIe link, into one page-locked frame buffer in POSIX shared memory.  Same clock at every N: CUDA events
    # around the call on the rank's stream, summed over the steps, max over ranks.
    shm = None
    sharded_ok = None
    if world == 1:
        state = host.RendererState(wl.pipeline, width, height, keep_hits=False)
    else:
        from multiprocessing import shared_memory
        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=npix * 4)
            name = [shm.name]
        dist.broadcast_object_list(name, 0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
            try:                                         # attaching ranks must not let their resource tracker unlink the segment
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        host_frame = np.ndarray(npix, dtype=np.uint32, buffer=shm.buf)
        eng.host_register(host_frame)
        state = host.RendererState(wl.pipeline, width, height, frame=host_frame)
    for _ in range(max(2, args.warmup)):
        wl.advance()
        renderer.render(state, scene)
    barrier()
    h2d0, d2h0 = renderer.stats()["h2d_bytes"], renderer.stats()["d2h_bytes"]
    ev = []
    e2e_first_frame = wl.frame + 1
    for _ in range(args.steps):
        wl.advance()                                     # host-side scene update (not part of render(), as in the reference)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        renderer.render(state, scene)                    # uploads + refit + trace + shade + D2H of the (rank's rows of the) frame
        e1.record(stream)
        ev.append((e0, e1))
    barrier()
    e2e_ms = [a.elapsed_time(b) for a, b in ev]
    t_e2e = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device="cuda")
    io = torch.tensor([renderer.stats()["h2d_bytes"] - h2d0, renderer.stats()["d2h_bytes"] - d2h0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
    e2e_total_ms = float(t_e2e.item())
    e2e = {"value": npix * args.steps / (e2e_total_ms * 1e-3) / 1e6, "unit": "Mrays/s",
           "h2d_bytes_per_step": int(io[0].item()) // args.steps, "d2h_bytes_per_step": int(io[1].item()) // args.steps,
           "ms_per_step": e2e_total_ms / args.steps,
           "call": "Renderer::render -> CudaPathTracer::evaluate (bvht_blas_update_vertices/refit when vertices moved, bvht_tlas_set, bvht_render_frame), "
                   "Rgba<u8> frame buffer to page-locked host memory" + ("" if world == 1 else
                   "; per rank: its tile rows over its own PCIe link into ONE frame in POSIX shared memory; sum of per-step CUDA-event intervals, max over ranks")}
    if world == 1:
        frame_checksum = int(np.bitwise_xor.reduce(state.frame_buffer()))
        # the SAME frames of the animation with two in flight (Renderer::render_begin / render_end): frame n's copies run under
        # frame n+1's kernels.  One interval around the whole loop -- scene updates, L2 flushes and the last frame's copies
        # included -- closed after the last render_end has returned.
        state_b = host.RendererState(wl.pipeline, width, height, keep_hits=False)
        pair = [state, state_b]
        wl.goto(max(e2e_first_frame - 1 - 4, 0))
        for i in range(min(4, e2e_first_frame - 1)):
            wl.advance()
            renderer.render_begin(pair[i & 1], scene)
            if i >= 1:
                renderer.render_end()
        renderer.render_end()
        barrier()
        wl.precompute(args.steps)
        d2h1 = renderer.stats()["d2h_bytes"]
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fl = []
        p0.record(stream)
        for i in range(args.steps):
            wl.advance()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream); flush.zero_(); f1.record(stream)
            fl.append((f0, f1))
            renderer.render_begin(pair[i & 1], scene)
            if i >= 1:
                renderer.render_end()
        renderer.render_end()
        p1.record(stream)
        torch.cuda.synchronize()
        loop_ms = p0.elapsed_time(p1)
        flush_ms = sum(a.elapsed_time(b) for a, b in fl)
        pipelined_checksum = int(np.bitwise_xor.reduce(pair[(args.steps - 1) & 1].frame_buffer()))
        e2e["two_frames_in_flight"] = {
            "value": npix * args.steps / (loop_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": loop_ms / args.steps,
            "l2_flush_ms_per_step": flush_ms / args.steps,
            "d2h_bytes_per_step": (renderer.stats()["d2h_bytes"] - d2h1) // args.steps,
            "frames": [e2e_first_frame, e2e_first_frame + args.steps - 1],
            "call": "Renderer::render_begin / render_end over two RendererStates, the frames of the e2e loop above; ONE CUDA-event "
                    "interval around the whole loop (scene updates, the L2 flush before every frame and the last frame's copies inside it)",
            "last_frame_checksum_xor": pipelined_checksum, "last_frame_equals_render": pipelined_checksum == frame_checksum}
        del state_b
        if wl.kind != "bigben":
            # the same call returning the 16-byte hit records as well
            state_h = host.RendererState(wl.pipeline, width, height, keep_hits=True)
            renderer.render(state_h, scene)
            evh = []
            for _ in range(max(3, args.steps // 4)):
                wl.advance()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); renderer.render(state_h, scene); e1.record(stream)
                evh.append((e0, e1))
            torch.cuda.synchronize()
            hm = [a.elapsed_time(b) for a, b in evh]
            e2e["with_hit_records"] = {"value": npix / (float(np.mean(hm)) * 1e-3) / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": npix * 20}
            del state_h
    else:
        frame_checksum = int(np.bitwise_xor.reduce(host_frame)) if rank == 0 else 0
        # the assembled frame must equal the same frame rendered by rank 0 alone (outside every timed region)
        barrier()
        if rank == 0:
            eng.set_shard(0, 1)
            d_single = eng.device_alloc(npix * 4)
            renderer.sync_scene(scene)
            eng.render_frame_device(cam, width, height, wl.shade_params(), tile, None, d_single, None)
            eng.sync()
            single = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_single)
            eng.device_free(d_single)
            eng.set_shard(rank, world)
            sharded_ok = bool(np.array_equal(single, host_frame))
        barrier()
    clocks = sampler.stop()

    # ---- work counters of sampled timed frames (instrumented build of the same kernels, launched the same way): the
    # lane-instructions the frame PERFORMED, for the roofline
    performed = None
    weights, weights_src = lane_op_weights()
    if rank == 0 and weights is not None and args.mode == "strict-accel":
        frames = sorted({first_timed_frame, (first_timed_frame + last_timed_frame) // 2, last_timed_frame})
        # duration of the trace kernel ALONE (bvht_stats.last_k1_ms: CUDA events around K1 on its stream), frame by frame, under the
        # timed loop's own conditions -- same frames, L2 flushed before each -- but with a synchronisation per frame to read it
        d_probe = eng.device_alloc(npix * 16) if world > 1 and not gather_frame else d_hits
        k1_of, bake_log = {}, [(renderer.stats()["rebakes"], round(renderer.stats()["bake_d_max"], 3), round(renderer.stats()["bake_o_max"], 3))]
        wl.goto(first_timed_frame - 1)
        for f in range(first_timed_frame, last_timed_frame + 1):
            wl.advance()
            renderer.sync_scene(scene)
            flush.zero_()
            eng.render_frame_device(cam, width, height, None, tile, None, None, d_probe)
            eng.sync()
            k1_of[f] = float(renderer.stats()["last_k1_ms"])
            bake_log.append((renderer.stats()["rebakes"], round(renderer.stats()["bake_d_max"], 3), round(renderer.stats()["bake_o_max"], 3)))
        per_frame = []
        for f in frames:
            wl.goto(f)
            renderer.sync_scene(scene)
            c = eng.debug_trace_stats(cam, width, height, tile)
            per_frame.append((sum(weights[k] * c[k] for k in weights), c, k1_of[f]))
        if d_probe is not d_hits:
            eng.device_free(d_probe)
        performed = {"lane_ops_per_frame": float(np.mean([p for p, _, _ in per_frame])), "frames": frames,
                     "k1_ms_per_frame": [t for _, _, t in per_frame], "k1_ms_mean_all_timed_frames": float(np.mean(list(k1_of.values()))),
                     "k1_ms_all_timed_frames": [round(k1_of[f], 4) for f in sorted(k1_of)], "bake_log": bake_log[:3] + bake_log[-2:],
                     "lane_ops_per_ms": float(np.mean([p / t for p, _, t in per_frame])),
                     "counters_per_ray": {k: float(np.mean([c[k] / max(c["rays"], 1) for _, c, _ in per_frame])) for k in per_frame[0][1] if k != "rays"}}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the box's host cores, bounded sample of the same frame
    cpu = None
    roof = None
    peak, peak_src = measured_peak()
    if rank == 0:
        counters = None
        if world == 1 and not args.no_cpu_baseline:
            ow = OracleWorkload(args.workload)
            threads = ow.O.max_threads()
            frac = args.cpu_baseline_fraction if args.cpu_baseline_fraction > 0 else min(1.0, 10 * CPU_FRACTION[args.workload])
            rows = oracle_sample_rows(height, tile, frac)      # ~10 s of CPU work on 16 cores
            r, s, counters = oracle_time_frame(ow, first_timed_frame, width, height, rows, threads)
            r1, s1, _ = oracle_time_frame(ow, first_timed_frame, width, height, rows[::16] or rows[:1], 1)
            cpu = {"value": r / s / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"frame {first_timed_frame}: every {max(1, int(round(1.0 / frac)))}th 8-pixel tile row ({len(rows)} rows, {r} rays, {s:.1f} s)",
                   "single_thread": {"value": r1 / s1 / 1e6, "unit": "Mrays/s", "rays": r1, "seconds": s1,
                                     "note": "the reference itself is single-threaded (renderer.rs:353-368)"}}
        rays_per_launch = npix / world
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        lane_peak = 148 * 128 * sm_mhz * 1e6 / 1e12          # T lane-instructions/s: 148 SMs x 4 schedulers x 32 lanes x sampled SM clock
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"{args.workload}:{args.mode}", {}).get("dram_bytes_per_frame")
            except Exception:
                traffic = None
        k1_ms = performed["k1_ms_mean_all_timed_frames"] if performed is not None else None
        roof = {"bound": "issue", "kernel": "trace_primary_kernel", "launch_ms": k1_ms if k1_ms is not None else kernel_ms_local,
                "frame_ms": kernel_ms_local, "traffic": traffic,
                "unit": "T lane-instructions/s", "peak": lane_peak,
                "peak_source": f"148 SMs x 4 schedulers x 32 lanes x {sm_mhz:.0f} MHz (SM clock sampled during the timed region); the path is "
                               "instruction-issue bound, not HBM or tensor bound (DESIGN.md 4)",
                "traffic_note": "DRAM bytes of all kernels of one resident frame (K7 + K0 + K1) from an ncu pass of the same workload (profiles/traffic.json)",
                "hbm": {"compulsory_gbs": 16.0 * rays_per_launch / (kernel_ms_local * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src,
                        "frac": 16.0 * rays_per_launch / (kernel_ms_local * 1e-3) / 1e9 / peak,
                        "note": "the only compulsory HBM traffic is the 16 B/ray hit record; the scene is L1/L2 resident"}}
        if performed is not None:
            achieved = performed["lane_ops_per_ms"] * 1e3 / 1e12
            roof.update({"achieved": achieved, "frac": achieved / lane_peak, "performed": performed, "weights": weights_src,
                         "note": "achieved = lane-instructions the frame PERFORMED (work counters of the instrumented build x per-event "
                                 "instruction counts calibrated against an ncu source-level capture, tools/lane_op_weights.py) / duration of the "
                                 "trace kernel alone (launch_ms: CUDA events around K1 on its stream, bvht_stats.last_k1_ms, sampled frames; "
                                 "frame_ms adds the coverage raster and K0); "
                                 "frac = ncu's issue slots busy per ELAPSED cycle x lanes per instruction / 32 (profiles/r02_lane_op_model_validation.txt)"})
        else:
            roof.update({"achieved": None, "frac": None,
                         "note": "no performed-work estimate (needs --mode strict-accel and profiles/lane_op_weights.json)"})
        if counters is not None:
            b_per_ray = algorithmic_bytes(counters) / counters["rays"]
            f_per_ray = algorithmic_flops(counters) / counters["rays"]
            alg_gbs = b_per_ray * rays_per_launch / (kernel_ms_local * 1e-3) / 1e9
            roof["algorithmic"] = {
                "bytes_per_ray": b_per_ray, "fp32_ops_per_ray": f_per_ray, "gbs": alg_gbs, "vs_hbm_peak": alg_gbs / peak,
                "source": f"oracle counters on the cpu_baseline sample of frame {first_timed_frame}",
                "note": "SURVEY.md 8(d): what the REFERENCE's brute-force leaf traversal fetches / computes for these rays; the leaf "
                        "accelerator answers the same question without doing that work, hence >> 1 against any peak"}
            if performed is not None:
                roof["algorithmic_speedup"] = f_per_ray * rays_per_launch / performed["lane_ops_per_frame"]

    if rank == 0:
        value = npix * args.steps / (total_ms * 1e-3) / 1e6
        n_tris = [int(m.primitives().shape[0]) for m in wl.models]
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (" + ", ".join(wl.spec.meshes) + "), example camera and animation",
            "config": {"workload": args.workload, "baseline_config": wl.tag, "width": width, "height": height, "tile": tile, "mode": args.mode,
                       "frames": f"{first_timed_frame}..{last_timed_frame}", "instances": len(wl.spec.objects), "triangles_per_blas": n_tris,
                       "shading": wl.shading,
                       "l2": "flushed between timed iterations (256 MiB memset outside the per-step CUDA-event intervals)",
                       "timing": "sum over steps of CUDA-event intervals around the frame's device work on the launching stream, max over ranks"
                                 + ("; C5: vertex upload + sub-BVH refit/re-bake + Bvh::refit (K2) + trace" if wl.kind == "bigben" else ""),
                       "sharding": ("single GPU" if world == 1 else
                                    "tile rows interleaved over ranks; the Rgba<u8> frame buffer is assembled in rank 0's HBM by P2P stores over NVLink, "
                                    "the 16-byte hit records of a rank's rows stay in its own HBM" if gather_frame else
                                    "tile rows interleaved over ranks; the 16-byte hit records are assembled in rank 0's HBM by P2P stores over NVLink"),
                       "parity": "strict modes are bit-identical to the CPU oracle (tests/test_gpu_parity.py)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "frame_checksum": frame_checksum, "wall_s_resident_loop": wall_resident,
            "sharded_frame_equals_single_gpu": sharded_ok, "gathered_device_frame_equals_single_gpu": gathered_ok,
        }
    del state
    if world > 1:
        barrier()
        eng.host_unregister(host_frame)
        del host_frame
        try:
            shm.close()
            if rank == 0:
                shm.unlink()
        except Exception:
            pass
        if rank != 0:
            eng.ipc_close(shared)
        barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stderr.flush()
        print(json.dumps(line), flush=True)              # the ONE JSON line, last thing on stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="strict-accel", choices=sorted(MODES))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--cpu-fraction", type=float, default=0.0, help="fraction of tile rows the CPU oracle renders per frame (0 = per-workload default)")
    ap.add_argument("--cpu-baseline-fraction", type=float, default=0.0, help="fraction of tile rows of ONE frame for the cpu_baseline leg (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=float, default=2.0, help="NVML sampling period of the SM clock / throttle reasons during the timed regions")
    ap.add_argument("--gather", default="frame", choices=["frame", "hits"],
                    help="N > 1: what is assembled on rank 0 over NVLink P2P (frame = Rgba<u8> frame buffer, hits = the 16-byte records)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
