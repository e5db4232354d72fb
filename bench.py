#!/usr/bin/env python3
"""bench.py -- primary closest-hit throughput (Mrays/s) of the B200 engine on the reference's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode strict-accel|...]

Workload (BASELINE.json configs[2], the one the metric's target is quoted on): `sixteen_armadillos` at
3840x2160 -- 16 animated instances of one 30,001-triangle armadillo BLAS, TLAS rebuilt per frame on the host.
One STEP = one animation frame: every primary ray of the frame traced to its closest hit.
  value  : rays / device time of the trace (scene, TLAS, instances and camera already resident in HBM; the
           16-byte hit records stay in HBM).
  e2e    : the same frame through the reference-facing plugin call `Renderer::render` of the C++ host mirror
           (CudaPathTracer::evaluate: per-frame TLAS/instance upload from pinned memory + trace + on-device
           accumulator/pixel shader + frame buffer copied back to pinned host memory), timed with CUDA events.
Multi-GPU (torchrun, one rank per GPU): the scene is replicated, tile rows are interleaved across ranks, every
rank stores its pixels (Rgba<u8>, the example's Depth accumulator + shader fused into the trace kernel) straight
into rank 0's frame buffer over NVLink P2P (CUDA IPC mapping), no collective in the data path; the 16-byte hit
records of a rank's rows stay in that rank's HBM, as at N = 1.  `--gather hits` assembles the 16-byte records on
rank 0 instead (at N = 8 that saturates rank 0's NVLink ingress: 0.93 GB per step).  Scaling is WEAK: the frame
grows to 3840 x (2160 * N) so per-GPU work is constant.
`--impl reference` times the CPU oracle (the C restatement of the reference's Rust path; the reference itself
cannot be compiled here: no cargo/rustc) with all host threads on a bounded sample of the same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

W4K, H4K = 3840, 2160
METRIC = "Mrays/s closest-hit (primary)"
MODES = {"strict-brute": 0x0, "strict-accel": 0x2, "fast-brute": 0x1, "fast-accel": 0x3}


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

    def __init__(self, uuid=None, index=0):
        self.uuid, self.index = uuid, index
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
                time.sleep(0.002)
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
                "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "source": "nvml, 2 ms period, whole timed region"}

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


# ------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_sample_rows(height, tile, fraction):
    """Every k-th tile row of the frame: a spatially uniform sample (top/sky and model rows in proportion)."""
    n_rows = (height + tile - 1) // tile
    k = max(1, int(round(1.0 / fraction)))
    return [r for r in range(k // 2, n_rows, k)]


def oracle_time_frame(frame_index, width, height, rows, threads, tile=8):
    """Render the sampled tile rows of one frame with the CPU oracle; -> (rays, seconds, counters dict)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import scene_build as SB
    from bvhtracer_b200 import examples
    scene, cam = SB.oracle_scene(examples.sixteen_armadillos(frame_index))
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    height = H4K * (world if args.scaling == "weak" else 1)
    threads = O.max_threads()
    rows = oracle_sample_rows(height, 8, args.cpu_fraction)
    for w in range(args.warmup):
        oracle_time_frame(w, W4K, height, rows[:max(1, len(rows) // 8)], threads)
    rays = secs = 0
    t_begin = time.perf_counter()
    for k in range(args.steps):
        r, s, _ = oracle_time_frame(args.warmup + k, W4K, height, rows, threads)
        rays += r; secs += s
    wall = time.perf_counter() - t_begin
    value = rays / secs / 1e6
    sample = f"every {max(1, int(round(1.0 / args.cpu_fraction)))}th 8-pixel tile row of each frame ({len(rows)} rows, {rays // args.steps} rays/step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "reference assets (armadillo.tri), example camera/animation",
        "config": {"workload": "sixteen_armadillos", "width": W4K, "height": height, "tile": 8, "frames": f"{args.warmup}..{args.warmup + args.steps - 1}",
                   "note": "CPU oracle = C restatement of the reference's Rust path (no cargo/rustc here); bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from bvhtracer_b200 import _ffi, examples, host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    flags = MODES[args.mode]
    width = W4K
    height = H4K * (world if args.scaling == "weak" else 1)
    npix = width * height
    tile = 8

    # ---- scene through the C++ host mirror (the reference-facing API)
    anim = examples.GridAnimation()
    spec = examples.sixteen_armadillos(0)
    scene, models = host.build_scene(spec)
    renderer = host.Renderer(flags=flags, device=local_rank, tile=tile)
    stream = torch.cuda.Stream()                         # a real (non-NULL) stream: the library launches on it, so
    torch.cuda.set_stream(stream)                        # torch.cuda.Event sees exactly the kernels we time
    renderer.set_stream(stream.cuda_stream)
    eng = renderer.engine()
    cam = scene.camera()
    shade = eng.shade_depth(80.0, 3.0)                   # DepthAccumulator + DepthMappingShader::new(80, 3)

    def advance_frame():
        """AppState::update (sixteen_armadillos.rs:132-163): 16 x set_transform + Scene::rebuild, on the host."""
        anim.update()
        for i, o in enumerate(anim.objects()):
            scene.set_transform(i, host.object_transform(o))
        scene.rebuild()

    # ---- resident output.  N = 1: the frame of 16-byte hit records in HBM.  N > 1, --gather frame (default): the FRAME BUFFER
    # (Rgba<u8>, DepthAccumulator + DepthMappingShader fused into the trace kernel) lives in rank 0's HBM, peers map it and
    # store into it over NVLink P2P; the hit records of a rank's tile rows stay in that rank's HBM.  --gather hits: the 16-byte
    # records themselves are gathered on rank 0 (at N = 8 that is 0.93 GB per 1.3 ms step into one GPU: NVLink-ingress bound).
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
        renderer.sync_scene(scene)                       # this frame's TLAS/instances -> HBM (outside the timed interval)
        flush.zero_()                                    # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.render_frame_device(cam, width, height, shade if gather_frame else None, tile, None, d_frame, d_hits)
        e1.record(stream)
        return e0, e1

    # ---- warm-up
    for _ in range(args.warmup):
        advance_frame()
        resident_step()
    barrier()

    # ---- timed: resident (value)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(gpu_uuid, local_rank)
    sampler.start()
    launches0 = renderer.stats()["kernel_launches"]
    barrier()
    wall0 = time.perf_counter()
    events = []
    for _ in range(args.steps):
        advance_frame()
        events.append(resident_step())
    barrier()
    wall_resident = time.perf_counter() - wall0
    launches = renderer.stats()["kernel_launches"] - launches0
    ms = [a.elapsed_time(b) for a, b in events]
    t_resident = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_resident, op=dist.ReduceOp.MAX)
    total_ms = float(t_resident.item())
    kernel_ms_local = float(np.mean(ms))

    # the frame assembled in rank 0's HBM by all ranks (last timed frame) must equal the same frame rendered by rank 0 alone
    gathered_ok = None
    if gather_frame:
        barrier()
        if rank == 0:
            got = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_frame)
            eng.set_shard(0, 1)
            d_single = eng.device_alloc(npix * 4)
            eng.render_frame_device(cam, width, height, shade, tile, None, d_single, None)
            eng.sync()
            single = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_single)
            eng.device_free(d_single)
            eng.set_shard(rank, world)
            gathered_ok = bool(np.array_equal(single, got))
            del got, single
        barrier()

    # ---- timed: end to end through Renderer::render (N = 1) / sharded render + rank-0 read-back (N > 1)
    e2e = None
    if world == 1:
        state = host.RendererState(host.depth_pipeline(80.0, 3.0), width, height, keep_hits=False)
        for _ in range(max(2, args.warmup)):
            advance_frame()
            renderer.render(state, scene)
        h2d0, d2h0 = renderer.stats()["h2d_bytes"], renderer.stats()["d2h_bytes"]
        ev = []
        torch.cuda.synchronize()
        for _ in range(args.steps):
            advance_frame()                              # host-side scene update (not part of render(), as in the reference)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            renderer.render(state, scene)                # upload TLAS/instances + trace + shade + D2H frame buffer
            e1.record(stream)
            ev.append((e0, e1))
        torch.cuda.synchronize()
        e2e_ms = [a.elapsed_time(b) for a, b in ev]
        st = renderer.stats()
        e2e = {"value": npix * args.steps / (sum(e2e_ms) * 1e-3) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": (st["h2d_bytes"] - h2d0) // args.steps, "d2h_bytes_per_step": (st["d2h_bytes"] - d2h0) // args.steps,
               "ms_per_step": float(np.mean(e2e_ms)),
               "call": "Renderer::render -> CudaPathTracer::evaluate (bvht_tlas_set + bvht_render_frame), Rgba<u8> frame buffer to pinned host memory"}
        frame_checksum = int(np.bitwise_xor.reduce(state.frame_buffer()))
        # the same call returning the 16-byte hit records as well
        state_h = host.RendererState(host.depth_pipeline(80.0, 3.0), width, height, keep_hits=True)
        renderer.render(state_h, scene)
        ev = []
        for _ in range(max(3, args.steps // 4)):
            advance_frame()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); renderer.render(state_h, scene); e1.record(stream)
            ev.append((e0, e1))
        torch.cuda.synchronize()
        hm = [a.elapsed_time(b) for a, b in ev]
        e2e["with_hit_records"] = {"value": npix / (float(np.mean(hm)) * 1e-3) / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": npix * 20}
        del state_h
    else:
        # every rank renders its tile rows (shaded frame) and copies them, over ITS OWN PCIe link, into one page-locked
        # host frame buffer in POSIX shared memory; the frame is complete on the host after the closing barrier
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
        host_frame = np.ndarray(npix, dtype="<u4", buffer=shm.buf)
        eng.host_register(host_frame)
        h2d0, d2h0 = renderer.stats()["h2d_bytes"], renderer.stats()["d2h_bytes"]
        for _ in range(2):
            advance_frame()
            renderer.sync_scene(scene)
            eng.render_frame(cam, width, height, shade, tile, None, frame_out=host_frame)
        h2d0, d2h0 = renderer.stats()["h2d_bytes"], renderer.stats()["d2h_bytes"]
        barrier()
        e2e_s = 0.0
        for _ in range(args.steps):
            advance_frame()                              # host-side scene update: outside the timed region, as for N = 1
            barrier()
            t0 = time.perf_counter()
            renderer.sync_scene(scene)                   # per-frame H2D (TLAS + instances)
            eng.render_frame(cam, width, height, shade, tile, None, frame_out=host_frame)   # trace + shade + D2H of the owned rows
            barrier()                                    # every rank's rows are in the shared host frame
            e2e_s += time.perf_counter() - t0
        barrier()
        e2e_wall = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(e2e_wall, op=dist.ReduceOp.MAX)
        st_ = renderer.stats()
        e2e = {"value": npix * args.steps / float(e2e_wall.item()) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": (st_["h2d_bytes"] - h2d0) // args.steps, "d2h_bytes_per_step": (st_["d2h_bytes"] - d2h0) // args.steps * world,
               "ms_per_step": float(e2e_wall.item()) / args.steps * 1e3,
               "call": "per rank: bvht_tlas_set + bvht_render_frame (tile-row shard; trace + shade + D2H of the owned rows over the rank's own "
                       "PCIe link into ONE page-locked frame buffer in POSIX shared memory); host wall clock between barriers, max over ranks"}
        frame_checksum = int(np.bitwise_xor.reduce(host_frame)) if rank == 0 else 0
        # the assembled frame must equal the same frame rendered by rank 0 alone (outside every timed region)
        sharded_ok = None
        barrier()
        if rank == 0:
            eng.set_shard(0, 1)
            d_single = eng.device_alloc(npix * 4)
            eng.render_frame_device(cam, width, height, shade, tile, None, d_single, None)
            eng.sync()
            single = eng.memcpy_d2h(np.zeros(npix, "<u4"), d_single)
            eng.device_free(d_single)
            eng.set_shard(rank, world)
            sharded_ok = bool(np.array_equal(single, host_frame))
        barrier()
    clocks = sampler.stop()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the box's host cores, bounded sample of the same frame
    cpu = None
    roof = None
    peak, peak_src = measured_peak()
    if rank == 0:
        counters = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            threads = O.max_threads()
            rows = oracle_sample_rows(height, tile, args.cpu_baseline_fraction)      # ~10 s of CPU work on 16 cores
            r, s, counters = oracle_time_frame(args.warmup + 1, width, height, rows, threads)
            r1, s1, _ = oracle_time_frame(args.warmup + 1, width, height, rows[::16], 1)
            cpu = {"value": r / s / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"frame {args.warmup + 1}: every {max(1, int(round(1.0 / args.cpu_baseline_fraction)))}th 8-pixel tile row ({len(rows)} rows, {r} rays, {s:.1f} s)",
                   "single_thread": {"value": r1 / s1 / 1e6, "unit": "Mrays/s", "rays": r1, "seconds": s1,
                                     "note": "the reference itself is single-threaded (renderer.rs:353-368)"}}
        # roofline of the dominant kernel (trace_primary_kernel): algorithmic bytes per launch / launch duration
        if counters is not None:
            b_per_ray = algorithmic_bytes(counters) / counters["rays"]
            f_per_ray = algorithmic_flops(counters) / counters["rays"]
            src = f"oracle counters on the cpu_baseline sample of frame {args.warmup + 1}"
        else:
            b_per_ray, f_per_ray = 94000.0, 78000.0       # DESIGN.md table (frame 3 sample), used when the oracle leg is skipped
            src = "DESIGN.md per-ray figure (oracle leg skipped)"
        rays_per_launch = npix / world
        achieved = b_per_ray * rays_per_launch / (kernel_ms_local * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.mode)
            except Exception:
                traffic = None
        # the same algorithmic bytes against the L2 -> SM read bandwidth measured here with the library's
        # own streaming kernel, and the algorithmic f32 operations against the non-FMA issue peak
        try:
            l2_gbs = max(eng.debug_read_bandwidth(64 << 20, 50) for _ in range(2))
        except Exception:
            l2_gbs = None
        fp32_peak_tops = 148 * 128 * 1.965e9 / 1e12
        fp32_tops = f_per_ray * rays_per_launch / (kernel_ms_local * 1e-3) / 1e12
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "l2": {"peak": l2_gbs, "unit": "GB/s", "frac": (achieved / l2_gbs) if l2_gbs else None,
                       "peak_source": "measured in this run: bvht_debug_read_bandwidth, 64 MiB buffer (L2-resident), 50 sweeps"},
                "fp32": {"achieved": fp32_tops, "peak": fp32_peak_tops, "unit": "Tops/s (one op per add/mul/div/min/max lane, no FMA credit)",
                         "frac": fp32_tops / fp32_peak_tops},
                "kernel": "trace_primary_kernel", "launch_ms": kernel_ms_local, "algorithmic_bytes_per_ray": b_per_ray,
                "algorithmic_fp32_ops_per_ray": f_per_ray, "source": src, "peak_source": peak_src,
                "compulsory_hbm_gbs": 16.0 * rays_per_launch / (kernel_ms_local * 1e-3) / 1e9,
                "note": "algorithmic bytes are defined on the reference's brute-force leaf traversal (SURVEY.md 8d); the scene (1.4 MB) is "
                        "L1/L2-resident and the leaf accelerator skips most of those fetches, so frac >> 1 is expected for *-accel modes; "
                        "the compulsory HBM traffic is the 16 B/ray hit record (compulsory_hbm_gbs)"}

    if rank == 0:
        value = npix * args.steps / (total_ms * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (armadillo.tri 30,001 triangles), example camera and closed-form animation",
            "config": {"workload": "sixteen_armadillos", "width": width, "height": height, "tile": tile, "mode": args.mode,
                       "frames": f"{args.warmup + 1}..{args.warmup + args.steps}", "instances": 16, "triangles_per_blas": 30001,
                       "l2": "flushed between timed iterations (256 MiB memset outside the per-step CUDA-event intervals)",
                       "timing": "sum over steps of CUDA-event intervals around the trace launch on the launching stream, max over ranks",
                       "sharding": ("single GPU" if world == 1 else
                                    "tile rows interleaved over ranks; the Rgba<u8> frame buffer is assembled in rank 0's HBM by P2P stores over NVLink, "
                                    "the 16-byte hit records of a rank's rows stay in its own HBM" if gather_frame else
                                    "tile rows interleaved over ranks; the 16-byte hit records are assembled in rank 0's HBM by P2P stores over NVLink"),
                       "parity": "strict modes are bit-identical to the CPU oracle (tests/test_gpu_parity.py)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "frame_checksum": frame_checksum, "wall_s_resident_loop": wall_resident,
            "sharded_frame_equals_single_gpu": (sharded_ok if world > 1 else None),
            "gathered_device_frame_equals_single_gpu": gathered_ok,
        }
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
    ap.add_argument("--mode", default="strict-accel", choices=sorted(MODES))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-fraction", type=float, default=0.05, help="fraction of tile rows the CPU oracle renders per frame")
    ap.add_argument("--cpu-baseline-fraction", type=float, default=0.5, help="fraction of tile rows of ONE frame for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
