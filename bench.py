#!/usr/bin/env python
"""bench.py -- throughput of the rasterize() hot path on N B200s (driver contract in the task brief).

A step = one pass of the path over one batch of synthetic input: `--frames` (default 8) camera
positions of the workload scene rendered by one launch sequence (`rxc_rasterize_batch`), each frame a
full `Rasterizer::setup(..).rasterize(..)` (projection, clipping, binning, raster, shading, 2D pass).
Default workload: BASELINE.json's textured map scene at 3840x2160 (the config its target is quoted on).

  value     whole-job Mpixel/s with the scene resident in HBM and frames written to device memory
  e2e       same metric through the public API with pinned HOST pixel buffers (D2H inside the timing)
  roofline  k_raster: algorithmic bytes / CUDA-event time of that kernel, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a port of the reference algorithm) on this box's host cores

`--impl reference` times only the CPU port on the same workload (rank 0 only).
Multi-GPU: frames are sharded across ranks (one process per GPU, torchrun), no data-path collective.
  value              render-only: every rank's frames stay in its own HBM (weak scaling, F frames per GPU and step)
  value_with_gather  the same step with every rank's raster kernel writing its tiles straight into rank 0's
                     buffer over NVLink (rxc_mgpu_*: peer mapping, completion flags; no copy or collective after
                     the render), timed until rank 0's stream has seen every rank's flag
  sweep4096          BASELINE.json config E as specified: 4096 frames of the map scene at 1920x1080 in contiguous
                     blocks per rank, STRONG scaling, total-job seconds render-only and delivered to rank 0
  band_split         config D: one 7680x4320 frame of the 1M-triangle scene split into bands, render-only and
                     written in place into rank 0's frame
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (builder, kwargs, description)
    "map4k": ("sweep", dict(width=3840, height=2160, tile_size=40, real=True), "minigame map scene (world.rxm walls/floor/fence/sky with the minigame's own PNG tiles, 1 point light, logo rect), 3840x2160, Nearest, tile 40"),
    "teapot1080": ("teapot", dict(width=1920, height=1080, tile_size=60, real=True), "examples/teapot.obj (1202 vertices, 2256 triangles) textured with images/logo.png, 1920x1080, Linear, tile 60, orbit sweep"),
    "sweep1080": ("sweep", dict(width=1920, height=1080, tile_size=40, real=True), "camera sweep of the map scene, 1920x1080, Nearest, tile 40"),
    "dense8k": ("dense", dict(width=7680, height=4320, tile_size=40), "991,232-triangle heightfield in 1024 batches, 7 lights, 7680x4320, Linear"),
    "cube800": ("cube", dict(width=800, height=600, tile_size=200, real=True), "cube textured with images/logo.png, 800x600, Nearest, tile 200"),
    "cube2000": ("cube", dict(width=2000, height=2000, tile_size=40, real=True), "cube textured with images/logo.png, 2000x2000, tile 40 (the shape of benches/rasterize_cube.rs)"),
    # the rows SURVEY 8f marks "next", measured like the others
    "chunked1080": ("chunked", dict(width=1920, height=1080, tile_size=40), "chunked map (4 chunks: opacity panes with surface ids, terrain textures, occluded sectors, chunk lights, entity/item tiles, 2D overlay with lines), 1920x1080, Nearest"),
    "shaded1080": ("shaded", dict(width=1920, height=1080, tile_size=40), "batch shaders (Rusteria VM programs on 3D, chunk, opacity-pass and 2D batches; one program cuts holes through opacity), 1920x1080, Nearest"),
    "game2d1080": ("game2d", dict(width=1920, height=1280), "2D game screen (834 2D records in sorted per-tile lists, translucent decals, sprites, 2 point lights with line of sight, sector occlusion, lines), 1920x1280, 2D-only render mode"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name, frames_per_step, rank, world):
    from rusterix_b200 import scenes

    builder, kw, desc = WORKLOADS[name]
    cfg = scenes.BUILDERS[builder](**kw)
    # rank r renders frames r, r+world, ... of the camera path (frame sharding)
    n_path = max(cfg.n_frames, 1)
    if cfg.cameras is not None:
        stride = max(1, n_path // (frames_per_step * world)) if n_path >= frames_per_step * world else 1
        frame_ids = [((rank + i * world) * stride) % n_path for i in range(frames_per_step)]
    else:
        frame_ids = [0] * frames_per_step
    return cfg, frame_ids, desc


def kernels_note(ctx):
    """Which k_raster ran: the library's generic instantiation or the one recompiled for the scene (rx_jit.cu)."""
    info = ctx.vm_jit_info()
    return {"k_raster": "recompiled for the scene with NVRTC (scene / frame constants folded in, rx_jit.cu)" if info["launches"] else "generic instantiation of librxcuda.so",
            "scene_kernels_loaded": info["kernels"], "launches_of_scene_kernels": info["launches"], "compile_log": info["log"][:300]}


def make_config(desc, cfg, world):
    """The `config` object of the JSON line: the same keys and values in the native and the reference arm."""
    return {"workload": desc, "width": cfg.width, "height": cfg.height, "triangles": cfg.counts()[1], "tile_size": cfg.tile_size,
            "sample_mode": getattr(cfg.sample_mode, "name", str(cfg.sample_mode)), "lights": len(cfg.scene.all_lights())}


def cpu_port_time(cfg, frame_ids, budget_s=20.0, max_frames=10, warm=1):
    """Times the oracle (CPU port of the reference algorithm, all host threads) on whole frames of the
    same workload.  Returns (Mpixel/s, cores, sample description, seconds per frame)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi

    lib = oracle_ffi.load()
    cores = int(lib.rxo_hardware_threads())
    times = []
    t_start = time.time()
    k = 0
    while True:
        r = cfg.rasterizer(frame_ids[k % len(frame_ids)])
        t0 = time.perf_counter()
        oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, want_planes=False, n_threads=0)
        dt = time.perf_counter() - t0
        if k >= warm:
            times.append(dt)
        k += 1
        if len(times) >= max_frames or (time.time() - t_start > budget_s and len(times) >= 1):
            break
    med = statistics.median(times)
    mpix = cfg.width * cfg.height / med / 1e6
    return mpix, cores, f"{len(times)} whole frames of the workload after {warm} warm-up, median", med


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, frame_ids, desc = build_workload(args.workload, args.frames, 0, 1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi

    lib = oracle_ffi.load()
    cores = int(lib.rxo_hardware_threads())
    # a step of the reference arm = ONE frame of the workload (bounded sample of the GPU arm's step)
    def step(i):
        r = cfg.rasterizer(frame_ids[i % len(frame_ids)])
        oracle_ffi.rasterize(r, cfg.scene, cfg.assets, cfg.width, cfg.height, cfg.tile_size, want_planes=False, n_threads=0)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = cfg.width * cfg.height * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mpixels/s shaded", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "frames_per_s": args.steps / dt,
        "config": make_config(desc, cfg, 1),
        "step": "1 frame (a bounded sample of the GPU arm's multi-frame step; the metric is a rate)",
        "note": "the Rust reference cannot be built in this image (no cargo); this is the C++ port of its algorithm (oracle/rx_oracle.cpp), all host threads",
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": f"{args.steps} frames, 1 per step"},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


_JSON_FD = None


def capture_stdout():
    """Library chatter (e.g. NCCL's version banner) must not share stdout with the one JSON line the
    driver parses: everything written to fd 1 from here on goes to stderr, emit_line() writes the result
    to the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="map4k", choices=list(WORKLOADS))
    ap.add_argument("--frames", type=int, default=32, help="frames per step (camera batch); 32 views of the camera path = 1.06 GB of 4K frames per call: the per-call front end and the raster grid's ramp / tail are paid once per call (8 per call: 55.5 Gpixel/s device-resident / 13.5 end to end, 16: 58.2 / 13.8, 32: 58.5 / 14.0, 64: 58.9 / 14.0)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from rusterix_b200 import mgpu
    numa_cpus = mgpu.bind_to_gpu_numa_node(local_rank)  # before any pinned host allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from rusterix_b200 import DeviceContext, Rasterizer
    from rusterix_b200._abi import RXC_N_KERNELS

    cfg, frame_ids, desc = build_workload(args.workload, args.frames, rank, world)
    F = len(frame_ids)
    W, H = cfg.width, cfg.height
    frame_bytes = W * H * 4
    rasts = [cfg.rasterizer(i).on_device(local_rank) for i in frame_ids]
    ctx = DeviceContext.get(local_rank)
    # The raster kernel is recompiled for the scene (NVRTC: its constants folded in, batch shaders as straight-line code;
    # rusterix_b200/csrc/rx_jit.cu).  The library does that in the background and renders with its generic kernel meanwhile; the
    # bench asks for the compilation BEFORE the first frame (a few seconds inside the warm-up), so that the timed steps measure
    # the kernel a running application ends up with.  RXC_VM_JIT=0 benches the generic kernels.
    ctx.set_vm_jit(int(os.environ.get("RXC_VM_JIT", "2")))
    # a dedicated (non-default) torch stream: the kernels, the L2 flush and the torch.cuda.Events
    # all live on it, so the events bracket exactly the launches of a step
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    out_dev = torch.empty((F, H, W, 4), dtype=torch.uint8, device=dev)
    out_host = torch.empty((F, H, W, 4), dtype=torch.uint8, pin_memory=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    batch = Rasterizer.prepare_batch(rasts, cfg.scene, W, H, cfg.tile_size, cfg.assets, device=local_rank)

    def step_device():
        batch.run(out_dev, sync=False)

    def step_e2e():
        batch.run(out_host, sync=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
            flush.zero_()
        barrier()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        wall0 = time.perf_counter()
        for i in range(steps):
            e0[i].record()
            step_fn()
            e1[i].record()
            flush.zero_()  # evict the frames and the scene from L2 between timed steps (not timed)
        barrier()
        wall = time.perf_counter() - wall0
        ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
        timed.per_rank_ms = [ms / steps]
        if world > 1:
            every = torch.zeros(world, dtype=torch.float64, device=dev)
            every[rank] = ms / steps
            dist.all_reduce(every, op=dist.ReduceOp.SUM)
            timed.per_rank_ms = [round(float(x), 4) for x in every.tolist()]   # diagnostics: the value is computed from the slowest
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    # ---- device-resident throughput (the headline `value`)
    sampler = ClockSampler(local_rank)
    ctx.synchronize()
    step_device(); ctx.synchronize()   # first call uploads the scene / sizes the workspace
    ctx.reset_stats()
    if rank == 0:
        sampler.start()
    # warm-up is inside timed(); stats are reset after it by measuring launches per step separately
    ms_total, _ = timed(step_device, args.steps, args.warmup)
    per_rank_ms = list(timed.per_rank_ms)
    clocks = sampler.stop() if rank == 0 else None
    st = ctx.stats()
    launches_per_step = st.kernel_launches // (args.steps + args.warmup)
    gpu_launches = launches_per_step * args.steps
    ms_per_step = ms_total / args.steps
    pixels_per_step = F * W * H * world
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API with host pixel buffers
    e2e_steps = max(3, min(args.steps, 10))
    ctx.reset_stats()
    ms_e2e, wall_e2e = timed(step_e2e, e2e_steps, 3)
    st2 = ctx.stats()
    # synchronous API: host wall-clock is the honest end-to-end time (device events miss host staging)
    e2e_ms_step = max(ms_e2e, wall_e2e * 1e3) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms_step = float(t.item())
    e2e_value = pixels_per_step / (e2e_ms_step * 1e-3) / 1e6
    h2d_step = st2.h2d_bytes // (e2e_steps + 3)
    d2h_step = st2.d2h_bytes // (e2e_steps + 3)

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline
    ctx.set_profiling(True)
    ctx.reset_stats()
    prof_steps = 5
    for _ in range(prof_steps):
        step_device()
        flush.zero_()
    ctx.synchronize()
    stp = ctx.stats()
    ctx.set_profiling(False)
    names = ctx.kernel_names()
    kernel_ms = {names[i]: (stp.kernel_ms[i] / max(1, stp.launches[i])) for i in range(RXC_N_KERNELS) if stp.launches[i]}
    step_kernel_ms = sum(stp.kernel_ms[i] for i in range(RXC_N_KERNELS)) / prof_steps
    raster_ms = kernel_ms.get("k_raster", float("nan"))
    peak, peak_src = load_peaks()
    # algorithmic bytes of one k_raster launch: F frames written once + scene/textures/lights read once
    b_alg_frame = cfg.algorithmic_bytes()
    b_alg_launch = F * frame_bytes + (b_alg_frame - frame_bytes)
    achieved = b_alg_launch / (raster_ms * 1e-3) / 1e9
    traffic = None
    ncu_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_json):
        try:
            ncu_row = dict(json.load(open(ncu_json)).get(args.workload, {}))
            # the capture's per-launch counts, scaled to this run's frames per launch when they differ
            k_f = F / float(ncu_row.get("frames_per_launch") or F)
            for key in ("k_raster_dram_bytes_per_launch", "warp_instructions", "dram_read_bytes"):
                if ncu_row.get(key) is not None and k_f != 1.0:
                    ncu_row[key] = int(round(ncu_row[key] * k_f))
            traffic = ncu_row.get("k_raster_dram_bytes_per_launch")
        except Exception:
            ncu_row, traffic = {}, None
    else:
        ncu_row = {}
    roofline = {"bound": "hbm", "kernel": "k_raster", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg_launch,
                "kernel_ms_per_launch": raster_ms, "kernel_share_of_step": stp.kernel_ms[7] / prof_steps / step_kernel_ms if step_kernel_ms else None,
                "all_kernels_ms_per_launch": kernel_ms,
                # what actually bounds the kernel (from the committed ncu capture of this workload, profiles/): the SM
                # issue slots; `live` re-derives the instruction rate from this run's kernel time
                "sm_issue": {"issue_active_pct_ncu": ncu_row.get("smsp_issue_active_pct"),
                             "thread_instructions_per_pixel_ncu": ncu_row.get("thread_instructions_per_pixel"),
                             "warp_instructions_per_launch_ncu": ncu_row.get("warp_instructions"),
                             "live_warp_inst_per_clk_per_smsp": (ncu_row.get("warp_instructions") / (raster_ms * 1e-3 * (clocks or {}).get("sm_mhz", 1965.0) * 1e6 * 148 * 4)) if ncu_row.get("warp_instructions") and rank == 0 else None,
                             "profile_version": ncu_row.get("version")},
                "note": "the kernel is SM-issue bound, not HBM bound: about %s thread-instructions per pixel (exact divisions, no FMA contraction, per-light BRDF) keep the issue slots ~80 %% busy; see DESIGN.md section 5" % (ncu_row.get("thread_instructions_per_pixel") or 500)}

    # ---- the same step with delivery to rank 0 fused into the raster kernel's write-back (rxc_mgpu_*)
    delivered = None
    dl = None
    if world > 1:
        dl = mgpu.Delivery(ctx, rank, world)
        slots = 2
        dl.target(slots * world * F * frame_bytes)
        regions = [mgpu.frame_regions(world, F, frame_bytes, s_ * world * F * frame_bytes) for s_ in range(slots)]
        k_step = [0]

        def step_delivered():
            k = k_step[0]; k_step[0] += 1
            slot = k % slots
            if k >= slots:
                dl.release()                    # rank 0 is done with what this slot held two steps ago
            dl.render(batch, (slot * world + rank) * F * frame_bytes)
            dl.deliver(regions[slot])           # ranks > 0: flag behind the raster kernel; rank 0: wait for every flag

        ms_del, _ = timed(step_delivered, args.steps, args.warmup)
        ms_del /= args.steps
        stt = dl.status()
        nv_bytes = (world - 1) * F * frame_bytes
        delivered = {"value": pixels_per_step / (ms_del * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": ms_del,
                     "bytes_to_rank0_per_step": nv_bytes, "rank0_ingest_GBps": nv_bytes / (ms_del * 1e-3) / 1e9,
                     "mode": "peer writes over NVLink from k_raster's tile write-back (cudaIpc mapping of rank 0's buffer), completion flags" if stt["mode"] != "nccl" else "local staging + one ncclGroup of send/recv (no peer mapping)",
                     "timeouts": stt["timeouts"] if rank == 0 else None,
                     "timing": "CUDA events around render + deliver (+ release) on every rank's stream, max over ranks; rank 0's end event sits behind its wait for every rank's flag"}

    line = None
    if rank == 0:
        line = {
            "metric": "Mpixels/s shaded", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": F * world / (ms_per_step * 1e-3),
            "config": make_config(desc, cfg, world),
            "frames_per_step_per_gpu": F,
            "notes": {"sharding": "frames sharded across ranks, no data-path collective" if world > 1 else "single GPU",
                      "l2": "256 MiB flush between timed steps; each step also writes %.0f MB of frames (> 126 MB L2)" % (F * frame_bytes / 1e6),
                      "timing": "CUDA events per step on the launching stream, summed over steps, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                    "ms_per_step": e2e_ms_step, "steps": e2e_steps, "api": "Rasterizer.rasterize_batch -> rxc_rasterize_batch, pinned host pixels",
                    "host_numa_binding": ("rank pinned to the %d cores next to its GPU" % len(numa_cpus)) if numa_cpus else "none"},
            "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step),
            "roofline": roofline, "clocks": clocks,
            "kernels": kernels_note(ctx), "ms_per_step_per_rank": per_rank_ms,
        }
        if delivered:
            line["value_with_gather"] = delivered["value"]
            line["delivered"] = delivered
    if not args.no_extras:
        sw = sweep4096(rank, world, local_rank, dev, barrier, dl)
        if rank == 0:
            line["sweep4096"] = sw
    if world > 1 and not args.no_extras:
        bs = band_split(rank, world, local_rank, dev, flush, barrier, dl)
        if rank == 0:
            line["band_split"] = bs

    # ---- CPU baseline and secondary workloads: rank 0, N=1 only
    if rank == 0 and world == 1 and not args.no_extras:
        mpix, cores, sample, spf = cpu_port_time(cfg, frame_ids)
        line["cpu_baseline"] = {"value": mpix, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample, "s_per_frame": spf}
        also = {}
        for wname in ("cube800", "cube2000", "teapot1080", "dense8k", "sweep1080", "chunked1080", "game2d1080", "shaded1080"):
            if wname == args.workload:
                continue
            try:
                also[wname] = secondary(wname, local_rank, dev, flush)
            except Exception as e:  # a secondary workload never hides the headline
                also[wname] = {"error": repr(e)}
        try:
            also["frame_loop_dense4k"] = frame_loop(local_rank, dev)
        except Exception as e:
            also["frame_loop_dense4k"] = {"error": repr(e)}
        line["also"] = also
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        emit_line(line)
    if dl is not None:
        ctx.synchronize()
        dl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sweep4096(rank, world, local_rank, dev, barrier, dl, total_frames=4096, per_launch=int(os.environ.get("RXB_SWEEP_PER_LAUNCH", "128"))):
    """BASELINE.json config E as specified: 4096 frames of the map scene at 1920x1080, sharded by frame across the
    ranks in contiguous blocks -- STRONG scaling (the job is the same 4096 frames at every N).  A rank renders its
    block `per_launch` frames per launch sequence (128 frames = 1.06 GB of frames: more than the L2 holds, so no flush
    is needed between launches; one front-end pass per launch, so its ~60 us are paid 32 times per 4096 frames instead of
    128 times with 32 frames per launch: 0.1698 -> 0.1640 s on one GPU, 0.1631 with 256, 0.1627 with 512).  Two totals: render-only (frames stay in the rank's HBM, a ring of one launch) and delivered
    (every launch written straight into a two-slot ring in rank 0's memory, rxc_mgpu_*, release hand-shake included)."""
    import torch
    import torch.distributed as dist
    from rusterix_b200 import DeviceContext, Rasterizer, mgpu

    cfg, _ids, desc = build_workload("sweep1080", 1, 0, 1)
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    per_rank = total_frames // world
    first = rank * per_rank
    n_launch = (per_rank + per_launch - 1) // per_launch
    t_prep = time.perf_counter()
    batches = []
    for k in range(n_launch):
        ids = range(first + k * per_launch, min(first + (k + 1) * per_launch, first + per_rank))
        batches.append(Rasterizer.prepare_batch([cfg.rasterizer(i).on_device(local_rank) for i in ids], cfg.scene, W, H, cfg.tile_size, cfg.assets, device=local_rank))
    t_prep = time.perf_counter() - t_prep
    ctx = DeviceContext.get(local_rank)
    out = torch.empty((per_launch, H, W, 4), dtype=torch.uint8, device=dev)

    def job_render():
        for b in batches:
            b.run(out, sync=False)

    own = None
    if dl is None:   # one GPU: the delivery buffer is local memory, the path through rxc_mgpu_* is the same
        own = dl = mgpu.Delivery(ctx, rank, world)
    slots = 2
    dl.target(slots * world * per_launch * fb)
    regions = [mgpu.frame_regions(world, per_launch, fb, s_ * world * per_launch * fb) for s_ in range(slots)]

    def job_delivered():
        for k, b in enumerate(batches):
            slot = k % slots
            if k >= slots:
                dl.release()
            dl.render(b, (slot * world + rank) * per_launch * fb)
            dl.deliver(regions[slot])
        for _ in range(min(slots, len(batches))):   # leave the ring drained for the next job
            dl.release()

    def timed_job(job):
        for b in batches[:3]:
            b.run(out, sync=False)          # warm-up: scene upload, workspace, clocks
        barrier()
        a = torch.cuda.Event(enable_timing=True); z = torch.cuda.Event(enable_timing=True)
        a.record(); job(); z.record()
        barrier()
        ms = a.elapsed_time(z)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_r = min(timed_job(job_render) for _ in range(2))
    ms_d = min(timed_job(job_delivered) for _ in range(2))
    st = dl.status()
    if own is not None:
        ctx.synchronize()
        own.close()
    pix = per_rank * world * W * H
    return {"workload": "camera sweep: %d frames of the map scene at %dx%d, contiguous blocks of %d frames per rank, %d frames per launch" % (per_rank * world, W, H, per_rank, per_launch),
            "scaling": "strong", "frames": per_rank * world, "n_gpus": world,
            "render_only": {"total_job_s": ms_r * 1e-3, "frames_per_s": per_rank * world / (ms_r * 1e-3), "Mpixel_per_s": pix / (ms_r * 1e-3) / 1e6},
            "delivered_to_rank0": {"total_job_s": ms_d * 1e-3, "frames_per_s": per_rank * world / (ms_d * 1e-3), "Mpixel_per_s": pix / (ms_d * 1e-3) / 1e6,
                                   "bytes_to_rank0": (world - 1) * per_rank * fb, "rank0_ingest_GBps": (world - 1) * per_rank * fb / (ms_d * 1e-3) / 1e9 if world > 1 else None,
                                   "mode": ("peer writes over NVLink (ranks > 0), local stores (rank 0)" if st["mode"] != "nccl" else "nccl send/recv") if world > 1 else "local", "timeouts": st["timeouts"] if rank == 0 else None},
            "host_prepare_s": t_prep,
            "timing": "one pair of CUDA events around the whole job on every rank's stream (best of 2 jobs after a 3-launch warm-up), max over ranks; every launch writes %d MB of frames (> L2)" % (per_launch * fb // 1000000)}


def band_split(rank, world, local_rank, dev, flush, barrier, dl, steps=5, warmup=3, balance_iters=12):
    """BASELINE.json config D: one 7680x4320 frame of the 1M-triangle scene split by screen band over the ranks
    (SURVEY 8e).  Strong scaling of a single frame, time = max over ranks.  Every rank drops the batches that cannot
    touch its rectangle before it loads a vertex of them (k_frame_setup), so the front end is sharded too.  Four splits
    are timed render-only (the band stays in the rank's HBM): equal-height row bands, row bands whose boundaries
    `mgpu.BandBalancer` moved from the ranks' measured times, equal-width column bands and column bands balanced the same
    way.  The balanced ones and the equal columns are then timed
    DELIVERED: every rank's raster kernel writes its band in place into the full frame in rank 0's memory
    (rxc_mgpu_rasterize with the frame's row pitch), until rank 0 has seen every rank's completion flag."""
    import torch
    import torch.distributed as dist
    from rusterix_b200 import Rasterizer, mgpu

    cfg, frame_ids, desc = build_workload("dense8k", 1, 0, 1)
    W, H = cfg.width, cfg.height
    rast = cfg.rasterizer(frame_ids[0]).on_device(local_rank)
    out = torch.empty((1, H, W, 4), dtype=torch.uint8, device=dev)   # any band of this rank fits

    def prepare(band):
        y0, y1 = band[0], band[1]
        x0, x1 = (band[2], band[3]) if len(band) == 4 else (0, W)
        return Rasterizer.prepare_batch([rast], cfg.scene, W, H, cfg.tile_size, cfg.assets, band=(y0, y1, x0, x1), device=local_rank) if (y1 > y0 and x1 > x0) else None

    def timed_frame(run):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record(); flush.zero_()
        barrier()
        return a.elapsed_time(b)

    def measure(run):
        for _ in range(warmup):
            timed_frame(run)
        ms = 0.0
        for _ in range(steps):
            t = torch.tensor([timed_frame(run)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms += float(t.item())
        return ms / steps

    def local_run(batch):
        return (lambda: batch.run(out, sync=False)) if batch is not None else (lambda: None)

    bal = mgpu.BandBalancer(H, world, 32)
    equal_ms = measure(local_run(prepare(bal.band(rank))))
    for _ in range(balance_iters):
        run = local_run(prepare(bal.band(rank)))
        timed_frame(run)                       # the first frame with new bands may regrow the tile-list arenas
        bal.update(mgpu.all_gather_times(timed_frame(run), device=dev))
    row_bands = bal.use_best()
    rbatch = prepare(row_bands[rank])
    rows_ms = measure(local_run(rbatch))

    # column bands (rxc_frame.band_x0/x1): every rank gets an equal-width slice of every tile row, the cheap sky rows and
    # the expensive horizon rows alike, so the split is balanced by construction
    col_bands = [(0, H) + mgpu.column_band_for_rank(W, r, world) for r in range(world)]
    cbatch = prepare(col_bands[rank])
    cols_ms = measure(local_run(cbatch))
    # ... up to the triangles a band holds: the same balancer over the tile columns evens that out
    cbal = mgpu.BandBalancer(W, world, 32)
    for _ in range(balance_iters):
        run = local_run(prepare((0, H) + cbal.band(rank)))
        timed_frame(run)
        cbal.update(mgpu.all_gather_times(timed_frame(run), device=dev))
    bcol_bands = [(0, H) + b for b in cbal.use_best()]
    bcbatch = prepare(bcol_bands[rank])
    bcols_ms = measure(local_run(bcbatch))

    # delivered: the band written in place into rank 0's full frame
    dl.target(W * H * 4)

    def delivered_run(batch, band, regions):
        y0, y1 = band[0], band[1]
        x0 = band[2] if len(band) == 4 else 0

        def run():
            if batch is not None:
                dl.render(batch, (y0 * W + x0) * 4, pitch_bytes=W * 4)
            dl.deliver(regions)
            dl.release()
        return run

    row_regions = mgpu.band_regions([(y0, y1, 0, W) for y0, y1 in row_bands], W)
    col_regions = mgpu.band_regions(col_bands, W)
    bcol_regions = mgpu.band_regions(bcol_bands, W)
    rows_del = measure(delivered_run(rbatch, row_bands[rank], row_regions))
    cols_del = measure(delivered_run(cbatch, col_bands[rank], col_regions))
    bcols_del = measure(delivered_run(bcbatch, bcol_bands[rank], bcol_regions))
    st = dl.status()
    names = ("column bands", "cost-balanced row bands", "cost-balanced column bands")
    ms, split = min(zip((cols_ms, rows_ms, bcols_ms), names))
    ms_del, split_del = min(zip((cols_del, rows_del, bcols_del), names))
    own0 = {"column bands": (col_bands[0][3] - col_bands[0][2]) * H, "cost-balanced row bands": (row_bands[0][1] - row_bands[0][0]) * W,
            "cost-balanced column bands": (bcol_bands[0][3] - bcol_bands[0][2]) * H}[split_del]   # rank 0's own pixels do not cross NVLink
    nv_bytes = (W * H - own0) * 4
    return {"workload": desc + ", split into %d bands" % world, "scaling": "strong", "ms_per_frame": ms,
            "Mpixel_per_s": W * H / (ms * 1e-3) / 1e6, "frames_per_s": 1.0 / (ms * 1e-3),
            "split": split,
            "column_bands_ms_per_frame": cols_ms, "balanced_column_bands_ms_per_frame": bcols_ms, "balanced_row_bands_ms_per_frame": rows_ms,
            "equal_row_bands_ms_per_frame": equal_ms,
            "row_band_edges": [b[0] for b in row_bands] + [H], "column_band_edges": [b[2] for b in bcol_bands] + [W],
            "row_bands": "cost-balanced from the ranks' measured times (mgpu.BandBalancer, %d frames)" % balance_iters,
            "delivered": {"ms_per_frame": ms_del, "Mpixel_per_s": W * H / (ms_del * 1e-3) / 1e6, "frames_per_s": 1.0 / (ms_del * 1e-3),
                          "split": split_del,
                          "column_bands_ms_per_frame": cols_del, "balanced_column_bands_ms_per_frame": bcols_del, "balanced_row_bands_ms_per_frame": rows_del,
                          "bytes_to_rank0": nv_bytes, "rank0_ingest_GBps": nv_bytes / (ms_del * 1e-3) / 1e9,
                          "mode": "peer writes over NVLink (ranks > 0), local stores (rank 0)" if st["mode"] != "nccl" else "nccl send/recv",
                          "timeouts": st["timeouts"] if rank == 0 else None,
                          "what": "every rank's k_raster writes its band in place into the 7680x4320 frame in rank 0's memory over NVLink; timed on every rank from the first front-end kernel to its completion flag (rank 0: until it has seen all flags), max over ranks"}}


def frame_loop(local_rank, dev, frames=10):
    """An engine's frame loop: the 991k-triangle world stays resident, one entity (a dynamic 3D batch) moves every frame.  Scene
    hand-over per frame through rxc_set_scene (everything again) and through rxc_update_scene (what follows the unchanged batches),
    the C-ABI call alone, median; plus the 4K render of the frame."""
    import ctypes as C

    import torch
    from rusterix_b200 import Batch3D, CullMode, DeviceContext, PixelSource, marshal, scenes

    cfg = scenes.dense(3840, 2160, 40)
    ctx = DeviceContext.get(local_rank)
    out = torch.empty((cfg.height, cfg.width, 4), dtype=torch.uint8, device=dev)

    def entity(k):
        return (Batch3D.from_box(26.0 + 0.1 * k, 2.0, 10.0, 3.0, 4.5, 3.0).source(PixelSource.StaticTileIndex(0)).cull_mode(CullMode.Off)
                .with_computed_normals())

    cfg.scene.d3_dynamic = [entity(0)]
    r = cfg.rasterizer(0).on_device(local_rank)
    r.rasterize(cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
    n_static = len(marshal.submission_order(cfg.scene)[0]) - 1
    res = {}
    for mode in ("rxc_set_scene", "rxc_update_scene"):
        t_call, t_frame = [], []
        for k in range(1, frames + 1):
            cfg.scene.d3_dynamic = [entity(k)]
            m = marshal.marshal_scene(cfg.scene, 4, cfg.assets)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = ctx.lib.rxc_set_scene(ctx.handle, C.byref(m.struct)) if mode == "rxc_set_scene" else ctx.lib.rxc_update_scene(ctx.handle, C.byref(m.struct), n_static)
            t1 = time.perf_counter()
            if st != 0:
                raise RuntimeError(ctx.lib.rxc_last_error(ctx.handle).decode())
            ctx._scene_key = (cfg.scene._uid, cfg.scene._generation, 4, cfg.scene.structure_key())   # resident: the wrapper need not upload again
            ctx._geometry_keys = marshal.geometry_keys(cfg.scene)
            r.rasterize(cfg.scene, out, cfg.width, cfg.height, cfg.tile_size, cfg.assets)
            torch.cuda.synchronize()
            t_call.append(t1 - t0); t_frame.append(time.perf_counter() - t1)
        res[mode] = {"scene_hand_over_ms": statistics.median(t_call) * 1e3, "render_ms": statistics.median(t_frame) * 1e3}
    ctx._scene_key = None
    return {"workload": "991,232-triangle world resident + one moving entity (dynamic batch), 3840x2160, one frame per hand-over", **res,
            "kept_batches": n_static}


def secondary(wname, local_rank, dev, flush, steps=5, warmup=3):
    """Device-resident Mpixel/s of another BASELINE.json config (not a headline; same timing rules)."""
    import torch
    from rusterix_b200 import DeviceContext, Rasterizer

    # config E is a 4096-frame batch: 32 cameras per call (265 MB of frames) amortise the per-call front end
    F = 1 if wname == "dense8k" else 32 if wname == "sweep1080" else 8
    cfg, frame_ids, desc = build_workload(wname, F, 0, 1)
    rasts = [cfg.rasterizer(i).on_device(local_rank) for i in frame_ids]
    out = torch.empty((F, cfg.height, cfg.width, 4), dtype=torch.uint8, device=dev)
    ctx = DeviceContext.get(local_rank)

    batch = Rasterizer.prepare_batch(rasts, cfg.scene, cfg.width, cfg.height, cfg.tile_size, cfg.assets, device=local_rank)

    def step():
        batch.run(out, sync=False)

    for _ in range(warmup):
        step(); flush.zero_()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); flush.zero_()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    ms /= steps
    ctx.set_profiling(True); ctx.reset_stats()
    step(); ctx.synchronize()
    s = ctx.stats(); ctx.set_profiling(False)
    names = ctx.kernel_names()
    extra = {}
    if wname == "shaded1080":
        info = ctx.vm_jit_info()
        extra["vm"] = {"programs_translated": info["translated"], "jit_kernels": info["kernels"], "jit_launches": info["launches"],
                       "mode": "NVRTC-compiled straight-line programs" if info["launches"] else "interpreter" + (": " + info["log"][:200] if info["log"] else "")}
    return {**extra, "workload": desc, "frames_per_step": F, "ms_per_step": ms, "Mpixel_per_s": F * cfg.width * cfg.height / (ms * 1e-3) / 1e6,
            "frames_per_s": F / (ms * 1e-3), "triangles": cfg.counts()[1],
            "kernel_ms": {names[i]: s.kernel_ms[i] for i in range(len(names)) if s.launches[i]},
            "binned_refs": int(s.last_binned_refs), "large_tris": int(s.last_large_tris), "visible_tris": int(s.last_visible_tris)}


if __name__ == "__main__":
    main()
