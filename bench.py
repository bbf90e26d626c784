#!/usr/bin/env python
"""bench.py — headline benchmark of mort-b200 (contract: see the task's bench section / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mort|reference]

Workload = BASELINE.json configs[2] (config 3): the book-2 final scene (mort scene 8: 2401 quads + 1007 spheres behind a
4-wide SAH BVH, instanced sub-list, two constant media, earth image texture, Perlin noise, quad light) at 800x800, max
depth 40.  It is the configuration that exercises EVERYTHING on the hot path — tree traversal, instance transforms, media
probes, every material and texture, light sampling — where round 1's headline (Cornell, 13 primitives) never entered the
traversal kernel.  A "step" is one 1024-spp pass (32x32 strata) over the whole frame = a quarter of the 4096-spp frame the
config names (samples/s does not depend on how the 4096 samples are cut into passes; a 4096-spp step would make the
default run take minutes).
`value` = camera samples (paths) per second over the whole job, frames device-resident, CUDA events on the launching stream;
`e2e`   = the same metric through the reference-facing host-buffer call (mort_render: kernels + tone map + device->host
          copy of the RGBA8 frame into pinned memory), wall clock around every call;
`per_config` (N = 1) = short runs of BASELINE configs 1, 2 and 4 for continuity with round 1.
N > 1: one process per GPU (torchrun), the frame is SAMPLE-split across ranks (strong scaling: total work is fixed), the
exact partial frames are combined by ONE NCCL uint64 sum-reduce per frame issued through the C ABI
(mort_comm_reduce_exact; torch.distributed only hands the 128-byte NCCL id around), rank 0 resolves and tone-maps.

`--impl reference` times the UNMODIFIED reference renderer (oracle/_ref/mort_ref = /root/reference/mort.cu rebuilt for
sm_100a behind the headless harness) on the same scene / resolution / depth.  The reference has no CPU renderer (every
hit/scatter is __device__-only), so per BASELINE.json's north_star its baseline arm is its own CUDA kernel on ONE B200, timed
as the reference times itself (CUDA events around renderKernel, mort.cu:96-114) on a bounded sample: 1 spp per frame (the
reference is thread-per-pixel: 640 000 resident threads whatever the spp, so Msamples/s does not depend on spp; one
1024-spp pass would take it ~1 hour).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"scene": 8, "width": 800, "spp": 1024, "depth": 40}
SCENE_NAMES = {1: "random_spheres", 6: "cornell_box", 8: "final_scene (book 2)", 9: "final_scene (small)"}
# algorithmic FLOP per path segment = 2*ceil(log2 n)*F_box*1.5 + 2*F_prim + F_rec + F_shade, evaluated there per config
FLOP_PER_RAY = {"config 1": 692.0, "config 2": 482.0, "config 3": 860.0, "config 4": 1286.0}
REF_SPP = 1                   # bounded sample for the reference arm (see the module docstring)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mort", choices=["mort", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "mega", "wave", "pool"])
    ap.add_argument("--scene", type=int, default=WORKLOAD["scene"])
    ap.add_argument("--width", type=int, default=WORKLOAD["width"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"])
    ap.add_argument("--depth", type=int, default=WORKLOAD["depth"])
    ap.add_argument("--aspect", type=float, default=0.0)
    ap.add_argument("--stage", type=int, default=0)
    ap.add_argument("--bps", type=int, default=0, help="blocks per SM (selects the register-capped kernel variant)")
    ap.add_argument("--tpb", type=int, default=0, help="threads per block")
    ap.add_argument("--pool", type=int, default=0, help="block wavefront: paths per block pool")
    ap.add_argument("--refill", type=int, default=0, help="block wavefront: lane refill threshold")
    ap.add_argument("--split", default="sample", choices=["sample", "tile"], help="how the frame is sharded over GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--field", type=int, default=0, help="BASELINE config 4: sphere field over cells [-G,G)^2 instead of --scene (G = 500: 1 M spheres)")
    ap.add_argument("--fieldcam", type=int, default=0, choices=[0, 1], help="0 = book view, 1 = aerial")
    return ap.parse_args()


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  NVML in-process (one nvmlInit before the timed region, then a
    few microseconds per query); spawning nvidia-smi every 200 ms instead re-initialises the driver's management interface
    each time and was measured to slow the running kernel by ~2 % (profiles/r01_clock_sampler_ab.jsonl).  Falls back to
    nvidia-smi when the NVML binding is missing.  MORT_BENCH_CLOCKS=smi|nvml|off overrides (experiments)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.rows, self.stop, self.gpu = [], threading.Event(), gpu_index
        self.how = os.environ.get("MORT_BENCH_CLOCKS", "nvml")
        self.nv = self.handle = None
        if self.how == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(ids[gpu_index] if ids and gpu_index < len(ids) else gpu_index)
                self.nv = pynvml
            except Exception:
                self.how = "smi"
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        nv, h = self.nv, self.handle
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        row = [str(self.gpu), str(sm), str(mx), "", hex(mask)] + ["Active" if mask & b else "Not Active" for b in self.BITS.values()]
        self.rows.append(row)

    def _run(self):
        if self.how == "off":
            return
        while not self.stop.is_set():
            try:
                if self.how == "nvml":
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop.wait(0.05 if self.how == "nvml" else 0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": self.how}


def scene_label(a):
    return SCENE_NAMES.get(a.scene, "scene")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    exe = os.path.join(ROOT, "oracle", "_ref", "mort_ref")
    cfg = {"workload": f"mort scene {a.scene} ({scene_label(a)}) {a.width}x{a.width} depth {a.depth}; reference arm at {REF_SPP} spp per frame (bounded sample: "
                       "thread-per-pixel kernel, Msamples/s independent of spp)",
           "scene": a.scene, "width": a.width, "spp": REF_SPP, "depth": a.depth, "l2": "the reference's recursion scratch (depth x W x H x 32 B = 0.8 GB) exceeds L2; launches are seconds long"}
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mort_ref not built (needs /root/reference at build time)"}))
        return 0
    cmd = [exe, "--scene", str(a.scene), "--width", str(a.width), "--spp", str(REF_SPP), "--depth", str(a.depth),
           "--frames", str(a.steps), "--warmup", str(a.warmup)]
    if a.aspect > 0:
        cmd += ["--aspect", str(a.aspect)]
    # The unmodified reference kernel dies now and then on B200 with "invalid program counter" (profiles/r02_reference_crash.md: 2 of ~12
    # launches in round 1, 3 of 3 on one box of the driver's scaling run, none of 15 probe launches in round 2).  It calls virtual functions
    # through pointers device-side `new` returned without checking them and recurses through hitDispatch, so the two harness-side settings
    # that could matter are the device heap and the stack limit: a crashed attempt is repeated with both raised (the kernel itself is never
    # touched), up to 3 attempts, and the attempts and the settings of the one that ran are reported.
    ladder = [[], ["--heap-mb", "4096"], ["--heap-mb", "4096", "--stack", "16384"]]
    line, attempts, failures = None, 0, []
    while line is None and attempts < 3:
        attempts += 1
        with ClockSampler(0) as cs:
            out = subprocess.run(cmd + ladder[attempts - 1], capture_output=True, text=True, cwd=os.path.dirname(exe), timeout=1500)
        for l in out.stdout.splitlines():
            if '"timing":"renderKernel"' in l:
                line = json.loads(l)
        if line is None:
            failures.append(f"rc={out.returncode}: {out.stderr.strip()[-160:]}")
            print(f"bench.py: reference attempt {attempts} failed: {failures[-1]}", file=sys.stderr)
    if line is None:
        print(json.dumps({"impl": "reference", "unavailable": f"reference harness failed {attempts} times: {failures[-1]}"}))
        return 0
    total_ms = line["ms_total_timed"]
    samples = line["samples_per_frame"] * line["frames_timed"]
    value = samples / (total_ms * 1e3)
    res = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": line["frames_timed"], "warmup": line["warmup"],
           "ms_per_step": total_ms / line["frames_timed"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference", "config": cfg, "clocks": cs.summary(),
           "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": 0, "kind": "reference",
                            "sample": f"the reference's own CUDA renderKernel on 1 B200 (it has no CPU renderer): {line['frames_timed']} frames of "
                                      f"{a.width}x{a.width} at {REF_SPP} spp"},
           "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "attempts": attempts, "harness_settings": " ".join(ladder[attempts - 1]) or "default (heap 1 GiB, stack 8192 B as mort.cu:703)"}
    print(json.dumps(res))
    return 0


def cpu_baseline(a):
    """The oracle (CPU restatement) timed on this box's host cores on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as O
    from mort_b200.api import Renderer  # scene dump comes from the product's host scene builder
    tmp = "/tmp/mort_bench_scene.mscn"
    r = Renderer(int(os.environ.get("LOCAL_RANK", "0")))
    r.build_scene(a.scene)
    r.dump_scene(tmp)
    r.close()
    from mort_b200 import formats as F
    earth = F.read_ppm(os.path.join(ROOT, "mort_b200", "assets", "earthmap.ppm"))
    osc = O.OracleScene(tmp, earth)
    # ~10-20 s of CPU work: the oracle scans scene 8's 3408 primitives linearly per segment, like the reference
    w, spp = (96, 36) if a.scene in (8, 9) else (192, 1024)
    osc.override(width=w, spp=spp, depth=a.depth)
    cores = os.cpu_count() or 1
    t0 = time.time()
    _, _, st = osc.render(seed=1, threads=cores, want_rgba8=False)
    dt = time.time() - t0
    return {"value": st["samples"] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"oracle/mort_oracle.c, scene {a.scene} at {w}x{w}, {spp} spp, depth {a.depth}: {st['samples']} samples in {dt:.1f} s on {cores} threads",
            "mrays_per_s": st["segments"] / dt / 1e6}


def pick_mode(a, linear_scan):
    """auto = the scheduler measured fastest for the scene class (profiles/r02_*): the block wavefront everywhere it was measured."""
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL, MODE_WAVEFRONT
    if a.mode == "mega":
        return MODE_MEGAKERNEL, "mega"
    if a.mode == "wave":
        return MODE_WAVEFRONT, "wave"
    return MODE_POOL, "pool"


def traffic_for(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel variant bench.py launches for `key`, from this
    round's ncu --set full capture (profiles/r02_traffic.json; null when that variant was not captured)."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        e = tj.get(key)
        if e and "dram_bytes_const" in e:              # capture of a shorter launch of the same kernel: const + per_sample x samples of THIS launch
            return e, None
        return (e["dram_bytes_per_launch"], e.get("samples_per_launch")) if e else (None, None)
    except Exception:
        return None, None


def short_run(torch, Renderer, dev, stream, setup, spp_label, frames, warmup, mode, kw, flop_key, sm_count, sm_clock_hz, batch=False):
    """A few device-resident frames of another BASELINE config (per_config extra); CUDA events on the launching stream.
    batch=True: the frames are ALSO rendered as one launch (opts.n_frames: the reference's frame loop without a kernel tail per frame)."""
    r = Renderer(dev.index)
    r.set_stream(stream.cuda_stream)
    setup(r)
    r.commit()
    st = r.stats
    H, W, n_spp = st["height"], st["width"], st["sqrt_spp"] ** 2
    accum = torch.zeros(frames if batch else 1, H, W, 4, dtype=torch.float32, device=dev)

    def timed(fn, reps):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        segs, ker = 0, 0.0
        e0.record(stream)
        for i in range(reps):
            fn(warmup + i)
            s = r.stats
            segs += s["last_segments"]; ker += s["last_render_ms"]
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), segs, ker

    ms, segs, ker = timed(lambda i: r.render_device(accum.data_ptr(), seed=69420, frame=i, mode=mode, **kw), frames)
    samples = H * W * n_spp * frames
    one_per_launch = samples / (ms * 1e3)
    if batch:
        bms, bsegs, bker = timed(lambda i: r.render_device(accum.data_ptr(), seed=69420, frame=i * frames, mode=mode, n_frames=frames, **kw), 3)
        ms, segs, ker, samples = bms, bsegs, bker, 3 * samples
    st2 = r.stats
    r.close()
    achieved = segs / (ker * 1e-3) * FLOP_PER_RAY[flop_key] / 1e12
    peak = sm_count * 128 * 2 * sm_clock_hz / 1e12
    res = {"workload": f"{spp_label} {W}x{H}, {n_spp} spp effective, depth {st['bounce_limit']}" + (f", {frames} frames per launch" if batch else ""),
           "value": samples / (ms * 1e3), "unit": "Msamples/s",
           "mrays_per_s": segs / (ms * 1e3), "ms_per_frame": ms / (frames * (3 if batch else 1)), "frames": frames,
           "roofline_frac_fp32": achieved / peak, "flop_per_ray": FLOP_PER_RAY[flop_key], "value_one_launch_per_frame": one_per_launch,
           "kernel": {"regs": st2["regs_per_thread"], "threads_per_block": st2["threads_per_block"], "blocks_per_sm": st2["blocks_per_sm"],
                      "bvh_nodes": st2["n_nodes"], "leaves": st2["n_leaves"]}}
    return res


def run_mort(a):
    import torch
    import torch.distributed as dist
    from mort_b200 import dist as D
    from mort_b200.api import MODE_MEGAKERNEL, MODE_WAVEFRONT, Renderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — mort_b200 has no CPU path")
    torch.cuda.set_device(local)
    json_fd = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the communicator is created (with
        # NCCL_DEBUG set on the box, NCCL_DEBUG_FILE notwithstanding), so fd 1 points at stderr until the result is written
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    r = Renderer(local)
    if a.field > 0:
        r.build_sphere_field(a.field, 69420, a.fieldcam)
    else:
        r.build_scene(a.scene)
    r.override_camera(width=a.width, aspect=a.aspect, spp=a.spp, depth=a.depth).commit()
    st = r.stats
    H, W, n_spp = st["height"], st["width"], st["sqrt_spp"] ** 2
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    mode, mode_name = pick_mode(a, st["n_nodes"] == 1)
    kw = dict(stage_nodes=a.stage, blocks_per_sm=a.bps, threads_per_block=a.tpb, pool_paths=a.pool, pool_refill=a.refill)
    if world > 1:
        # the data-path collective lives behind the C ABI; torch.distributed only carries the 128-byte NCCL id to the ranks
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Renderer.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        r.comm_attach(bytes(idt.cpu().numpy().tobytes()), world, rank)
    accum = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
    # N > 1: ranks exchange EXACT partial frames (4 x uint64 fixed-point words per pixel): integer sums are associative, so the
    # combined frame is bit-identical to the single-GPU frame whatever the reduction order
    exact = torch.zeros(H, W, 4, dtype=torch.int64, device=dev) if world > 1 else None
    rgba = torch.zeros(H, W, 4, dtype=torch.uint8, device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    mod, rem = D.sample_split(rank, world)
    split = dict(sample_mod=mod, sample_rem=rem) if a.split == "sample" else dict(tile_mod=world, tile_rem=rank)
    seg_total, ker_ms, launches = 0, 0.0, 0

    def frame(i, host_rgba=None):
        """one step = one frame: every rank renders its share, one collective, rank 0 resolves + tone-maps"""
        if world > 1 and mode != MODE_WAVEFRONT:
            if a.split == "tile":
                exact.zero_()                                   # ranks only write their own bands
            r.render_device(exact.data_ptr(), seed=69420, frame=i, mode=mode, exact_accum=1, **kw, **split)
            s = r.stats
            r.comm_reduce_exact(exact.data_ptr(), 0)
            if rank == 0:
                r.resolve_exact_device(exact.data_ptr(), accum.data_ptr())
        else:
            r.render_device(accum.data_ptr(), seed=69420, frame=i, mode=mode, **kw, **split)
            s = r.stats
            if world > 1:
                D.combine(accum, how="reduce")
        if rank == 0:
            r.tonemap_device(accum.data_ptr(), n_spp, rgba.data_ptr())
            if host_rgba is not None:
                host_rgba.copy_(rgba, non_blocking=False)
        return s

    def step(i, timed):
        nonlocal seg_total, ker_ms, launches
        flush.zero_()                                                       # L2 flush between iterations
        s = frame(i)
        if timed:
            seg_total += s["last_segments"]; ker_ms += s["last_render_ms"]
            launches += s["last_kernel_launches"] + (1 if rank == 0 else 0) + (2 if world > 1 and rank == 0 else (1 if world > 1 else 0))

    for i in range(a.warmup):
        step(i, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        e0.record(stream)
        for i in range(a.steps):
            step(a.warmup + i, True)
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, float(seg_total), ker_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, seg_all, ker_ms_max = float(tmax[0]), float(tsum[1]), float(tmax[2])
    else:
        seg_all, ker_ms_max = float(seg_total), ker_ms
    samples_per_frame = H * W * n_spp
    value = samples_per_frame * a.steps / (ms * 1e3)                         # Msamples/s, whole job
    mrays = seg_all / (ms * 1e3)

    # ---- end-to-end through the reference-facing host-buffer call ----
    host_rgba = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()
    if world == 1:
        r.render_into(host_rgba.data_ptr(), 0, seed=69420, frame=0, mode=mode, **kw)      # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            r.render_into(host_rgba.data_ptr(), 0, seed=69420, frame=a.warmup + i, mode=mode, **kw)
            _ = int(host_rgba[0, 0, 3])                                        # touch the result on the host
        e2e_ms = (time.perf_counter() - t0) * 1e3
    else:
        # N > 1: the same frame function + the RGBA8 frame copied to pinned host memory on rank 0
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            frame(a.warmup + i, host_rgba if rank == 0 else None)
            if rank == 0:
                _ = int(host_rgba[0, 0, 3])
        dist.barrier(); torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_ms = float(tt[0])

    if rank == 0:
        st2 = r.stats
        clocks = cs.summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_clock = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        fp32_peak = st["sm_count"] * 128 * 2 * sm_clock / 1e12                    # TFLOP/s at the clock actually sustained
        kernel_ms_per_launch = ker_ms_max / a.steps
        per_gpu_rays_per_s = (seg_all / world) / a.steps / (kernel_ms_per_launch * 1e-3)
        flop_cfg = "config 4" if a.field > 0 else {1: "config 1", 6: "config 2"}.get(a.scene, "config 3")
        flop_per_ray = FLOP_PER_RAY[flop_cfg]
        achieved = per_gpu_rays_per_s * flop_per_ray / 1e12
        tkey = f"field{a.field}:{mode_name}" if a.field > 0 else f"scene{a.scene}:{mode_name}"
        traffic, traffic_samples = traffic_for(tkey) if world == 1 else (None, None)
        if isinstance(traffic, dict):
            traffic = traffic["dram_bytes_const"] + traffic["dram_bytes_per_sample"] * samples_per_frame
        elif traffic and traffic_samples:                                       # the capture ran a shorter launch of the same kernel: scale to this launch
            traffic = traffic * (samples_per_frame / traffic_samples)
        kname = {"mega": "mega_kernel", "pool": "pool_kernel", "wave": "wavefront kernels"}[mode_name]
        workload = (f"sphere field G={a.field} ({st2['n_leaves']} leaves, camera {a.fieldcam})" if a.field > 0 else
                    f"BASELINE config 3 scene: mort scene {a.scene} ({scene_label(a)})" if a.scene == 8 else f"mort scene {a.scene} ({scene_label(a)})")
        res = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload + f" {W}x{H}, max depth {a.depth}; one step = one {a.spp}-spp pass ({n_spp} effective)"
                                   + (" = 1/4 of the 4096-spp frame the config names" if (a.scene, a.spp) == (8, 1024) else ""),
                       "scene": None if a.field > 0 else a.scene, "width": W, "height": H, "spp": a.spp, "depth": a.depth, "mode": mode_name,
                       "parallelism": f"{a.split}-split x{world} + 1 NCCL uint64 SUM reduce of the exact partial frames per frame (C ABI: mort_comm_reduce_exact)" if world > 1 else "single GPU",
                       "l2": "192 MiB buffer written between timed iterations (L2 flush)"},
            "mrays_per_s": mrays, "segments_per_sample": seg_all / (samples_per_frame * a.steps),
            "clocks": clocks,
            "e2e": {"value": samples_per_frame * a.steps / (e2e_ms * 1e3), "unit": "Msamples/s",
                    "h2d_bytes_per_step": int(st["reserved0"]) or 512, "d2h_bytes_per_step": H * W * 4,
                    "note": "per step: kernel-parameter block up (the scene is resident, as in the reference's frame loop), RGBA8 frame down to pinned host memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": traffic, "traffic_unit": "bytes of DRAM per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel variant, this round; extrapolated from a shorter launch: profiles/r02_traffic.json)", "kernel": kname,
                         "kernel_ms_per_launch": kernel_ms_per_launch,
                         "how": f"algorithmic {flop_per_ray:.0f} FLOP per path segment (SURVEY.md §8d, {flop_cfg}) x segments per launch / CUDA-event kernel time; "
                                f"peak = {st['sm_count']} SMs x 128 lanes x 2 x median SM clock under load (the path is neither HBM- nor tensor-bound)",
                         "algorithmic_bytes_per_launch": H * W * 16,
                         "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
                         # the same kernel against the HBM roofline (the contract's other bound): measured DRAM bytes per launch / kernel time
                         "hbm": ({"achieved_gbs": traffic / (kernel_ms_per_launch * 1e6), "peak_gbs": peaks.get("hbm_gbs"),
                                  "frac": traffic / (kernel_ms_per_launch * 1e6) / peaks["hbm_gbs"]}
                                 if traffic and kernel_ms_per_launch and peaks.get("hbm_gbs") else None)},
            "kernel": {"regs": st2["regs_per_thread"], "threads_per_block": st2["threads_per_block"], "blocks_per_sm": st2["blocks_per_sm"],
                       "staged_nodes": st2["staged_nodes"], "bvh_nodes": st2["n_nodes"], "leaves": st2["n_leaves"], "linear_scan": st2["n_nodes"] == 1},
        }
        r.close()
        if world == 1 and not a.no_per_config and a.field == 0 and (a.scene, a.width, a.depth) == (WORKLOAD["scene"], WORKLOAD["width"], WORKLOAD["depth"]):
            # continuity with round 1 (whose headline was config 2): short device-resident runs of the other single-GPU configs
            pc = {}
            try:
                args = (torch, Renderer, dev, stream)
                pc["config 1"] = short_run(*args, lambda q: q.build_scene(1).override_camera(width=400, aspect=16 / 9, spp=32, depth=50), "mort scene 1 (random_spheres)", 20, 3, mode, kw, "config 1", st["sm_count"], sm_clock, batch=(mode_name == "pool"))
                pc["config 2"] = short_run(*args, lambda q: q.build_scene(6).override_camera(width=600, spp=1024, depth=50), "mort scene 6 (cornell_box)", 3, 1, mode, kw, "config 2", st["sm_count"], sm_clock)
                pc["config 4"] = short_run(*args, lambda q: q.build_sphere_field(500, 69420, 0).override_camera(width=1920, aspect=16 / 9, spp=256, depth=50), "1 M-sphere field, book view", 2, 1, mode, kw, "config 4", st["sm_count"], sm_clock)
            except Exception as ex:  # an extra, never a gate
                pc["error"] = str(ex)
            res["per_config"] = pc
        if world == 1 and not a.no_cpu_baseline and a.field == 0:
            try:
                res["cpu_baseline"] = cpu_baseline(a)
            except Exception as ex:  # the baseline is a report, never a gate
                res["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        if json_fd is not None:
            sys.stdout.flush()
            os.write(json_fd, (json.dumps(res) + "\n").encode())
        else:
            print(json.dumps(res))
    else:
        r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_mort(args))
