#!/usr/bin/env python
"""bench.py — headline benchmark of mort-b200 (contract: see the task's bench section / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mort|reference]

A "step" is one full frame of the workload BASELINE.json's metric is quoted on:
  configs[1] = Cornell box (scene 6) at 600x600, 1024 spp (32x32 strata), max depth 50.
`value` = camera samples (paths) per second over the whole job, frames rendered device-resident;
`e2e`   = the same metric through the reference-facing host-buffer call (mort_render: kernels + tone map +
          device->host copy of the RGBA8 frame into pinned memory) with the time of every call included.
N > 1: one process per GPU (torchrun), the frame is SAMPLE-split across ranks (strong scaling: total work is
fixed), partial accumulation buffers are combined with one NCCL reduce per frame, rank 0 tone-maps.

`--impl reference` times the UNMODIFIED reference renderer (oracle/_ref/mort_ref = /root/reference/mort.cu
rebuilt for sm_100a behind the headless harness).  The reference has no CPU renderer (every hit/scatter is
__device__-only), so per BASELINE.json's north_star its baseline arm is its own CUDA kernel on ONE B200; it
is timed as the reference times itself (CUDA events around renderKernel, mort.cu:96-114) on a bounded sample
of the same workload (same scene / resolution / depth at 16 spp — the reference needs ~100 s for one
1024-spp frame); Msamples/s does not depend on spp once every SM has resident work.
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

WORKLOAD = {"scene": 6, "width": 600, "spp": 1024, "depth": 50}
FLOP_PER_RAY = 482.0          # algorithmic FLOP per path segment for config 2 (n ~ 20 primitives)
REF_SPP = 16                  # bounded sample for the reference arm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mort", choices=["mort", "reference"])
    ap.add_argument("--mode", default="mega", choices=["mega", "wave"])
    ap.add_argument("--scene", type=int, default=WORKLOAD["scene"])
    ap.add_argument("--width", type=int, default=WORKLOAD["width"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"])
    ap.add_argument("--depth", type=int, default=WORKLOAD["depth"])
    ap.add_argument("--aspect", type=float, default=0.0)
    ap.add_argument("--stage", type=int, default=-1)
    ap.add_argument("--bps", type=int, default=0, help="megakernel blocks per SM (selects the register-capped variant)")
    ap.add_argument("--split", default="sample", choices=["sample", "tile"], help="how the frame is sharded over GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    exe = os.path.join(ROOT, "oracle", "_ref", "mort_ref")
    cfg = {"workload": f"mort scene {a.scene} (cornell_box) {a.width}x{a.width} depth {a.depth}; reference arm at {REF_SPP} spp per frame (bounded sample)",
           "scene": a.scene, "width": a.width, "spp": REF_SPP, "depth": a.depth, "l2": "working set << L2; reference launches are seconds long"}
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mort_ref not built (needs /root/reference at build time)"}))
        return 0
    cmd = [exe, "--scene", str(a.scene), "--width", str(a.width), "--spp", str(REF_SPP), "--depth", str(a.depth),
           "--frames", str(a.steps), "--warmup", str(a.warmup)]
    if a.aspect > 0:
        cmd += ["--aspect", str(a.aspect)]
    # The unmodified reference kernel dies now and then with "invalid program counter" (2 of ~12 launches of this very command
    # on B200; its device-side new/delete, recursion and unchecked indices are SURVEY.md App. A material): a crashed attempt is
    # repeated, up to 3 times, and the number of attempts is reported.
    line, attempts, failures = None, 0, []
    while line is None and attempts < 3:
        attempts += 1
        with ClockSampler(0) as cs:
            out = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(exe), timeout=1500)
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
           "gpu_launches": 0, "attempts": attempts}
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
    import numpy as np
    from mort_b200 import formats as F
    earth = F.read_ppm(os.path.join(ROOT, "mort_b200", "assets", "earthmap.ppm"))
    osc = O.OracleScene(tmp, earth)
    w, spp = 192, 1024                                         # ~10 s of CPU work on 16 threads
    osc.override(width=w, spp=spp, depth=a.depth)
    cores = os.cpu_count() or 1
    t0 = time.time()
    _, _, st = osc.render(seed=1, threads=cores, want_rgba8=False)
    dt = time.time() - t0
    return {"value": st["samples"] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"oracle/mort_oracle.c, scene {a.scene} at {w}x{w}, {spp} spp, depth {a.depth}: {st['samples']} samples in {dt:.1f} s on {cores} threads",
            "mrays_per_s": st["segments"] / dt / 1e6}


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
    accum = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
    # N > 1: ranks exchange EXACT partial frames (4 x int64 fixed-point words per pixel): integer sums are associative, so the
    # combined frame is bit-identical to the single-GPU frame whatever the reduction order
    exact = torch.zeros(H, W, 4, dtype=torch.int64, device=dev) if world > 1 else None
    rgba = torch.zeros(H, W, 4, dtype=torch.uint8, device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    mode = MODE_MEGAKERNEL if a.mode == "mega" else MODE_WAVEFRONT
    mod, rem = D.sample_split(rank, world)
    split = dict(sample_mod=mod, sample_rem=rem) if a.split == "sample" else dict(tile_mod=world, tile_rem=rank)
    seg_total, ker_ms, launches = 0, 0.0, 0

    def step(i, timed):
        nonlocal seg_total, ker_ms, launches
        flush.zero_()                                                       # L2 flush between iterations
        if world > 1 and mode == MODE_MEGAKERNEL:
            if a.split == "tile":
                exact.zero_()                                   # ranks only write their own bands
            r.render_device(exact.data_ptr(), seed=69420, frame=i, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps, exact_accum=1, **split)
            s = r.stats
            D.combine(exact, how="reduce")
            if rank == 0:
                r.resolve_exact_device(exact.data_ptr(), accum.data_ptr())
        else:
            r.render_device(accum.data_ptr(), seed=69420, frame=i, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps, **split)
            s = r.stats
            if world > 1:
                D.combine(accum, how="reduce")
        if rank == 0:
            r.tonemap_device(accum.data_ptr(), n_spp, rgba.data_ptr())
        if timed:
            seg_total += s["last_segments"]; ker_ms += s["last_render_ms"]; launches += s["last_kernel_launches"] + (1 if rank == 0 else 0)

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

    # ---- end-to-end through the reference-facing host-buffer call (N = 1 semantics on every rank's share) ----
    host_rgba = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()
    e2e_ms = None
    if world == 1:
        r.render_into(host_rgba.data_ptr(), 0, seed=69420, frame=0, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps)      # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            r.render_into(host_rgba.data_ptr(), 0, seed=69420, frame=a.warmup + i, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps)
            _ = int(host_rgba[0, 0, 3])                                        # touch the result on the host
        e2e_ms = (time.perf_counter() - t0) * 1e3
    else:
        # N > 1: same call per rank on its sample share + the one collective + tone map + D2H on rank 0
        part = torch.empty(H, W, 4, dtype=torch.float32).pin_memory()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            if mode == MODE_MEGAKERNEL:
                if a.split == "tile":
                    exact.zero_()
                r.render_device(exact.data_ptr(), seed=69420, frame=a.warmup + i, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps, exact_accum=1, **split)
                D.combine(exact, how="reduce")
                if rank == 0:
                    r.resolve_exact_device(exact.data_ptr(), accum.data_ptr())
            else:
                r.render_device(accum.data_ptr(), seed=69420, frame=a.warmup + i, mode=mode, stage_nodes=a.stage, blocks_per_sm=a.bps, **split)
                D.combine(accum, how="reduce")
            if rank == 0:
                r.tonemap_device(accum.data_ptr(), n_spp, rgba.data_ptr())
                host_rgba.copy_(rgba, non_blocking=False)
                _ = int(host_rgba[0, 0, 3])
        dist.barrier(); torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_ms = float(tt[0])
        del part

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
        # FLOP_ray = 2*ceil(log2 n)*F_box*1.5 + 2*F_prim + F_rec + F_shade, evaluated there for the four configs
        flop_per_ray, flop_cfg = (1286.0, "config 4") if a.field > 0 else {1: (692.0, "config 1"), 8: (860.0, "config 3"), 9: (860.0, "config 3")}.get(a.scene, (FLOP_PER_RAY, "config 2"))
        achieved = per_gpu_rays_per_s * flop_per_ray / 1e12
        traffic = None
        try:    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this config, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if (a.scene, a.width, a.spp, a.depth, a.mode, a.field) == (6, 600, 1024, 50, "mega", 0) and world == 1:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        res = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"sphere field G={a.field} ({st2['n_leaves']} leaves, camera {a.fieldcam})" if a.field > 0 else
                                    f"mort scene {a.scene} ({'cornell_box' if a.scene == 6 else 'scene'})") + f" {W}x{H}, {a.spp} spp ({n_spp} effective), max depth {a.depth}",
                       "scene": None if a.field > 0 else a.scene, "width": W, "height": H, "spp": a.spp, "depth": a.depth, "mode": a.mode,
                       "parallelism": f"{a.split}-split x{world} + 1 NCCL int64 SUM reduce of the exact partial frames per frame" if world > 1 else "single GPU",
                       "l2": "192 MiB buffer written between timed iterations (L2 flush)"},
            "mrays_per_s": mrays, "segments_per_sample": seg_all / (samples_per_frame * a.steps),
            "clocks": clocks,
            "e2e": {"value": samples_per_frame * a.steps / (e2e_ms * 1e3), "unit": "Msamples/s",
                    "h2d_bytes_per_step": int(st["reserved0"]) or 512, "d2h_bytes_per_step": H * W * 4,
                    "note": "per step: kernel-parameter block up (the scene is resident, as in the reference's frame loop), RGBA8 frame down to pinned host memory"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": traffic, "traffic_unit": "bytes of DRAM per launch (ncu)", "kernel": "mega_kernel" if a.mode == "mega" else "wavefront kernels",
                         "kernel_ms_per_launch": kernel_ms_per_launch,
                         "how": f"algorithmic {flop_per_ray:.0f} FLOP per path segment (SURVEY.md §8d, {flop_cfg}) x segments per launch / CUDA-event kernel time; "
                                f"peak = {st['sm_count']} SMs x 128 lanes x 2 x median SM clock under load (the path is neither HBM- nor tensor-bound)",
                         "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
                         # the same kernel against the HBM roofline (the contract's other bound): measured DRAM bytes per launch / kernel time
                         "hbm": ({"achieved_gbs": traffic / (kernel_ms_per_launch * 1e6), "peak_gbs": peaks.get("hbm_gbs"),
                                  "frac": traffic / (kernel_ms_per_launch * 1e6) / peaks["hbm_gbs"]}
                                 if traffic and kernel_ms_per_launch and peaks.get("hbm_gbs") else None)},
            "kernel": {"regs": st2["regs_per_thread"], "threads_per_block": st2["threads_per_block"], "blocks_per_sm": st2["blocks_per_sm"],
                       "staged_nodes": st2["staged_nodes"], "bvh_nodes": st2["n_nodes"], "leaves": st2["n_leaves"], "linear_scan": st2["n_nodes"] == 1},
        }
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
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_mort(args))
