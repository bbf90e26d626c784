#!/usr/bin/env python
"""BASELINE config 5's load-balance matrix on ONE GPU: every rank's share of an N-GPU frame (sample split and tile split) is rendered
one after the other on the same device and timed with CUDA events.  What an N-GPU run adds to max-over-ranks kernel time is one
ncclReduce of the exact partial frames (measured on real GPUs: profiles/r02_m2_*.json, r01_scale8_*.json), so
    predicted efficiency(N) = t(1) / (N * max_r t_r(N))
isolates the part of the scaling loss that is load imbalance + the per-launch tail — the part VERDICT r01 weak #7 asks about —
without 8 GPUs.  Also checks that the ranks' exact partial frames add up to the 1-GPU frame bit for bit.
  python scripts/r02/virtual_ranks.py [--width 3840] [--spp 64] [--scenes 1,2,...] [--out file.jsonl]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--aspect", type=float, default=16 / 9)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--scenes", default="8,1,2,3,4,5,6,7,9,10")
    ap.add_argument("--field", type=int, default=0, help="G of the sphere field instead of a shipped scene (config 4)")
    ap.add_argument("--ns", default="2,4,8")
    ap.add_argument("--tile-rows", type=int, default=2)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    from mort_b200.api import MODE_POOL, Renderer
    out = open(a.out, "a") if a.out else None
    with Renderer(0) as r:
        scenes = [0] if a.field else [int(s) for s in a.scenes.split(",")]
        for sc in scenes:
            if a.field:
                r.build_sphere_field(a.field, 69420, 0)
            else:
                r.build_scene(sc)
            r.override_camera(width=a.width, aspect=a.aspect, spp=a.spp).commit()
            st = r.stats
            H, W, q = st["height"], st["width"], st["sqrt_spp"]
            full = torch.zeros((H, W, 4), dtype=torch.int64, device="cuda")
            part = torch.zeros_like(full)
            torch.cuda.synchronize()

            def run(buf, **kw):
                r.render_device(buf.data_ptr(), mode=MODE_POOL, exact_accum=1, seed=69420, frame=0, **kw)
                return r.stats["last_render_ms"]
            run(full)                                                       # warm-up
            t1 = min(run(full) for _ in range(2))
            samples = W * H * q * q
            row = {"scene": sc if not a.field else f"field{a.field}", "width": W, "height": H, "spp_eff": q * q, "t1_ms": t1, "msamples_per_s_n1": samples / t1 / 1e3}
            for split in ("sample", "tile"):
                for n in [int(x) for x in a.ns.split(",")]:
                    if split == "sample" and n > q:
                        continue
                    part.zero_()
                    torch.cuda.synchronize()
                    ts = []
                    for rank in range(n):
                        kw = dict(sample_mod=n, sample_rem=rank, accumulate=1) if split == "sample" else dict(tile_mod=n, tile_rem=rank, tile_rows=a.tile_rows, accumulate=1)
                        ts.append(run(part, **kw))
                    same = bool(torch.equal(part, full))
                    row[f"{split}_n{n}"] = {"max_ms": max(ts), "min_ms": min(ts), "sum_ms": sum(ts), "efficiency": t1 / (n * max(ts)), "bit_identical_to_n1": same}
            line = json.dumps(row)
            print(line, flush=True)
            if out:
                out.write(line + "\n"); out.flush()


if __name__ == "__main__":
    main()
