"""Multi-GPU path on real devices: N processes (one per GPU, NCCL), sample-split + the one per-frame reduce,
against the single-GPU frame.  Skipped when fewer than 2 GPUs are visible (the sample-split partition itself is
covered on one GPU by test_gpu_render.py::test_deterministic_and_sample_split and on CPU by test_dist_cpu.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MORT_ROOT"])
from mort_b200 import dist as D
from mort_b200.api import Renderer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r = Renderer(local)
r.build_scene(6).override_camera(width=64, spp=64).commit()
st = r.stats
r.set_stream(torch.cuda.current_stream().cuda_stream)
acc = torch.zeros(st["height"], st["width"], 4, dtype=torch.int64, device="cuda")
mod, rem = D.sample_split(rank, world)
r.render_device(acc.data_ptr(), seed=5, frame=2, sample_mod=mod, sample_rem=rem, exact_accum=1)
how = os.environ["MORT_HOW"]
if how == "cabi":          # the collective behind the C ABI (mort_comm_*): torch.distributed only carries the 128-byte NCCL id
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(Renderer.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    r.comm_attach(idt.cpu().numpy().tobytes(), world, rank)
    r.comm_reduce_exact(acc.data_ptr(), 0)
    torch.cuda.synchronize()
    tot = acc if rank == 0 else None
else:
    tot = D.combine(acc, how=how)
if rank == 0:
    full = torch.zeros_like(acc)
    r.render_device(full.data_ptr(), seed=5, frame=2, exact_accum=1)
    torch.cuda.synchronize()
    print("RESULT", int(torch.equal(tot, full)), 0.0, flush=True)     # exact partial frames: bit-identical to the single-GPU frame
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("how", ["reduce", "gather", "cabi"])
def test_two_gpu_sample_split_matches_single_gpu(how, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MORT_ROOT=ROOT, MORT_HOW=how)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(w)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "1" and float(line[2]) <= 1e-3
