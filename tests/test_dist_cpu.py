"""The N > 1 path on the CPU: two gloo ranks run the sample-split partition + the one per-frame exchange
(mort_b200.dist) on partial frames, and the result equals the single-rank frame.  The partial frames come
from the oracle here (no GPU); on the GPU box the same code runs over NCCL with the CUDA renderer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, golden_scene_path


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, how, scene_path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle_binding as O
    from mort_b200 import dist as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    osc = O.OracleScene(scene_path)
    osc.override(width=32, spp=16)
    mod, rem = D.sample_split(rank, world)
    part, _, st = osc.render(seed=7, sj_mod=mod, sj_rem=rem, threads=2, want_rgba8=False)
    total = D.combine(torch.from_numpy(part), how=how)
    if rank == 0:
        q.put((total.numpy(), st["samples"]))
    else:
        assert total is None
        q.put((None, st["samples"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("how", ["reduce", "gather"])
def test_two_rank_sample_split_equals_single_rank(how, tmp_path):
    import oracle_binding as O
    path = golden_scene_path(6, str(tmp_path))
    osc = O.OracleScene(path)
    osc.override(width=32, spp=16)
    full, _, st = osc.render(seed=7, threads=2, want_rgba8=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, how, path, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    totals = [g[0] for g in got if g[0] is not None]
    assert len(totals) == 1 and sum(g[1] for g in got) == st["samples"]
    tot = totals[0]
    assert np.array_equal(tot[..., 3], full[..., 3])
    assert np.allclose(tot[..., :3], full[..., :3], rtol=1e-5, atol=1e-5, equal_nan=True)


def test_partition_helpers():
    from mort_b200 import dist as D
    for world in (1, 2, 3, 4, 8):
        for sq in (1, 5, 8, 32, 64):
            rows = [D.rows_of_rank(sq, r, world) for r in range(world)]
            assert sum(rows) == sq and max(rows) - min(rows) <= 1
    with pytest.raises(ValueError):
        D.sample_split(2, 2)
    t = torch.ones(2, 2, 4)
    assert D.combine(t) is t                     # no process group: identity
    assert torch.equal(D.combine_virtual([t, t, t]), 3 * t)
