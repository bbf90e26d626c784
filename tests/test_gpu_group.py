"""Multi-GPU rendering through the C ABI (mort_group_*, mort_comm_*; mort_b200/csrc/group.cu): the N-GPU frame must equal
the single-GPU frame bit for bit, for the sample split and the tile split, for both exact schedulers.  The 1-rank cases run
on any GPU box (they still go through NCCL: ncclCommInitRank / ncclReduce with one rank); the 2-rank cases need 2 GPUs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _single(scene, width, spp, **opts):
    from mort_b200.api import Renderer
    with Renderer(0) as r:
        r.build_scene(scene).override_camera(width=width, spp=spp).commit()
        return r.render(seed=11, frame=1, **opts)


@pytest.mark.parametrize("n", [1, 2])
@pytest.mark.parametrize("mode", ["mega", "pool"])
def test_group_frame_equals_single_gpu_frame(n, mode):
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL, SPLIT_SAMPLE, SPLIT_TILE, Group
    if _n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    m = MODE_POOL if mode == "pool" else MODE_MEGAKERNEL
    ref = _single(8, 48, 36, mode=MODE_MEGAKERNEL)
    with Group(n) as g:
        g.for_each(lambda r: r.build_scene(8).override_camera(width=48, spp=36).commit())
        for split in (SPLIT_SAMPLE, SPLIT_TILE):
            fr = g.render(split=split, seed=11, frame=1, mode=m)
            assert np.array_equal(fr.accum, ref.accum, equal_nan=True), f"{n} GPUs, split {split}, {mode}"
            assert np.array_equal(fr.rgba8, ref.rgba8)
            st = fr.stats
            assert st["n_gpus"] == n and st["samples"] == ref.stats["last_samples"] and st["segments"] == ref.stats["last_segments"]
            assert (st["collective_bytes"] > 0) == (n > 1)


def test_group_refuses_mismatched_scenes():
    from mort_b200.api import Group, MortError
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    with Group(2) as g:
        g.ranks[0].build_scene(6).override_camera(width=32, spp=4).commit()
        g.ranks[1].build_scene(5).override_camera(width=32, spp=4).commit()
        with pytest.raises(MortError):
            g.render()


def test_comm_reduce_with_one_rank_is_the_identity():
    """mort_comm_*: the per-process form of the collective (what bench.py uses under torchrun), here with world = 1."""
    import torch
    from mort_b200.api import Renderer
    with Renderer(0) as r:
        r.build_scene(6).override_camera(width=32, spp=16).commit()
        st = r.stats
        buf = torch.zeros(st["height"], st["width"], 4, dtype=torch.int64, device="cuda")
        r.render_device(buf.data_ptr(), seed=2, exact_accum=1)
        before = buf.clone()
        r.comm_attach(Renderer.comm_unique_id(), 1, 0)
        r.comm_reduce_exact(buf.data_ptr(), 0)
        torch.cuda.synchronize()
        assert torch.equal(buf, before)
        r.comm_detach()
