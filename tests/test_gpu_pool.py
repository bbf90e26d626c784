"""GPU parity of the block wavefront (MORT_MODE_POOL, mort_b200/csrc/pool.cu) through the C ABI.

The pool kernel runs the same per-ray code, Philox stream and estimator as the megakernel and both accumulate finished
samples as integers, so their EXACT frames must agree bit for bit — on every shipped scene, for any pool size / block
shape, for sample and tile splits, and for progressive accumulation.  (The megakernel itself is pinned to the oracle and
to the reference's fixtures in test_gpu_render.py / test_gpu_trace.py.)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL = {1: (96, 64), 2: (96, 64), 3: (96, 64), 4: (96, 36), 5: (64, 64), 6: (64, 64), 7: (64, 64), 8: (40, 16), 9: (48, 36), 10: (96, 16)}


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _exact(renderer, H, W, **kw):
    import torch
    buf = torch.zeros(H, W, 4, dtype=torch.int64, device="cuda")
    renderer.render_device(buf.data_ptr(), exact_accum=1, **kw)
    torch.cuda.synchronize()
    return buf, renderer.stats


@pytest.mark.parametrize("sc", list(range(1, 11)))
def test_pool_frame_is_bit_identical_to_the_megakernel(renderer, sc):
    import torch
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL
    w, spp = SMALL[sc]
    renderer.build_scene(sc).override_camera(width=w, spp=spp).commit()
    st = renderer.stats
    H, W = st["height"], st["width"]
    mega, sm = _exact(renderer, H, W, seed=99, frame=2, mode=MODE_MEGAKERNEL)
    pool, sp = _exact(renderer, H, W, seed=99, frame=2, mode=MODE_POOL)
    other, _ = _exact(renderer, H, W, seed=99, frame=2, mode=MODE_POOL, pool_refill=-1 if st["n_nodes"] > 1 else 8)
    assert torch.equal(other, pool), "fixed 32-ray chunks vs lane refill + classify pass"
    assert sp["last_samples"] == sm["last_samples"] == H * W * st["sqrt_spp"] ** 2
    assert sp["last_segments"] == sm["last_segments"]
    assert torch.equal(pool, mega), f"scene {sc}: {(pool != mega).any(-1).float().mean().item():.4f} of pixels differ"
    # the float4 request goes through a context-owned exact frame + one resolve pass
    facc = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    renderer.render_device(facc.data_ptr(), seed=99, frame=2, mode=MODE_POOL)
    ref = torch.zeros_like(facc)
    renderer.resolve_exact_device(mega.data_ptr(), ref.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(facc, nan=-1.0), torch.nan_to_num(ref, nan=-1.0))


@pytest.mark.parametrize("refill", [-1, 6, 20])
@pytest.mark.parametrize("shape", [(256, 2, 256), (256, 3, 544), (384, 2, 1024), (512, 1, 2048), (512, 2, 96), (640, 1, 2048), (768, 1, 1536), (1024, 1, 2304)])
def test_pool_shape_does_not_change_the_frame(renderer, shape, refill):
    """any block shape, any pool size, fixed 32-ray trace chunks (-1) or lane refill at any threshold: the same exact frame"""
    import torch
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL
    threads, bps, paths = shape
    renderer.build_scene(8).override_camera(width=40, spp=16).commit()       # tree, instances, media, noise + image textures
    st = renderer.stats
    H, W = st["height"], st["width"]
    mega, _ = _exact(renderer, H, W, seed=5, mode=MODE_MEGAKERNEL)
    pool, sp = _exact(renderer, H, W, seed=5, mode=MODE_POOL, threads_per_block=threads, blocks_per_sm=bps, pool_paths=paths, pool_refill=refill)
    assert torch.equal(pool, mega)
    assert sp["threads_per_block"] == threads


def test_pool_splits_and_progressive_accumulation(renderer):
    import torch
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL, MortError
    renderer.build_scene(1).override_camera(width=100, spp=16, depth=20).commit()      # 100 x 56: the last 8-row band is partial
    st = renderer.stats
    H, W = st["height"], st["width"]
    full, _ = _exact(renderer, H, W, seed=8, mode=MODE_MEGAKERNEL)
    for world in (2, 3):
        tot = torch.zeros_like(full)
        for r in range(world):
            part, _ = _exact(renderer, H, W, seed=8, mode=MODE_POOL, sample_mod=world, sample_rem=r)
            tot += part
        assert torch.equal(tot, full), f"{world}-way sample split"
        tiles = torch.zeros_like(full)
        for r in range(world):                                   # ranks write only their own 8-row bands of a shared frame
            renderer.render_device(tiles.data_ptr(), seed=8, mode=MODE_POOL, exact_accum=1, tile_mod=world, tile_rem=r)
        torch.cuda.synchronize()
        assert torch.equal(tiles, full), f"{world}-way tile split"
    # progressive: frame 0, then frame 1 added in place == the two frames summed
    f0, _ = _exact(renderer, H, W, seed=8, frame=0, mode=MODE_MEGAKERNEL)
    f1, _ = _exact(renderer, H, W, seed=8, frame=1, mode=MODE_MEGAKERNEL)
    run = torch.full((H, W, 4), 7, dtype=torch.int64, device="cuda")            # stale contents must be overwritten by frame 0
    renderer.render_device(run.data_ptr(), seed=8, frame=0, mode=MODE_POOL, exact_accum=1)
    renderer.render_device(run.data_ptr(), seed=8, frame=1, mode=MODE_POOL, exact_accum=1, accumulate=1)
    torch.cuda.synchronize()
    assert torch.equal(run, f0 + f1)
    with pytest.raises(MortError):
        renderer.render_device(run.data_ptr(), mode=MODE_POOL, pool_paths=16)


def test_pool_through_the_host_buffer_call(renderer):
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL
    renderer.build_scene(6).override_camera(width=64, spp=64).commit()
    a = renderer.render(seed=3, mode=MODE_MEGAKERNEL)
    b = renderer.render(seed=3, mode=MODE_POOL)
    assert np.array_equal(a.accum, b.accum, equal_nan=True) and np.array_equal(a.rgba8, b.rgba8)
    assert b.stats["last_kernel_launches"] >= 3


def test_a_batch_of_frames_in_one_launch_equals_the_frames_one_by_one(renderer):
    """opts.n_frames: the reference's frame loop (mort.cu:93-120) as ONE launch; frame k of the batch = the frame with key frame + k"""
    import torch
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL, MortError
    renderer.build_scene(1).override_camera(width=80, spp=9, depth=20).commit()
    st = renderer.stats
    H, W = st["height"], st["width"]
    batch = torch.zeros(3, H, W, 4, dtype=torch.int64, device="cuda")
    renderer.render_device(batch.data_ptr(), seed=4, frame=5, mode=MODE_POOL, exact_accum=1, n_frames=3)
    torch.cuda.synchronize()
    assert renderer.stats["last_samples"] == 3 * H * W * 9
    for k in range(3):
        one, _ = _exact(renderer, H, W, seed=4, frame=5 + k, mode=MODE_MEGAKERNEL)
        assert torch.equal(batch[k], one), f"frame {k} of the batch"
    fl = torch.zeros(3, H, W, 4, dtype=torch.float32, device="cuda")
    renderer.render_device(fl.data_ptr(), seed=4, frame=5, mode=MODE_POOL, n_frames=3)
    ref = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    renderer.resolve_exact_device(batch[2].contiguous().data_ptr(), ref.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(fl[2], nan=-1.0), torch.nan_to_num(ref, nan=-1.0))
    with pytest.raises(MortError):
        renderer.render_device(batch.data_ptr(), mode=MODE_MEGAKERNEL, exact_accum=1, n_frames=3)
