"""GPU: progressive accumulation and checkpoint / resume through the C ABI.

The reference re-renders the same frame forever and keeps nothing (mort.cu:99-119).  Here frames with different
Philox `frame` keys are independent sample sets whose exact (integer) sums add in any order, so
  * accumulating in the kernel == adding separately rendered frames afterwards, bit for bit;
  * render k frames, save, resume in a NEW context, render the rest == render them all at once, bit for bit;
  * a checkpoint is refused for another scene, camera, frame size or seed.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, luminance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _exact(torch, r, **opts):
    st = r.stats
    t = torch.zeros((st["height"], st["width"], 4), dtype=torch.int64, device="cuda:0")
    r.render_device(t.data_ptr(), exact_accum=1, **opts)
    torch.cuda.synchronize()
    return t


@pytest.mark.parametrize("sc,w,spp", [(6, 64, 16), (1, 96, 16), (9, 48, 16)])
def test_in_kernel_accumulation_equals_sum_of_frames(renderer, sc, w, spp):
    import torch
    renderer.build_scene(sc).override_camera(width=w, spp=spp).commit()
    frames = [_exact(torch, renderer, seed=7, frame=f) for f in range(3)]
    assert not torch.equal(frames[0], frames[1])                      # different sample sets
    # (a) accumulate flag of the render kernel
    run = torch.zeros_like(frames[0])
    for f in range(3):
        renderer.render_device(run.data_ptr(), exact_accum=1, accumulate=1 if f else 0, seed=7, frame=f)
    torch.cuda.synchronize()
    want = frames[0] + frames[1] + frames[2]
    assert torch.equal(run, want)
    # (b) the merge kernel, in the opposite order
    s = frames[2].clone()
    renderer.accumulate_exact_device(s.data_ptr(), frames[1].data_ptr())
    renderer.accumulate_exact_device(s.data_ptr(), frames[0].data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(s, want)


def test_accumulate_needs_exact_sums(renderer):
    import torch
    from mort_b200.api import MortError
    renderer.build_scene(6).override_camera(width=32, spp=4).commit()
    t = torch.zeros((32, 32, 4), dtype=torch.float32, device="cuda:0")
    with pytest.raises(MortError):
        renderer.render_device(t.data_ptr(), accumulate=1)


def test_resume_in_a_new_context_is_bit_identical(tmp_path):
    from mort_b200.api import Renderer
    ck = str(tmp_path / "cornell.mckp")

    def fresh():
        return Renderer(0).build_scene(6).override_camera(width=64, spp=16).commit()

    with fresh() as r:
        whole, n = r.render_progressive(4, seed=99)
        assert n == 4
        fp = r.scene_fingerprint
    with fresh() as r:
        part, n = r.render_progressive(2, checkpoint=ck, seed=99)
        assert n == 2 and os.path.getsize(ck) == 64 + 64 * 64 * 4 * 8
        assert not np.array_equal(part.accum, whole.accum)
    with fresh() as r:
        assert r.scene_fingerprint == fp
        rest, n = r.render_progressive(2, checkpoint=ck, resume=True, seed=99)
        assert n == 4
        assert np.array_equal(rest.accum.view(np.uint32), whole.accum.view(np.uint32))
        assert np.array_equal(rest.rgba8, whole.rgba8)
        # nothing to render: the stored image comes back unchanged
        again, n = r.render_progressive(0, checkpoint=ck, resume=True, seed=99)
        assert n == 4 and np.array_equal(again.rgba8, whole.rgba8)
        # same context keeps adding; reset starts over
        more, n = r.render_progressive(1, seed=99)
        assert n == 5
        r.reset_progressive()
        one, n = r.render_progressive(1, seed=99)
        assert n == 1


def test_progressive_mean_is_the_mean_of_the_frames(renderer):
    renderer.build_scene(6).override_camera(width=64, spp=16).commit()
    renderer.reset_progressive()
    singles = [renderer.render(seed=5, frame=f).accum[..., :3].astype(np.float64) for f in range(3)]
    fr, n = renderer.render_progressive(3, seed=5)
    assert n == 3
    want = sum(singles)
    ok = np.isfinite(want).all(-1) & np.isfinite(fr.accum[..., :3]).all(-1)
    assert ok.mean() > 0.99
    assert np.allclose(fr.accum[..., :3][ok], want[ok], rtol=1e-6, atol=1e-6)
    # 8-bit frame = tone pipeline over the total sample count
    mean = fr.accum[..., :3] / (16 * 3)
    lin = np.sqrt(np.clip(np.nan_to_num(mean, nan=0.0), 0, None))
    q = (256 * np.clip(lin, 0.0, 0.999)).astype(np.uint8)
    assert (np.abs(q.astype(int) - fr.rgba8[..., :3].astype(int)) > 1).mean() < 0.01


def test_checkpoint_is_refused_for_another_scene_camera_or_seed(tmp_path):
    from mort_b200.api import MortError, Renderer
    ck = str(tmp_path / "a.mckp")
    with Renderer(0) as r:
        r.build_scene(6).override_camera(width=48, spp=4).commit()
        r.render_progressive(1, checkpoint=ck, seed=1)
        with pytest.raises(MortError):                                   # other seed
            r.render_progressive(1, checkpoint=ck, resume=True, seed=2)
        r.override_camera(width=48, spp=9)                               # other samples per frame
        with pytest.raises(MortError):
            r.render_progressive(1, checkpoint=ck, resume=True, seed=1)
        r.override_camera(width=40, spp=4)                               # other frame size
        with pytest.raises(MortError):
            r.render_progressive(1, checkpoint=ck, resume=True, seed=1)
        r.build_scene(7).override_camera(width=48, spp=4).commit()       # other scene, same frame size
        with pytest.raises(MortError):
            r.render_progressive(1, checkpoint=ck, resume=True, seed=1)
        open(ck, "wb").write(b"nope")
        r.build_scene(6).override_camera(width=48, spp=4).commit()
        with pytest.raises(MortError):
            r.render_progressive(1, checkpoint=ck, resume=True, seed=1)


def test_cli_accumulate_checkpoint_resume(tmp_path):
    exe = os.path.join(ROOT, "mort_b200", "mort")
    base = [exe, "6", "--width", "64", "--spp", "16", "--accumulate"]
    a, b, ck = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "c.mckp")
    out = subprocess.run(base + ["--frames", "4", "--out", a], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    assert '"frames_accumulated":4' in out and out.count("Avg. time per frame") == 4
    subprocess.run(base + ["--frames", "1", "--checkpoint", ck], check=True, capture_output=True, cwd=ROOT)
    out = subprocess.run(base + ["--frames", "3", "--checkpoint", ck, "--resume", "--out", b], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    assert '"frames_accumulated":4' in out
    assert open(a, "rb").read() == open(b, "rb").read()


def test_cli_dump_then_load_renders_the_same_frame(tmp_path):
    exe = os.path.join(ROOT, "mort_b200", "mort")
    sc, a, b = str(tmp_path / "s.mscn"), str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
    subprocess.run([exe, "7", "--width", "64", "--spp", "16", "--dump", sc, "--out", a], check=True, capture_output=True, cwd=ROOT)
    subprocess.run([exe, "0", "--load", sc, "--width", "64", "--spp", "16", "--out", b], check=True, capture_output=True, cwd=ROOT)
    assert open(a, "rb").read() == open(b, "rb").read()
