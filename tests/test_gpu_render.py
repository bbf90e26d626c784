"""GPU parity, part 2: rendered frames through the C ABI.
  (a) product vs oracle with the SAME Philox stream: agreement far below Monte-Carlo noise, identical NaN
      pixel sets (the estimator, every material / texture / pdf, media, light sampling, tie rules);
  (b) product vs the reference's own converged frames (tests/golden/conv_*.npz, 4096 spp): per-channel RMSE
      within 2x the reference's seed-to-seed noise floor and mean luminance within 0.5 % (BASELINE.json);
  (c) determinism, sample-split equivalence (N virtual ranks on one GPU), tone pipeline, staging on/off.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bin8, golden_scene_path, luminance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


SMALL = {1: (96, 64), 2: (96, 64), 3: (96, 64), 4: (96, 36), 5: (64, 64), 6: (64, 64), 7: (64, 64), 8: (40, 16), 9: (48, 36), 10: (96, 16)}


@pytest.mark.parametrize("sc", list(range(1, 11)))
def test_frame_matches_oracle_same_stream(renderer, earth, sc, tmp_path):
    import oracle_binding as O
    w, spp = SMALL[sc]
    renderer.build_scene(sc).override_camera(width=w, spp=spp).commit()
    fr = renderer.render(seed=4242, frame=3)
    osc = O.OracleScene(golden_scene_path(sc, str(tmp_path)), earth)
    osc.override(width=w, spp=spp)
    hdr, rgba, st = osc.render(seed=4242, frame=3)
    acc = fr.accum
    assert acc.shape == hdr.shape
    n = fr.stats["sqrt_spp"] ** 2
    nan_p, nan_o = acc[..., 3], hdr[..., 3]
    # NaN samples are a deterministic function of the path; rare diverged paths may move single counts
    assert (nan_p != nan_o).mean() <= 0.003, f"scene {sc}: NaN-sample counts differ on {(nan_p != nan_o).mean():.4f} of pixels"
    ok = (nan_p == 0) & (nan_o == 0) & np.isfinite(acc[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
    a, b = acc[..., :3][ok] / n, hdr[..., :3][ok] / n
    rel = np.abs(a - b).max(-1) / (np.abs(b).max(-1) + 1e-2)
    # the two implementations share the RNG stream; they differ by libm ulps and FMA contraction in shading,
    # which only matters for the few paths that sit on a decision boundary
    assert (rel > 1e-3).mean() <= 0.05, f"scene {sc}: {(rel > 1e-3).mean():.4f} of pixels differ by more than 1e-3"
    assert abs(a.mean() - b.mean()) <= 2e-3 * max(b.mean(), 1e-3), f"scene {sc}: mean {a.mean()} vs oracle {b.mean()}"
    seg_p, seg_o = fr.stats["last_segments"], st["segments"]
    assert abs(seg_p - seg_o) <= 0.002 * seg_o + 16, f"scene {sc}: segments {seg_p} vs oracle {seg_o}"
    assert fr.stats["last_samples"] == st["samples"]
    # 8-bit frame: same tone pipeline, same sums up to summation order -> at most 1 code value apart on nearly all pixels
    d8 = np.abs(fr.rgba8[..., :3].astype(int) - rgba[..., :3].astype(int))
    assert (d8 > 1).mean() <= 0.05 and (fr.rgba8[..., 3] == 255).all()


CONV = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]       # scenes 8/9 at reduced size and spp (the reference runs them at 0.03-0.7 Msamples/s)


@pytest.mark.parametrize("sc", CONV)
def test_converged_frame_matches_reference(renderer, sc):
    g = np.load(f"{GOLDEN}/conv_{sc}.npz")
    ref = g["mean_a"].astype(np.float32)
    H, W = ref.shape[:2]
    spp = int(g["spp"])
    renderer.build_scene(sc).override_camera(width=W, spp=spp).commit()
    st = renderer.stats
    assert (st["height"], st["width"]) == (H, W)
    fr = renderer.render(seed=777)
    n = st["sqrt_spp"] ** 2
    mine = fr.accum[..., :3] / n
    ok = (fr.accum[..., 3] == 0) & (g["nan_a"] == 0) & np.isfinite(mine).all(-1) & np.isfinite(ref).all(-1)
    # NaN-flushed pixels (App. A-Q6): same population size as the reference's own two seeds show
    nan_ref = float((g["nan_a"] > 0).mean()); nan_ref_b = float((g["nan_b"] > 0).mean()); nan_mine = float((fr.accum[..., 3] > 0).mean())
    assert abs(nan_mine - nan_ref) <= 3 * abs(nan_ref - nan_ref_b) + 0.01, f"scene {sc}: NaN pixel fraction {nan_mine} vs reference {nan_ref}/{nan_ref_b}"
    if ok.sum() < 100:
        return                                          # scene 7: almost everything is NaN-flushed in the reference too
    rmse = np.sqrt(((mine[ok] - ref[ok]) ** 2).mean(axis=0))
    floor = g["rmse_ab"]
    # tolerance = 2 x the reference's own seed-to-seed RMSE (+ half-precision storage of the fixture)
    assert (rmse <= 2.0 * floor + 2e-3 * np.abs(ref[ok]).mean(axis=0) + 1e-5).all(), f"scene {sc}: RMSE {rmse} vs noise floor {floor}"
    la, lb = float(luminance(mine[ok]).mean()), float(luminance(ref[ok]).mean())
    assert abs(la - lb) <= 0.005 * lb + 1e-6, f"scene {sc}: mean luminance {la} vs reference {lb}"
    if "rmse_ab_bin8" in g.files:
        # the same check on 8x8-binned frames (64x the samples per value): a bias hidden below per-pixel noise shows up here
        bm, nm = bin8(mine, ok); br, _ = bin8(ref, ok)
        fullb = nm >= 48
        rb = np.sqrt(((bm[fullb] - br[fullb]) ** 2).mean(axis=0))
        assert (rb <= 2.0 * g["rmse_ab_bin8"] + 2e-3 * np.abs(br[fullb]).mean(axis=0) + 1e-5).all(), f"scene {sc}: binned RMSE {rb} vs floor {g['rmse_ab_bin8']}"


def test_deterministic_and_sample_split(renderer):
    import torch
    from mort_b200 import dist as D
    renderer.build_scene(6).override_camera(width=64, spp=64).commit()
    a = renderer.render(seed=1, frame=0).accum
    b = renderer.render(seed=1, frame=0).accum
    assert np.array_equal(a, b, equal_nan=True), "same (seed, frame) must give the same bits"
    c = renderer.render(seed=1, frame=1).accum
    assert not np.array_equal(a, c, equal_nan=True)
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            mod, rem = D.sample_split(r, world)
            parts.append(torch.from_numpy(renderer.render(seed=1, frame=0, sample_mod=mod, sample_rem=rem, want_rgba8=False).accum.copy()))
        tot = D.combine_virtual(parts).numpy()
        assert np.array_equal(tot[..., 3], a[..., 3]), "NaN-sample counts must add up exactly"
        ok = np.isfinite(a[..., :3]).all(-1) & np.isfinite(tot[..., :3]).all(-1)
        assert np.allclose(tot[..., :3][ok], a[..., :3][ok], rtol=2e-5, atol=1e-5), f"{world}-way sample split differs from the single-GPU frame"
        assert sum(D.rows_of_rank(8, r, world) for r in range(world)) == 8


def test_exact_accumulation_makes_any_sample_split_bit_identical(renderer):
    """N virtual ranks, exact int64 partial frames summed in any order == the single-GPU frame, bit for bit"""
    import torch
    renderer.build_scene(8).override_camera(width=40, spp=64).commit()       # media, noise, image texture, instances
    st = renderer.stats
    H, W = st["height"], st["width"]
    full = torch.zeros(H, W, 4, dtype=torch.int64, device="cuda")
    renderer.render_device(full.data_ptr(), seed=3, frame=1, exact_accum=1)
    facc = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    renderer.resolve_exact_device(full.data_ptr(), facc.data_ptr())
    plain = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    renderer.render_device(plain.data_ptr(), seed=3, frame=1)
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(facc, nan=-1.0), torch.nan_to_num(plain, nan=-1.0)), "resolve(exact) must equal the float frame"
    for world in (2, 3, 8):
        parts = []
        for r in reversed(range(world)):                     # reversed on purpose: the order must not matter
            p = torch.zeros(H, W, 4, dtype=torch.int64, device="cuda")
            renderer.render_device(p.data_ptr(), seed=3, frame=1, exact_accum=1, sample_mod=world, sample_rem=r)
            parts.append(p)
        tot = sum(parts[1:], parts[0].clone())
        torch.cuda.synchronize()
        assert torch.equal(tot, full), f"{world}-way exact sample split differs from the single-GPU frame"
    from mort_b200.api import MortError, MODE_WAVEFRONT
    with pytest.raises(MortError):
        renderer.render_device(full.data_ptr(), exact_accum=1, mode=MODE_WAVEFRONT)


def test_tile_split_assembles_the_same_frame(renderer):
    """8-row bands dealt round-robin to N virtual ranks; each pixel is rendered by exactly one rank -> bit-identical frame"""
    import torch
    from mort_b200 import dist as D
    renderer.build_scene(1).override_camera(width=100, spp=16, depth=20).commit()      # 100 x 56: the last band is partial
    st = renderer.stats
    H, W = st["height"], st["width"]
    full = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    renderer.render_device(full.data_ptr(), seed=8)
    for world in (2, 3, 8):
        tot = torch.zeros_like(full)
        seen = torch.zeros(H, dtype=torch.int32)
        for r in range(world):
            part = torch.zeros_like(full)
            mod, rem = D.tile_split(r, world)
            renderer.render_device(part.data_ptr(), seed=8, tile_mod=mod, tile_rem=rem)
            torch.cuda.synchronize()
            rows = (part.abs().sum(dim=(1, 2)) > 0).cpu()
            want = torch.tensor([(y // 8) % world == r for y in range(H)])
            assert bool((rows <= want).all()), "a rank wrote outside its bands"
            seen += rows.int()
            tot += part
        torch.cuda.synchronize()
        assert torch.equal(torch.nan_to_num(tot, nan=-1.0), torch.nan_to_num(full, nan=-1.0)), f"{world}-way tile split differs"
    from mort_b200.api import MortError, MODE_WAVEFRONT
    with pytest.raises(MortError):
        renderer.render_device(full.data_ptr(), tile_mod=2, tile_rem=2)
    with pytest.raises(MortError):
        renderer.render_device(full.data_ptr(), tile_mod=2, tile_rem=1, mode=MODE_WAVEFRONT)


def test_staging_and_launch_shapes_do_not_change_the_image(renderer):
    renderer.build_scene(1).override_camera(width=96, spp=25, depth=50).commit()
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL
    base = renderer.render(seed=9).accum
    for kw in ({"stage_nodes": 64, "mode": MODE_MEGAKERNEL}, {"stage_nodes": 100000, "mode": MODE_MEGAKERNEL}, {"blocks_per_sm": 1, "mode": MODE_MEGAKERNEL},
               {"threads_per_block": 64, "mode": MODE_MEGAKERNEL}, {"mode": MODE_MEGAKERNEL}, {"mode": MODE_POOL, "threads_per_block": 256, "blocks_per_sm": 3, "pool_paths": 512}):
        other = renderer.render(seed=9, **kw).accum
        assert np.array_equal(base, other, equal_nan=True), f"{kw} changed the frame"


def test_tone_pipeline_matches_reference_formula(renderer):
    renderer.build_scene(2).override_camera(width=64, spp=16).commit()
    fr = renderer.render(seed=5)
    mean = fr.accum[..., :3] * np.float32(1.0 / 16)
    mean = np.where(np.isnan(mean), np.float32(0), mean)
    want = (np.float32(256) * np.clip(np.sqrt(mean), np.float32(0), np.float32(0.999))).astype(np.int32)
    assert np.abs(want - fr.rgba8[..., :3].astype(np.int32)).max() <= 1 and (want != fr.rgba8[..., :3]).mean() < 1e-3


def test_errors_are_reported_not_fatal(renderer):
    from mort_b200.api import MortError
    renderer.build_scene(5)
    with pytest.raises(MortError):
        renderer.render()                                # not committed
    renderer.commit()
    with pytest.raises(MortError):
        renderer.render(sample_mod=2, sample_rem=5)
    with pytest.raises(MortError):
        renderer.render(mode=7)


def test_custom_scene_through_builder_calls(renderer):
    """the scene-builder surface (world::add / constructors) end to end: cornell-like box, checked against brute force + oracle dump"""
    import oracle_binding as O
    r = renderer
    r.clear_scene()
    white = r.add_lambertian(r.add_solid(.73, .73, .73))
    light = r.add_diffuse_light(r.add_solid(15, 15, 15))
    lamp = r.add_quad((343, 554, 332), (-130, 0, 0), (0, 0, -105), light)
    r.add_quad((0, 0, 0), (555, 0, 0), (0, 0, 555), white)
    r.add_quad((0, 0, 555), (555, 0, 0), (0, 555, 0), white)
    r.add_rotated_box((165, 330, 165), (265, 0, 295), 15, white)
    ball = r.add_moving_sphere((190, 90, 190), (190, 120, 190), 90, r.add_metal(0.8, 0.85, 0.88, 0.1))
    smoke = r.add_sphere((400, 100, 150), 80, r.add_dielectric(1.5))
    r.add_constant_medium(smoke, 0.01, r.add_isotropic(r.add_solid(0.2, 0.4, 0.9)))
    # a visible top-level list next to a medium: world::hit tests it AFTER the media (world.cuh:154-168), which the
    # product reproduces with a second, visit-order-windowed traversal pass
    lst = r.add_list(False)
    r.list_add(lst, r.add_sphere((120, 60, 120), 60, r.add_lambertian(r.add_checker(40.0, r.add_solid(.9, .2, .2), r.add_solid(.2, .9, .2))), skip=True))
    r.list_add(lst, r.add_quad((300, 0.5, 300), (150, 0, 0), (0, 0, 150), white, skip=True))
    cam = r.get_camera()
    cam.aspect_ratio = 1.0; cam.image_width = 48; cam.samples_per_pixel = 16; cam.bounce_limit = 12; cam.vfov = 40
    cam.background[:] = (0.02, 0.02, 0.02); cam.lookfrom[:] = (278, 278, -800); cam.lookat[:] = (278, 278, 0); cam.vup[:] = (0, 1, 0)
    cam.light_obj_type, cam.light_obj_idx = lamp.type, lamp.idx
    r.set_camera(cam)
    r.commit()
    path = "/tmp/mort_custom_scene.mscn"
    r.dump_scene(path)
    osc = O.OracleScene(path)
    rng = np.random.default_rng(4)
    rays = np.concatenate([np.tile([278, 278, -800], (4000, 1)), rng.normal(size=(4000, 3)) * [0.35, 0.35, 0.0] + [0, 0, 1], rng.random((4000, 1))], 1).astype(np.float32)
    hits, probes = r.trace(rays)
    ref_hits, ref_probes = osc.trace(rays)
    b = ref_hits["hit"] == 1
    assert (hits["hit"] == ref_hits["hit"]).all() and (hits["t"].view(np.uint32)[b] == ref_hits["t"].view(np.uint32)[b]).all()
    assert ((hits["leaf_type"] == ref_hits["leaf_type"]) & (hits["leaf_idx"] == ref_hits["leaf_idx"]))[b].all()
    assert (probes["hit2"] == ref_probes["hit2"]).all() and (probes["t2"].view(np.uint32) == ref_probes["t2"].view(np.uint32)).all()
    fr = r.render(seed=11)
    hdr, _, st = osc.render(seed=11)
    ok = (fr.accum[..., 3] == 0) & (hdr[..., 3] == 0) & np.isfinite(fr.accum[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
    rel = np.abs(fr.accum[..., :3][ok] - hdr[..., :3][ok]).max(-1) / (np.abs(hdr[..., :3][ok]).max(-1) + 0.16)
    assert (rel > 1e-3).mean() <= 0.05
    assert abs(fr.stats["last_segments"] - st["segments"]) <= 0.003 * st["segments"] + 16


@pytest.mark.parametrize("sc,w,spp", [(1, 96, 36), (6, 64, 64), (7, 48, 16), (9, 48, 16), (5, 64, 25)])
def test_wavefront_renders_the_same_frame_as_the_megakernel(renderer, sc, w, spp):
    """same Philox stream, same per-ray code, different scheduling: only the order of the per-pixel float sums differs"""
    from mort_b200.api import MODE_MEGAKERNEL, MODE_WAVEFRONT
    renderer.build_scene(sc).override_camera(width=w, spp=spp).commit()
    a = renderer.render(seed=21, mode=MODE_MEGAKERNEL)
    for paths in (0, w * w):                              # default slot count, and the minimum (one slot per pixel)
        b = renderer.render(seed=21, mode=MODE_WAVEFRONT, wavefront_paths=paths)
        # the two kernels inline the shared per-ray code separately, so FMA contraction in the shading arithmetic may
        # differ: identical up to the few paths that sit on a decision boundary
        assert (a.accum[..., 3] != b.accum[..., 3]).mean() <= 0.003
        ok = (a.accum[..., 3] == 0) & (b.accum[..., 3] == 0) & np.isfinite(a.accum[..., :3]).all(-1) & np.isfinite(b.accum[..., :3]).all(-1)
        rel = np.abs(a.accum[..., :3][ok] - b.accum[..., :3][ok]).max(-1) / (np.abs(a.accum[..., :3][ok]).max(-1) + 0.01 * spp)
        assert (rel > 1e-3).mean() <= 0.03, f"{(rel > 1e-3).mean():.4f} of pixels differ"
        assert abs(b.stats["last_segments"] - a.stats["last_segments"]) <= 0.002 * a.stats["last_segments"] + 16
        assert b.stats["last_samples"] == a.stats["last_samples"]
        assert (np.abs(a.rgba8.astype(int) - b.rgba8.astype(int)) > 1).mean() <= 0.03
    c = renderer.render(seed=21, mode=MODE_WAVEFRONT)
    d = renderer.render(seed=21, mode=MODE_WAVEFRONT)
    assert np.array_equal(c.accum, d.accum, equal_nan=True), "wavefront frames must be bit-reproducible"


def test_bright_emitters_do_not_wrap_the_fixed_point_sums(renderer):
    """1024 samples of radiance 3e6 sum to 3e9 per pixel: beyond a Q31.32 accumulator, well inside the Q39.24 one"""
    import oracle_binding as O
    r = renderer
    r.clear_scene()
    lamp = r.add_diffuse_light(r.add_solid(3e6, 2e6, 1e6))
    r.add_quad((-1, -1, 0), (2, 0, 0), (0, 2, 0), lamp)            # faces the camera (normal +z), fills the view
    cam = r.get_camera()
    cam.aspect_ratio = 1.0; cam.image_width = 16; cam.samples_per_pixel = 1024; cam.bounce_limit = 4; cam.vfov = 20
    cam.background[:] = (0, 0, 0); cam.lookfrom[:] = (0, 0, 3); cam.lookat[:] = (0, 0, 0); cam.vup[:] = (0, 1, 0)
    cam.light_obj_type = -1
    r.set_camera(cam)
    r.commit()
    fr = r.render(seed=3)
    assert np.isfinite(fr.accum[..., :3]).all() and (fr.accum[..., 3] == 0).all()
    centre = fr.accum[4:12, 4:12, :3] / 1024
    assert np.allclose(centre, [3e6, 2e6, 1e6], rtol=1e-6)
    assert (fr.rgba8[4:12, 4:12, :3] == 255).all()
    path = "/tmp/mort_bright_scene.mscn"
    r.dump_scene(path)
    hdr, _, _ = O.OracleScene(path).render(seed=3, want_rgba8=False)
    assert np.allclose(fr.accum[..., :3], hdr[..., :3], rtol=1e-4, atol=1.0)
