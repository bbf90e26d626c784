"""Reference fixtures for the paths none of the reference's ten scene functions reaches.

Scenes 101-104 are built inside oracle/ref_harness.cu from the reference's OWN classes through its own builder calls and
rendered / traced by its unmodified device code on a B200 (scripts/gpu_ref_fixtures2.sh -> scripts/curate_golden.py):
  101  two media scattering with `isotropic` (materials.cuh:169-202, sphere_pdf pdf.cuh:28-42) behind a sphere and behind a
       translate(rotate_y(list)) boundary; light handle = one quad sampled directly (objects.cuh:217-235)
  102  light handle = one emissive sphere: hittable_pdf -> sphere::pdf_value / sphere::random (objects.cuh:110-145)
  103  scene 1 with defocus_angle = 0.6 (lens sampling, camera.cuh:222-242)
  104  a medium + a VISIBLE top-level list (tested after the media, world.cuh:154-168) + translate(translate(rotate_y(list)))
The CPU half pins the ORACLE to them, the GPU half pins the PRODUCT (both schedulers) to the very same records.
"""
import numpy as np
import pytest

import oracle_binding as O
from conftest import GOLDEN, bits, golden_scene_path, luminance, trimmed_rmse
from mort_b200 import formats as F

EXTRA = [101, 102, 103, 104]


def _check_hits(out, probes, g, kind):
    ref, rp = g[f"{kind}_hits"], g[f"{kind}_probes"]
    b = ref["hit"] == 1
    assert (out["hit"] == ref["hit"]).all()
    assert (bits(out["t"])[b] == bits(ref["t"])[b]).all(), "t must match the reference bit for bit"
    for k in ("leaf_type", "leaf_idx", "top_type", "top_idx", "mat_type", "mat_idx", "front_face"):
        assert (out[k][b] == ref[k][b]).all(), k
    assert (bits(out["p"])[b] == bits(ref["p"])[b]).all() and (bits(out["normal"])[b] == bits(ref["normal"])[b]).all()
    if rp.size:
        assert (probes["hit1"] == rp["hit1"]).all() and (probes["hit2"] == rp["hit2"]).all()
        assert (bits(probes["t1"]) == bits(rp["t1"])).all() and (bits(probes["t2"]) == bits(rp["t2"])).all()
    assert (ref["flags"][b] & 1).sum() == 0, "the harness identified every reference hit by bit-equal t"


# ------------------------------------------------------------------------------------------------------------------
# CPU: the oracle against the reference
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sc", EXTRA)
def test_oracle_primary_hits_bit_exact(sc):
    g = np.load(f"{GOLDEN}/hits_{sc}.npz")
    osc = O.OracleScene(golden_scene_path(sc))
    for kind in ("grid", "rnd"):
        out, probes = osc.trace(g[f"{kind}_rays"])
        _check_hits(out, probes, g, kind)


@pytest.mark.parametrize("sc", EXTRA)
def test_oracle_camera_initialize(sc):
    path = golden_scene_path(sc)
    ref = F.read_scene(path)["camera"]
    osc = O.OracleScene(path)
    osc.override()
    assert osc.camera.tobytes() == ref.tobytes()            # 103: includes the defocus disk basis


@pytest.mark.parametrize("sc", EXTRA)
def test_oracle_noisy_frame_statistics(sc):
    """Oracle (Philox) vs reference (XORWOW) at 96 px / 64 spp: RMSE at the reference's own seed-to-seed level."""
    g = np.load(f"{GOLDEN}/small_{sc}.npz")
    ra, rb = g["smallhdr_a"], g["smallhdr_b"]
    osc = O.OracleScene(golden_scene_path(sc))
    osc.override(width=96, spp=64)
    assert (int(osc.camera["image_height"]), int(osc.camera["image_width"])) == ra.shape[:2]
    hdr, _, _ = osc.render(seed=2024, want_rgba8=False)
    ok = (ra[..., 3] == 0) & (rb[..., 3] == 0) & (hdr[..., 3] == 0)
    ok &= np.isfinite(ra[..., :3]).all(-1) & np.isfinite(rb[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
    assert ok.mean() > 0.97
    a, b, o = ra[..., :3][ok] / 64, rb[..., :3][ok] / 64, hdr[..., :3][ok] / 64
    floor = np.sqrt(((a - b) ** 2).mean())
    mine = np.sqrt(((o - a) ** 2).mean())
    assert mine <= 1.25 * floor + 1e-4, f"scene {sc}: RMSE vs reference {mine:.5f}, floor {floor:.5f}"
    la, lb, lo = luminance(a).mean(), luminance(b).mean(), luminance(o).mean()
    assert abs(lo - la) <= 4 * abs(la - lb) + 0.02 * la + 1e-4, f"scene {sc}: mean luminance oracle {lo:.5f} vs reference {la:.5f} / {lb:.5f}"


@pytest.mark.parametrize("sc", EXTRA)
def test_oracle_converged_frame(sc):
    """160 px / 1024 spp against the reference's frame: per-channel RMSE within 2x its seed-to-seed floor, luminance 0.5 %."""
    g = np.load(f"{GOLDEN}/conv_{sc}.npz")
    ref = g["mean_a"].astype(np.float32)
    H, W = ref.shape[:2]
    osc = O.OracleScene(golden_scene_path(sc))
    spp = 256                                              # CPU budget: a quarter of the fixture's samples (noise scaled below)
    osc.override(width=W, spp=spp)
    hdr, _, _ = osc.render(seed=777, want_rgba8=False, threads=8)
    mine = hdr[..., :3] / spp
    ok = (hdr[..., 3] == 0) & (g["nan_a"] == 0) & np.isfinite(mine).all(-1) & np.isfinite(ref).all(-1)
    rmse = trimmed_rmse(mine[ok], ref[ok])                 # firefly-robust on both sides (scene 101: 5 pixels carry 70 % of the squared error)
    floor = g["rmse_ab_trim"] * np.sqrt(0.5 + 0.5 * int(g["spp"]) / spp)      # sqrt(sigma_1024^2 + sigma_256^2) in units of sqrt(2) sigma_1024
    assert (rmse <= 1.3 * floor + 2e-3 * np.abs(ref[ok]).mean(axis=0) + 1e-5).all(), f"scene {sc}: RMSE {rmse} vs expected {floor}"
    la, lb = float(luminance(mine[ok]).mean()), float(luminance(ref[ok]).mean())
    assert abs(la - lb) <= 0.01 * lb + 1e-6, f"scene {sc}: mean luminance {la} vs reference {lb}"


# ------------------------------------------------------------------------------------------------------------------
# GPU: the product against the same records
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sc", EXTRA)
def test_product_primary_hits_bit_exact(renderer, sc):
    g = np.load(f"{GOLDEN}/hits_{sc}.npz")
    renderer.load_scene(golden_scene_path(sc)).commit()
    for kind in ("grid", "rnd"):
        out, probes = renderer.trace(g[f"{kind}_rays"])
        _check_hits(out, probes, g, kind)
        brute, _ = renderer.trace(g[f"{kind}_rays"], brute_force=True)
        assert (bits(out["t"]) == bits(brute["t"])).all() and (out["leaf_idx"] == brute["leaf_idx"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["mega", "pool"])
@pytest.mark.parametrize("sc", EXTRA)
def test_product_converged_frame(renderer, sc, mode):
    from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL
    g = np.load(f"{GOLDEN}/conv_{sc}.npz")
    ref = g["mean_a"].astype(np.float32)
    H, W = ref.shape[:2]
    spp = int(g["spp"])
    renderer.load_scene(golden_scene_path(sc)).override_camera(width=W, spp=spp).commit()
    st = renderer.stats
    assert (st["height"], st["width"]) == (H, W)
    fr = renderer.render(seed=777, mode=MODE_POOL if mode == "pool" else MODE_MEGAKERNEL)
    n = st["sqrt_spp"] ** 2
    mine = fr.accum[..., :3] / n
    ok = (fr.accum[..., 3] == 0) & (g["nan_a"] == 0) & np.isfinite(mine).all(-1) & np.isfinite(ref).all(-1)
    assert ok.mean() > 0.97
    rmse = trimmed_rmse(mine[ok], ref[ok])
    assert (rmse <= 2.0 * g["rmse_ab_trim"] + 2e-3 * np.abs(ref[ok]).mean(axis=0) + 1e-5).all(), f"scene {sc}: RMSE {rmse} vs noise floor {g['rmse_ab_trim']}"
    la, lb = float(luminance(mine[ok]).mean()), float(luminance(ref[ok]).mean())
    assert abs(la - lb) <= 0.005 * lb + 1e-6, f"scene {sc}: mean luminance {la} vs reference {lb}"


@pytest.mark.gpu
@pytest.mark.parametrize("sc", EXTRA)
def test_product_frame_matches_oracle_same_stream(renderer, sc):
    renderer.load_scene(golden_scene_path(sc)).override_camera(width=64, spp=36).commit()
    fr = renderer.render(seed=4242, frame=3)
    osc = O.OracleScene(golden_scene_path(sc))
    osc.override(width=64, spp=36)
    hdr, _, st = osc.render(seed=4242, frame=3, want_rgba8=False)
    assert (fr.accum[..., 3] != hdr[..., 3]).mean() <= 0.003
    ok = (fr.accum[..., 3] == 0) & (hdr[..., 3] == 0) & np.isfinite(fr.accum[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
    a, b = fr.accum[..., :3][ok] / 36, hdr[..., :3][ok] / 36
    rel = np.abs(a - b).max(-1) / (np.abs(b).max(-1) + 1e-2)
    assert (rel > 1e-3).mean() <= 0.05, f"scene {sc}: {(rel > 1e-3).mean():.4f} of pixels differ by more than 1e-3"
    assert abs(fr.stats["last_segments"] - st["segments"]) <= 0.002 * st["segments"] + 16
