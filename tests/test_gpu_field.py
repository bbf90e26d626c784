"""Synthetic sphere field (BASELINE.json config 4) and scene-file ingestion on the GPU: sizes far beyond the
reference's fixed capacities (objects.cuh:451,521,746), checked against brute force, the oracle and itself."""
import numpy as np
import pytest

from conftest import GOLDEN, bits, golden_scene_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _rays(n, seed, extent):
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-extent, extent, n), rng.uniform(0.05, 3.0, n), rng.uniform(-extent, extent, n)], 1)
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.3          # mostly grazing downwards: many candidates per ray
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d, rng.random((n, 1))], 1).astype(np.float32)


def test_field_40k_spheres_bvh_vs_brute_force_vs_oracle(renderer, tmp_path):
    import oracle_binding as O
    renderer.build_sphere_field(100, seed=7).commit()                      # ~40 000 spheres, 36 x the reference's capacity
    st = renderer.stats
    assert st["n_leaves"] > 38_000 and st["n_nodes"] > 1000
    rays = _rays(20_000, 3, 100.0)
    out, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    assert (out["hit"] == brute["hit"]).all() and (bits(out["t"]) == bits(brute["t"])).all()
    assert ((out["leaf_type"] == brute["leaf_type"]) & (out["leaf_idx"] == brute["leaf_idx"])).all()
    assert out["hit"].mean() > 0.9
    path = str(tmp_path / "field.mscn")
    renderer.dump_scene(path)
    osc = O.OracleScene(path)
    ref, _ = osc.trace(rays[:1500])                                        # the oracle scans 40 k spheres per ray
    b = ref["hit"] == 1
    assert (out["hit"][:1500] == ref["hit"]).all() and (bits(out["t"][:1500])[b] == bits(ref["t"])[b]).all()
    assert (out["leaf_idx"][:1500][b] == ref["leaf_idx"][b]).all() and (bits(out["p"][:1500])[b] == bits(ref["p"])[b]).all()
    # frames: same stream as the oracle on a small crop of the workload
    renderer.override_camera(width=48, spp=4, depth=6)
    fr = renderer.render(seed=3)
    osc.override(width=48, spp=4, depth=6)
    hdr, _, stt = osc.render(seed=3)
    ok = (fr.accum[..., 3] == 0) & (hdr[..., 3] == 0) & np.isfinite(hdr[..., :3]).all(-1) & np.isfinite(fr.accum[..., :3]).all(-1)
    rel = np.abs(fr.accum[..., :3][ok] - hdr[..., :3][ok]).max(-1) / (np.abs(hdr[..., :3][ok]).max(-1) + 0.04)
    assert (rel > 1e-3).mean() <= 0.05
    assert abs(fr.stats["last_segments"] - stt["segments"]) <= 0.003 * stt["segments"] + 16


def test_field_one_million_spheres_commits_and_renders(renderer):
    renderer.build_sphere_field(500, seed=69420, camera_kind=1).commit()
    st = renderer.stats
    assert st["n_leaves"] > 950_000 and st["bvh_depth"] < 24
    rays = _rays(4_000, 5, 500.0)
    out, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)                      # 4 k rays x 1 M spheres
    assert (out["hit"] == brute["hit"]).all() and (bits(out["t"]) == bits(brute["t"])).all() and (out["leaf_idx"] == brute["leaf_idx"]).all()
    renderer.override_camera(width=320, spp=16, depth=50)
    a = renderer.render(seed=1).accum
    b = renderer.render(seed=1).accum
    assert np.array_equal(a, b, equal_nan=True) and np.isfinite(a[..., :3]).mean() > 0.99
    assert 0.05 < (a[..., :3] / 16).mean() < 1.0


@pytest.mark.parametrize("sc", [1, 6, 8])
def test_scene_file_round_trip(renderer, sc, tmp_path):
    """mort_load_scene on the REFERENCE's own dump renders the same bits as the restated scene function"""
    renderer.build_scene(sc).override_camera(width=64, spp=16).commit()
    a = renderer.render(seed=2).accum
    renderer.load_scene(golden_scene_path(sc, str(tmp_path))).override_camera(width=64, spp=16).commit()
    b = renderer.render(seed=2).accum
    assert np.array_equal(a, b, equal_nan=True)
