"""GPU parity, part 1: primary-ray hits through the C ABI (mort_trace) against (a) the reference's own
records committed under tests/golden/, (b) the oracle on fresh ray sets, (c) brute force over all leaves
(the padded SAH BVH must never cull a hit the exact primitive tests accept)."""
import numpy as np
import pytest

from conftest import GOLDEN, bits, golden_scene_path

pytestmark = pytest.mark.gpu

SCENES = list(range(1, 11))


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _check(out, ref, what, probes=None, ref_probes=None, check_top=True):
    assert (out["hit"] == ref["hit"]).all(), f"{what}: hit flags differ on {(out['hit'] != ref['hit']).sum()} rays"
    b = ref["hit"] == 1
    assert (bits(out["t"])[b] == bits(ref["t"])[b]).all(), f"{what}: t not bit-equal on {(bits(out['t'])[b] != bits(ref['t'])[b]).sum()} rays"
    assert ((out["leaf_type"] == ref["leaf_type"]) & (out["leaf_idx"] == ref["leaf_idx"]))[b].all(), f"{what}: primitive ids differ"
    if check_top:
        assert ((out["top_type"] == ref["top_type"]) & (out["top_idx"] == ref["top_idx"]))[b].all(), f"{what}: top-level objects differ"
    assert (out["front_face"] == ref["front_face"])[b].all(), f"{what}: front_face differs"
    assert ((out["mat_type"] == ref["mat_type"]) & (out["mat_idx"] == ref["mat_idx"]))[b].all(), f"{what}: materials differ"
    assert (bits(out["p"])[b] == bits(ref["p"])[b]).all(), f"{what}: hit points not bit-equal"
    assert (bits(out["normal"])[b] == bits(ref["normal"])[b]).all(), f"{what}: normals not bit-equal"
    # u,v go through acosf/atan2f (libm): a few ulp
    for k in ("u", "v"):                               # acosf(-y) is NaN on both sides when rounding pushes |y| past 1
        both_nan = np.isnan(out[k]) & np.isnan(ref[k])
        assert np.abs(np.where(both_nan, 0, out[k] - ref[k]))[b].max(initial=0) <= 4e-7, k
    if probes is not None and ref_probes is not None and ref_probes.size:
        for k in ("hit1", "hit2"):
            assert (probes[k] == ref_probes[k]).all(), f"{what}: medium probe {k} differs"
        assert (bits(probes["t1"]) == bits(ref_probes["t1"])).all() and (bits(probes["t2"]) == bits(ref_probes["t2"])).all()


@pytest.mark.parametrize("sc", SCENES)
def test_trace_matches_reference_records(renderer, sc):
    g = np.load(f"{GOLDEN}/hits_{sc}.npz")
    renderer.build_scene(sc).commit()
    for kind in ("grid", "rnd"):
        out, probes = renderer.trace(g[f"{kind}_rays"])
        _check(out, g[f"{kind}_hits"], f"scene {sc} {kind}", probes, g[f"{kind}_probes"])


def _random_rays(cam, n, seed, scale):
    rng = np.random.default_rng(seed)
    centre = np.asarray(cam["lookat"], dtype=np.float64)
    o = centre + (rng.random((n, 3)) - 0.5) * 2 * scale
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d *= rng.uniform(0.5, 2.0, (n, 1))
    # half the rays leave from the camera towards random scene points (coherent-ish), half are incoherent
    o[: n // 2] = np.asarray(cam["lookfrom"], dtype=np.float64)
    d[: n // 2] = (centre + (rng.random((n // 2, 3)) - 0.5) * 2 * scale) - o[: n // 2]
    return np.concatenate([o, d, rng.random((n, 1))], axis=1).astype(np.float32)


@pytest.mark.parametrize("sc", SCENES)
def test_trace_matches_oracle_and_brute_force(renderer, earth, sc, tmp_path):
    import oracle_binding as O
    renderer.build_scene(sc).commit()
    osc = O.OracleScene(golden_scene_path(sc, str(tmp_path)), earth)
    cam = osc.camera
    scale = {1: 12.0, 2: 12.0, 3: 4.0, 4: 8.0, 5: 6.0, 6: 400.0, 7: 400.0, 8: 700.0, 9: 700.0, 10: 40.0}[sc]
    n = 200_000 if sc in (8, 9) else 400_000
    rays = _random_rays(cam, n, 1000 + sc, scale)
    out, probes = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True, want_probes=False)
    assert (out["hit"] == brute["hit"]).all() and (bits(out["t"]) == bits(brute["t"])).all(), f"scene {sc}: BVH and brute force disagree"
    assert ((out["leaf_type"] == brute["leaf_type"]) & (out["leaf_idx"] == brute["leaf_idx"])).all(), f"scene {sc}: BVH and brute force pick different primitives"
    m = 40_000 if sc in (8, 9) else 100_000          # the oracle scans linearly: bound its share
    ref, ref_probes = osc.trace(rays[:m])
    _check(out[:m], ref, f"scene {sc} vs oracle", probes[:m], ref_probes)


def test_trace_edge_cases(renderer):
    """empty world, zero rays, axis-parallel and degenerate rays"""
    renderer.build_scene(0).commit()
    out, _ = renderer.trace(np.array([[0, 0, 0, 0, 0, -1, 0.5]], dtype=np.float32))
    assert out["hit"][0] == 0
    out, _ = renderer.trace(np.zeros((0, 7), dtype=np.float32))
    assert len(out) == 0
    renderer.build_scene(5).commit()
    rays = np.array([[0, 0, 9, 0, 0, -1, 0], [0, 0, 9, 0, 0, 1, 0], [0, 0, 9, 0, 0, 0, 0], [-3, 0, 3, 0, 0, 0, 0], [0, 0, 9, 1e-30, 0, -1, 0]], dtype=np.float32)
    out, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    assert (out["hit"] == brute["hit"]).all() and (bits(out["t"]) == bits(brute["t"])).all()
    assert out["hit"][0] == 1 and out["leaf_type"][0] == 2 and out["hit"][1] == 0
