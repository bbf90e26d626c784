"""GPU: the fuzz parity of tests/test_fuzz_parity.py through the C ABI — random scene files loaded with mort_load_scene_text,
closest hits and medium boundary probes from mort_trace (tree and brute force) against the oracle, bit for bit."""
import numpy as np
import pytest

from conftest import bits
from mort_b200 import formats as F
from test_fuzz_parity import REFUSALS
from test_scene_text import _random_scene_text

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(300, 324)) + list(range(1300, 1316)))
def test_random_scene_through_the_abi(seed, tmp_path):
    """seeds >= 1000: the generator nests media in wrappers and lists most of the time (media_stages in rt_core.cuh, all three schedulers)"""
    import oracle_binding as O
    from mort_b200.api import MODE_MEGAKERNEL, MortError, Renderer
    rng = np.random.default_rng(seed)
    txt, dump = tmp_path / "s.txt", str(tmp_path / "s.mscn")
    txt.write_text(_random_scene_text(rng, 0.7 if seed >= 1000 else 0.1))
    with Renderer(0) as r:
        r.load_scene_text(str(txt))
        r.dump_scene(dump)
        try:
            r.commit()
        except MortError as e:
            assert any(x in str(e) for x in REFUSALS), str(e)
            pytest.skip("scene refused: " + str(e)[-90:])
        n = 4000
        o, tgt = rng.uniform(-8, 8, (n, 3)), rng.uniform(-5, 5, (n, 3))
        rays = np.concatenate([o, tgt - o, rng.random((n, 1))], 1).astype(np.float32)
        hits, probes = r.trace(rays)
        brute, _ = r.trace(rays, brute_force=True)
        ref, ref_probes = O.OracleScene(dump).trace(rays)
        b = ref["hit"] == 1
        assert (hits["hit"] == ref["hit"]).all() and (bits(hits["t"])[b] == bits(ref["t"])[b]).all()
        for k in ("leaf_type", "leaf_idx", "mat_type", "mat_idx", "front_face"):
            assert (hits[k][b] == ref[k][b]).all(), k
        assert (bits(hits["p"])[b] == bits(ref["p"])[b]).all() and (bits(hits["normal"])[b] == bits(ref["normal"])[b]).all()
        assert (hits["hit"] == brute["hit"]).all() and (bits(hits["t"]) == bits(brute["t"])).all() and (hits["leaf_idx"] == brute["leaf_idx"]).all()
        # the product only keeps the media world::hit can reach: top-level (not hidden) ones, none at all once a bvh exists
        sc = F.read_scene(dump)
        vis = [] if len(sc["bvhs"]) else [i for i, m in enumerate(sc["media"]) if int(m["skip"]) == 0]
        assert probes.shape[1] == len(vis)
        for j, m in enumerate(vis):
            mine, want = probes[:, j], ref_probes[:, m]
            assert (mine["hit1"] == want["hit1"]).all() and (mine["hit2"] == want["hit2"]).all()
            b1, b2 = want["hit1"] == 1, want["hit2"] == 1
            assert (bits(mine["t1"])[b1] == bits(want["t1"])[b1]).all() and (bits(mine["t2"])[b2] == bits(want["t2"])[b2]).all()
        # a small frame against the oracle with the same stream
        r.override_camera(width=32, spp=9)
        fr = r.render(seed=5)
        if seed >= 1000:                                  # the schedulers share the per-ray code: same exact frame from the pool and the megakernel
            other = r.render(seed=5, mode=MODE_MEGAKERNEL)        # a scene with nested media goes to the general-media kernel whatever the mode
            assert np.array_equal(other.accum, fr.accum, equal_nan=True)
        osc = O.OracleScene(dump)
        osc.override(width=32, spp=9)
        hdr, _, st = osc.render(seed=5, want_rgba8=False)
        assert (fr.accum[..., 3] != hdr[..., 3]).mean() <= 0.02
        ok = (fr.accum[..., 3] == 0) & (hdr[..., 3] == 0) & np.isfinite(fr.accum[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
        if ok.any():
            rel = np.abs(fr.accum[..., :3][ok] - hdr[..., :3][ok]).max(-1) / (np.abs(hdr[..., :3][ok]).max(-1) + 0.16)
            assert (rel > 1e-3).mean() <= 0.08
