"""SURVEY.md section 8f-4 on the GPU: the binned-SAH tree built on the device (gpu_build.cu) against the host builder
(bvh_build.cpp, the CPU statement of the same rule, bvh_sah.hpp), refit after edits against brute force and a fresh commit, and
motion-aware node boxes against brute force.  The reference side of all this is a host median-split build capped at 1024 nodes
(objects.cuh:521,528-661) and union boxes for moving spheres (objects.cuh:46-55); what must not change is every hit."""
import json

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from mort_b200.api import Renderer
    r = Renderer(0)
    yield r
    r.close()


def _rays(n, seed, extent, cam=None):
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-extent, extent, n), rng.uniform(0.05, 3.0, n), rng.uniform(-extent, extent, n)], 1)
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) * 0.3
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d, rng.random((n, 1))], 1).astype(np.float32)


def _same_hits(a, b, what):
    assert (a["hit"] == b["hit"]).all(), f"{what}: hit flags differ on {(a['hit'] != b['hit']).sum()} rays"
    assert (bits(a["t"]) == bits(b["t"])).all(), f"{what}: t differs"
    assert ((a["leaf_type"] == b["leaf_type"]) & (a["leaf_idx"] == b["leaf_idx"])).all(), f"{what}: primitive ids differ"


def _exact(r, **kw):
    import torch
    from mort_b200.api import MODE_POOL
    st = r.stats
    buf = torch.zeros((st["height"], st["width"], 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    r.render_device(buf.data_ptr(), mode=MODE_POOL, exact_accum=1, **kw)
    torch.cuda.synchronize()
    return buf.cpu().numpy()


@pytest.mark.parametrize("what", ["scene1", "scene8", "field64", "field64_plain_atomics", "field64_small16"])
def test_gpu_build_is_the_host_tree(renderer, what):
    """same fingerprint (it hashes the node array and the records in leaf order), same statistics, same hits"""
    from mort_b200.api import BUILD_GPU, BUILD_HOST
    def build(builder, **kw):
        if what.startswith("scene"):
            renderer.build_scene(int(what[5:]))
        else:
            renderer.build_sphere_field(64, seed=11)
        renderer.set_build_opts(builder=builder, **kw).commit()
        return renderer.stats, renderer.build_info, renderer.scene_fingerprint
    kw = {"gpu_flags": 1} if what.endswith("plain_atomics") else {"gpu_small": 16} if what.endswith("small16") else {}
    hs, hi, hf = build(BUILD_HOST)
    rays = _rays(20_000, 5, 8.0 if what == "scene1" else 64.0)
    if what == "scene8":
        rays[:, :3] = rays[:, :3] * 4 + np.array([200, 200, 200], dtype=np.float32)
    host_hits, _ = renderer.trace(rays)
    gs, gi, gf = build(BUILD_GPU, **kw)
    assert hi["built_on_gpu"] == 0 and gi["built_on_gpu"] == 1 and gi["gpu_levels"] >= 1
    assert gs["n_nodes"] == hs["n_nodes"] and gs["bvh_depth"] == hs["bvh_depth"] and gi["bvh2_nodes"] == hi["bvh2_nodes"]
    assert abs(gs["sah_cost"] - hs["sah_cost"]) <= 1e-4 * hs["sah_cost"]
    assert gf == hf, "the GPU-built tree differs from the host-built tree"
    gpu_hits, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    _same_hits(gpu_hits, host_hits, what + ": GPU tree vs host tree")
    _same_hits(gpu_hits, brute, what + ": GPU tree vs brute force")
    print(json.dumps({"what": what, "leaves": gs["n_leaves"], "nodes": gs["n_nodes"], "host_build_ms": hs["build_ms"], "gpu_build_ms": gs["build_ms"],
                      "gpu_stream_ms": gi["gpu_stream_ms"], "levels": gi["gpu_levels"], "small_subtrees": gi["gpu_small_subtrees"], "workspace_mb": gi["gpu_workspace_bytes"] / 2**20}))
    renderer.set_build_opts()


def test_gpu_build_one_million_spheres(renderer):
    """BASELINE config 4's field: AUTO picks the GPU builder; the tree is the host's; commit time is reported"""
    from mort_b200.api import BUILD_HOST
    renderer.set_build_opts()
    renderer.build_sphere_field(500, seed=69420, camera_kind=1).commit()
    gs, gi, gf = renderer.stats, renderer.build_info, renderer.scene_fingerprint
    assert gi["built_on_gpu"] == 1 and gs["n_leaves"] > 950_000
    rays = _rays(4_000, 5, 500.0)
    out, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    _same_hits(out, brute, "1 M field, GPU-built tree vs brute force")
    renderer.set_build_opts(builder=BUILD_HOST).commit()
    hs, hf = renderer.stats, renderer.scene_fingerprint
    assert hf == gf and hs["n_nodes"] == gs["n_nodes"]
    print(json.dumps({"what": "field500", "leaves": gs["n_leaves"], "nodes": gs["n_nodes"], "host_build_ms": hs["build_ms"], "gpu_build_ms": gs["build_ms"],
                      "gpu_stream_ms": gi["gpu_stream_ms"], "levels": gi["gpu_levels"], "small_subtrees": gi["gpu_small_subtrees"],
                      "flatten_ms": gi["flatten_ms"], "upload_ms": gs["upload_ms"], "workspace_mb": gi["gpu_workspace_bytes"] / 2**20}))
    renderer.set_build_opts()


def test_refit_after_moving_spheres(renderer):
    """spheres move, the topology stays: hits equal brute force and a fresh commit of the edited scene; frames equal bit for bit"""
    from mort_b200.api import Handle
    from mort_b200 import formats as F
    renderer.set_build_opts()
    renderer.build_sphere_field(40, seed=3).commit()                        # ~6400 spheres
    n_sph = renderer.stats["n_spheres"]
    rays = _rays(30_000, 9, 40.0)
    before, _ = renderer.trace(rays)
    rng = np.random.default_rng(4)
    moved = rng.choice(np.arange(1, n_sph - 4), size=600, replace=False)     # slot 0 is the ground
    for k, slot in enumerate(moved):
        c0 = np.array([rng.uniform(-38, 38), rng.uniform(0.2, 2.5), rng.uniform(-38, 38)])
        c1 = c0 + np.array([0, rng.uniform(0, 0.8), 0]) if k % 2 else None
        renderer.update_sphere(Handle(F.OBJ_SPHERE, int(slot)), c0, c1, float(rng.uniform(0.1, 0.6)))
    renderer.refit()
    info = renderer.build_info
    assert info["refits"] == 1
    after, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    _same_hits(after, brute, "refitted tree vs brute force")
    assert (bits(after["t"]) != bits(before["t"])).mean() > 0.01, "the edit changed nothing?"
    renderer.override_camera(width=96, spp=4, depth=8)
    a = _exact(renderer, seed=5)
    renderer.commit()                                                      # fresh build of the edited scene
    fresh, _ = renderer.trace(rays)
    _same_hits(after, fresh, "refitted tree vs fresh commit")
    b = _exact(renderer, seed=5)
    assert np.array_equal(a, b)
    print(json.dumps({"what": "refit", "spheres": n_sph, "moved": int(len(moved)), "refit_ms": info["refit_ms"]}))


def test_refit_of_a_light_or_boundary_sphere_needs_a_commit(renderer):
    """scene 8's media are bounded by spheres whose geometry is copied into the boundary records: editing one invalidates the commit"""
    from mort_b200.api import Handle, MortError
    from mort_b200 import formats as F
    renderer.set_build_opts()
    refused = 0
    for slot in range(8):
        renderer.build_scene(8).commit()
        renderer.update_sphere(Handle(F.OBJ_SPHERE, slot), (0, 150, 145), None, 50.0)
        try:
            renderer.refit()
        except MortError:
            refused += 1
            renderer.commit()                                              # the documented way out
            out, _ = renderer.trace(_rays(2000, 1, 300.0))
            brute, _ = renderer.trace(_rays(2000, 1, 300.0), brute_force=True)
            _same_hits(out, brute, "scene 8 re-committed after an edit")
    assert 1 <= refused < 8


@pytest.mark.parametrize("sc", ["scene1", "field48"])
def test_motion_bounds_change_no_hit(renderer, sc):
    """node boxes interpolated at the ray's time: hits equal brute force and the union-box tree; frames equal bit for bit"""
    def build(motion):
        if sc == "scene1":
            renderer.build_scene(1)
        else:
            renderer.build_sphere_field(48, seed=2)
        renderer.set_build_opts(motion_bounds=motion).commit()
    rays = _rays(40_000, 13, 10.0 if sc == "scene1" else 48.0)
    rays[::7, 6] = 0.0
    rays[3::7, 6] = np.float32(1.0) - np.float32(2.0 ** -24)
    build(0)
    union_hits, _ = renderer.trace(rays)
    assert renderer.build_info["motion_nodes"] == 0
    renderer.override_camera(width=120, spp=9, depth=12)
    a = _exact(renderer, seed=8)
    ms_union = renderer.stats["last_render_ms"]
    build(1)
    assert renderer.build_info["motion_nodes"] == renderer.stats["n_nodes"]
    motion_hits, _ = renderer.trace(rays)
    brute, _ = renderer.trace(rays, brute_force=True)
    _same_hits(motion_hits, brute, sc + ": motion boxes vs brute force")
    _same_hits(motion_hits, union_hits, sc + ": motion boxes vs union boxes")
    renderer.override_camera(width=120, spp=9, depth=12)
    b = _exact(renderer, seed=8)
    assert np.array_equal(a, b)
    print(json.dumps({"what": "motion " + sc, "ms_union": ms_union, "ms_motion": renderer.stats["last_render_ms"]}))
    renderer.set_build_opts()
