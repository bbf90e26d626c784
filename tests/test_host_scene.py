"""Host logic of the product on the CPU: scene builder vs the reference's dumps, flattening + SAH wide BVH +
per-ray code (through the test-only host simulation) vs the reference's hit records and the oracle's frames."""
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as O
from conftest import ASSETS, GOLDEN, bits, golden_scene_path
from mort_b200 import formats as F

SCENES = list(range(1, 11))


def run(hostsim, *args):
    return subprocess.run([hostsim, *map(str, args)], check=True, capture_output=True, text=True).stderr


@pytest.mark.parametrize("sc", SCENES)
def test_scene_dump_is_byte_identical_to_the_reference(hostsim, sc, tmp_path):
    out = tmp_path / f"s{sc}.mscn"
    run(hostsim, sc, ASSETS, "dump", out)
    assert open(out, "rb").read() == open(golden_scene_path(sc, str(tmp_path)), "rb").read()


def test_host_rand_is_glibc_rand(hostsim):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    want = [libc.rand() for _ in range(1000)]
    got = [int(x) for x in run(hostsim, 0, ASSETS, "rand", 1000).split()]
    assert got == want


@pytest.mark.parametrize("sc", SCENES)
def test_flatten_bvh_and_traversal_reproduce_reference_hits(hostsim, sc, tmp_path):
    g = np.load(f"{GOLDEN}/hits_{sc}.npz")
    for kind in ("grid", "rnd"):
        fin, fout, fbr = tmp_path / "in.mhit", tmp_path / "out.mhit", tmp_path / "brute.mhit"
        F.write_hits(fin, g[f"{kind}_rays"], g[f"{kind}_hits"])
        run(hostsim, sc, ASSETS, "trace", fin, fout)
        run(hostsim, sc, ASSETS, "trace", fin, fbr, "brute")
        res = F.read_hits(fout)
        ref, out, br = g[f"{kind}_hits"], res["hits"], F.read_hits(fbr)["hits"]
        if f"{kind}_probes" in g.files and g[f"{kind}_probes"].size:          # medium boundary probes recorded from the reference
            want, got = g[f"{kind}_probes"], res["probes"]
            assert got.shape == want.shape
            assert (got["hit1"] == want["hit1"]).all() and (got["hit2"] == want["hit2"]).all()
            b1, b2 = want["hit1"] == 1, want["hit2"] == 1
            assert (bits(got["t1"])[b1] == bits(want["t1"])[b1]).all() and (bits(got["t2"])[b2] == bits(want["t2"])[b2]).all()
        b = ref["hit"] == 1
        assert (out["hit"] == ref["hit"]).all() and (bits(out["t"])[b] == bits(ref["t"])[b]).all()
        for k in ("leaf_type", "leaf_idx", "mat_type", "mat_idx", "front_face"):
            assert (out[k][b] == ref[k][b]).all(), k
        assert (bits(out["p"])[b] == bits(ref["p"])[b]).all() and (bits(out["normal"])[b] == bits(ref["normal"])[b]).all()
        assert (out["hit"] == br["hit"]).all() and (bits(out["t"]) == bits(br["t"])).all() and (out["leaf_idx"] == br["leaf_idx"]).all()


@pytest.mark.parametrize("sc,w,spp", [(1, 48, 16), (4, 48, 9), (6, 40, 16), (7, 40, 16), (9, 32, 9)])
def test_frames_match_oracle_same_stream(hostsim, earth, sc, w, spp, tmp_path):
    out = tmp_path / "f.mimg"
    run(hostsim, sc, ASSETS, "render", w, spp, 0, 99, out)
    mine = F.read_mimg(out)
    osc = O.OracleScene(golden_scene_path(sc, str(tmp_path)), earth)
    osc.override(width=w, spp=spp)
    hdr, _, _ = osc.render(seed=99, want_rgba8=False)
    assert (mine[..., 3] != hdr[..., 3]).mean() <= 0.003
    ok = (mine[..., 3] == 0) & (hdr[..., 3] == 0) & np.isfinite(mine[..., :3]).all(-1) & np.isfinite(hdr[..., :3]).all(-1)
    rel = np.abs(mine[..., :3][ok] - hdr[..., :3][ok]).max(-1) / (np.abs(hdr[..., :3][ok]).max(-1) + 0.16)
    assert (rel > 1e-3).mean() <= 0.05


def test_bvh_builder_invariants(hostsim, tmp_path):
    """every leaf appears exactly once, children boxes lie inside the node they hang from, leaves are type-homogeneous"""
    txt = run(hostsim, 8, ASSETS, "checkbvh")
    assert "bvh ok" in txt, txt
    txt = run(hostsim, 1, ASSETS, "checkbvh")
    assert "bvh ok" in txt, txt
    txt = run(hostsim, "field:40", ASSETS, "checkbvh")
    assert "bvh ok" in txt, txt


def test_parallel_bvh_build_gives_the_serial_tree(hostsim, tmp_path):
    """subtrees are built on several threads above 16 k primitives; the tree must not depend on the thread count"""
    import re
    outs = []
    for threads in ("1", "8"):
        env = dict(os.environ, MORT_BUILD_THREADS=threads)
        p = subprocess.run([hostsim, "field:100", ASSETS, "checkbvh"], check=True, capture_output=True, text=True, env=env)
        assert "bvh ok" in p.stderr and "tree hash" in p.stderr
        outs.append(re.sub(r"build [0-9.]+ ms", "build X ms", p.stderr))
    assert outs[0] == outs[1]


def test_edited_spheres_flatten_and_trace_like_brute_force(hostsim, tmp_path):
    """mort_update_sphere's host side (Scene::update_sphere): spheres of scene 1 — whose world is a reference bvh — are moved far outside
    the boxes that bvh holds for them; the edited scene must still flatten (an edited scene has left the reference's build behind)
    and the new tree must agree with brute force and with the oracle working from the edited arrays"""
    import oracle_binding as O
    from conftest import bits
    from mort_b200 import formats as F
    env = dict(os.environ, MORT_EDIT_SPHERES="120")
    dump = tmp_path / "edited.mscn"
    subprocess.run([hostsim, "1", ASSETS, "dump", str(dump)], check=True, capture_output=True, env=env)
    rng = np.random.default_rng(3)
    n = 3000
    o, tgt = rng.uniform(-12, 12, (n, 3)), rng.uniform(-10, 10, (n, 3)); o[:, 1] = np.abs(o[:, 1]) + 0.3; tgt[:, 1] = np.abs(tgt[:, 1]) * 0.3
    rays = np.concatenate([o, tgt - o, rng.random((n, 1))], 1).astype(np.float32)
    fin, fout, fbr = tmp_path / "in.mhit", tmp_path / "out.mhit", tmp_path / "br.mhit"
    F.write_hits(fin, rays, np.zeros(n, dtype=F.hit_dt))
    p = subprocess.run([hostsim, "1", ASSETS, "trace", str(fin), str(fout)], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    subprocess.run([hostsim, "1", ASSETS, "trace", str(fin), str(fbr), "brute"], check=True, capture_output=True, env=env)
    out, br = F.read_hits(fout)["hits"], F.read_hits(fbr)["hits"]
    assert (out["hit"] == br["hit"]).all() and (bits(out["t"]) == bits(br["t"])).all() and (out["leaf_idx"] == br["leaf_idx"]).all()
    # unedited scene: the same rays must hit something else somewhere (the edit is visible)
    subprocess.run([hostsim, "1", ASSETS, "trace", str(fin), str(tmp_path / "orig.mhit")], check=True, capture_output=True)
    assert (bits(F.read_hits(tmp_path / "orig.mhit")["hits"]["t"]) != bits(out["t"])).mean() > 0.01
