"""Scene text files: one builder call per statement, so scenes need no recompile.

CPU-only, through tests/hostsim (the host builder is the same code libmort_b200.so links):
  * every shipped scene, written out as text from the journal of its builder calls and loaded back, dumps to the
    reference's own scene bytes (tests/golden/scene_N.mscn) — slots, list orders, the reference's BVH, Perlin
    tables and the camera all survive the text form;
  * a hand-written file with aliases and the colour shorthands builds the scene the equivalent calls build;
  * malformed files fail with file:line diagnostics and leave no half-built scene behind.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ASSETS, golden_scene_path

SCENES = list(range(1, 11))


def run(hostsim, *args, ok=True):
    p = subprocess.run([hostsim, *map(str, args)], capture_output=True, text=True)
    if ok:
        assert p.returncode == 0, p.stderr
    return p


@pytest.mark.parametrize("sc", SCENES)
def test_shipped_scene_round_trips_through_text(hostsim, sc, tmp_path):
    txt, out = tmp_path / f"s{sc}.txt", tmp_path / f"s{sc}.mscn"
    run(hostsim, sc, ASSETS, "dumptext", txt)
    run(hostsim, f"text:{txt}", ASSETS, "dump", out)
    assert open(out, "rb").read() == open(golden_scene_path(sc, str(tmp_path)), "rb").read()
    # and the text is stable: dumping the reloaded scene gives the same file
    txt2 = tmp_path / f"s{sc}b.txt"
    run(hostsim, f"text:{txt}", ASSETS, "dumptext", txt2)
    assert open(txt).read() == open(txt2).read()


HAND_WRITTEN = """
# a small room: aliases, colour shorthands, a box helper, an instance chain and a medium
white = lambertian 0.73 0.73 0.73
red   = lambertian 0.65 0.05 0.05          # implicit solid texture
lamp  = light 15 15 15
glass = dielectric 1.5
chk   = checker 0.5 sol0 sol1               # canonical spelling of the first two solids
floor = lambertian chk
quad 0 0 0  4 0 0  0 0 4  floor
quad 0 0 0  0 4 0  0 0 4  red
top = quad 1 3.99 1  2 0 0  0 0 2  lamp
steel = metal 0.8 0.85 0.88 0.05
ball = sphere 3 0.7 2.5 0.7 steel
box 0.2 0 0.2  1.0 0.8 1.0  white
crate = rotated_box 1 1 1  2.5 0 0.5  30 white
fog_shell = sphere 1.2 1.2 2.5 0.6 glass hidden
smoke = isotropic 0.9 0.9 0.9
medium fog_shell 0.8 smoke
lights = list hidden
add lights top
camera width 64 aspect 1 spp 16 depth 8 vfov 40
camera lookfrom 2 2 -6 lookat 2 2 0 vup 0 1 0 background 0 0 0 light lights
"""


def test_hand_written_scene_builds_expected_arrays(hostsim, tmp_path):
    from mort_b200 import formats as F
    txt, out = tmp_path / "room.txt", tmp_path / "room.mscn"
    txt.write_text(HAND_WRITTEN)
    run(hostsim, f"text:{txt}", ASSETS, "dump", out)
    sc = F.read_scene(str(out))
    assert len(sc["spheres"]) == 2 and len(sc["quads"]) == 3 + 6 + 6
    assert len(sc["lambertians"]) == 3 and len(sc["solids"]) == 2 + 1 + 1     # white, red, lamp colour, smoke colour
    assert len(sc["diffuse_lights"]) == 1 and len(sc["dielectrics"]) == 1 and len(sc["metals"]) == 1 and len(sc["isotropics"]) == 1 and len(sc["checkers"]) == 1
    assert len(sc["media"]) == 1 and len(sc["translates"]) == 1 and len(sc["rotates"]) == 1
    assert [len(l["items"]) for l in sc["lists"]] == [6, 1]                    # the crate's sides, then the light list
    assert sc["spheres"]["skip"].tolist() == [0, 1]
    cam = sc["camera"]
    assert int(cam["image_width"]) == 64 and int(cam["image_height"]) == 64 and int(cam["sqrt_spp"]) == 4
    assert (int(cam["light_obj_type"]), int(cam["light_obj_idx"])) == (6, 1)
    # the same scene renders on the CPU build of the product's per-ray code (no NaNs, light reaches the floor)
    img = tmp_path / "room.mimg"
    run(hostsim, f"text:{txt}", ASSETS, "render", 32, 4, 0, 7, img)
    hdr = F.read_mimg(str(img))
    assert np.isfinite(hdr[..., :3]).all() and hdr[..., :3].mean() > 0.01


@pytest.mark.parametrize("body,needle", [
    ("sphere 0 0 0 1 lam0\n", ":1: 'lam0' does not exist yet"),
    ("m = lambertian 1 1 1\nsphere 0 0 0 m\n", ":2: number expected"),
    ("m = lambertian 1 1 1\nsphere 0 0 0 1 m extra\n", ":2: unexpected 'extra'"),
    ("t = solid 1 1 1\nsphere 0 0 0 1 t\n", ":2: 't' is not a material"),
    ("frobnicate 1 2 3\n", ":1: unknown statement 'frobnicate'"),
    ("l = list\ns = add l l\n", ":2: this statement returns no handle"),
    ("camera width 0\n", ":1: camera width"),
    ("image nothere.ppm\n", ":1: cannot read image"),
])
def test_malformed_files_fail_with_line_numbers(hostsim, tmp_path, body, needle):
    txt = tmp_path / "bad.txt"
    txt.write_text(body)
    p = run(hostsim, f"text:{txt}", ASSETS, "dump", tmp_path / "x.mscn", ok=False)
    assert p.returncode != 0 and needle in p.stderr, p.stderr
    assert not os.path.exists(tmp_path / "x.mscn")


def test_scene_loaded_from_a_binary_dump_has_no_text_form(hostsim, tmp_path):
    p = run(hostsim, golden_scene_path(1), ASSETS, "dumptext", tmp_path / "s.txt", ok=False)
    assert p.returncode != 0 and "no text form" in p.stderr


def _random_scene_text(rng, nested_media=0.1):
    """A random but valid scene file: every statement kind, forward references only to things that exist.
    nested_media: how often a wrapper / list / medium may pick ANY object (media included) instead of a plain one."""
    lines, tex, mat, obj, lists = [], [], [], [], []
    f = lambda lo=-5.0, hi=5.0: f"{rng.uniform(lo, hi):.6g}"
    v3 = lambda lo=-5.0, hi=5.0: " ".join(f(lo, hi) for _ in range(3))
    for i in range(rng.integers(2, 5)):
        lines.append(f"t{i} = solid {v3(0, 1)}"); tex.append(f"t{i}")
    lines.append(f"tc = checker {f(0.1, 2)} {rng.choice(tex)} {rng.choice(tex)}"); tex.append("tc")
    if rng.random() < 0.5:
        lines.append(f"tn = noise {f(0.5, 4)} at {int(rng.integers(0, 5000))}"); tex.append("tn")
    kinds = ["lambertian", "metal", "dielectric", "light", "isotropic"]
    for i in range(rng.integers(3, 7)):
        k = kinds[i] if i < 5 else str(rng.choice(kinds))
        if k == "metal":
            lines.append(f"m{i} = metal {v3(0, 1)} {f(0, 0.5)}")
        elif k == "dielectric":
            lines.append(f"m{i} = dielectric {f(1.1, 2.0)}")
        else:
            arg = str(rng.choice(tex)) if rng.random() < 0.6 else v3(0, 1)
            lines.append(f"m{i} = {k} {arg}")
        mat.append(f"m{i}")
    depth, plain = {}, []                      # wrapper nesting per object; objects that are not media (media only exist at top level)
    for i in range(rng.integers(4, 12)):
        m, hid = rng.choice(mat), (" hidden" if rng.random() < 0.3 else "")
        k = rng.integers(0, 6)
        shallow = [o for o in plain if depth[o] < 7] if rng.random() < 1.0 - nested_media else list(obj)   # mostly stay inside what the product supports
        name, d, is_medium = f"o{i}", 0, False
        if k == 0:
            lines.append(f"{name} = sphere {v3()} {f(0.1, 2)} {m}{hid}")
        elif k == 1:
            lines.append(f"{name} = moving_sphere {v3()} {v3()} {f(0.1, 2)} {m}{hid}")
        elif k == 2:
            lines.append(f"{name} = quad {v3()} {f(0.5, 3)} 0 0  0 {f(0.5, 3)} 0 {m}{hid}")
        elif k == 3 and shallow:
            t = str(rng.choice(shallow)); d = depth.get(t, 0) + 1
            lines.append(f"{name} = translate {t} {v3()}{hid}")
        elif k == 4 and shallow:
            t = str(rng.choice(shallow)); d = depth.get(t, 0) + 1
            lines.append(f"{name} = rotate_y {t} {f(-90, 90)}{hid}")
        elif k == 5 and shallow:
            t = str(rng.choice(shallow)); is_medium = True
            lines.append(f"{name} = medium {t} {f(0.01, 2)} {m}{hid}")
        else:
            lines.append(f"{name} = rotated_box {v3(0.5, 2)} {v3()} {f(-45, 45)} {m}"); d = 2
        obj.append(name); depth[name] = d
        if not is_medium:
            plain.append(name)
    if rng.random() < 0.7:
        lines.append(f"box {v3()} {v3()} {rng.choice(mat)}")
    for i in range(rng.integers(1, 3)):
        lines.append(f"l{i} = list" + (" hidden" if rng.random() < 0.5 else ""))
        pool = plain if rng.random() < 1.0 - nested_media else obj
        for o in rng.choice(pool, size=min(len(pool), int(rng.integers(1, 5))), replace=False):
            lines.append(f"add l{i} {o}")
        lists.append(f"l{i}")
    if rng.random() < 0.4:
        lines.append(f"b0 = bvh {lists[0]}")
    if rng.random() < 0.3:
        lines.append(f"host_rand_skip {int(rng.integers(0, 100))}")
    lines.append(f"camera width {int(rng.integers(16, 200))} aspect {f(0.5, 2)} spp {int(rng.integers(1, 64))} depth {int(rng.integers(0, 30))} vfov {int(rng.integers(10, 90))}")
    lines.append(f"camera lookfrom {v3()} lookat {v3()} vup 0 1 0 background {v3(0, 1)} defocus_angle {f(0, 2)} focus_dist {f(1, 20)}")
    lines.append("camera light " + (str(rng.choice(obj + lists)) if rng.random() < 0.6 else "none"))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(12))
def test_random_scene_files_round_trip(hostsim, seed, tmp_path):
    """text -> arrays -> journal text -> arrays: same bytes, and the journal text is a fixed point"""
    rng = np.random.default_rng(1000 + seed)
    t0, a, t1, b, t2 = (tmp_path / n for n in ("s0.txt", "a.mscn", "s1.txt", "b.mscn", "s2.txt"))
    t0.write_text(_random_scene_text(rng))
    run(hostsim, f"text:{t0}", ASSETS, "dump", a)
    run(hostsim, f"text:{t0}", ASSETS, "dumptext", t1)
    run(hostsim, f"text:{t1}", ASSETS, "dump", b)
    run(hostsim, f"text:{t1}", ASSETS, "dumptext", t2)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert open(t1).read() == open(t2).read()


@pytest.mark.parametrize("sc", [2, 6, 7])
def test_example_files_are_the_shipped_scenes(hostsim, sc, tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "e.mscn"
    run(hostsim, "text:" + os.path.join(root, "examples", f"scene_{sc}.txt"), ASSETS, "dump", out)
    assert open(out, "rb").read() == open(golden_scene_path(sc, str(tmp_path)), "rb").read()
    assert open(os.path.join(root, "examples", "room_handwritten.txt")).read().strip() == HAND_WRITTEN.strip()
