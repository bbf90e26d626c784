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
