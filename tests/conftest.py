import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "mort_b200", "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_scene_path(sc, tmpdir=None):
    """Path of the reference's scene dump; scene 9 is scene 8's arrays + its own camera record."""
    if sc != 9:
        return os.path.join(GOLDEN, f"scene_{sc}.mscn")
    from mort_b200 import formats as F
    raw8 = open(os.path.join(GOLDEN, "scene_8.mscn"), "rb").read()
    cam9 = open(os.path.join(GOLDEN, "camera_9.bin"), "rb").read()
    out = os.path.join(tmpdir or "/tmp", "mort_golden_scene_9.mscn")
    open(out, "wb").write(raw8[:-F.camera_dt.itemsize] + cam9)
    return out


@pytest.fixture(scope="session")
def earth():
    from mort_b200 import formats as F
    return F.read_ppm(os.path.join(ASSETS, "earthmap.ppm"))


@pytest.fixture(scope="session")
def hostsim():
    """Test-only host build of the product's per-ray code (tests/hostsim)."""
    out = os.path.join(ROOT, "tests", "hostsim", "hostsim.bin")
    src = [os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")] + [os.path.join(ROOT, "mort_b200", "csrc", f) for f in
                                                                      ("scene.cpp", "scenes.cpp", "scene_text.cpp", "flatten.cpp", "bvh_build.cpp")]
    deps = src + [os.path.join(ROOT, "mort_b200", "csrc", f) for f in ("rt_core.cuh", "device_types.h", "flatten.hpp", "scene.hpp", "bvh_sah.hpp")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-DMORT_GENERAL_MEDIA", "-I" + os.path.join(ROOT, "mort_b200", "csrc"),
                               "-I" + os.path.join(ROOT, "include")] + src + ["-o", out, "-lpthread"])
    return out


@pytest.fixture(scope="session")
def buildsim():
    """Test-only serial run of the GPU tree builder's and the refit pass's per-thread bodies (tests/hostsim/buildsim.cpp)."""
    out = os.path.join(ROOT, "tests", "hostsim", "buildsim.bin")
    src = [os.path.join(ROOT, "tests", "hostsim", "buildsim.cpp")] + [os.path.join(ROOT, "mort_b200", "csrc", f) for f in
                                                                       ("scene.cpp", "scenes.cpp", "scene_text.cpp", "flatten.cpp", "bvh_build.cpp")]
    deps = src + [os.path.join(ROOT, "mort_b200", "csrc", f) for f in ("bvh_sah.hpp", "gpu_build_core.cuh", "gpu_build_driver.hpp", "refit_core.cuh", "device_types.h", "flatten.hpp", "scene.hpp")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "mort_b200", "csrc"),
                               "-I" + os.path.join(ROOT, "include")] + src + ["-o", out, "-lpthread"])
    return out


def luminance(rgb):
    return 0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def trimmed_rmse(a, b, trim=0.005):
    """Per-channel RMSE of two (N, 3) images without the `trim` share of pixels with the largest squared difference
    (scripts/curate_golden.py computes the fixtures' rmse_ab_trim the same way)."""
    d2 = ((a - b) ** 2).sum(axis=1)
    keep = d2 <= np.quantile(d2, 1.0 - trim)
    return np.sqrt(((a[keep] - b[keep]) ** 2).mean(axis=0))


def bin8(img, ok):
    """8x8-binned means of an (H, W, 3) image over the pixels `ok` marks valid -> (means, valid-pixel counts) (as scripts/curate_golden.py)."""
    H8, W8 = img.shape[0] // 8 * 8, img.shape[1] // 8 * 8
    v = np.where(ok[:H8, :W8, None], img[:H8, :W8, :3], 0.0).reshape(H8 // 8, 8, W8 // 8, 8, 3).sum(axis=(1, 3))
    n = ok[:H8, :W8].reshape(H8 // 8, 8, W8 // 8, 8).sum(axis=(1, 3))
    return v / np.maximum(n, 1)[..., None], n
