"""The image files the CLI writes against the REFERENCE's own frame buffer.

`mort --out x.ppm` must hold the reference's 8-bit frame (camera.cuh:194-207: mean, NaN flush, gamma 2, 256 * clamp(.., 0.999))
with the rows flipped to the file's top-down order; `--hdr x.pfm` the linear radiance in the frame's own bottom-up order.
tests/golden/conv_<n>.npz holds `rgba8_a`: the bytes the unmodified reference wrote for its 4096-spp frame (bottom-up, as it hands
them to glDrawPixels).  The two renderers use different random streams, so the comparison is statistical — but the ROW ORDER is
not: against the flipped file the mean absolute difference is a few code values, against the unflipped one it is tens."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _read_ppm(path):
    raw = open(path, "rb").read()
    assert raw[:2] == b"P6"
    parts = raw.split(b"\n", 3)
    w, h = (int(x) for x in parts[1].split())
    assert parts[2] == b"255"
    return np.frombuffer(parts[3], dtype=np.uint8).reshape(h, w, 3)


def _read_pfm(path):
    raw = open(path, "rb").read()
    parts = raw.split(b"\n", 3)
    assert parts[0] == b"PF" and float(parts[2]) < 0          # little endian
    w, h = (int(x) for x in parts[1].split())
    return np.frombuffer(parts[3], dtype="<f4").reshape(h, w, 3)


@pytest.mark.parametrize("sc", [1, 6])
def test_ppm_and_pfm_match_the_reference_frame_buffer(sc, tmp_path):
    g = np.load(f"{GOLDEN}/conv_{sc}.npz")
    ref8 = g["rgba8_a"]                                     # bottom-up, RGBA
    H, W = ref8.shape[:2]
    ppm, pfm = str(tmp_path / "f.ppm"), str(tmp_path / "f.pfm")
    subprocess.run([os.path.join(ROOT, "mort_b200", "mort"), str(sc), "--width", str(W), "--spp", "1024", "--out", ppm, "--hdr", pfm],
                   check=True, capture_output=True, cwd=ROOT)
    img = _read_ppm(ppm)
    assert img.shape == (H, W, 3) and (ref8[..., 3] == 255).all()
    ok = g["nan_a"] == 0                                    # pixels the reference flushed from NaN are black there, by chance here
    top_down_ref = ref8[::-1, :, :3].astype(np.int32)
    d_flipped = np.abs(img.astype(np.int32) - top_down_ref)[ok[::-1]].mean()
    d_unflipped = np.abs(img.astype(np.int32) - ref8[..., :3].astype(np.int32))[ok].mean()
    assert d_flipped < 4.0, f"scene {sc}: PPM differs from the reference's flipped frame by {d_flipped:.2f} code values on average"
    assert d_unflipped > 3 * d_flipped, "row order: the file must be top-down, the frame bottom-up"
    # PFM: linear radiance, bottom-up like the frame; the 8-bit file is its tone-mapped, flipped image
    hdr = _read_pfm(pfm)
    assert hdr.shape == (H, W, 3)
    mean = np.where(np.isnan(hdr), 0.0, hdr)
    tone = (256.0 * np.clip(np.sqrt(np.maximum(mean, 0.0)), 0.0, 0.999)).astype(np.int32)
    assert np.abs(tone[::-1] - img.astype(np.int32)).max() <= 1
    ref_mean = g["mean_a"].astype(np.float32)
    rel = abs(float(mean[ok].mean()) - float(ref_mean[ok].mean())) / float(ref_mean[ok].mean())
    assert rel < 0.01, f"scene {sc}: PFM mean radiance off the reference's by {rel:.4f}"
