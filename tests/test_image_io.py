"""Image files: mort_write_image / mort_write_pfm take host buffers, so this runs without a GPU.
The PNG is decoded here with zlib + the PNG chunk rules (CRC of every chunk checked) and compared with the frame."""
import struct
import zlib

import numpy as np
import pytest


def _decode_png(raw):
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(raw):
        n, = struct.unpack(">I", raw[pos:pos + 4]); typ = raw[pos + 4:pos + 8]; data = raw[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) & 0xFFFFFFFF == crc, f"bad CRC in {typ}"
        chunks.append((typ, data)); pos += 12 + n
    assert [c[0] for c in chunks][0] == b"IHDR" and chunks[-1][0] == b"IEND"
    w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, flt, inter) == (8, 2, 0, 0, 0)
    px = zlib.decompress(b"".join(d for t, d in chunks if t == b"IDAT"))
    rows = np.frombuffer(px, dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3)


@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (225, 400), (300, 173)])
def test_png_ppm_pfm_round_trip(tmp_path, shape):
    from mort_b200 import api, formats as F
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    frame = rng.integers(0, 256, (H, W, 4), dtype=np.uint8); frame[..., 3] = 255
    api.write_image(tmp_path / "a.png", frame)
    api.write_image(tmp_path / "a.ppm", frame)
    png = _decode_png(open(tmp_path / "a.png", "rb").read())
    assert np.array_equal(png, frame[::-1, :, :3]), "PNG rows must be the frame's rows top-down"
    assert np.array_equal(F.read_ppm(str(tmp_path / "a.ppm")), frame[::-1, :, :3])
    acc = rng.random((H, W, 4)).astype(np.float32) * 40; acc[0, 0, 1] = np.nan
    api.write_pfm(tmp_path / "a.pfm", acc, 0.25)
    raw = open(tmp_path / "a.pfm", "rb").read()
    head = f"PF\n{W} {H}\n-1.0\n".encode()
    assert raw.startswith(head)
    body = np.frombuffer(raw[len(head):], dtype="<f4").reshape(H, W, 3)
    assert np.array_equal(body, acc[..., :3] * np.float32(0.25), equal_nan=True)


def test_writer_rejects_unknown_extension(tmp_path):
    from mort_b200 import api
    with pytest.raises(api.MortError):
        api.write_image(tmp_path / "a.jpg", np.zeros((2, 2, 4), dtype=np.uint8))
