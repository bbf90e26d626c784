"""ctypes binding of oracle/_ref/liboracle.so — the CPU restatement of the reference.  TEST
INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mort_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/liboracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "mort_oracle.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_scene_load.restype = C.c_void_p
        L.oracle_scene_load.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_scene_free.argtypes = [C.c_void_p]
        L.oracle_scene_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_scene_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_camera_override.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int]
        L.oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_camera_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_stream_uniforms.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class OracleScene:
    def __init__(self, mscn_path, image_rgb: np.ndarray | None = None):
        L = lib()
        if image_rgb is not None:
            image_rgb = np.ascontiguousarray(image_rgb, dtype=np.uint8)
            h, w = image_rgb.shape[:2]
            self._h = L.oracle_scene_load(str(mscn_path).encode(), image_rgb.ctypes.data, w, h)
        else:
            self._h = L.oracle_scene_load(str(mscn_path).encode(), None, 0, 0)
        if not self._h:
            raise RuntimeError(f"oracle could not load {mscn_path}")

    def close(self):
        if self._h:
            lib().oracle_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def camera(self):
        c = np.zeros(1, dtype=F.camera_dt)
        lib().oracle_scene_camera(self._h, c.ctypes.data)
        return c[0]

    @property
    def counts(self):
        h = np.zeros(1, dtype=F.header_dt)
        lib().oracle_scene_counts(self._h, h.ctypes.data)
        return h[0]

    def override(self, width=0, aspect=0.0, spp=0, depth=0):
        lib().oracle_camera_override(self._h, int(width), float(aspect), int(spp), int(depth))

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        n = rays.shape[0]
        nm = int(self.counts["n_medium"])
        out = np.zeros(n, dtype=F.hit_dt)
        probes = np.zeros((n, nm), dtype=F.probe_dt)
        lib().oracle_trace(self._h, rays.ctypes.data, n, out.ctypes.data, probes.ctypes.data if nm else None)
        return out, probes

    def render(self, seed=69420, frame=0, sj_mod=1, sj_rem=0, threads=None, want_rgba8=True):
        cam = self.camera
        W, H = int(cam["image_width"]), int(cam["image_height"])
        hdr = np.zeros((H, W, 4), dtype=np.float32)
        rgba = np.zeros((H, W, 4), dtype=np.uint8) if want_rgba8 else None
        counters = np.zeros(2, dtype=np.uint64)
        threads = threads or (os.cpu_count() or 1)
        lib().oracle_render(self._h, seed, frame, sj_mod, sj_rem, threads, hdr.ctypes.data,
                            rgba.ctypes.data if rgba is not None else None, counters.ctypes.data)
        return hdr, rgba, {"segments": int(counters[0]), "samples": int(counters[1])}

    def camera_ray(self, x, y, s_i, s_j, seed=69420, frame=0):
        out = np.zeros(7, dtype=np.float32)
        lib().oracle_camera_ray(self._h, seed, frame, x, y, s_i, s_j, out.ctypes.data)
        return out


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def stream_uniforms(seed, frame, pixel, sample, n):
    o = np.zeros(n, dtype=np.float32)
    lib().oracle_stream_uniforms(seed, frame, pixel, sample, n, o.ctypes.data)
    return o
