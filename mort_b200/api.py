"""ctypes host layer over libmort_b200.so (include/mort_b200.h).

Mirrors the reference's scene-builder / camera / render surface (world.cuh:27-102, camera.cuh:12-84,
mort.cu:633-689) one call per C-ABI entry.  There is no CPU rendering path: if the CUDA library is missing
or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import formats as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MORT_B200_LIB") or os.path.join(_HERE, "libmort_b200.so")     # the override is for A/B builds of the same library (experiments)
ASSET_DIR = os.path.join(_HERE, "assets")

MODE_MEGAKERNEL, MODE_WAVEFRONT, MODE_POOL, MODE_AUTO = 0, 1, 2, 3
TRACE_BVH, TRACE_BRUTE_FORCE = 0, 1


class MortError(RuntimeError):
    pass


class Handle(C.Structure):
    _fields_ = [("type", C.c_int32), ("idx", C.c_int32)]

    def __iter__(self):
        return iter((self.type, self.idx))

    def __repr__(self):
        return f"Handle({self.type},{self.idx})"


class CameraDesc(C.Structure):
    _fields_ = [("aspect_ratio", C.c_float), ("image_width", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("bounce_limit", C.c_int32), ("vfov", C.c_int32), ("background", C.c_float * 3),
                ("lookfrom", C.c_float * 3), ("lookat", C.c_float * 3), ("vup", C.c_float * 3),
                ("defocus_angle", C.c_float), ("focus_dist", C.c_float),
                ("light_obj_type", C.c_int32), ("light_obj_idx", C.c_int32)]


class RenderOpts(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("frame", C.c_uint32), ("mode", C.c_int32), ("sample_mod", C.c_int32),
                ("sample_rem", C.c_int32), ("stage_nodes", C.c_int32), ("threads_per_block", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("wavefront_paths", C.c_int32), ("exact_accum", C.c_int32), ("tile_mod", C.c_int32), ("tile_rem", C.c_int32), ("accumulate", C.c_int32), ("pool_paths", C.c_int32), ("pool_refill", C.c_int32), ("n_frames", C.c_int32), ("tile_rows", C.c_int32), ("pool_flags", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("sqrt_spp", C.c_int32), ("bounce_limit", C.c_int32),
                ("n_leaves", C.c_int32), ("n_spheres", C.c_int32), ("n_quads", C.c_int32), ("n_nodes", C.c_int32),
                ("bvh_depth", C.c_int32), ("n_media", C.c_int32), ("n_instances", C.c_int32), ("n_materials", C.c_int32),
                ("n_textures", C.c_int32),
                ("sah_cost", C.c_double), ("build_ms", C.c_double), ("upload_ms", C.c_double), ("last_render_ms", C.c_double),
                ("last_segments", C.c_uint64), ("last_samples", C.c_uint64), ("last_kernel_launches", C.c_uint64),
                ("sm_count", C.c_int32), ("staged_nodes", C.c_int32), ("threads_per_block", C.c_int32), ("blocks_per_sm", C.c_int32),
                ("regs_per_thread", C.c_int32), ("reserved0", C.c_int32), ("device_bytes", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BuildOpts(C.Structure):
    _fields_ = [("builder", C.c_int32), ("max_leaf", C.c_int32), ("k_trav", C.c_float), ("host_threads", C.c_int32),
                ("gpu_small", C.c_int32), ("gpu_flags", C.c_int32), ("motion_bounds", C.c_int32), ("reserved", C.c_int32)]


class BuildInfo(C.Structure):
    _fields_ = [("built_on_gpu", C.c_int32), ("gpu_levels", C.c_int32), ("gpu_small_subtrees", C.c_int32), ("bvh2_nodes", C.c_int32),
                ("motion_nodes", C.c_int32), ("refits", C.c_int32), ("reserved", C.c_int32 * 2),
                ("flatten_ms", C.c_double), ("build_ms", C.c_double), ("gpu_stream_ms", C.c_double), ("refit_ms", C.c_double),
                ("gpu_workspace_bytes", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


BUILD_AUTO, BUILD_HOST, BUILD_GPU = 0, 1, 2


class GroupStats(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("split", C.c_int32), ("kernel_ms_max", C.c_double), ("kernel_ms_min", C.c_double),
                ("collective_ms", C.c_double), ("collective_bytes", C.c_uint64), ("segments", C.c_uint64), ("samples", C.c_uint64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SPLIT_SAMPLE, SPLIT_TILE = 0, 1

# every symbol include/mort_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "mort_create", "mort_destroy", "mort_last_error", "mort_set_stream",
    "mort_build_scene", "mort_build_sphere_field", "mort_load_scene", "mort_dump_scene", "mort_load_scene_text", "mort_dump_scene_text", "mort_clear_scene",
    "mort_add_solid", "mort_add_checker", "mort_add_image", "mort_add_noise",
    "mort_add_lambertian", "mort_add_metal", "mort_add_dielectric", "mort_add_diffuse_light", "mort_add_isotropic",
    "mort_add_sphere", "mort_add_moving_sphere", "mort_add_quad", "mort_add_translate", "mort_add_rotate_y",
    "mort_add_constant_medium", "mort_add_list", "mort_list_add", "mort_add_bvh", "mort_add_box", "mort_add_rotated_box",
    "mort_host_rand",
    "mort_get_camera", "mort_set_camera", "mort_override_camera", "mort_get_camera_record",
    "mort_commit", "mort_set_build_opts", "mort_get_build_info", "mort_update_sphere", "mort_refit", "mort_default_render_opts", "mort_render_device", "mort_resolve_exact_device", "mort_tonemap_device", "mort_render",
    "mort_accumulate_exact_device", "mort_scene_fingerprint", "mort_save_checkpoint", "mort_load_checkpoint",
    "mort_render_progressive", "mort_reset_progressive",
    "mort_trace", "mort_get_stats", "mort_write_image", "mort_write_pfm",
    "mort_comm_unique_id", "mort_comm_attach", "mort_comm_detach", "mort_comm_reduce_exact",
    "mort_group_create", "mort_group_destroy", "mort_group_size", "mort_group_ctx", "mort_group_last_error", "mort_group_render", "mort_group_get_stats",
]

_lib = None


def load_library():
    """dlopen libmort_b200.so; raises MortError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MortError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(mort_b200 has no CPU rendering path)")
    L = C.CDLL(LIB_PATH)
    P, H, I, Fl = C.c_void_p, Handle, C.c_int, C.c_float
    HP, F3 = C.POINTER(Handle), C.POINTER(C.c_float)
    sig = {
        "mort_create": [I, C.POINTER(P)], "mort_destroy": [P], "mort_set_stream": [P, P],
        "mort_build_scene": [P, I, C.c_char_p], "mort_build_sphere_field": [P, I, C.c_uint64, I],
        "mort_load_scene": [P, C.c_char_p, C.c_char_p], "mort_dump_scene": [P, C.c_char_p], "mort_clear_scene": [P],
        "mort_load_scene_text": [P, C.c_char_p, C.c_char_p], "mort_dump_scene_text": [P, C.c_char_p],
        "mort_add_solid": [P, Fl, Fl, Fl, HP], "mort_add_checker": [P, Fl, H, H, HP], "mort_add_image": [P, P, I, I, HP],
        "mort_add_noise": [P, Fl, HP], "mort_add_lambertian": [P, H, HP], "mort_add_metal": [P, Fl, Fl, Fl, Fl, HP],
        "mort_add_dielectric": [P, Fl, HP], "mort_add_diffuse_light": [P, H, HP], "mort_add_isotropic": [P, H, HP],
        "mort_add_sphere": [P, F3, Fl, H, I, HP], "mort_add_moving_sphere": [P, F3, F3, Fl, H, I, HP],
        "mort_add_quad": [P, F3, F3, F3, H, I, HP], "mort_add_translate": [P, H, F3, I, HP], "mort_add_rotate_y": [P, H, Fl, I, HP],
        "mort_add_constant_medium": [P, H, Fl, H, I, HP], "mort_add_list": [P, I, HP], "mort_list_add": [P, H, H],
        "mort_add_bvh": [P, H, I, HP], "mort_add_box": [P, F3, F3, H], "mort_add_rotated_box": [P, F3, F3, Fl, H, HP],
        "mort_host_rand": [P],
        "mort_get_camera": [P, C.POINTER(CameraDesc)], "mort_set_camera": [P, C.POINTER(CameraDesc)],
        "mort_override_camera": [P, I, Fl, I, I], "mort_get_camera_record": [P, P],
        "mort_commit": [P], "mort_default_render_opts": [C.POINTER(RenderOpts)],
        "mort_set_build_opts": [P, C.POINTER(BuildOpts)], "mort_get_build_info": [P, C.POINTER(BuildInfo)],
        "mort_update_sphere": [P, H, F3, F3, Fl], "mort_refit": [P],
        "mort_render_device": [P, C.POINTER(RenderOpts), P], "mort_resolve_exact_device": [P, P, P], "mort_tonemap_device": [P, P, I, P],
        "mort_accumulate_exact_device": [P, P, P], "mort_scene_fingerprint": [P, C.POINTER(C.c_uint64)],
        "mort_save_checkpoint": [P, C.c_char_p, P, C.c_uint32, C.c_uint32],
        "mort_load_checkpoint": [P, C.c_char_p, P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
        "mort_render_progressive": [P, C.POINTER(RenderOpts), I, C.c_char_p, I, P, P, C.POINTER(C.c_uint32)], "mort_reset_progressive": [P],
        "mort_render": [P, C.POINTER(RenderOpts), P, P], "mort_trace": [P, P, I, P, P, I], "mort_get_stats": [P, C.POINTER(Stats)],
    }
    sig.update({"mort_comm_unique_id": [P], "mort_comm_attach": [P, P, I, I], "mort_comm_detach": [P], "mort_comm_reduce_exact": [P, P, I],
                "mort_group_create": [I, C.POINTER(C.c_int), C.POINTER(P)], "mort_group_destroy": [P], "mort_group_size": [P],
                "mort_group_render": [P, C.POINTER(RenderOpts), I, P, P], "mort_group_get_stats": [P, C.POINTER(GroupStats)]})
    sig.update({"mort_write_image": [C.c_char_p, P, I, I], "mort_write_pfm": [C.c_char_p, P, I, I, Fl]})
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = None if name == "mort_default_render_opts" else C.c_int
    L.mort_last_error.argtypes = [P]
    L.mort_last_error.restype = C.c_char_p
    L.mort_group_last_error.argtypes = [P]
    L.mort_group_last_error.restype = C.c_char_p
    L.mort_group_ctx.argtypes = [P, I]
    L.mort_group_ctx.restype = P
    _lib = L
    return L


def write_image(path, rgba8: np.ndarray):
    """.ppm / .png from a bottom-up (H, W, 4) uint8 frame (mort_write_image; no device involved)."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    rc = load_library().mort_write_image(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0])
    if rc != 0:
        raise MortError(f"mort_write_image({path}) failed with {rc}")


def write_pfm(path, accum: np.ndarray, scale: float):
    a = np.ascontiguousarray(accum, dtype=np.float32)
    rc = load_library().mort_write_pfm(str(path).encode(), a.ctypes.data, a.shape[1], a.shape[0], float(scale))
    if rc != 0:
        raise MortError(f"mort_write_pfm({path}) failed with {rc}")


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


@dataclass
class Frame:
    """One rendered frame: `accum` (H, W, 4) float32 rows bottom-up — xyz = sum of sample colours, w = #NaN
    samples — and `rgba8` (H, W, 4) uint8 exactly as the reference's frame buffer (camera.cuh:194-207)."""
    accum: np.ndarray | None
    rgba8: np.ndarray | None
    stats: dict


class Renderer:
    """Owns one mort_ctx on one CUDA device."""

    def __init__(self, device: int = 0, _borrowed=None):
        self._L = load_library()
        self._owned = _borrowed is None
        if _borrowed is not None:                      # a context owned by a Group
            self._h = C.c_void_p(_borrowed)
            self.device = device
            return
        h = C.c_void_p()
        rc = self._L.mort_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise MortError(f"mort_create(device={device}) failed with {rc}: a CUDA device is required (no CPU fallback)")
        self._h = h
        self.device = device

    # -- plumbing --
    def _ck(self, rc):
        if rc != 0:
            raise MortError(f"{self._L.mort_last_error(self._h).decode()} (status {rc})")

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._L.mort_destroy(self._h)
            self._h = None

    # -- multi-process collective (one process per GPU) --
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if load_library().mort_comm_unique_id(buf) != 0:
            raise MortError("mort_comm_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def comm_attach(self, id128: bytes, world: int, rank: int):
        self._ck(self._L.mort_comm_attach(self._h, C.c_char_p(id128), int(world), int(rank)))

    def comm_detach(self):
        self._ck(self._L.mort_comm_detach(self._h))

    def comm_reduce_exact(self, d_exact_ptr: int, root: int = 0):
        """One NCCL uint64 SUM reduce of the (H, W, 4) exact partial frames to `root`, in place, on the context's stream."""
        self._ck(self._L.mort_comm_reduce_exact(self._h, C.c_void_p(d_exact_ptr), int(root)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self._L.mort_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    # -- scenes --
    def build_scene(self, scene_id: int, asset_dir: str = ASSET_DIR):
        self._ck(self._L.mort_build_scene(self._h, int(scene_id), asset_dir.encode()))
        return self

    def build_sphere_field(self, G: int, seed: int = 69420, camera_kind: int = 0):
        self._ck(self._L.mort_build_sphere_field(self._h, int(G), int(seed), int(camera_kind)))
        return self

    def load_scene(self, path, asset_dir: str = ASSET_DIR):
        self._ck(self._L.mort_load_scene(self._h, str(path).encode(), asset_dir.encode()))
        return self

    def load_scene_text(self, path, asset_dir: str = ASSET_DIR):
        """Scene text file: one builder call per statement (grammar in mort_b200/csrc/scene_text.cpp)."""
        self._ck(self._L.mort_load_scene_text(self._h, str(path).encode(), asset_dir.encode()))
        return self

    def dump_scene_text(self, path):
        self._ck(self._L.mort_dump_scene_text(self._h, str(path).encode()))

    def dump_scene(self, path):
        self._ck(self._L.mort_dump_scene(self._h, str(path).encode()))

    def clear_scene(self):
        self._ck(self._L.mort_clear_scene(self._h))

    def _add(self, fn, *args):
        out = Handle()
        self._ck(fn(self._h, *args, C.byref(out)))
        return out

    def add_solid(self, r, g, b): return self._add(self._L.mort_add_solid, r, g, b)
    def add_checker(self, scale, even, odd): return self._add(self._L.mort_add_checker, scale, even, odd)
    def add_image(self, rgb: np.ndarray):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        return self._add(self._L.mort_add_image, rgb.ctypes.data, rgb.shape[1], rgb.shape[0])
    def add_noise(self, scale): return self._add(self._L.mort_add_noise, scale)
    def add_lambertian(self, tex): return self._add(self._L.mort_add_lambertian, tex)
    def add_metal(self, r, g, b, fuzz): return self._add(self._L.mort_add_metal, r, g, b, fuzz)
    def add_dielectric(self, ior): return self._add(self._L.mort_add_dielectric, ior)
    def add_diffuse_light(self, tex): return self._add(self._L.mort_add_diffuse_light, tex)
    def add_isotropic(self, tex): return self._add(self._L.mort_add_isotropic, tex)
    def add_sphere(self, c, r, mat, skip=False): return self._add(self._L.mort_add_sphere, _f3(c), r, mat, int(skip))
    def add_moving_sphere(self, c0, c1, r, mat, skip=False): return self._add(self._L.mort_add_moving_sphere, _f3(c0), _f3(c1), r, mat, int(skip))
    def add_quad(self, Q, u, v, mat, skip=False): return self._add(self._L.mort_add_quad, _f3(Q), _f3(u), _f3(v), mat, int(skip))
    def add_translate(self, obj, offset, skip=False): return self._add(self._L.mort_add_translate, obj, _f3(offset), int(skip))
    def add_rotate_y(self, obj, degrees, skip=False): return self._add(self._L.mort_add_rotate_y, obj, degrees, int(skip))
    def add_constant_medium(self, boundary, density, mat, skip=False): return self._add(self._L.mort_add_constant_medium, boundary, density, mat, int(skip))
    def add_list(self, skip=False): return self._add(self._L.mort_add_list, int(skip))
    def list_add(self, lst, obj): self._ck(self._L.mort_list_add(self._h, lst, obj))
    def add_bvh(self, lst, skip=False): return self._add(self._L.mort_add_bvh, lst, int(skip))
    def add_box(self, a, b, mat): self._ck(self._L.mort_add_box(self._h, _f3(a), _f3(b), mat))
    def add_rotated_box(self, size, translation, degrees, mat): return self._add(self._L.mort_add_rotated_box, _f3(size), _f3(translation), degrees, mat)
    def host_rand(self): return self._L.mort_host_rand(self._h)

    # -- camera --
    def get_camera(self) -> CameraDesc:
        d = CameraDesc()
        self._ck(self._L.mort_get_camera(self._h, C.byref(d)))
        return d

    def set_camera(self, desc: CameraDesc):
        self._ck(self._L.mort_set_camera(self._h, C.byref(desc)))

    def override_camera(self, width=0, aspect=0.0, spp=0, depth=0):
        self._ck(self._L.mort_override_camera(self._h, int(width), float(aspect), int(spp), int(depth)))
        return self

    def camera_record(self):
        c = np.zeros(1, dtype=F.camera_dt)
        self._ck(self._L.mort_get_camera_record(self._h, c.ctypes.data))
        return c[0]

    def commit(self):
        self._ck(self._L.mort_commit(self._h))
        return self

    # -- tree build on the GPU / refit / motion-aware bounds --
    def set_build_opts(self, **kw):
        """builder=BUILD_AUTO|BUILD_HOST|BUILD_GPU, max_leaf, k_trav, host_threads, gpu_small, gpu_flags, motion_bounds; applies to the next commit."""
        o = BuildOpts()
        for k, v in kw.items():
            if not hasattr(o, k):
                raise MortError(f"unknown build option {k}")
            setattr(o, k, v)
        self._ck(self._L.mort_set_build_opts(self._h, C.byref(o)))
        return self

    @property
    def build_info(self) -> dict:
        b = BuildInfo()
        self._ck(self._L.mort_get_build_info(self._h, C.byref(b)))
        return b.asdict()

    def update_sphere(self, handle, center0, center1, radius):
        """Move / resize a sphere of the committed scene (center1 = None: static); call refit() (or commit()) before rendering."""
        self._ck(self._L.mort_update_sphere(self._h, handle, _f3(center0), _f3(center1) if center1 is not None else None, float(radius)))
        return self

    def refit(self):
        self._ck(self._L.mort_refit(self._h))
        return self

    # -- render --
    def opts(self, **kw) -> RenderOpts:
        o = RenderOpts()
        self._L.mort_default_render_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    @property
    def stats(self) -> dict:
        s = Stats()
        self._ck(self._L.mort_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def render(self, want_rgba8=True, want_accum=True, **opts) -> Frame:
        """Host-buffer frame (mort_render): kernels + tone map + device->host copies."""
        st = self.stats
        H, W = st["height"], st["width"]
        rgba = np.empty((H, W, 4), dtype=np.uint8) if want_rgba8 else None
        acc = np.empty((H, W, 4), dtype=np.float32) if want_accum else None
        o = self.opts(**opts)
        self._ck(self._L.mort_render(self._h, C.byref(o), rgba.ctypes.data if rgba is not None else None,
                                     acc.ctypes.data if acc is not None else None))
        return Frame(acc, rgba, self.stats)

    def render_into(self, host_rgba8_ptr: int, host_accum_ptr: int = 0, **opts):
        o = self.opts(**opts)
        self._ck(self._L.mort_render(self._h, C.byref(o), C.c_void_p(host_rgba8_ptr) if host_rgba8_ptr else None,
                                     C.c_void_p(host_accum_ptr) if host_accum_ptr else None))

    def render_device(self, d_accum_ptr: int, **opts):
        """Device-resident frame into a caller-owned float4 buffer (e.g. a torch tensor's data_ptr())."""
        o = self.opts(**opts)
        self._ck(self._L.mort_render_device(self._h, C.byref(o), C.c_void_p(d_accum_ptr)))

    def resolve_exact_device(self, d_exact_ptr: int, d_accum_ptr: int):
        """(H, W, 4) uint64 exact sums (render_device(..., exact_accum=1)) -> (H, W, 4) float32 accumulation image."""
        self._ck(self._L.mort_resolve_exact_device(self._h, C.c_void_p(d_exact_ptr), C.c_void_p(d_accum_ptr)))

    # -- progressive accumulation + checkpoint / resume --
    def accumulate_exact_device(self, d_sum_ptr: int, d_frame_ptr: int):
        """d_sum += d_frame over (H, W, 4) uint64 exact sums (merging partial frames; order never matters)."""
        self._ck(self._L.mort_accumulate_exact_device(self._h, C.c_void_p(d_sum_ptr), C.c_void_p(d_frame_ptr)))

    @property
    def scene_fingerprint(self) -> int:
        v = C.c_uint64(0)
        self._ck(self._L.mort_scene_fingerprint(self._h, C.byref(v)))
        return v.value

    def save_checkpoint(self, path, d_sum_ptr: int, seed: int, frames_done: int):
        self._ck(self._L.mort_save_checkpoint(self._h, os.fsencode(path), C.c_void_p(d_sum_ptr), int(seed), int(frames_done)))

    def load_checkpoint(self, path, d_sum_ptr: int):
        """-> (seed, frames_done); MortError if the file belongs to another scene, camera or frame size."""
        seed, done = C.c_uint32(0), C.c_uint32(0)
        self._ck(self._L.mort_load_checkpoint(self._h, os.fsencode(path), C.c_void_p(d_sum_ptr), C.byref(seed), C.byref(done)))
        return seed.value, done.value

    def render_progressive(self, n_frames: int, checkpoint=None, resume=False, want_rgba8=True, want_accum=True, **opts):
        """Adds n_frames frames to the context's running image (mort_render_progressive).
        -> (Frame, frames_total): Frame.accum holds float SUMS over all frames so far (divide by spp * frames_total)."""
        st = self.stats
        H, W = st["height"], st["width"]
        rgba = np.empty((H, W, 4), dtype=np.uint8) if want_rgba8 else None
        acc = np.empty((H, W, 4), dtype=np.float32) if want_accum else None
        total = C.c_uint32(0)
        o = self.opts(**opts)
        self._ck(self._L.mort_render_progressive(self._h, C.byref(o), int(n_frames), os.fsencode(checkpoint) if checkpoint else None, int(bool(resume)),
                                                 rgba.ctypes.data if rgba is not None else None, acc.ctypes.data if acc is not None else None, C.byref(total)))
        return Frame(acc, rgba, self.stats), total.value

    def reset_progressive(self):
        self._ck(self._L.mort_reset_progressive(self._h))

    def tonemap_device(self, d_accum_ptr: int, samples_per_pixel_total: int, d_rgba8_ptr: int):
        self._ck(self._L.mort_tonemap_device(self._h, C.c_void_p(d_accum_ptr), int(samples_per_pixel_total), C.c_void_p(d_rgba8_ptr)))

    # -- parity hook --
    def trace(self, rays: np.ndarray, brute_force=False, want_probes=True):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 7)
        n = rays.shape[0]
        nm = self.stats["n_media"]
        out = np.zeros(n, dtype=F.hit_dt)
        probes = np.zeros((n, nm), dtype=F.probe_dt)
        self._ck(self._L.mort_trace(self._h, rays.ctypes.data, n, out.ctypes.data,
                                    probes.ctypes.data if (nm and want_probes) else None,
                                    TRACE_BRUTE_FORCE if brute_force else TRACE_BVH))
        return out, probes


class Group:
    """One process driving n GPUs of the box (mort_group_*): a Renderer view per rank for scene building, one render call."""

    def __init__(self, n_devices: int, devices=None):
        self._L = load_library()
        h = C.c_void_p()
        arr = (C.c_int * n_devices)(*devices) if devices is not None else None
        rc = self._L.mort_group_create(int(n_devices), arr, C.byref(h))
        if rc != 0 or not h:
            raise MortError(f"mort_group_create({n_devices}) failed with {rc}: needs {n_devices} CUDA devices and NCCL")
        self._h = h
        self.n = n_devices
        self.ranks = [Renderer(devices[i] if devices is not None else i, _borrowed=self._L.mort_group_ctx(h, i)) for i in range(n_devices)]

    def for_each(self, fn):
        for r in self.ranks:
            fn(r)
        return self

    def render(self, split=SPLIT_SAMPLE, want_rgba8=True, want_accum=True, **opts) -> Frame:
        st = self.ranks[0].stats
        H, W = st["height"], st["width"]
        rgba = np.empty((H, W, 4), dtype=np.uint8) if want_rgba8 else None
        acc = np.empty((H, W, 4), dtype=np.float32) if want_accum else None
        o = self.ranks[0].opts(**opts)
        rc = self._L.mort_group_render(self._h, C.byref(o), int(split), rgba.ctypes.data if rgba is not None else None, acc.ctypes.data if acc is not None else None)
        if rc != 0:
            raise MortError(f"{self._L.mort_group_last_error(self._h).decode()} (status {rc})")
        return Frame(acc, rgba, self.stats)

    @property
    def stats(self) -> dict:
        s = GroupStats()
        self._L.mort_group_get_stats(self._h, C.byref(s))
        return s.asdict()

    def close(self):
        if getattr(self, "_h", None):
            for r in self.ranks:
                r.close()
            self._L.mort_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
