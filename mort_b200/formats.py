"""numpy views of the on-disk layouts in include/mort_scene_format.h (scene dumps, hit records,
harness images).  Pure data plumbing used by the Python host layer and the tests."""
from __future__ import annotations

import numpy as np

MSCN_MAGIC = 0x4E43534D
MHIT_MAGIC = 0x5449484D
MIMG_MAGIC = 0x474D494D

OBJ_SPHERE, OBJ_QUAD, OBJ_TRANSLATE, OBJ_ROTATE_Y, OBJ_CONSTANT_MEDIUM, OBJ_HITTABLE_LIST, OBJ_BVH = range(1, 8)
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC = range(1, 6)
TEX_SOLID, TEX_CHECKER, TEX_IMAGE, TEX_NOISE = range(1, 5)

f4, i4, u4 = np.dtype("<f4"), np.dtype("<i4"), np.dtype("<u4")

header_dt = np.dtype([("magic", u4), ("version", u4),
                      ("n_sphere", i4), ("n_quad", i4), ("n_translate", i4), ("n_rotate_y", i4), ("n_medium", i4),
                      ("n_list", i4), ("n_bvh", i4),
                      ("n_lambertian", i4), ("n_metal", i4), ("n_dielectric", i4), ("n_diffuse_light", i4), ("n_isotropic", i4),
                      ("n_solid", i4), ("n_checker", i4), ("n_image", i4), ("n_noise", i4),
                      ("bvh_mode", i4), ("reserved", i4, 7)])
sphere_dt = np.dtype([("center", f4, 3), ("radius", f4), ("moves", i4), ("center_vec", f4, 3),
                      ("mat_type", i4), ("mat_idx", i4), ("skip", i4), ("bbox", f4, 6)])
quad_dt = np.dtype([("Q", f4, 3), ("u", f4, 3), ("v", f4, 3), ("normal", f4, 3), ("w", f4, 3), ("D", f4), ("area", f4),
                    ("mat_type", i4), ("mat_idx", i4), ("skip", i4), ("bbox", f4, 6)])
translate_dt = np.dtype([("obj_type", i4), ("obj_idx", i4), ("offset", f4, 3), ("skip", i4)])
rotate_dt = np.dtype([("obj_type", i4), ("obj_idx", i4), ("sin_theta", f4), ("cos_theta", f4), ("skip", i4)])
medium_dt = np.dtype([("obj_type", i4), ("obj_idx", i4), ("neg_inv_density", "<f8"), ("mat_type", i4), ("mat_idx", i4),
                      ("skip", i4), ("pad", i4)])
bvh_node_dt = np.dtype([("left_type", i4), ("left_idx", i4), ("right_type", i4), ("right_idx", i4), ("is_internal", i4),
                        ("bbox", f4, 6)])
tex_ref_dt = np.dtype([("tex_type", i4), ("tex_idx", i4)])
metal_dt = np.dtype([("albedo", f4, 3), ("fuzz", f4)])
dielectric_dt = np.dtype([("ior", f4), ("inv_ior", f4), ("albedo", f4, 3)])
solid_dt = np.dtype([("color", f4, 3)])
checker_dt = np.dtype([("inv_scale", f4), ("even_type", i4), ("even_idx", i4), ("odd_type", i4), ("odd_idx", i4)])
image_dt = np.dtype([("width", i4), ("height", i4), ("fnv1a", u4)])
noise_dt = np.dtype([("scale", f4), ("ranvec", f4, (256, 3)), ("perm_x", i4, 256), ("perm_y", i4, 256), ("perm_z", i4, 256)])
camera_dt = np.dtype([("aspect_ratio", f4), ("image_width", i4), ("image_height", i4), ("samples_per_pixel", i4),
                      ("pixel_samples_scale", f4), ("sqrt_spp", i4), ("recip_sqrt_spp", f4), ("bounce_limit", i4), ("vfov", i4),
                      ("background", f4, 3), ("light_obj_type", i4), ("light_obj_idx", i4),
                      ("center", f4, 3), ("pixel00_loc", f4, 3), ("pixel_delta_u", f4, 3), ("pixel_delta_v", f4, 3),
                      ("lookfrom", f4, 3), ("lookat", f4, 3), ("vup", f4, 3), ("v", f4, 3), ("u", f4, 3), ("w", f4, 3),
                      ("defocus_angle", f4), ("focus_dist", f4), ("defocus_disk_u", f4, 3), ("defocus_disk_v", f4, 3)])
hit_dt = np.dtype([("hit", i4), ("t", f4), ("leaf_type", i4), ("leaf_idx", i4), ("top_type", i4), ("top_idx", i4),
                   ("mat_type", i4), ("mat_idx", i4), ("front_face", i4), ("flags", i4),
                   ("p", f4, 3), ("normal", f4, 3), ("u", f4), ("v", f4)])
probe_dt = np.dtype([("hit1", i4), ("hit2", i4), ("t1", f4), ("t2", f4)])
assert header_dt.itemsize == 104 and sphere_dt.itemsize == 68 and quad_dt.itemsize == 104 and camera_dt.itemsize == 208
assert hit_dt.itemsize == 72 and medium_dt.itemsize == 32 and noise_dt.itemsize == 4 + 3072 + 3072


def read_scene(path) -> dict:
    """Parse a .mscn scene dump into a dict of structured arrays."""
    buf = memoryview(open(path, "rb").read())
    off = 0

    def take(dt, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=dt, count=n, offset=off)
        off += dt.itemsize * n
        return a

    h = take(header_dt, 1)[0]
    if int(h["magic"]) != MSCN_MAGIC:
        raise ValueError(f"{path}: not a scene dump")
    s = {"header": h}
    s["spheres"] = take(sphere_dt, int(h["n_sphere"]))
    s["quads"] = take(quad_dt, int(h["n_quad"]))
    s["translates"] = take(translate_dt, int(h["n_translate"]))
    s["rotates"] = take(rotate_dt, int(h["n_rotate_y"]))
    s["media"] = take(medium_dt, int(h["n_medium"]))
    s["lists"] = []
    for _ in range(int(h["n_list"])):
        skip, num = take(i4, 2)
        s["lists"].append({"skip": int(skip), "items": take(i4, 2 * int(num)).reshape(-1, 2)})
    s["bvhs"] = []
    for _ in range(int(h["n_bvh"])):
        skip, n = take(i4, 2)
        s["bvhs"].append({"skip": int(skip), "nodes": take(bvh_node_dt, int(n))})
    s["lambertians"] = take(tex_ref_dt, int(h["n_lambertian"]))
    s["metals"] = take(metal_dt, int(h["n_metal"]))
    s["dielectrics"] = take(dielectric_dt, int(h["n_dielectric"]))
    s["diffuse_lights"] = take(tex_ref_dt, int(h["n_diffuse_light"]))
    s["isotropics"] = take(tex_ref_dt, int(h["n_isotropic"]))
    s["solids"] = take(solid_dt, int(h["n_solid"]))
    s["checkers"] = take(checker_dt, int(h["n_checker"]))
    s["images"] = take(image_dt, int(h["n_image"]))
    s["noises"] = take(noise_dt, int(h["n_noise"]))
    s["camera"] = take(camera_dt, 1)[0]
    if off != len(buf):
        raise ValueError(f"{path}: {len(buf) - off} trailing bytes")
    return s


def read_hits(path) -> dict:
    """Parse a .mhit primary-hit file: rays (n,7), records, medium probes (n, n_medium)."""
    buf = memoryview(open(path, "rb").read())
    magic, n, nm, nleaves = np.frombuffer(buf, dtype=u4, count=4)
    if int(magic) != MHIT_MAGIC:
        raise ValueError(f"{path}: not a hit file")
    n, nm = int(n), int(nm)
    off = 16
    rays = np.frombuffer(buf, dtype=f4, count=7 * n, offset=off).reshape(n, 7)
    off += 28 * n
    rec = np.frombuffer(buf, dtype=hit_dt, count=n, offset=off)
    off += hit_dt.itemsize * n
    probes = np.frombuffer(buf, dtype=probe_dt, count=n * nm, offset=off).reshape(n, nm) if nm else np.zeros((n, 0), probe_dt)
    return {"rays": rays, "hits": rec, "probes": probes, "n_leaves": int(nleaves)}


def write_hits(path, rays, hits, probes=None, n_leaves=0):
    nm = 0 if probes is None or probes.size == 0 else probes.shape[1]
    with open(path, "wb") as f:
        f.write(np.array([MHIT_MAGIC, len(rays), nm, n_leaves], dtype=u4).tobytes())
        f.write(np.ascontiguousarray(rays, dtype=f4).tobytes())
        f.write(np.ascontiguousarray(hits, dtype=hit_dt).tobytes())
        if nm:
            f.write(np.ascontiguousarray(probes, dtype=probe_dt).tobytes())


def read_mimg(path) -> np.ndarray:
    """Harness image: (H, W, C) uint8 or float32, rows bottom-up exactly as rendered."""
    buf = open(path, "rb").read()
    magic, w, h, c, dt = np.frombuffer(buf, dtype=u4, count=5)
    if int(magic) != MIMG_MAGIC:
        raise ValueError(f"{path}: not a harness image")
    dtype = np.uint8 if int(dt) == 0 else f4
    return np.frombuffer(buf, dtype=dtype, offset=20).reshape(int(h), int(w), int(c))


def read_ppm(path):
    with open(path, "rb") as f:
        toks = []
        while len(toks) < 4:
            line = f.readline()
            if not line.startswith(b"#"):
                toks += line.split()
        assert toks[0] == b"P6" and toks[3] == b"255"
        w, h = int(toks[1]), int(toks[2])
        return np.frombuffer(f.read(w * h * 3), dtype=np.uint8).reshape(h, w, 3)
