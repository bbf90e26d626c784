"""The C-ABI shared library loads on a machine without a GPU, exports every symbol include/mort_b200.h
declares, and refuses — loudly, without a CPU fallback — to create a context when no CUDA device exists."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mort_b200 import api


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mort_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mort_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = api.load_library()
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors vs the sizes a C compiler derives from include/mort_b200.h (a mismatch would corrupt every call)"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mort_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mort_render_opts),'
                   ' sizeof(mort_camera_desc), sizeof(mort_handle), sizeof(mort_stats), sizeof(mhit_record), sizeof(mscn_camera), sizeof(mscn_header),'
                   ' sizeof(mort_build_opts), sizeof(mort_build_info), sizeof(mort_group_stats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    from mort_b200 import formats as F
    assert sizes == [ctypes.sizeof(api.RenderOpts), ctypes.sizeof(api.CameraDesc), ctypes.sizeof(api.Handle), ctypes.sizeof(api.Stats),
                     F.hit_dt.itemsize, F.camera_dt.itemsize, F.header_dt.itemsize,
                     ctypes.sizeof(api.BuildOpts), ctypes.sizeof(api.BuildInfo), ctypes.sizeof(api.GroupStats)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.MortError):
        api.Renderer(0)
    L = api.load_library()
    h = ctypes.c_void_p()
    assert L.mort_create(0, ctypes.byref(h)) == -2 and not h          # MORT_ERR_CUDA
    o = api.RenderOpts()
    L.mort_default_render_opts(ctypes.byref(o))
    assert (o.seed, o.mode, o.sample_mod, o.sample_rem, o.stage_nodes, o.pool_paths) == (69420, api.MODE_POOL, 1, 0, 0, 0)


def test_product_sources_never_touch_the_oracle():
    """the product path must not route through oracle/ (or any CPU renderer)"""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "mort_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r'#include\s*[<"][^>"]*oracle|import\s+oracle|from\s+oracle|liboracle|oracle_binding|CDLL\([^)]*oracle|oracle_scene_|oracle_render|oracle_trace', txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
