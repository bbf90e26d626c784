"""Why is the megakernel ~3 % slower inside bench.py than inside the CLI?  Same library, same scene, kernel time from the library's own CUDA events."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mort_b200.api import Renderer

def run(tag, r, n=4, **kw):
    ms = []
    for f in range(n):
        r.render(want_accum=False, seed=69420, frame=f, **kw)
        ms.append(r.stats["last_render_ms"])
    print(json.dumps({"case": tag, "kernel_ms": [round(x, 2) for x in ms]}), flush=True)

import ctypes
_cu = ctypes.CDLL("libcuda.so.1")
def limits(tag):
    names = {0: "stack", 1: "printf_fifo", 2: "malloc_heap", 5: "l2_fetch_granularity", 6: "persisting_l2"}
    out = {}
    for k, n in names.items():
        v = ctypes.c_size_t(0)
        rc = _cu.cuCtxGetLimit(ctypes.byref(v), k)
        out[n] = v.value if rc == 0 else f"rc{rc}"
    cfg = ctypes.c_int(0); _cu.cuCtxGetCacheConfig(ctypes.byref(cfg)); out["cache_config"] = cfg.value
    print(json.dumps({"limits": tag, **out}), flush=True)
def set_stack(n):
    return _cu.cuCtxSetLimit(0, ctypes.c_size_t(n))

which = sys.argv[1]
if which == "plain":
    with Renderer(0) as r:
        r.build_scene(6).override_camera(width=600, spp=1024, depth=50).commit()
        run("ctypes only, library's own stream", r)
        limits("plain")
elif which == "torch":
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda:0"); torch.cuda.synchronize()
    with Renderer(0) as r:
        r.build_scene(6).override_camera(width=600, spp=1024, depth=50).commit()
        run("torch context alive, library's own stream", r)
        limits("torch")
        for n in (1024, 2048, 0):
            if n:
                print(json.dumps({"set_stack": n, "rc": set_stack(n)}), flush=True)
                run(f"torch context alive, stack limit {n}", r, n=3)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        run("torch context alive, torch's current (legacy default) stream", r)
        st = r.stats
        acc = torch.zeros(st["height"], st["width"], 4, dtype=torch.float32, device="cuda:0")
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        ms = []
        for f in range(4):
            r.render_device(acc.data_ptr(), seed=69420, frame=f); ms.append(r.stats["last_render_ms"])
        print(json.dumps({"case": "render_device into a torch tensor, legacy default stream", "kernel_ms": [round(x, 2) for x in ms]}), flush=True)
