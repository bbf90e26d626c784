"""Diagnostic: where and by how much do the exact frames of the megakernel and the block wavefront differ?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mort_b200.api import MODE_MEGAKERNEL, MODE_POOL, Renderer

r = Renderer(0)
for sc, w, spp in ((6, 64, 64), (1, 96, 64), (8, 40, 16), (2, 96, 64)):
    r.build_scene(sc).override_camera(width=w, spp=spp).commit()
    st = r.stats
    H, W = st["height"], st["width"]
    out = {}
    for name, mode, kw in (("mega", MODE_MEGAKERNEL, {}), ("pool", MODE_POOL, {}), ("pool_refill", MODE_POOL, dict(pool_refill=8))):
        buf = torch.zeros(H, W, 4, dtype=torch.int64, device="cuda")
        r.render_device(buf.data_ptr(), seed=99, frame=2, mode=mode, exact_accum=1, **kw)
        torch.cuda.synchronize()
        out[name] = (buf.cpu().numpy(), r.stats["last_segments"], r.stats["last_samples"])
    a = out["mega"][0]
    for name in ("pool", "pool_refill"):
        b = out[name][0]
        d = (a != b).any(-1)
        diff = np.abs(a[..., :3] - b[..., :3]).astype(np.float64) / 2 ** 24
        print(f"scene {sc} {name}: segments {out['mega'][1]} vs {out[name][1]}, samples {out['mega'][2]} vs {out[name][2]}, pixels differing {d.sum()} of {d.size}, "
              f"max |diff| {diff.max():.3e} (pixel sums ~{np.abs(a[..., :3]).mean() / 2 ** 24:.3f}), flags differ {(a[..., 3] != b[..., 3]).sum()}, total sum diff {(a[..., :3].sum() - b[..., :3].sum()) / 2 ** 24:.3e}")
r.close()
