#!/usr/bin/env python
"""profiles/r02_call_*_ab.jsonl (raw `mort` CLI lines of the round-2 A/B GPU calls, each followed by `# tag :: arguments`)
-> profiles/r02_pool_ab.md: one table per call, one row per build/shape tag, one column per workload."""
import collections
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
README = {}
for l in open(os.path.join(ROOT, "scripts", "r02", "README.md")):
    if l.startswith("| `gpu_r2_"):
        for name in l.split("|")[1].split(","):
            README[name.strip().strip("`")] = l.split("|")[2].strip()


def workload(d, args):
    if "--field" in args:
        g = args.split()[args.split().index("--field") + 1]
        cam = args.split()[args.split().index("--fieldcam") + 1] if "--fieldcam" in args else "0"
        return f"field G={g} cam {cam} {d['width']}x{d['height']} {d['spp_eff']} spp"
    return f"scene {d['scene']} {d['width']}x{d['height']} {d['spp_eff']} spp"


def main():
    out = ["# Round 2: A/B runs of the block wavefront (pool.cu) on one B200", "",
           "Msamples/s from the `mort` CLI (CUDA-event kernel time, second of two frames), every row of a table measured in the same `gpurun` call",
           "on the same box.  Raw lines: `profiles/r02_call_*_ab.jsonl`; what each call built: `scripts/r02/gpu_r2_*.sh`.  Tags: `mega` = round-1",
           "megakernel; `pool_TxB_P` = T threads per block, B blocks per SM, P paths per block pool; `refillN` = lanes refill when N of a warp's lanes",
           "are idle; `xN` = `pool_flags` N; a prefix such as `old_` / `new_` / `cold_` / `top_` names the A/B build of that call.", ""]
    extra = {"final": "closing call L: leaf postponing (`ab_postpone` = `-DMORT_POSTPONE`) against the shipped build", "n_auto": "call N: `MORT_MODE_AUTO` against `--mode pool` and `--mode mega`"}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_call_*_ab.jsonl"))) + [os.path.join(ROOT, "profiles", "r02_final_ab.jsonl"), os.path.join(ROOT, "profiles", "r02_n_auto_ab.jsonl")]
    for f in files:
        call = os.path.basename(f)[len("r02_"):-len("_ab.jsonl")]
        call = call[len("call_"):] if call.startswith("call_") else call
        README.update({"gpu_r2_" + k + ".sh": v for k, v in extra.items()})
        lines = open(f).read().splitlines()
        table, cols = collections.OrderedDict(), []
        for i, l in enumerate(lines):
            if not l.startswith("{"):
                continue
            try:
                d = json.loads(l)
            except ValueError:                          # a line cut short by the call's `cut -c`: the leading fields are all this needs
                d = {k: (float(v) if "." in v else int(v)) for k, v in re.findall(r'"(scene|width|height|spp_eff|msamples_per_s|regs|bps)":([0-9.]+)', l)}
            if "msamples_per_s" not in d or i + 1 >= len(lines) or not lines[i + 1].lstrip().startswith("#"):
                continue
            tag, _, args = lines[i + 1].strip()[2:].partition("::")
            w = workload(d, args)
            if w not in cols:
                cols.append(w)
            table.setdefault(tag.strip(), {})[w] = (d["msamples_per_s"], d.get("regs"), d.get("bps"))
        if not table:
            continue
        out += [f"## call {call} — {README.get('gpu_r2_' + call + '.sh', '')}", "", "| tag | " + " | ".join(cols) + " | regs / blocks per SM |", "|---|" + "---:|" * len(cols) + "---|"]
        best = {w: max(v[w][0] for v in table.values() if w in v) for w in cols}
        for tag, v in table.items():
            cells = [(f"**{v[w][0]:.0f}**" if v[w][0] == best[w] else f"{v[w][0]:.0f}") if w in v else "" for w in cols]
            rb = next(iter(v.values()))
            out.append(f"| `{tag}` | " + " | ".join(cells) + f" | {rb[1]} / {rb[2]} |")
        out.append("")
    open(os.path.join(ROOT, "profiles", "r02_pool_ab.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()
