#!/usr/bin/env python
"""Inclusive per-function profile of one kernel from an ncu source-page export.

usage: ncu_inclusive_profile.py <ncu --page source --csv export> <kernel name fragment, e.g. 'pool_kernelILi640ELi1E'>
                                <nvdisasm -gi dump of the SAME cubin> <dir with the .cu/.cuh of the SAME build> [--lines N]

`nvdisasm -gi` prints, before every SASS instruction, the chain of source lines it was inlined through; every instruction is
charged to each function in its chain (inclusive).  ncu lists the kernel's own instructions first and then the out-of-line
device functions it calls; those are matched to their nvdisasm sections by opcode sequence.
lanes = thread instructions / warp instructions (32 = fully converged); samp = share of warp stall samples.
With --lines N the N hottest source lines are printed as well."""
import collections
import csv
import os
import re
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
prof, variant, dis, srcdir = args[:4]
n_lines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
src = {}
for f in os.listdir(srcdir):
    if f.endswith((".cu", ".cuh", ".hpp", ".h")):
        src[f] = open(os.path.join(srcdir, f), errors="replace").read().split("\n")

FUNC_RE = re.compile(r"^(?:MORT_HD(?:_NOINLINE)?|__global__|__device__|static|inline|template|MORT_HD_NOINLINE).*?\b([A-Za-z_0-9]+)\s*\(")


def func_of(f_, ln):
    if f_ not in src:
        return f_
    L = src[f_]
    for k in range(min(ln, len(L)) - 1, -1, -1):
        line = L[k]
        if line.startswith(" ") or line.startswith("\t") or line.startswith("//") or line.startswith("#"):
            continue
        m = FUNC_RE.match(line)
        if m:
            name = m.group(1)
            if name == "__launch_bounds__":
                m2 = re.search(r"\)\s*([A-Za-z_0-9]+)\s*\(", line)
                return m2.group(1) if m2 else "kernel"
            return name
        if line.startswith("}"):
            return f_            # between functions
    return f_


# ---- nvdisasm: sections -> [(offset, opcode text, chain)]
sections = collections.OrderedDict()
fn = None
chain, pending = [], []
for l in open(dis, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1); sections[fn] = []; chain = []; pending = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and fn:
        if pending:
            chain = pending; pending = []
        sections[fn].append((int(m.group(1), 16), m.group(2), list(chain)))

kern = [k for k in sections if variant in k]
assert kern, f"no section matches {variant}"
kern = kern[0]


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t.strip())
    return t.split()[0] if t.split() else ""


rows = list(csv.reader(open(prof)))
for i, r in enumerate(rows[:6]):
    if "Source" in r:
        hdr = r; data = rows[i + 1:]
        break
ix = {h: i for i, h in enumerate(hdr)}
ops = [opcode(r[ix["Source"]]) for r in data]

# kernel's own instructions first
assign = [None] * len(data)
nk = len(sections[kern])
for k in range(min(nk, len(data))):
    assign[k] = sections[kern][k][2]
# callees: find each remaining section by its opcode sequence
pos = nk
others = {k: v for k, v in sections.items() if "kernel" not in k and len(v) > 4}
while pos < len(data):
    best = None
    for name, ins in others.items():
        n = len(ins)
        if pos + n <= len(data) and all(opcode(ins[j][1]) == ops[pos + j] for j in range(min(n, 24))):
            if best is None or n > len(others[best]):
                best = name
    if best is None:
        pos += 1
        continue
    for j, (_, _, ch) in enumerate(others[best]):
        assign[pos + j] = ch if ch else [(best, 0)]
    pos += len(others[best])

incl = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
lines = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = stot = 0.0
unassigned = 0.0
for r, ch in zip(data, assign):
    ie = float(r[ix["Instructions Executed"]] or 0); te = float(r[ix["Thread Instructions Executed"]] or 0); sm = float(r[ix["# Samples"]] or 0)
    tot += ie; stot += sm
    if ch is None:
        unassigned += ie
        continue
    fns = []
    for (f_, ln) in ch:
        f2 = func_of(f_, ln) if ln else f_
        if f2 not in fns:
            fns.append(f2)
    for f2 in fns:
        a = incl[f2]; a[0] += ie; a[1] += te; a[2] += sm
    if ch:
        f_, ln = ch[-1] if False else ch[0]          # innermost frame is printed first by nvdisasm -gi
        a = lines[(f_, ln)]; a[0] += ie; a[1] += te; a[2] += sm
print(f"kernel section {kern[:80]}: {nk} instructions; listing {len(data)}; unassigned {100 * unassigned / max(tot, 1):.1f}% of executed")
print("total warp instr %.4g, lanes %.1f" % (tot, sum(float(r[ix['Thread Instructions Executed']] or 0) for r in data) / max(tot, 1)))
for k, a in sorted(incl.items(), key=lambda x: -x[1][0])[:48]:
    print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[2] / max(stot, 1):5.1f}% samp lanes {a[1] / max(a[0], 1):5.1f}  {k}")
if n_lines:
    print("-- hottest source lines (innermost frame)")
    for (f_, ln), a in sorted(lines.items(), key=lambda x: -x[1][0])[:n_lines]:
        text = src.get(f_, [""] * (ln + 1))[ln - 1].strip()[:110] if f_ in src and 0 < ln <= len(src[f_]) else ""
        print(f"{100 * a[0] / tot:5.2f}% inst {100 * a[2] / max(stot, 1):5.2f}% samp lanes {a[1] / max(a[0], 1):5.1f}  {f_}:{ln}  {text}")
