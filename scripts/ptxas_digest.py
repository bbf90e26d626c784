#!/usr/bin/env python
"""Compiles one .cu of mort_b200/csrc with -Xptxas -v and prints, per kernel: registers, stack frame, spill bytes and
(with --sass) counts of the SASS mnemonics that matter here (LDG.E.128, LDL/STL, LDS/STS, FFMA, DFMA/DMUL/DADD, MUFU, BAR, ATOMS, RED).
  python scripts/ptxas_digest.py pool.cu [-DFOO ...] [--sass]"""
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mort_b200", "csrc")


def main():
    args = sys.argv[1:]
    sass = "--sass" in args
    args = [a for a in args if a != "--sass"]
    src, extra = args[0], args[1:]
    obj = "/tmp/ptxas_digest_" + os.path.basename(src) + ".o"
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xptxas", "-v", "-Xcompiler", "-fPIC",
           "-I" + CSRC, "-I" + os.path.join(ROOT, "include")] + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    out = subprocess.run(cmd, capture_output=True, text=True).stderr
    cur = None
    rows = {}
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", "").replace("void ", "")).replace("mort::", "")
            rows[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "stack" not in rows[cur]:
            rows[cur].update(stack=int(m.group(1)), st=int(m.group(2)), ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            rows[cur]["regs"] = int(m.group(1))
            cur = None
    if "error" in out:
        print(out)
    counts = {}
    if sass:
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name = None
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", "")).replace("mort::", "")
                counts[name] = {}
                continue
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and name:
                op = m.group(1)
                c = counts[name]
                c["total"] = c.get("total", 0) + 1
                for key, pat in (("LDG.128", r"^LDG\.E\.128"), ("LDG", r"^LDG"), ("LDL", r"^LDL"), ("STL", r"^STL"), ("LDS", r"^LDS"), ("STS", r"^STS"),
                                 ("FFMA", r"^FFMA"), ("FP64", r"^D(FMA|MUL|ADD|SETP)"), ("MUFU", r"^MUFU"), ("BAR", r"^BAR"), ("ATOMS", r"^ATOMS"),
                                 ("RED/ATOMG", r"^(RED|ATOMG)"), ("CALL", r"^CALL")):
                    if re.search(pat, op):
                        c[key] = c.get(key, 0) + 1
    for k, v in rows.items():
        line = f"{k:60s} regs {v.get('regs', 0):4d}  stack {v.get('stack', 0):5d} B  spill st/ld {v.get('st', 0):4d}/{v.get('ld', 0):4d} B"
        if k in counts:
            line += "  | " + " ".join(f"{a} {b}" for a, b in counts[k].items())
        print(line)


if __name__ == "__main__":
    main()
