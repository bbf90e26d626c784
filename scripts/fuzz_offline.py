#!/usr/bin/env python
"""Offline run of the CPU fuzz parity (tests/test_fuzz_parity.py) over many more seeds: random scene files through the product's
builder + flattening + per-ray code (tests/hostsim) against the oracle.  Prints how many scenes agree, how many the product refuses
(by reason) and every disagreement.   python scripts/fuzz_offline.py [first_seed] [n_seeds] [--frames]"""
import collections
import os
import pathlib
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pytest  # noqa: E402

import test_fuzz_parity as T  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 500
frames = "--frames" in sys.argv
hostsim = os.path.join(ROOT, "tests", "hostsim", "hostsim.bin")
res = collections.Counter()
fn = T.test_random_scene_frames_match_the_oracle if frames else T.test_random_scene_closest_hits_match_the_oracle
for seed in range(first, first + count):
    with tempfile.TemporaryDirectory() as d:
        try:
            fn.__wrapped__(hostsim, seed, pathlib.Path(d)) if hasattr(fn, "__wrapped__") else fn(hostsim, seed, pathlib.Path(d))
            res["agree"] += 1
        except pytest.skip.Exception as e:
            res["refused: " + str(e)[-70:]] += 1
        except AssertionError as e:
            res["MISMATCH"] += 1
            print("seed", seed, "MISMATCH", str(e)[:300], flush=True)
for k, v in sorted(res.items(), key=lambda kv: -kv[1]):
    print(f"{v:6d}  {k}")
