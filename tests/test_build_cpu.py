"""The GPU tree builder and the refit / motion-box pass without a GPU: their per-thread bodies (gpu_build_core.cuh, refit_core.cuh)
and the builder's level loop (gpu_build_driver.hpp) run serially on the CPU (tests/hostsim/buildsim.cpp) — positions visited in a
shuffled order, standing in for the arbitrary order atomics resolve in — and must give exactly the host builder's tree
(bvh_build.cpp): same 4-wide nodes, same leaf order, same level table.  The refit pass must reproduce the builder's boxes bit for
bit on an unchanged scene, and the interpolated motion boxes must contain every primitive at every time and nest.
The CUDA side of the same checks is tests/test_gpu_build.py."""
import subprocess

import pytest

from conftest import ASSETS


def _run(buildsim, what, k_small):
    p = subprocess.run([buildsim, what, ASSETS, str(k_small)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-600:]
    return p.stderr


@pytest.mark.parametrize("k_small", [4, 16, 64])
@pytest.mark.parametrize("what", ["1", "8", "9", "field:40", "dup:300"])
def test_emulated_gpu_build_equals_host_build(buildsim, what, k_small):
    err = _run(buildsim, what, k_small)
    assert "order equal, word/box differences 0" in err
    if not what.startswith("dup"):
        assert "refit of the unchanged scene: 0 box words differ" in err
        assert " 0 violations" in err and " 0 edit-violations" in err


@pytest.mark.parametrize("seed", range(12))
def test_emulated_gpu_build_on_random_boxes(buildsim, seed):
    n = [37, 300, 2000, 9000][seed % 4]
    err = _run(buildsim, f"rand:{n}:{seed}", [1, 4, 64][seed % 3])
    assert "order equal, word/box differences 0" in err


def test_linear_scenes_have_no_tree(buildsim):
    assert "linear-scan scene: no tree" in _run(buildsim, "6", 64)
