"""CPU test of the tile enumeration of the symmetric kernels (csrc/sym_tc_dev.cuh Tile5Iter): the host-side program
tests/native/tile_iter_check.cu is compiled with nvcc and run here -- no GPU needed (the enumeration is __host__ __device__)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_run_based_tile_iterator_visits_what_next_live_visits(tmp_path):
    exe = str(tmp_path / "tile_iter_check")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr",
           "-I", os.path.join(ROOT, "randomly-projected-additive-gps_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "tests", "native", "tile_iter_check.cu")]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and run.stdout.startswith("ok "), run.stdout[-2000:]
    assert int(run.stdout.split()[1]) > 100000        # every (size, row block, split, tile) combination was walked
