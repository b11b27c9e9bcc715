"""CPU: the grid nearest-neighbour search logic shared by host and device (gennbv_b200/csrc/nn_grid.cuh) returns, for every
query, exactly the minimum a scan over all points returns -- surface-like, volumetric, degenerate (single / coincident /
planar / collinear / empty) clouds, coarse and over-fine grids (brute-force escape), queries far outside the box.
The CUDA kernels (chamfer.cu) call the same functions; their parallel build is covered by the GPU tests."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_grid_search_equals_scan_on_host(tmp_path):
    exe = str(tmp_path / "nn_grid_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe,
                           os.path.join(HERE, "host", "nn_grid_host_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "OK (0 mismatches)" in out.stdout
