"""examples/replay_city.cpp — the call sequence INTEGRATION.md describes (readers, undistortion, voxel grid, pose table, merged
scan, iterated update, map deltas) as one C++ program against the header, the library and the REAL reference tree type.
Here: it must compile and link (the reference's ikd_Tree.h/.cpp resolved from /root/reference through the oracle's PCL shim),
its host-only part must run, and without a GPU the full program must fail loudly at malio_create."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IKD = "/root/reference/MA_LIO/include/ikd-Tree"
LIBDIR = os.path.join(ROOT, "ma-lio_b200", "malio_b200")


@pytest.mark.skipif(not os.path.exists(os.path.join(IKD, "ikd_Tree.cpp")), reason="reference tree not present (GPU box)")
@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libmalio_b200.so")), reason="library not built")
def test_cpp_example_compiles_links_and_its_host_side_runs(tmp_path):
    exe = str(tmp_path / "replay_city")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++14", "-fopenmp", "-pthread", "-w", "-I" + os.path.join(ROOT, "oracle", "pcl_shim"), "-I" + IKD,
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "replay_city.cpp"), "-L" + LIBDIR, "-lmalio_b200",
           "-Wl,-rpath," + LIBDIR, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe, "--host-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    line = [l for l in r.stdout.splitlines() if l.startswith("host-only:")][0]
    assert "table entries 4 (offsets 0 2 4)" in line and "spline ok 1 p.x 0.350000" in line and "outside 0" in line
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "malio_create:" in r.stderr      # no silent CPU fallback
