"""CPU test of the host-side staging helpers of the C-ABI library (threaded non-temporal copy, PointCloud2 field
extraction for layouts with adjacent and separate x / y / z fields): compiled with the host compiler and run."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_copy_pool_and_pointcloud2_extraction(tmp_path):
    exe = tmp_path / "host_pool_check"
    csrc = os.path.join(ROOT, "gvom_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "host_pool_check.cpp"),
                           os.path.join(csrc, "gvom_host.cpp"), "-o", str(exe), "-lpthread"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr
