"""The model writers' number formatter (svinet_b200/host/fixed_fmt.hh) must produce printf's bytes: the reference
writes gamma.txt / groups.txt with fprintf("%.5f") / ("%.3f") (linksampling.cc:805-837, :1453-1476)."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_append_fixed_equals_printf(tmp_path):
    exe = str(tmp_path / "fixed_fmt_check")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-I", os.path.join(REPO, "svinet_b200", "host"),
                           "-o", exe, os.path.join(REPO, "tests", "cc", "fixed_fmt_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 mismatches" in out.stdout
