"""MT19937 jump-ahead (svinet_b200/host/mt_jump.hh): the parallel producers of init_gamma2 must hand out exactly the
reference's stream (src/linksampling.cc:374-401 draws K uniforms per link from one gsl_rng)."""
import os
import subprocess

import numpy as np

from golden_util import MANIFEST, Scratch, input_path
from svinet_b200 import build as svbuild

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jump_ahead_equals_sequential_generation(tmp_path):
    exe = str(tmp_path / "mt_jump_check")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-I", os.path.join(REPO, "svinet_b200", "host"),
                           "-o", exe, os.path.join(REPO, "tests", "cc", "mt_jump_check.cc"), "-lpthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == "0 mismatches", out.stdout


def test_parallel_producers_give_the_serial_start_state():
    """The CLI's start-up state (-dump-init) with 1, 3 and 4 producers: bit-identical gamma, on ca-AstroPh (K = 20,
    4e6 uniforms in 241 chunks) and on LFR K = 28."""
    svbuild.build_lib()
    cli = svbuild.build_cli()
    for case in ("c2_m12", "lfr_k28_m20"):
        ent = MANIFEST[case]
        got = []
        for producers in ("1", "3", "4"):
            with Scratch() as d:
                inp, local = input_path(ent["input"], d), os.path.join(d, ent["input"])
                if not os.path.exists(local):
                    os.symlink(inp, local)
                dump = os.path.join(d, "dump")
                os.makedirs(dump)
                subprocess.check_call([cli, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"]), "-link-sampling",
                                       "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL,
                                      env=dict(os.environ, SVINET_INIT_PRODUCERS=producers))
                got.append(np.fromfile(os.path.join(dump, "gamma.f64")))
        assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[2]), case
