"""End-to-end drop-in check on a B200: the C++ CLI (svinet_b200/lib/svinet) is run with the same flags the
UNMODIFIED reference was run with when tests/golden/ was generated, in a scratch cwd, and its output
directory is compared file by file with the reference's: communities.txt and validation-edges.txt
byte-for-byte, the numeric files field by field with +-1 unit of the last printed digit (SURVEY.md
Appendix F), wall-clock columns excluded."""
import os
import subprocess

import pytest

from golden_util import MANIFEST, Scratch, compare_numeric_text, golden_text, input_path
from svinet_b200 import build as svbuild

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    svbuild.build_lib()
    path = svbuild.build_cli()
    assert path and os.path.exists(path)
    return path


def run_case(cli, case, d, extra=(), env=None):
    ent = MANIFEST[case]
    for name in [ent["input"]] + ent.get("extra_inputs", []):
        inp, local = input_path(name, d), os.path.join(d, name)
        if not os.path.exists(local):
            os.symlink(inp, local)
    cmd = [cli, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"]), "-link-sampling"] + ent["flags"] + list(extra)
    p = subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr.decode()
    return ent, os.path.join(d, ent["outdir"])


@pytest.mark.parametrize("case", ["c1_m30", "c1_natural", "c1_m1", "c1_seed7_m12", "c1_accuracy_m8", "c1_k7_m15",
                                  "lfr_k28_m20", "c2_m12", "c2_m25", "c2_natural", "c1_etasparse_m10",
                                  "c1_initcomm_m10"])  # the -link-sampling fixtures
def test_cli_output_directory_matches_reference(cli, case):
    with Scratch() as d:
        ent, out = run_case(cli, case, d)
        flips = {}
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "validation.txt", "max.txt"):
            want = golden_text(case, fname)
            if want is None:
                continue
            got = open(os.path.join(out, fname)).read()
            flips[fname] = compare_numeric_text(got, want, skip_cols=(1,) if fname in ("validation.txt", "max.txt") else ())
        for fname in ("communities.txt", "validation-edges.txt", "param.txt"):
            want = golden_text(case, fname)
            if want is not None:
                assert open(os.path.join(out, fname)).read() == want, fname
        if "gamma.txt" not in flips:
            # c2_natural (ca-AstroPh run to its validation stop) keeps the small files only: ending on the
            # reference's iteration (max.txt, compared above) with its communities.txt is the point
            assert golden_text(case, "max.txt").split("\t")[0] == "30"
            return
        nf, noff = flips["gamma.txt"]
        # last-digit (1e-5 absolute) flips: a state that agrees to ~1e-9 absolute flips about 2e-4 of the
        # printed fields; 1e-3 of the fields is the ceiling (the compare above already bounds every field)
        print(case, flips)
        assert noff <= max(2, nf // 1000), flips
        for f in ("infer.log", "logl.txt", "test-edges.txt", "network.dat"):
            assert os.path.lexists(os.path.join(out, f)), f


@pytest.mark.parametrize("case,gpus", [("c1_m30", 3), ("c1_natural", 2), ("lfr_k28_m20", 4), ("c2_m25", 2)])
def test_cli_gpus_n_writes_the_reference_directory(cli, case, gpus):
    """`svinet -link-sampling -gpus N` (node-block shards exchanging rows over peer memory, one host thread; the seam
    is src/main.cc:337-341) against the same reference fixtures as the single-GPU run.  On a one-GPU box the shards
    share the device (SVINET_SHARDS_ON_ONE_GPU=1): same code path, copies stay on the GPU."""
    import torch
    env = {} if torch.cuda.device_count() >= gpus else {"SVINET_SHARDS_ON_ONE_GPU": "1", "CUDA_DEVICE_MAX_CONNECTIONS": "32"}
    with Scratch() as d:
        ent, out = run_case(cli, case, d, extra=["-gpus", str(gpus)], env=env)
        flips = {}
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "validation.txt", "max.txt"):
            got = open(os.path.join(out, fname)).read()
            flips[fname] = compare_numeric_text(got, golden_text(case, fname),
                                                skip_cols=(1,) if fname in ("validation.txt", "max.txt") else ())
        for fname in ("communities.txt", "validation-edges.txt"):
            assert open(os.path.join(out, fname)).read() == golden_text(case, fname), fname
        nf, noff = flips["gamma.txt"]
        assert noff <= max(2, nf // 1000), flips


def test_cli_nmi_on_lfr_benchmark(cli):
    """the LFR example run to its validation stop with -nmi <ground truth>.  The reference's
    recorded run ends at NMI 0.897 (example/n1000-k28-LFR-linksampling.tgz: mutual.txt, from the external `mutual`
    binary); the CLI computes the same measure itself (host/nmi.hh), one `mutual3:` line per report."""
    import nmi_lfk
    from test_nmi import as_matrix, lfr_ground_truth
    ent = MANIFEST["lfr_k28_m20"]
    with Scratch() as d:
        for name in (ent["input"], "LFR-ground-truth-n1000-k28.txt"):
            inp, local = input_path(name, d), os.path.join(d, name)
            if not os.path.exists(local):
                os.symlink(inp, local)
        cmd = [cli, "-file", ent["input"], "-n", "1000", "-k", "28", "-link-sampling", "-nmi", "LFR-ground-truth-n1000-k28.txt"]
        p = subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600)
        assert p.returncode == 0, p.stderr.decode()
        out = os.path.join(d, ent["outdir"])
        lines = open(os.path.join(out, "mutual.txt")).read().strip().split("\n")
        assert all(ln.startswith("mutual3:\t") for ln in lines) and len(lines) > 20
        final = float(lines[-1].split("\t")[1])
        assert final >= 0.85, lines[-5:]
        index, gt = lfr_ground_truth()
        found = [[index[int(t)] for t in ln.split()] for ln in open(os.path.join(out, "communities.txt")) if ln.strip()]
        assert abs(nmi_lfk.nmi_lfk(as_matrix(1000, gt), as_matrix(1000, found)) - final) < 2e-6      # printed with %g
        gt_lines = open(os.path.join(out, "ground_truth.txt")).read().strip().split("\n")
        assert len(gt_lines) == 28


@pytest.mark.parametrize("case", ["c1_m30", "lfr_k28_m20"])
def test_reference_gml_consumer_reads_our_output(cli, case):
    """The consumer of this path's files: `svinet -gml` (MMSBGen::gml, src/mmsbgen.cc:74-149,911) loads gamma.txt /
    lambda.txt from its cwd.  The UNMODIFIED reference binary (oracle/_ref/svinet_ref) is run once on the reference's
    own files (the fixture) and once on the files the B200 CLI wrote: the GML and the statistics it derives must agree
    (same link-community group of every node, same integer counts; decimals within the last printed digits)."""
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "svinet_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/svinet_ref not built (needs /root/reference at build time)")
    with Scratch() as d:
        ent, out = run_case(cli, case, d)
        results = {}
        for tag in ("reference", "ours"):
            w = os.path.join(d, "gml_" + tag)
            os.makedirs(w)
            os.symlink(input_path(ent["input"], d), os.path.join(w, ent["input"]))
            for f in ("gamma.txt", "lambda.txt"):
                text = golden_text(case, f) if tag == "reference" else open(os.path.join(out, f)).read()
                open(os.path.join(w, f), "w").write(text)
            p = subprocess.run([ref, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"]), "-gml"], cwd=w,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
            assert p.returncode == 0 and b"Done writing GML file" in p.stdout, p.stdout.decode()[-400:]
            results[tag] = {f: open(os.path.join(w, "gml", f)).read() for f in
                            ("network.gml", "community_stats.txt", "node_bridgeness.txt", "node_influence.txt",
                             "number_of_memberships.txt")}
        for f in results["ours"]:
            a, b = results["ours"][f].split("\n"), results["reference"][f].split("\n")
            assert len(a) == len(b), f
            for la, lb in zip(a, b):
                ta, tb = la.split(), lb.split()
                assert len(ta) == len(tb), (f, la, lb)
                for x, y in zip(ta, tb):
                    if x == y:
                        continue
                    assert "." in y and abs(float(x) - float(y)) <= 2e-4 * max(1.0, abs(float(y))), (f, la, lb)


@pytest.mark.parametrize("case", ["c1_m30", "c1_natural", "lfr_k28_m20"])
def test_reference_host_code_over_the_library(cli, case):
    """INTEGRATION.md option B, executed (oracle/ref_b200.py, `make -C oracle ref_b200`): the reference's OWN
    LinkSampling -- its constructor, RNG, held-out draw, stop machine, writers -- with the loop body replaced by
    svi_ls_step and the three consumers fed from the device.  Its output directory must equal the stock reference's
    (the committed fixture)."""
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "svinet_ref_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/svinet_ref_b200 not built (needs /root/reference at build time)")
    ent = MANIFEST[case]
    with Scratch() as d:
        inp, local = input_path(ent["input"], d), os.path.join(d, ent["input"])
        if not os.path.exists(local):
            os.symlink(inp, local)
        cmd = [exe, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"]), "-link-sampling"] + ent["flags"]
        p = subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=900)
        assert p.returncode == 0, p.stderr.decode()[-500:]
        out = os.path.join(d, ent["outdir"])
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "validation.txt", "max.txt"):
            got = open(os.path.join(out, fname)).read()
            nf, noff = compare_numeric_text(got, golden_text(case, fname),
                                            skip_cols=(1,) if fname in ("validation.txt", "max.txt") else ())
            assert noff <= max(2, nf // 1000), (fname, nf, noff)
        for fname in ("communities.txt", "validation-edges.txt"):
            assert open(os.path.join(out, fname)).read() == golden_text(case, fname), fname


def test_cli_load_validation_with_repeated_pairs(cli):
    """-load-validation <file>: the reference keeps the held-out pairs in a std::map, so a file that repeats a pair
    counts it once.  Both binaries run on the same file (the reference's own validation-edges.txt with a
    few pairs repeated); validation.txt and the model must agree."""
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "svinet_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/svinet_ref not built (needs /root/reference at build time)")
    ent = MANIFEST["c1_m30"]
    pairs = [ln.split("\t")[:2] for ln in golden_text("c1_m30", "validation-edges.txt").strip().split("\n")]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        text = "".join("%s\t%s\n" % (a, b) for a, b in pairs + pairs[:5] + pairs[3:4])
        outs = {}
        for tag, exe in (("ref", ref), ("ours", cli)):
            w = os.path.join(d, tag)
            os.makedirs(w)
            os.symlink(inp, os.path.join(w, ent["input"]))
            open(os.path.join(w, "held.txt"), "w").write(text)
            p = subprocess.run([exe, "-file", ent["input"], "-n", "75", "-k", "4", "-link-sampling", "-max-iterations", "12",
                                "-no-stop", "-load-validation", "held.txt"], cwd=w, stdout=subprocess.DEVNULL,
                               stderr=subprocess.PIPE, timeout=600)
            assert p.returncode == 0, p.stderr.decode()[-400:]
            outs[tag] = os.path.join(w, ent["outdir"])
        for fname in ("validation.txt", "gamma.txt", "lambda.txt"):
            compare_numeric_text(open(os.path.join(outs["ours"], fname)).read(), open(os.path.join(outs["ref"], fname)).read(),
                                 skip_cols=(1,) if fname == "validation.txt" else ())
        assert open(os.path.join(outs["ours"], "communities.txt")).read() == open(os.path.join(outs["ref"], "communities.txt")).read()


FA2_CASES = [c for c in MANIFEST if MANIFEST[c].get("mode") == "-rnode -stratified"]


@pytest.mark.parametrize("case", FA2_CASES)
def test_fa2_cli_output_directory_matches_reference(cli, case):
    """`svinet -rnode -stratified` (class FastAMM2): same flags as the fixture, host replays the reference's
    mt19937 minibatch draws, the device runs every iteration."""
    ent = MANIFEST[case]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        local = os.path.join(d, ent["input"])
        if not os.path.exists(local):
            os.symlink(inp, local)
        cmd = [cli, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"])] + ent["mode"].split() + ent["flags"]
        p = subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=900)
        assert p.returncode == 0, p.stderr.decode()
        out = os.path.join(d, ent["outdir"])
        flips = {}
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "heldout.txt"):
            got = open(os.path.join(out, fname)).read()
            flips[fname] = compare_numeric_text(got, golden_text(case, fname), skip_cols=(1,) if fname == "heldout.txt" else ())
        for fname in ("communities.txt", "communities_size.txt", "summary.txt", "heldout-pairs.txt", "param.txt"):
            assert open(os.path.join(out, fname)).read() == golden_text(case, fname), fname
        print(case, flips)
        nf, noff = flips["gamma.txt"]
        assert noff <= max(2, nf // 1000), flips
        for f in ("infer.log", "cmap.txt", "precision.txt", "validation-pairs.txt", "mcount.txt", "aggregate.txt", "network.dat"):
            assert os.path.lexists(os.path.join(out, f)), f


def test_fa2_cli_device_draw(cli):
    """-device-draw: minibatches from the device's Philox stream (svi_fa2_run); no reference fixture can exist
    for it, so this checks the run end to end: the report cadence, finite positive state, rows that sum to 1."""
    import numpy as np
    ent = MANIFEST["fa2_lfr_k28_m300"]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        if not os.path.exists(os.path.join(d, ent["input"])):
            os.symlink(inp, os.path.join(d, ent["input"]))
        cmd = [cli, "-file", ent["input"], "-n", "1000", "-k", "28", "-rnode", "-stratified", "-max-iterations", "250",
               "-rfreq", "100", "-seed", "5", "-device-draw"]
        p = subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=900)
        assert p.returncode == 0, p.stderr.decode()
        out = os.path.join(d, "n1000-k28-mmsb-seed5-Srnode")
        rows = [l.split("\t") for l in open(os.path.join(out, "heldout.txt")).read().strip().split("\n")]
        assert [r[0] for r in rows] == ["0", "100", "200"]
        assert int(rows[2][-1]) > int(rows[1][-1]) > 0                      # pairs sampled keeps growing
        gam = np.array([[float(x) for x in l.split("\t")[2:]] for l in open(os.path.join(out, "gamma.txt")).read().strip().split("\n")])
        assert gam.shape == (1000, 28) and np.all(np.isfinite(gam)) and np.all(gam > 0)
        grp = np.array([[float(x) for x in l.split("\t")[2:-1]] for l in open(os.path.join(out, "groups.txt")).read().strip().split("\n")])
        assert np.allclose(grp.sum(axis=1), 1.0, atol=0.02)
        # held-out likelihood of links improves over the initial state
        assert float(rows[2][6]) > float(rows[0][6])


def test_cli_resume_from_saved_model(cli):
    """-load <dir/> (checkpoint/resume, linksampling.cc:1267-1352): a run resumed from a saved model's %.5f
    text starts from exactly that gamma/lambda."""
    with Scratch() as d:
        ent, out = run_case(cli, "c1_m30", d)
        os.rename(out, os.path.join(d, "saved"))
        cmd = [cli, "-file", ent["input"], "-n", "75", "-k", "4", "-link-sampling", "-load", "saved/", "-label", "resumed",
               "-dump-init", d]
        subprocess.check_call(cmd, cwd=d, stdout=subprocess.DEVNULL)
        import numpy as np
        gam = np.fromfile(os.path.join(d, "gamma.f64")).reshape(75, 4)
        want = np.array([[float(x) for x in l.split("\t")[2:]] for l in
                         open(os.path.join(d, "saved", "gamma.txt")).read().strip().split("\n")])
        assert np.array_equal(gam, want)
