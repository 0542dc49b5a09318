"""Host-side logic of the C++ drop-in CLI (svinet_b200/host) without a GPU.

`svinet ... -dump-init DIR` runs everything the reference's constructor does on the host -- ingest,
id mapping, MT19937 held-out draw, init_gamma2, training-link assignment, param.txt -- writes the result
as raw arrays and exits before the first device call.  Integer/RNG work must match the oracle (and,
through it, the reference) bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as orc
from golden_util import MANIFEST, Scratch, golden_text, input_path
from svinet_b200 import build as svbuild


@pytest.fixture(scope="module")
def cli():
    svbuild.build_lib()
    path = svbuild.build_cli()
    assert path and os.path.exists(path)
    return path


def run_dump(cli, case, d, extra=()):
    ent = MANIFEST[case]
    for name in [ent["input"]] + ent.get("extra_inputs", []):
        inp, local = input_path(name, d), os.path.join(d, name)
        if not os.path.exists(local):
            os.symlink(inp, local)
    dump = os.path.join(d, "dump")
    os.makedirs(dump, exist_ok=True)
    cmd = [cli, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"]), "-link-sampling"] + ent["flags"] + \
          list(extra) + ["-dump-init", dump]
    subprocess.check_call(cmd, cwd=d, stdout=subprocess.DEVNULL)
    k = ent["k"]
    out = {
        "gamma": np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, k),
        "lambda": np.fromfile(os.path.join(dump, "lambda.f64")).reshape(k, 2),
        "validation": np.fromfile(os.path.join(dump, "validation.u32"), dtype=np.uint32).reshape(-1, 2),
        "links": np.fromfile(os.path.join(dump, "links.u32"), dtype=np.uint32).reshape(-1, 2),
        "tl": np.fromfile(os.path.join(dump, "tl.f64")),
        "outdir": os.path.join(d, ent["outdir"]),
    }
    return ent, out


def oracle_opts(flags):
    o = {}
    if "-seed" in flags:
        o["seed"] = float(flags[flags.index("-seed") + 1])
    if "-accuracy" in flags:
        o["accuracy"] = 1
    if "-eta-type" in flags:
        o["eta0"], o["eta1"] = {"sparse": (0.97, 6.33), "dense": (4700.59, 0.77)}[flags[flags.index("-eta-type") + 1]]
    return o


@pytest.mark.parametrize("case", ["c1_m30", "c1_seed7_m12", "c1_accuracy_m8", "c1_k7_m15", "lfr_k28_m20",
                                  "c1_etasparse_m10"])
def test_startup_state_is_bit_identical_to_oracle(cli, case):
    with Scratch() as d:
        ent, got = run_dump(cli, case, d)
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Model(g, ent["k"], **oracle_opts(ent["flags"]))
        st = m.state
        assert np.array_equal(got["gamma"], st.arr("gamma"))            # same RNG stream, same order
        assert np.array_equal(got["lambda"], st.arr("lambda_"))
        assert np.array_equal(got["validation"], m.validation_pairs())
        assert np.array_equal(got["links"], st.arr("links"))
        assert np.array_equal(got["tl"], st.arr("tl"))
        # files the constructor writes: identical to what the REFERENCE wrote for the same flags
        assert open(os.path.join(got["outdir"], "validation-edges.txt")).read() == \
            golden_text(case, "validation-edges.txt")
        want = golden_text(case, "param.txt").split("\n")
        have = open(os.path.join(got["outdir"], "param.txt")).read().split("\n")
        have = [l for l in have if l]
        assert have == want[:len(have)] and len(have) >= 50
        assert os.path.islink(os.path.join(got["outdir"], "network.dat"))
        m.close(); g.close()


def test_astroph_ingest_matches_oracle(cli):
    with Scratch() as d:
        ent, got = run_dump(cli, "c2_m12", d)
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Model(g, ent["k"])
        assert got["links"].shape == (195988, 2)                        # training links of config 2
        assert np.array_equal(got["links"], m.state.arr("links"))
        assert np.array_equal(got["gamma"], m.state.arr("gamma"))
        assert np.array_equal(got["validation"], m.validation_pairs())
        m.close(); g.close()


def test_ragged_input_selfloops_duplicates_and_single_nodes(cli):
    """Self-loops and repeated / reversed pairs are dropped, ids are arbitrary integers, and -n larger than
    the number of distinct ids pads the graph with 'single' nodes (network.cc:107-113) that inference skips."""
    with Scratch() as d:
        lines = ["10\t20", "20\t10", "10\t10", "30\t20", "10\t20", "7\t30", "7\t10", "99\t99", "5\t7"]
        open(os.path.join(d, "g.txt"), "w").write("\n".join(lines) + "\n")
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", "g.txt", "-n", "9", "-k", "3", "-link-sampling", "-accuracy",
                               "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        links = np.fromfile(os.path.join(dump, "links.u32"), dtype=np.uint32).reshape(-1, 2)
        gamma = np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, 3)
        g = orc.Graph.read(os.path.join(d, "g.txt"), 9)
        assert (g.n, g.singles, g.ones) == (6, 3, 5)
        m = orc.Model(g, 3, accuracy=1)
        assert gamma.shape == (6, 3)
        assert np.array_equal(links, m.state.arr("links")) and np.array_equal(gamma, m.state.arr("gamma"))
        # node 99 only has a self-loop: it exists, has no links and an all-zero initial gamma row
        assert np.all(gamma[4] == 0)                                    # ids in first-appearance order: 99 -> seq 4
        m.close(); g.close()


def test_init_communities_startup_state(cli):
    """-init-communities <file> (init_gamma_external, linksampling.cc:404-452; Network::load_init_communities,
    network.cc:374-437): gamma[p] = alpha + one normalised vector added once per adjacency entry; no RNG.
    init_memberships.txt must equal the reference's byte for byte, the start-up gamma a literal restatement."""
    with Scratch() as d:
        ent, got = run_dump(cli, "c1_initcomm_m10", d)
        assert open(os.path.join(got["outdir"], "init_memberships.txt")).read() == golden_text("c1_initcomm_m10", "init_memberships.txt")
        n, k, alpha = 75, 4, 0.25
        g = orc.Graph.read(input_path(ent["input"], d), n)
        id2seq = {int(v): i for i, v in enumerate(g.seq2id[:n])}
        member = [[] for _ in range(n)]
        for cid, line in enumerate(open(input_path("assort-75-4-init-communities.txt", d))):
            for tok in line.split():
                member[id2seq[int(tok)]].append(cid)
        want = np.full((n, k), alpha)
        for p_ in range(n):
            phi = np.full(k, alpha)
            for c in member[p_]:
                phi[c] += n / len(member[p_])
            s_ = 0.0
            for c in range(k):
                s_ += phi[c]
            phi = phi / s_
            for _ in range(int(g.adj_off[p_ + 1] - g.adj_off[p_])):
                want[p_] += phi
        assert np.array_equal(got["gamma"], want)
        assert np.array_equal(got["lambda"], np.ones((k, 2)))
        g.close()


def test_nmi_ground_truth_files_equal_the_references(cli):
    """-nmi <file>: Network::load_ground_truth + write_gt_communities (network.cc:254-307, 508-536) -- ground_truth.txt
    and ground_truth_community_sizes.txt as the reference wrote them for the LFR example (tests/golden/lfr_nmi/, from
    `svinet_ref ... -nmi LFR-ground-truth-n1000-k28.txt`)."""
    ent = MANIFEST["lfr_k28_m20"]
    with Scratch() as d:
        for name in (ent["input"], "LFR-ground-truth-n1000-k28.txt"):
            inp, local = input_path(name, d), os.path.join(d, name)
            if not os.path.exists(local):
                os.symlink(inp, local)
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", ent["input"], "-n", "1000", "-k", "28", "-link-sampling", "-nmi",
                               "LFR-ground-truth-n1000-k28.txt", "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        out = os.path.join(d, ent["outdir"])
        for f in ("ground_truth.txt", "ground_truth_community_sizes.txt"):
            assert open(os.path.join(out, f)).read() == golden_text("lfr_nmi", f), f


def test_cli_refuses_other_engines(cli):
    with Scratch() as d:
        p = subprocess.run([cli, "-file", "x", "-n", "5", "-k", "2", "-batch"], cwd=d, capture_output=True)
        assert p.returncode != 0 and b"are implemented here" in p.stderr


# ---- -rnode -stratified (class FastAMM2) ------------------------------------------------------------
@pytest.mark.parametrize("case", ["fa2_c1_m200", "fa2_c1_k6_seed9_m500", "fa2_lfr_k28_m300"])
def test_fa2_startup_state_is_bit_identical_to_oracle(cli, case):
    """shuffle_nodes, the held-out draw, init_gamma / init_lambda (fastamm2.cc:489-531) consume the same
    mt19937 stream as the oracle (and, through its fixtures, the reference)."""
    from test_oracle_fa2_golden import fa2_opts
    ent = MANIFEST[case]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        local = os.path.join(d, ent["input"])
        if not os.path.exists(local):
            os.symlink(inp, local)
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        cmd = [cli, "-file", ent["input"], "-n", str(ent["n"]), "-k", str(ent["k"])] + ent["mode"].split() + \
              ent["flags"] + ["-dump-init", dump]
        subprocess.check_call(cmd, cwd=d, stdout=subprocess.DEVNULL)
        k = ent["k"]
        g = orc.Graph.read(inp, ent["n"])
        m = orc.Fa2Model(g, k, **fa2_opts(ent["flags"]))
        assert np.array_equal(np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, k), m.gamma)
        assert np.array_equal(np.fromfile(os.path.join(dump, "lambda.f64")).reshape(k, 2), m.lambda_)
        assert np.array_equal(np.fromfile(os.path.join(dump, "shuffled.u32"), dtype=np.uint32), m.shuffled)
        assert np.array_equal(np.fromfile(os.path.join(dump, "heldout.u32"), dtype=np.uint32).reshape(-1, 2),
                              m.heldout_pairs(sorted_=False))
        out = os.path.join(d, ent["outdir"])
        assert open(os.path.join(out, "heldout-pairs.txt")).read() == golden_text(case, "heldout-pairs.txt")
        want = golden_text(case, "param.txt").split("\n")
        have = [l for l in open(os.path.join(out, "param.txt")).read().split("\n")]
        while have and not have[-1]:
            have.pop()
        # everything the constructor logs, i.e. all but the final "maxiterations reached" line
        assert have == want[:len(have)] and len(have) >= 60, (have[-3:], want[len(have) - 3:len(have)])
        m.close(); g.close()


def test_ingest_sort_dedup_matches_oracle_on_messy_input(cli):
    """The sort-based duplicate removal and the flat / hashed id tables keep the reference's semantics: first
    occurrence wins, either direction counts as a repeat, first-appearance sequence ids, insertion-ordered
    adjacency (oracle graph == reference Network::read, pinned by the fixtures)."""
    rng = np.random.default_rng(12)
    n_ids = 300
    ids = np.concatenate([rng.integers(0, 5000, n_ids - 20), rng.integers(3_000_000_000, 4_000_000_000, 20)])
    a = ids[rng.integers(0, n_ids, 6000)]
    b = ids[rng.integers(0, n_ids, 6000)]
    with Scratch() as d:
        with open(os.path.join(d, "g.txt"), "w") as f:
            for x, y in zip(a, b):
                f.write("%d%s%d\n" % (x, "\t" if (x + y) % 3 else " ", y))
        n_distinct = len(np.unique(np.concatenate([a, b])))
        for n_arg in (n_distinct, n_distinct + 7, n_distinct - 25):      # exact, padded with singles, too small
            dump = os.path.join(d, "dump%d" % n_arg); os.makedirs(dump)
            subprocess.check_call([cli, "-file", "g.txt", "-n", str(n_arg), "-k", "3", "-link-sampling", "-accuracy",
                                   "-label", "t%d" % n_arg, "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
            links = np.fromfile(os.path.join(dump, "links.u32"), dtype=np.uint32).reshape(-1, 2)
            gamma = np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, 3)
            g = orc.Graph.read(os.path.join(d, "g.txt"), n_arg)
            m = orc.Model(g, 3, accuracy=1)
            assert links.shape[0] == g.ones
            assert np.array_equal(links, m.state.arr("links")) and np.array_equal(gamma, m.state.arr("gamma"))
            m.close(); g.close()


def test_fa2_resume_from_saved_model(cli):
    """-rnode -stratified -load <dir/> (FastAMM2::load_model, fastamm2.cc:1717-1803): the run starts from exactly
    the %.5f text of a saved gamma.txt / lambda.txt."""
    from test_oracle_fa2_golden import fa2_opts
    ent = MANIFEST["fa2_c1_m200"]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        if not os.path.exists(os.path.join(d, ent["input"])):
            os.symlink(inp, os.path.join(d, ent["input"]))
        saved = os.path.join(d, "saved")
        os.makedirs(saved)
        for f in ("gamma.txt", "lambda.txt"):
            open(os.path.join(saved, f), "w").write(golden_text("fa2_c1_m200", f))
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", ent["input"], "-n", "75", "-k", "4", "-rnode", "-stratified", "-load", "saved/",
                               "-label", "resumed", "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        gam = np.fromfile(os.path.join(dump, "gamma.f64")).reshape(75, 4)
        lam = np.fromfile(os.path.join(dump, "lambda.f64")).reshape(4, 2)
        want_g = np.array([[float(x) for x in l.split("\t")[2:]] for l in golden_text("fa2_c1_m200", "gamma.txt").strip().split("\n")])
        want_l = np.array([[float(x) for x in l.split("\t")[1:]] for l in golden_text("fa2_c1_m200", "lambda.txt").strip().split("\n")])
        assert np.array_equal(gam, want_g) and np.array_equal(lam, want_l)
        assert os.path.isdir(os.path.join(d, "n75-k4-resumed-Srnode"))


def test_fa2_load_heldout_file_and_single_nodes(cli):
    """-rnode -stratified -load-validation <file> (FastAMM2::load_heldout, fastamm2.cc:267-299) on a graph padded
    with single nodes: the pairs of the file (external ids, either order) become the held-out set, and no RNG
    draw is spent on it (the shuffled order and the initial gamma stay those of a run without held-out draw)."""
    with Scratch() as d:
        lines = ["1\t2", "2\t3", "3\t4", "4\t5", "5\t1", "1\t3", "2\t5", "6\t7", "7\t8", "8\t6", "6\t1", "9\t2"]
        open(os.path.join(d, "g.txt"), "w").write("\n".join(lines) + "\n")
        open(os.path.join(d, "held.txt"), "w").write("3\t1\n7\t6\n4\t9\n")       # two links, one non-link
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", "g.txt", "-n", "11", "-k", "3", "-rnode", "-stratified", "-load-validation",
                               "held.txt", "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        held = np.fromfile(os.path.join(dump, "heldout.u32"), dtype=np.uint32).reshape(-1, 2)
        g = orc.Graph.read(os.path.join(d, "g.txt"), 11)
        assert (g.n, g.singles) == (9, 2)
        id2seq = {int(g.seq2id[i]): i for i in range(g.n)}
        want = [tuple(sorted((id2seq[a], id2seq[b]))) for a, b in ((3, 1), (7, 6), (4, 9))]
        assert [tuple(r) for r in held.tolist()] == want
        gamma = np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, 3)
        assert gamma.shape == (9, 3) and np.all(gamma > 0)
        out = os.path.join(d, "n11-k3-mmsb-Srnode")             # named after the -n ARGUMENT, like the reference
        seq2id = [int(x) for x in g.seq2id[:g.n]]
        want_txt = "".join("%d\t%d\n" % (seq2id[a], seq2id[b]) for a, b in want) + "\n"
        assert open(os.path.join(out, "heldout-pairs.txt")).read() == want_txt
        g.close()


def test_link_sampling_resume_from_saved_model_text(cli):
    """-link-sampling -load <dir/> (linksampling.cc:1267-1352) through the threaded loader: starts from exactly the
    %.5f text of the reference's own gamma.txt / lambda.txt (fixture c2_m25: 17 903 rows); a truncated file is an
    error, not a partial load."""
    ent = MANIFEST["c2_m25"]
    with Scratch() as d:
        inp = input_path(ent["input"], d)
        if not os.path.exists(os.path.join(d, ent["input"])):
            os.symlink(inp, os.path.join(d, ent["input"]))
        saved = os.path.join(d, "saved"); os.makedirs(saved)
        gtxt, ltxt = golden_text("c2_m25", "gamma.txt"), golden_text("c2_m25", "lambda.txt")
        open(os.path.join(saved, "gamma.txt"), "w").write(gtxt)
        open(os.path.join(saved, "lambda.txt"), "w").write(ltxt)
        dump = os.path.join(d, "dump"); os.makedirs(dump)
        base = [cli, "-file", ent["input"], "-n", "17903", "-k", "20", "-link-sampling", "-label", "resumed"]
        subprocess.check_call(base + ["-load", "saved/", "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        gam = np.fromfile(os.path.join(dump, "gamma.f64")).reshape(17903, 20)
        want = np.array([[float(x) for x in l.split("\t")[2:]] for l in gtxt.strip().split("\n")])
        assert np.array_equal(gam, want)
        lam = np.fromfile(os.path.join(dump, "lambda.f64")).reshape(20, 2)
        assert np.array_equal(lam, np.array([[float(x) for x in l.split("\t")[1:]] for l in ltxt.strip().split("\n")]))
        # truncated gamma.txt
        open(os.path.join(saved, "gamma.txt"), "w").write("\n".join(gtxt.strip().split("\n")[:-5]) + "\n")
        p = subprocess.run(base + ["-load", "saved/", "-dump-init", dump], cwd=d, capture_output=True)
        assert p.returncode != 0 and b"rows" in p.stderr


# ---- K beyond the register tiles (the device side: svi_ls_wide.cuh / svi_fa2_wide.cuh) --------------------------------
def test_startup_state_at_large_k_both_modes(cli):
    """The host side has no K limit of its own below the reference's 65 535: start-up state at K = 1500
    (-link-sampling) and K = 700 (-rnode -stratified) bit-identical to the oracle's."""
    from test_oracle_fa2_golden import fa2_opts
    with Scratch() as d:
        inp = input_path("assort-75-4.txt", d)
        local = os.path.join(d, "assort-75-4.txt")
        if not os.path.exists(local):
            os.symlink(inp, local)
        g = orc.Graph.read(inp, 75)
        k = 1500
        dump = os.path.join(d, "dump_ls"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", "assort-75-4.txt", "-n", "75", "-k", str(k), "-link-sampling", "-dump-init", dump],
                              cwd=d, stdout=subprocess.DEVNULL)
        m = orc.Model(g, k)
        st = m.state
        assert np.array_equal(np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, k), st.arr("gamma"))
        assert np.array_equal(np.fromfile(os.path.join(dump, "lambda.f64")).reshape(k, 2), st.arr("lambda_"))
        assert np.array_equal(np.fromfile(os.path.join(dump, "links.u32"), dtype=np.uint32).reshape(-1, 2), st.arr("links"))
        m.close()
        k = 700
        dump = os.path.join(d, "dump_fa2"); os.makedirs(dump)
        subprocess.check_call([cli, "-file", "assort-75-4.txt", "-n", "75", "-k", str(k), "-rnode", "-stratified",
                               "-dump-init", dump], cwd=d, stdout=subprocess.DEVNULL)
        m2 = orc.Fa2Model(g, k, **fa2_opts([]))
        assert np.array_equal(np.fromfile(os.path.join(dump, "gamma.f64")).reshape(-1, k), m2.gamma)
        assert np.array_equal(np.fromfile(os.path.join(dump, "lambda.f64")).reshape(k, 2), m2.lambda_)
        m2.close(); g.close()
