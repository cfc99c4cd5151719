"""The drop-in executables: argument conventions and error behaviour of the reference CLIs (CPU), byte-exact output
against the oracle for the same seed (GPU), statistical equivalence with the reference's own binaries (GPU + oracle/_ref)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth
from tests import util
from tests.conftest import ROOT

BIN = os.path.join(ROOT, "bin")


def run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    fa, fai = synth.make_genome(21, 4, 60000, prefix="chr", width=60, n_frac=0.01, n_run=(20, 300), lower_frac=0.1)
    (d / "g.fa").write_bytes(fa)
    synth.write_fai(str(d / "g.fa.fai"), fai)
    synth.write_sizefreq(str(d / "sf.size"))
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(d / "m-"), s5, s3)
    return d, fa, fai


@pytest.mark.parametrize("tool", ["fragSim", "deamSim", "adptSim", "gargammel-b200"])
def test_cli_usage_and_unknown_option(tool):
    """no args / -h => usage on stdout, exit 1; unknown option => message on stderr, exit 1 (fragSim.cpp:463-471,622-623)."""
    exe = os.path.join(BIN, tool)
    assert os.path.exists(exe), "run `make` first"
    r = run([exe]); assert r.returncode == 1 and len(r.stdout) > 100
    r = run([exe, "-h"]); assert r.returncode == 1 and len(r.stdout) > 100
    r = run([exe, "-zzz", "x", "input"])
    assert r.returncode == 1 and (b"unknown option -zzz" in r.stderr or b"Unknown option: -zzz" in r.stderr)


def test_cli_validation_errors(workdir):
    d, _, _ = workdir
    r = run([os.path.join(BIN, "fragSim"), "-l", "2000", str(d / "g.fa")])
    assert r.returncode == 1 and b"is longer than the maximum allowed fragment size" in r.stderr       # fragSim.cpp:638-643
    r = run([os.path.join(BIN, "fragSim"), "-l", "40", "-s", "x", str(d / "g.fa")])
    assert r.returncode == 1 and b"cannot specify -l and -s" in r.stderr
    r = run([os.path.join(BIN, "deamSim"), str(d / "g.fa")])
    assert r.returncode == 1 and b"Please specify the matrix to use or use the Briggs parameter model" in r.stderr   # deamSim.cpp:403
    r = run([os.path.join(BIN, "deamSim"), "-damage", "0.03,0.4,0.01", str(d / "g.fa")])
    assert r.returncode == 1 and b"Specify 4 comma-separated values" in r.stderr
    r = run([os.path.join(BIN, "deamSim"), "-mat", "triple", str(d / "g.fa")])
    assert r.returncode == 1 and b'Specify either "single" or "double"' in r.stderr
    r = run([os.path.join(BIN, "adptSim"), "-rr", "x.gz", str(d / "g.fa")])
    assert r.returncode == 1 and b"need to specify the output forward read" in r.stderr              # adptSim.cpp:179-182
    r = run([os.path.join(BIN, "gargammel-b200"), "--comp", "0.5,0.5", str(d)])
    assert r.returncode == 1 and b"3 comma-delimited fields" in r.stderr


def test_cli_circ_unknown_contig(workdir):
    d, _, _ = workdir
    r = run([os.path.join(BIN, "fragSim"), "-n", "5", "--circ", "nosuchchr", str(d / "g.fa")])
    assert r.returncode == 1 and b"Cannot find chromosome nosuchchr which is specified via --circ" in r.stderr   # fragSim.cpp:1175


def test_cli_without_gpu_fails_loudly(workdir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d, _, _ = workdir
    r = run([os.path.join(BIN, "fragSim"), "-n", "5", "-l", "30", str(d / "g.fa")])
    assert r.returncode == 1 and b"no CPU fallback" in r.stderr and r.stdout == b""


@pytest.mark.gpu
def test_cli_pipeline_matches_oracle(workdir, oracle_mod):
    """bin/fragSim | bin/deamSim | bin/adptSim with --seed == the oracle's stage composition, byte for byte."""
    d, fa, fai = workdir
    seed, n = 4242, 20000
    o = oracle_mod.OracleContext(seed)
    gid = o.upload_genome(fa, fai)
    o.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=0, maxlen=1000)
    util.set_damage(o, 0, "matrix")
    o.set_adapters(read_len=75)
    o.set_plan([api.Range(0, n, gid, 0, "e1_0")])
    r = run([os.path.join(BIN, "fragSim"), "--seed", str(seed), "-tag", "e1_0", "-n", str(n), "-m", "0", "-M", "1000", "-f", str(d / "sf.size"), str(d / "g.fa")])
    assert r.returncode == 0, r.stderr
    assert b"terminated succesfully, wrote %d sequences" % n in r.stderr
    util.assert_same(r.stdout, o.run_fused(0, n, api.EMIT_FRAG)[0], "bin/fragSim")
    (d / "f.fa").write_bytes(r.stdout)
    r2 = run([os.path.join(BIN, "deamSim"), "--seed", str(seed), "-matfile", str(d / "m-"), str(d / "f.fa")])
    assert r2.returncode == 0, r2.stderr
    util.assert_same(r2.stdout, o.run_fused(0, n, api.EMIT_DEAM)[0], "bin/deamSim")
    import gzip
    with gzip.open(str(d / "d.fa.gz"), "wb") as f:                 # the driver hands deamSim/adptSim gz files (gargammel.pl:1729,1847)
        f.write(r2.stdout)
    r3 = run([os.path.join(BIN, "adptSim"), "-l", "75", "-artp", str(d / "a.fa"), str(d / "d.fa.gz")])
    assert r3.returncode == 0, r3.stderr
    util.assert_same((d / "a.fa").read_bytes(), o.run_fused(0, n, api.EMIT_ARTP)[0], "bin/adptSim -artp")
    r4 = run([os.path.join(BIN, "adptSim"), "-l", "75", str(d / "d.fa.gz")])
    util.assert_same(r4.stdout, o.run_fused(0, n, api.EMIT_STDOUT4)[0], "bin/adptSim stdout")
    r5 = run([os.path.join(BIN, "adptSim"), "-l", "60", "-fr", str(d / "fr.gz"), "-rr", str(d / "rr.gz"), str(d / "d.fa.gz")])
    assert r5.returncode == 0, r5.stderr
    o.set_adapters(read_len=60)
    util.assert_same(gzip.open(str(d / "fr.gz")).read(), o.run_fused(0, n, api.EMIT_FR)[0], "-fr")
    util.assert_same(gzip.open(str(d / "rr.gz")).read(), o.run_fused(0, n, api.EMIT_RR)[0], "-rr")
    # -uniq: "_<count>" after every defline, then "_<tag>" (fragSim.cpp:119-135); the GPU stream carries no tag in this mode
    r7 = run([os.path.join(BIN, "fragSim"), "--seed", str(seed), "-tag", "e1_0", "-uniq", "-n", "3000", "-l", "25", "--norev", str(d / "g.fa")])
    assert r7.returncode == 0, r7.stderr
    o.set_sizes(api.SIZE_FIXED, fixed_len=25); o.set_norev(True); o.set_plan([api.Range(0, 3000, gid, -1, "")])
    seen, want = {}, []
    for h, s_ in util.records(o.run_fused(0, 3000, api.EMIT_FRAG)[0]):
        seen[h] = seen.get(h, 0) + 1
        want.append(h + b"_%d_e1_0\n" % seen[h] + s_ + b"\n")
    util.assert_same(r7.stdout, b"".join(want), "bin/fragSim -uniq")
    o.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=0, maxlen=1000); o.set_norev(False); o.set_plan([api.Range(0, n, gid, 0, "e1_0")])
    # Briggs through the CLI + gz output
    o.set_briggs(0, 0.03, 0.4, 0.01, 0.3)
    r6 = run([os.path.join(BIN, "deamSim"), "--seed", str(seed), "-damage", "0.03,0.4,0.01,0.3", "-o", str(d / "b.fa.gz"), str(d / "f.fa")])
    assert r6.returncode == 0, r6.stderr
    util.assert_same(gzip.open(str(d / "b.fa.gz")).read(), o.run_fused(0, n, api.EMIT_DEAM)[0], "bin/deamSim -damage -o")


@pytest.mark.gpu
def test_cli_fragsim_circ_matches_oracle(tmp_path, oracle_mod):
    """bin/fragSim --circ NAME == the oracle with that contig circular (fragSim.cpp:1171-1178,1407-1426)."""
    fa, fai = util.circ_genome()
    (tmp_path / "g.fa").write_bytes(fa)
    synth.write_fai(str(tmp_path / "g.fa.fai"), fai)
    o = oracle_mod.OracleContext(17)
    gid = o.upload_genome(fa, fai, circ=1)
    o.set_sizes(api.SIZE_FIXED, fixed_len=50)
    o.set_plan([api.Range(0, 5000, gid, -1, "")])
    r = run([os.path.join(BIN, "fragSim"), "--seed", "17", "-n", "5000", "-l", "50", "--circ", "m2", str(tmp_path / "g.fa")])
    assert r.returncode == 0, r.stderr
    util.assert_same(r.stdout, o.run_fused(0, 5000, api.EMIT_FRAG)[0], "bin/fragSim --circ")


@pytest.mark.gpu
def test_fused_driver_matches_oracle(tmp_path, oracle_mod):
    """bin/gargammel-b200 on an endo(x2)/cont/bact directory == oracle run of the same apportioned plan (1 and 2 workers)."""
    d = tmp_path / "in"
    files = {}
    for sub, names in (("endo", ["endo.1.fa", "endo.2.fa"]), ("cont", ["cont.1.fa", "cont.2.fa"]), ("bact", ["b1.fa", "b2.fa", "b3.fa"])):
        (d / sub).mkdir(parents=True)
        for k, nm in enumerate(names):
            fa, fai = synth.make_genome(hash((sub, k)) % 1000, 2 + k, 30000 + 7000 * k, prefix=sub[0] + "%d_" % k, n_frac=0.01, n_run=(10, 200))
            (d / sub / nm).write_bytes(fa)
            if k != 1:
                synth.write_fai(str(d / sub / (nm + ".fai")), fai)        # one file per dir has no .fai: built like `samtools faidx`
            files[(sub, nm)] = (fa, fai)
    (d / "bact" / "list").write_text("b2.fa\t0.5\nb1.fa\t0.2\nb3.fa\t0.3\n")
    synth.write_sizefreq(str(tmp_path / "sf.size"))
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(tmp_path / "m-"), s5, s3)
    seed, n = 777, 30000
    for gpus in (1,):
        r = run([os.path.join(BIN, "gargammel-b200"), "--comp", "0.3,0.1,0.6", "-n", str(n), "-f", str(tmp_path / "sf.size"), "-matfile", str(tmp_path / "m-"),
                 "-rl", "75", "--seed", str(seed), "--gpus", str(gpus), "--keep", "-o", str(tmp_path / "sim"), str(d)])
        assert r.returncode == 0, r.stderr
        lib = api.load_library()
        tot = (C.c_uint64 * 3)(); e = (C.c_uint64 * 2)(); c = (C.c_uint64 * 2)(); b = (C.c_uint64 * 3)(); err = C.create_string_buffer(256)
        clen = [sum(x.len for x in files[("cont", nm)][1]) for nm in ("cont.1.fa", "cont.2.fa")]
        assert lib.ggh_apportion(C.c_uint64(n), C.c_double(0.3), C.c_double(0.1), C.c_double(0.6), 2, (C.c_uint64 * 2)(*clen), 2,
                                 (C.c_double * 3)(0.2, 0.5, 0.3), 3, C.c_uint64(seed), tot, e, c, b, err, 256) == 0
        assert abs(b[1] / tot[2] - 0.5) < 0.03                       # weights follow the list by NAME
        o = oracle_mod.OracleContext(seed)
        o.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=0, maxlen=1000)
        util.set_damage(o, 0, "matrix"); util.set_damage(o, 1, "matrix")
        o.set_adapters(read_len=75)
        plan, first = [], 0
        for key, cnt, tag, slot in ([(("endo", "endo.1.fa"), e[0], "e1_0", 0), (("endo", "endo.2.fa"), e[1], "e2_0", 0)] +
                                    [(("bact", "b%d.fa" % (k + 1)), b[k], "b%d" % (k + 1), 1) for k in range(3)] +
                                    [(("cont", "cont.%d.fa" % (k + 1)), c[k], "c%d" % (k + 1), -1) for k in range(2)]):
            if cnt:
                plan.append(api.Range(first, int(cnt), o.upload_genome(*files[key]), slot, tag)); first += int(cnt)
        o.set_plan(plan)
        util.assert_same((tmp_path / "sim_a.fa").read_bytes(), o.run_fused(0, first, api.EMIT_ARTP)[0], "gargammel-b200 _a.fa")
        import gzip
        util.assert_same(gzip.open(str(tmp_path / "sim_d.fa.gz")).read(), o.run_fused(0, first, api.EMIT_DEAM)[0], "_d.fa.gz")
        frag = o.run_fused(0, first, api.EMIT_FRAG)[0]
        got = b"".join(gzip.open(str(tmp_path / ("sim.%s.fa.gz" % k))).read() for k in "ebc")
        util.assert_same(got, frag, ".e/.b/.c.fa.gz")
        # -uniq: "<defline>_<count>_<tag>" with one counter per source file (gargammel.pl:1506-1507; fragSim.cpp:119-135)
        r = run([os.path.join(BIN, "gargammel-b200"), "--comp", "0.3,0.1,0.6", "-n", str(n), "-f", str(tmp_path / "sf.size"), "-matfile", str(tmp_path / "m-"),
                 "-rl", "75", "--seed", str(seed), "-uniq", "-o", str(tmp_path / "simu"), str(d)])
        assert r.returncode == 0, r.stderr
        import re
        plain = util.records((tmp_path / "sim_a.fa").read_bytes()); uq = util.records((tmp_path / "simu_a.fa").read_bytes())
        assert len(plain) == len(uq)
        seen = {}
        for (h0, s0), (h1, s1) in zip(plain, uq):
            m = re.match(rb"^(>.*:\d+)(e1_0|e2_0|b\d+|c\d+)$", h0)
            key = (m.group(2), m.group(1)); seen[key] = seen.get(key, 0) + 1
            assert s0 == s1 and h1 == m.group(1) + b"_%d_" % seen[key] + m.group(2)


def _ref(name):
    from oracle import oracle
    return oracle.ref_bin(name)


@pytest.mark.gpu
@pytest.mark.skipif(_ref("deamSim") is None, reason="oracle/_ref not built")
def test_statistical_equivalence_with_reference_binaries(workdir, tmp_path):
    """Level 3 (north_star): our binaries vs the reference's own (its sources + our shim): per-position C->T / G->A rates
    within +-0.002 of the matrix (ours, n=2M) and of each other within sampling error; KS on lengths p > 0.01."""
    from scipy import stats
    d, _, _ = workdir
    rng = np.random.default_rng(1)
    L = 50
    s5, s3 = synth.golden_matrix("single-")
    want_ct = np.array([s5[i][5] if i < 25 else s3[49 - i][5] for i in range(L)])
    want_ga = np.array([s5[i][6] if i < 25 else s3[49 - i][6] for i in range(L)])

    def rates(exe, n, seed):
        seqs = synth._BASES[rng.integers(0, 4, size=(n, L))]
        p = tmp_path / ("in_%d.fa" % n)
        with open(p, "wb") as f:
            f.write(b"".join(b">chrS:+:%d:%d:50\n%s\n" % (i * 50, i * 50 + 50, seqs[i].tobytes()) for i in range(n)))
        r = run([exe, "--seed", str(seed), "-matfile", str(d / "m-"), str(p)])
        assert r.returncode == 0, r.stderr
        got = np.frombuffer(b"".join(r.stdout.split(b"\n")[1::2]), dtype=np.uint8).reshape(n, L)
        isC, isG = seqs == ord("C"), seqs == ord("G")
        return ((got == ord("T")) & isC).sum(0) / isC.sum(0), ((got == ord("A")) & isG).sum(0) / isG.sum(0), isC.sum(0)
    ct, ga, nC = rates(os.path.join(BIN, "deamSim"), 1_000_000, 42)              # BASELINE config 1
    assert np.abs(ct - want_ct).max() < 0.002 + 3 * np.sqrt(0.34 * 0.66 / nC.min()) * (want_ct.max() > 0.3)
    assert np.abs(ct[:-1] - want_ct[:-1]).max() < 0.002 and np.abs(ga - want_ga).max() < 0.002
    rct, rga, rnC = rates(_ref("deamSim"), 300_000, 42)
    sd = np.sqrt(want_ct * (1 - want_ct) * (1.0 / nC + 1.0 / rnC))
    assert (np.abs(ct - rct) < 5 * sd + 1e-4).all()
    # -damagess (single-strand model): ours vs the reference's binary, per-position C->T
    def rates_ss(exe, n, seed):
        seqs = synth._BASES[rng.integers(0, 4, size=(n, 30))]
        p = tmp_path / ("ss_%d.fa" % n)
        with open(p, "wb") as f:
            f.write(b"".join(b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n)))
        r = run([exe, "--seed", str(seed), "-damagess", "0.02,0.5,0.7,0.25,0.1", str(p)])
        assert r.returncode == 0, r.stderr
        got = np.frombuffer(b"".join(r.stdout.split(b"\n")[1::2]), dtype=np.uint8).reshape(n, 30)
        isC = seqs == ord("C")
        assert ((got != seqs) <= (isC & (got == ord("T")))).all()
        return ((got == ord("T")) & isC).sum(0) / isC.sum(0), isC.sum(0)
    a_ss, na = rates_ss(os.path.join(BIN, "deamSim"), 400_000, 42)
    b_ss, nb = rates_ss(_ref("deamSim"), 200_000, 42)
    assert (np.abs(a_ss - b_ss) < 5 * np.sqrt(np.maximum(a_ss * (1 - a_ss), 1e-4) * (1.0 / na + 1.0 / nb))).all(), (a_ss, b_ss)
    # lengths and strands: fragSim -f sizefreq
    ours = run([os.path.join(BIN, "fragSim"), "--seed", "5", "-n", "200000", "-f", str(d / "sf.size"), str(d / "g.fa")])
    ref = run([_ref("fragSim"), "--seed", "5", "-n", "60000", "-f", str(d / "sf.size"), str(d / "g.fa")])
    assert ours.returncode == 0 and ref.returncode == 0
    lo = np.array([len(s) for _, s in util.records(ours.stdout)]); lr = np.array([len(s) for _, s in util.records(ref.stdout)])
    assert stats.ks_2samp(lo, lr).pvalue > 0.01
    po = np.mean([h.split(b":")[1] == b"+" for h, _ in util.records(ours.stdout)])
    pr = np.mean([h.split(b":")[1] == b"+" for h, _ in util.records(ref.stdout)])
    assert abs(po - 0.5) < 0.01 and abs(pr - 0.5) < 0.02
    # --loc/--scale: ours (inverse-CDF table) vs the reference's std::lognormal_distribution
    ours_ln = run([os.path.join(BIN, "fragSim"), "--seed", "6", "-n", "100000", "--loc", "4.1", "--scale", "0.35", str(d / "g.fa")])
    ref_ln = run([_ref("fragSim"), "--seed", "6", "-n", "50000", "--loc", "4.1", "--scale", "0.35", str(d / "g.fa")])
    assert ours_ln.returncode == 0 and ref_ln.returncode == 0, ours_ln.stderr
    a = np.array([len(s) for _, s in util.records(ours_ln.stdout)]); b = np.array([len(s) for _, s in util.records(ref_ln.stdout)])
    assert stats.ks_2samp(a, b).pvalue > 0.01 and abs(a.mean() - np.exp(4.1 + 0.35 ** 2 / 2) + 0.5) < 1.0
    # --comp/--dist: base composition at the fragment ends, ours vs the reference (chi-square homogeneity of the first/last base)
    synth.write_dnacomp(str(d / "dnacomp.txt"), seed=11)
    oc = run([os.path.join(BIN, "fragSim"), "--seed", "8", "-n", "120000", "-l", "40", "--comp", str(d / "dnacomp.txt"), "--dist", "2", str(d / "g.fa")])
    rc = run([_ref("fragSim"), "--seed", "8", "-n", "40000", "-l", "40", "--comp", str(d / "dnacomp.txt"), "--dist", "2", str(d / "g.fa")])
    assert oc.returncode == 0 and rc.returncode == 0, oc.stderr + rc.stderr
    for pos in (0, 1, -2, -1):
        ta = np.array([[s[pos] == c for c in b"ACGT"] for _, s in util.records(oc.stdout)]).sum(0)
        tb = np.array([[s[pos] == c for c in b"ACGT"] for _, s in util.records(rc.stdout)]).sum(0)
        assert stats.chi2_contingency(np.array([ta, tb]))[1] > 0.001, (pos, ta, tb)
    # -gc: GC-biased survival, ours vs the reference (KS on the per-fragment GC fraction)
    og = run([os.path.join(BIN, "fragSim"), "--seed", "9", "-n", "60000", "-l", "60", "-gc", "8", str(d / "g.fa")])
    rg = run([_ref("fragSim"), "--seed", "9", "-n", "20000", "-l", "60", "-gc", "8", str(d / "g.fa")])
    assert og.returncode == 0 and rg.returncode == 0, og.stderr + rg.stderr
    assert b"GC content: mean:" in og.stderr
    fa_ = np.array([(s.count(b"G") + s.count(b"C")) / 60 for _, s in util.records(og.stdout)]); fb_ = np.array([(s.count(b"G") + s.count(b"C")) / 60 for _, s in util.records(rg.stdout)])
    assert stats.ks_2samp(fa_, fb_).pvalue > 0.001 and abs(fa_.mean() - fb_.mean()) < 0.004
    plain = run([os.path.join(BIN, "fragSim"), "--seed", "8", "-n", "120000", "-l", "40", str(d / "g.fa")])
    t0 = np.array([[s[0] == c for c in b"ACGT"] for _, s in util.records(plain.stdout)]).sum(0)
    t1 = np.array([[s[0] == c for c in b"ACGT"] for _, s in util.records(oc.stdout)]).sum(0)
    assert stats.chi2_contingency(np.array([t0, t1]))[1] < 1e-6            # the composition file really changes the ends
    # start positions are uniform over the accepted windows in both: compare coarse histograms of chr1 starts
    def starts(out):
        return np.array([int(h.split(b":")[2]) for h, _ in util.records(out) if h.startswith(b">chr1:")])
    assert stats.ks_2samp(starts(ours.stdout), starts(ref.stdout)).pvalue > 0.001


def _three_dirs(t, n_bact=2):
    d = t / "in"
    for sub, names in (("endo", ["endo.1.fa", "endo.2.fa"]), ("cont", ["cont.1.fa"]), ("bact", ["b%d.fa" % (k + 1) for k in range(n_bact)])):
        (d / sub).mkdir(parents=True)
        for k, nm in enumerate(names):
            fa, fai = synth.make_genome(3 * len(sub) + k, 2, 30000, prefix=sub[0] if sub == "endo" else sub[0] + str(k))
            (d / sub / nm).write_bytes(fa); synth.write_fai(str(d / sub / (nm + ".fai")), fai)
    (d / "bact" / "list").write_text("".join("b%d.fa\t%g\n" % (k + 1, 1.0 / n_bact) for k in range(n_bact)))
    return d


def test_fused_driver_mock_prints_plan_and_art_hand_off(tmp_path):
    """--mock (gargammel.pl:71-84): nothing runs, the plan and the commands after the kernel are printed; the ART command line and the
    final gzips are the reference's (gargammel.pl:1856-1889), -ss/-qs/-qs2 are passed through (:460-462).  Runs without a GPU."""
    d = _three_dirs(tmp_path)
    art = tmp_path / "art_illumina"; art.write_text("#!/bin/sh\nexit 0\n"); art.chmod(0o755)
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "--comp", "0.3,0.1,0.6", "-n", "1000", "-l", "40", "-damage", "0.03,0.4,0.01,0.3", "-ss", "MSv3", "-qs", "-3", "-qs2", "2",
             "--art", str(art), "-o", str(tmp_path / "sim"), str(d)])
    err = r.stderr.decode()
    assert r.returncode == 0, err
    assert "600\tendogenous fragments" in err and "100\tcontaminant fragments" in err and "300\tbacterial fragments" in err
    assert err.count("fused range ids") == 5 and "tag e1_0" in err and "tag c1" in err
    art_cmd = [l for l in err.splitlines() if l.startswith("running cmd " + str(art))]
    assert len(art_cmd) == 1
    toks = art_cmd[0].split()
    assert toks[3:] == ["-ss", "MSv3", "-amp", "-na", "-p", "-i", str(tmp_path / "sim") + "_a.fa", "-l", "75", "-c", "1", "-qs", "-3", "-qs2", "2", "-o", str(tmp_path / "sim") + "_s"]
    for f in ("_a.fa", "_s1.fq", "_s2.fq"):
        assert "running cmd gzip -f " + str(tmp_path / "sim") + f in err
    assert not os.path.exists(str(tmp_path / "sim") + "_a.fa")
    # single-end: no -p, one fastq
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "-se", "-n", "10", "--art", str(art), "-o", str(tmp_path / "s2"), str(d)])
    err = r.stderr.decode()
    assert r.returncode == 0 and " -p " not in err and "gzip -f " + str(tmp_path / "s2") + "_s.fq" in err
    # a missing --art is an error like the reference's fileExists (gargammel.pl:212-218,245-248); without one the run stops after _a.fa
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "-n", "10", "--art", str(tmp_path / "nope"), str(d)])
    assert r.returncode == 1 and b"Cannot find file" in r.stderr


def test_fused_driver_coverage_counts_follow_gargammel_pl(tmp_path):
    """-c: the endogenous count is int(coverage * sumE/2 / size) and stays that; contaminant and bacterial counts come from
    n = int(nE / compE) as int(n * comp / sum) (gargammel.pl:1236-1258)."""
    d = _three_dirs(tmp_path)
    cov, compB, compC, compE, L = 0.37, 0.3, 0.1, 0.6, 41
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "--noart", "--comp", "%g,%g,%g" % (compB, compC, compE), "-c", str(cov), "-l", str(L), str(d)])
    err = r.stderr.decode()
    assert r.returncode == 0, err
    sumE = 0
    for nm in ("endo.1.fa", "endo.2.fa"):
        sumE += sum(int(l.split("\t")[1]) for l in open(str(d / "endo" / (nm + ".fai"))))
    nE = int(cov * ((sumE / 2) / L)); n = int(nE / compE)
    nC = int(n * compC / (compE + compC + compB)); nB = int(n * compB / (compE + compC + compB))
    assert "%d\tendogenous fragments" % nE in err and "%d\tcontaminant fragments" % nC in err and "%d\tbacterial fragments" % nB in err


def _run_fused_driver(tmp_path, d, name, extra):
    synth.write_sizefreq(str(tmp_path / "sf.size"))
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(tmp_path / "m-"), s5, s3)
    r = run([os.path.join(BIN, "gargammel-b200"), "--comp", "0.3,0.1,0.6", "-n", "20000", "-f", str(tmp_path / "sf.size"), "-matfile", str(tmp_path / "m-"),
             "-rl", "75", "--seed", "4242", "-o", str(tmp_path / name)] + extra + [str(d)])
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r.stderr.decode()


@pytest.mark.gpu
def test_fused_driver_art_hand_off_final_names(tmp_path):
    """With an art_illumina at hand the fused driver finishes like gargammel.pl:1856-1889: ART on <prefix>_a.fa, then <prefix>_a.fa.gz,
    <prefix>_s1.fq.gz and <prefix>_s2.fq.gz (tests/art_stub.py stands in for ART, which this image does not have)."""
    import gzip
    d = _three_dirs(tmp_path)
    art = tmp_path / "art_illumina"
    art.write_text("#!/bin/sh\nexec python3 %s \"$@\"\n" % os.path.join(ROOT, "tests", "art_stub.py")); art.chmod(0o755)
    _run_fused_driver(tmp_path, d, "plain", ["--noart"])
    err = _run_fused_driver(tmp_path, d, "sim", ["--art", str(art)])
    assert "Program finished successfully" in err
    assert not os.path.exists(str(tmp_path / "sim_a.fa"))
    amplicons = gzip.open(str(tmp_path / "sim_a.fa.gz")).read()
    assert amplicons == (tmp_path / "plain_a.fa").read_bytes() and amplicons.count(b"\n") == 2 * 20000
    for f in ("sim_s1.fq.gz", "sim_s2.fq.gz"):
        assert gzip.open(str(tmp_path / f)).read().count(b"\n") == 4 * 20000


@pytest.mark.gpu
def test_fused_driver_two_gpus_byte_identical(tmp_path):
    """--gpus 2: every rank takes a contiguous slice of the fragment ids; the concatenated outputs are the single-GPU bytes."""
    import gzip
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = _three_dirs(tmp_path)
    _run_fused_driver(tmp_path, d, "one", ["--noart", "--keep", "--gpus", "1"])
    _run_fused_driver(tmp_path, d, "two", ["--noart", "--keep", "--gpus", "2"])
    assert (tmp_path / "one_a.fa").read_bytes() == (tmp_path / "two_a.fa").read_bytes()
    for suffix in ("_d.fa.gz", ".e.fa.gz", ".b.fa.gz", ".c.fa.gz"):
        assert gzip.open(str(tmp_path / ("one" + suffix))).read() == gzip.open(str(tmp_path / ("two" + suffix))).read(), suffix
    # with ART at hand every rank also makes its slice of <prefix>_a.fa.gz on the device; the parts are concatenated in rank order
    art = tmp_path / "art_illumina"
    art.write_text("#!/bin/sh\nexec python3 %s \"$@\"\n" % os.path.join(ROOT, "tests", "art_stub.py")); art.chmod(0o755)
    _run_fused_driver(tmp_path, d, "art2", ["--art", str(art), "--gpus", "2"])
    assert gzip.open(str(tmp_path / "art2_a.fa.gz")).read() == (tmp_path / "one_a.fa").read_bytes()
    assert not os.path.exists(str(tmp_path / "art2_a.fa")) and not [f for f in os.listdir(str(tmp_path)) if ".part" in f]


@pytest.mark.gpu
def test_cli_deamsim_table_modes_match_oracle(workdir, tmp_path, oracle_mod):
    """bin/deamSim -profile / -mat single|double / -mapdamage f single|double / -last / -name through the CLI == the oracle with the tables
    the reference's loaders give (deamSim.cpp:453-668,956-1094,1098-1241,1941-2002); -v (the per-read log) is rejected, not swallowed."""
    from oracle import parsers
    d, fa, fai = workdir
    seed, n = 99, 6000
    o = oracle_mod.OracleContext(seed)
    gid = o.upload_genome(fa, fai)
    o.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=0, maxlen=1000)
    o.set_plan([api.Range(0, n, gid, -1, "t")])
    frags = o.run_fused(0, n, api.EMIT_FRAG)[0]
    (tmp_path / "f.fa").write_bytes(frags)
    # the reference resolves -mat through <dir of the executable>/matrices/ (deamSim.cpp:1105-1106): a src/ directory like gargammel's
    src = tmp_path / "src"; (src / "matrices").mkdir(parents=True)
    os.symlink(os.path.join(BIN, "deamSim"), src / "deamSim")
    s5, s3 = synth.golden_matrix("single-"); d5, d3 = synth.golden_matrix("double-")
    synth.write_matrix_dat(str(src / "matrices" / "single-"), s5, s3)
    synth.write_matrix_dat(str(src / "matrices" / "double-"), d5, d3)
    prof = os.path.join(ROOT, "tests", "golden", "sample_")
    md = os.path.join(ROOT, "tests", "golden", "mapdamage_sample.txt")
    p5, p3 = parsers.load_profile(prof)
    m5, m3 = parsers.load_mapdamage(md)
    cases = [
        (["-profile", prof], lambda: o.set_matrix(0, np.array(p5), np.array(p3))),
        (["-profile", prof, "-last", "-name"], lambda: (o.set_matrix(0, np.array(p5), np.array(p3), clamp_last=False), o.set_deam_naming(True))),
        (["-mat", "single"], lambda: o.set_singledouble(0, api.DMG_SINGLE, s5, s3)),
        (["-mat", "double", "-name"], lambda: (o.set_singledouble(0, api.DMG_DOUBLE, d5, d3), o.set_deam_naming(True))),
        (["-mapdamage", md, "double"], lambda: o.set_singledouble(0, api.DMG_DOUBLE, np.array(m5), np.array(m3))),
        (["-mapdamage", md, "single"], lambda: o.set_singledouble(0, api.DMG_SINGLE, np.array(m5), np.array(m3))),
    ]
    for args, setup in cases:
        o.set_deam_naming(False)
        setup()
        r = run([str(src / "deamSim"), "--seed", str(seed)] + args + [str(tmp_path / "f.fa")])
        assert r.returncode == 0, (args, r.stderr)
        want = o.run_deam(frags, 0)[0]
        util.assert_same(r.stdout, want, "bin/deamSim " + " ".join(args))
        assert r.stdout != frags
    r = run([str(src / "deamSim"), "-v", "-mat", "single", str(tmp_path / "f.fa")])
    assert r.returncode == 1 and b"not supported" in r.stderr


@pytest.mark.gpu
def test_cli_streams_large_and_untidy_inputs(tmp_path, oracle_mod):
    """bin/deamSim and bin/adptSim read their input in chunks through a reader thread (bounded memory whatever the file size), normalise
    what FastQParser accepts (multi-line sequences, CRLF, concatenated gzip members: gargammel.pl appends with >>) and keep the fragment ids
    running across chunks."""
    import gzip
    seed = 12
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(tmp_path / "m-"), s5, s3)
    o = oracle_mod.OracleContext(seed); o.set_matrix(0, s5, s3); o.set_adapters(read_len=50)      # (50 < adapter length: no random padding, adptSim has no --seed)
    # ---- untidy: wrapped sequence lines, CRLF, an empty sequence, two gzip members
    rng = np.random.default_rng(3)
    seqs = [synth._BASES[rng.integers(0, 4, int(n))].tobytes() for n in rng.integers(0, 300, 400)]
    tidy = b"".join(b">u%d\n%s\n" % (i, s) for i, s in enumerate(seqs))
    wrapped = b"".join(b">u%d\r\n" % i + b"".join(s[k:k + 37] + b"\r\n" for k in range(0, len(s), 37)) for i, s in enumerate(seqs))
    half = wrapped.index(b">u200\r\n")
    with open(str(tmp_path / "u.fa.gz"), "wb") as f:
        f.write(gzip.compress(wrapped[:half]) + gzip.compress(wrapped[half:]))
    r = run([os.path.join(BIN, "deamSim"), "--seed", str(seed), "-matfile", str(tmp_path / "m-"), str(tmp_path / "u.fa.gz")])
    assert r.returncode == 0, r.stderr
    util.assert_same(r.stdout, o.run_deam(tidy, 0)[0], "bin/deamSim on wrapped CRLF gz members")
    r = run([os.path.join(BIN, "adptSim"), "-l", "50", "-fr", str(tmp_path / "f.gz"), "-rr", str(tmp_path / "r.gz"), str(tmp_path / "u.fa.gz")])
    assert r.returncode == 0, r.stderr
    util.assert_same(gzip.open(str(tmp_path / "f.gz")).read(), o.run_adpt(tidy, api.EMIT_FR)[0], "bin/adptSim -fr")
    util.assert_same(gzip.open(str(tmp_path / "r.gz")).read(), o.run_adpt(tidy, api.EMIT_RR)[0], "bin/adptSim -rr")
    # ---- large: 2.1 GB of records through bin/deamSim with a resident set far below the input size
    n, L = 30_000_000, 60
    rec = np.empty((n, 1 + 8 + 1 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord(">"); rec[:, 9] = 10; rec[:, -1] = 10
    idx = np.arange(n, dtype=np.int64)
    for d in range(8):
        rec[:, 8 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, 10:10 + L] = synth._BASES[np.random.default_rng(4).integers(0, 4, size=(n, L), dtype=np.uint8)]
    big = str(tmp_path / "big.fa")
    rec.tofile(big)
    head = rec[:150_000].tobytes()
    del rec, idx
    with open(str(tmp_path / "big.out"), "wb") as fo:           # GG_PRINT_RSS: the process reports its own peak resident set (VmHWM) when the stream is done
        p = subprocess.run([os.path.join(BIN, "deamSim"), "--seed", str(seed), "-matfile", str(tmp_path / "m-"), big], stdout=fo, stderr=subprocess.PIPE, env=dict(os.environ, GG_PRINT_RSS="1"))
    assert p.returncode == 0, p.stderr
    assert b"wrote %d sequences" % n in p.stderr
    import re
    peak_kb = int(re.search(rb"\[stream done\] VmHWM:\s+(\d+) kB", p.stderr).group(1))
    assert os.path.getsize(str(tmp_path / "big.out")) == os.path.getsize(big)
    assert peak_kb < 1_000_000, "resident set %d kB for a %.1f GB input" % (peak_kb, os.path.getsize(big) / 1e9)
    with open(str(tmp_path / "big.out"), "rb") as f:
        util.assert_same(f.read(len(head)), o.run_deam(head, 0)[0], "first records of the large run")
        f.seek(-71 * 1000, 2); tail = f.read()
    want_tail = open(big, "rb").read()[-71 * 1000:]
    assert [l for l in tail.split(b"\n")[0::2]] == [l for l in want_tail.split(b"\n")[0::2]]           # same deflines at the end: nothing lost or reordered


@pytest.mark.gpu
def test_methylation_through_the_clis_and_the_fused_driver(tmp_path, oracle_mod):
    """--methyl: `bin/fragSim --case`, `bin/deamSim -matfilenonmeth -matfilemeth` and `bin/gargammel-b200 --methyl` against the
    oracle's composition of the three loops on genomes with lower-case (methylated) bases."""
    import gzip
    from tests.test_driver_real import lower_some
    d = tmp_path / "in"
    files = {}
    for sub, nm, seed in (("endo", "endo.1.fa", 31), ("cont", "cont.1.fa", 32), ("bact", "b1.fa", 33)):
        (d / sub).mkdir(parents=True)
        fa, fai = synth.make_genome(seed, 2, 40000, prefix=sub[0] + "_", n_frac=0.01, n_run=(10, 200))
        fa = lower_some(fa, fai, seed)
        (d / sub / nm).write_bytes(fa); synth.write_fai(str(d / sub / (nm + ".fai")), fai)
        files[sub] = (fa, fai)
    (d / "bact" / "list").write_text("b1.fa\t1.0\n")
    no5, no3, wi5, wi3 = util.meth_tables()
    synth.write_matrix_dat(str(tmp_path / "no-"), no5, no3); synth.write_matrix_dat(str(tmp_path / "wi-"), wi5, wi3)
    synth.write_sizefreq(str(tmp_path / "sf.size"))
    seed, n = 99, 20000
    # ---- the three filters chained like gargammel.pl does it
    r = run([os.path.join(BIN, "fragSim"), "--seed", str(seed), "--case", "-n", "5000", "-f", str(tmp_path / "sf.size"), "-tag", "e", "-o", str(tmp_path / "f.fa.gz"), str(d / "endo" / "endo.1.fa")])
    assert r.returncode == 0, r.stderr
    r = run([os.path.join(BIN, "deamSim"), "--seed", str(seed), "-matfilenonmeth", str(tmp_path / "no-"), "-matfilemeth", str(tmp_path / "wi-"), "-name", str(tmp_path / "f.fa.gz")])
    assert r.returncode == 0, r.stderr
    o = oracle_mod.OracleContext(seed)
    o.set_case(True)
    o.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=0, maxlen=1000)
    o.set_matrix_meth(0, no5, no3, wi5, wi3)
    gid = o.upload_genome(*files["endo"])
    o.set_plan([api.Range(0, 5000, gid, 0, "e")])
    frag = o.run_fused(0, 5000, api.EMIT_FRAG)[0]
    util.assert_same(gzip.open(str(tmp_path / "f.fa.gz")).read(), frag, "bin/fragSim --case")
    o.set_deam_naming(True)
    util.assert_same(r.stdout, o.run_deam(frag, 0)[0], "bin/deamSim -matfilenonmeth -matfilemeth -name")
    o.set_deam_naming(False)
    for bad in (["-matfilemeth", str(tmp_path / "wi-")], ["-matfile", str(tmp_path / "no-"), "-matfilenonmeth", str(tmp_path / "no-"), "-matfilemeth", str(tmp_path / "wi-")]):
        assert run([os.path.join(BIN, "deamSim")] + bad + [str(tmp_path / "f.fa.gz")]).returncode == 1       # deamSim.cpp:414-430
    # ---- the fused driver
    r = run([os.path.join(BIN, "gargammel-b200"), "--comp", "0.2,0.2,0.6", "-n", str(n), "-f", str(tmp_path / "sf.size"), "--methyl", "-matfilenonmeth", str(tmp_path / "no-"),
             "-matfilemeth", str(tmp_path / "wi-"), "-rl", "75", "--seed", str(seed), "--keep", "--noart", "-o", str(tmp_path / "sim"), str(d)])
    assert r.returncode == 0, r.stderr
    o.set_adapters(read_len=75)
    plan, first = [], 0
    for sub, cnt, tag in (("endo", int(n * 0.6), "e"), ("bact", int(n * 0.2), "b1"), ("cont", int(n * 0.2), "c1")):
        plan.append(api.Range(first, cnt, o.upload_genome(*files[sub]), 0, tag)); first += cnt
    o.set_plan(plan)
    util.assert_same((tmp_path / "sim_a.fa").read_bytes(), o.run_fused(0, first, api.EMIT_ARTP)[0], "gargammel-b200 --methyl _a.fa")
    util.assert_same(gzip.open(str(tmp_path / "sim_d.fa.gz")).read(), o.run_fused(0, first, api.EMIT_DEAM)[0], "--methyl _d.fa.gz")
    got = b"".join(gzip.open(str(tmp_path / ("sim.%s.fa.gz" % k))).read() for k in "ebc")
    util.assert_same(got, o.run_fused(0, first, api.EMIT_FRAG)[0], "--methyl .e/.b/.c.fa.gz keep the case")
    assert any(ch in got for ch in b"acgt")
    r = run([os.path.join(BIN, "gargammel-b200"), "--methyl", "-matfilemeth", str(tmp_path / "wi-"), "-o", str(tmp_path / "x"), str(d)])
    assert r.returncode == 1 and b"both" in r.stderr
    o.close()


@pytest.mark.gpu
def test_cli_bam_modes(workdir, tmp_path):
    """Unaligned BAM in/out: fragSim -b [-u], deamSim -b (BAM in, XD:Z tag with -name), adptSim -bs / -bp (flags 4 / 77+141,
    XA:Z tag).  Every BAM run must hold exactly the names and sequences of the FASTA run with the same seed, in the fields the
    reference fills (fragSim.cpp:1589-1601, deamSim.cpp:2016-2024, adptSim.cpp:328-373)."""
    t = workdir[0]
    seed = "31"
    frag_args = ["--seed", seed, "-n", "3000", "-f", str(t / "sf.size"), "-tag", "e1_0"]
    fa = run([os.path.join(BIN, "fragSim")] + frag_args + [str(t / "g.fa")])
    assert fa.returncode == 0, fa.stderr
    want = util.records(fa.stdout)
    for extra, name in ((["-b"], "f.bam"), (["-u", "-b"], "fu.bam")):
        r = run([os.path.join(BIN, "fragSim")] + frag_args + extra + [str(tmp_path / name), str(t / "g.fa")])
        assert r.returncode == 0, r.stderr
        text, recs = util.read_bam(str(tmp_path / name))
        assert "@PG\tID:fragSim\tPN:fragSim\tCL:" in text
        assert [(b">" + n, s) for n, _, s, _, _ in recs] == want
        assert all(f == 4 and q == bytes(len(s)) and tg == {} for _, f, s, q, tg in recs)
    assert os.path.getsize(tmp_path / "fu.bam") > 2 * os.path.getsize(tmp_path / "f.bam")       # -u stores, -b deflates (on the device)
    assert run([os.path.join(BIN, "fragSim")] + frag_args + ["-o", str(tmp_path / "x.gz"), "-b", str(tmp_path / "x.bam"), str(t / "g.fa")]).returncode == 1
    # ---- deamSim: BAM in, BAM out
    (tmp_path / "f.fa").write_bytes(fa.stdout)
    dm_args = ["--seed", seed, "-matfile", str(t / "m-"), "-name"]
    dfa = run([os.path.join(BIN, "deamSim")] + dm_args + [str(tmp_path / "f.fa")])
    assert dfa.returncode == 0, dfa.stderr
    r = run([os.path.join(BIN, "deamSim")] + dm_args + ["-b", str(tmp_path / "d.bam"), str(tmp_path / "f.bam")])
    assert r.returncode == 0, r.stderr
    text, recs = util.read_bam(str(tmp_path / "d.bam"))
    assert "ID:fragSim" in text and "ID:deamSim" in text
    n_tagged = 0
    for (hdr, seq), (name, flag, bseq, qual, tags), (h0, _) in zip(util.records(dfa.stdout), recs, want):
        assert bseq == seq and flag == 4 and name == h0[1:]                      # the read name stays, the list goes into XD:Z
        assert hdr == (h0 + b"_DEAM:" + tags["XD"].encode() if tags else h0)
        n_tagged += bool(tags)
    assert len(recs) == len(want) and n_tagged > 500
    assert run([os.path.join(BIN, "deamSim")] + dm_args + ["-b", str(tmp_path / "e.bam"), str(tmp_path / "f.fa")]).returncode == 1       # not a BAM file
    # ---- adptSim: BAM in, single-end / paired-end BAM out
    ad_args = ["--seed", seed, "-l", "60"]
    r = run([os.path.join(BIN, "adptSim")] + ad_args + ["-name", "-tag", "lib7", "-bp", str(tmp_path / "p.bam"), str(tmp_path / "d.bam")])
    assert r.returncode == 0, r.stderr
    (tmp_path / "d.fa").write_bytes(b"".join(b">" + n + b"\n" + s + b"\n" for n, _, s, _, _ in recs))
    four = util.records(run([os.path.join(BIN, "adptSim")] + ad_args + [str(tmp_path / "d.fa")]).stdout)
    text, prec = util.read_bam(str(tmp_path / "p.bam"))
    assert "ID:adptSim" in text and len(prec) == 2 * len(recs) == len(four)
    for k, ((hdr, seq), (name, flag, bseq, qual, tags)) in enumerate(zip(four, prec)):
        src = recs[k // 2]
        assert bseq == seq and len(seq) == 60 and qual == bytes(60) and name == src[0]        # both mates carry the input's name (adptSim.cpp:273-279)
        assert flag == (141 if k & 1 else 77)
        L0 = len(src[2])
        assert tags["XA"] == ("AD" if L0 < 60 else "") + "lib7"
    r = run([os.path.join(BIN, "adptSim")] + ad_args + ["-u", "-bs", str(tmp_path / "s.bam"), str(tmp_path / "d.bam")])
    assert r.returncode == 0, r.stderr
    _, srec = util.read_bam(str(tmp_path / "s.bam"))
    assert [(n, f, s) for n, f, s, _, _ in srec] == [(n, 4, s) for (n, _, s, _, _) in prec[0::2]] and all(tg == {} for *_, tg in srec)


def _cell_dirs(tmp_path, n_cells, ploidy):
    d = tmp_path / "cells_in"
    for sub in ("cont", "bact"):
        (d / sub).mkdir(parents=True)
    fa, fai = synth.make_genome(99, 2, 30000, prefix="c")
    (d / "cont" / "cont.1.fa").write_bytes(fa); synth.write_fai(str(d / "cont" / "cont.1.fa.fai"), fai)
    for i in range(n_cells):
        (d / "endo" / ("C%d" % i)).mkdir(parents=True)
        for k in range(ploidy):
            fa, fai = synth.make_genome(10 * i + k + 1, 2, 30000, prefix="e")
            nm = "endo.%d.fa" % (k + 1)
            (d / "endo" / ("C%d" % i) / nm).write_bytes(fa); synth.write_fai(str(d / "endo" / ("C%d" % i) / (nm + ".fai")), fai)
    return d


@pytest.mark.parametrize("ploidy", [1, 2])
def test_fused_driver_multi_cell_plan_follows_gargammel_pl(tmp_path, ploidy):
    """endo/C0 .. CN (gargammel.pl:1017-1087,1363-1403,1464-1470): every endogenous fragment picks a cell uniformly, cell 0 gets the
    script's forced extra fragment (one in haploid mode, one more for its first file in diploid mode), cells run in order with the tags
    e1_<cell> / e2_<cell> (haploid: e for every cell), then the contaminant.  --mock: runs without a GPU."""
    import re
    n_cells = 3
    d = _cell_dirs(tmp_path, n_cells, ploidy)
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "--noart", "--comp", "0,0.1,0.9", "-n", "30000", "-l", "40", "--seed", "5", "-o", str(tmp_path / "mc"), str(d)])
    err = r.stderr.decode()
    assert r.returncode == 0, err
    assert "Found 3 directories: C[0..C2]" in err and "27000\tendogenous fragments" in err and "3000\tcontaminant fragments" in err
    per_cell = [int(x) for x in re.findall(r"Will extract (\d+) from cell: ", err)]
    assert len(per_cell) == n_cells and sum(per_cell) == 27000 + ploidy            # nE+1 (haploid) / nE+2 (diploid), as the script
    assert all(abs(c - 9000) < 5 * (27000 * (1 / 3) * (2 / 3)) ** 0.5 + 3 for c in per_cell)
    rng = [(int(a), int(b), tag, path) for a, b, tag, path in re.findall(r"fused range ids \[(\d+),(\d+)\) tag (\S+) file (\S+)", err)]
    assert [t[2] for t in rng] == (["e1_0", "e2_0", "e1_1", "e2_1", "e1_2", "e2_2", "c1"] if ploidy == 2 else ["e", "e", "e", "c1"])
    assert [os.path.basename(os.path.dirname(t[3])) for t in rng[:-1]] == [("C%d" % (i // ploidy)) for i in range(n_cells * ploidy)]
    assert rng[0][0] == 0 and all(rng[k][1] == rng[k + 1][0] for k in range(len(rng) - 1)) and rng[-1][1] == 30000 + ploidy
    for i in range(n_cells):
        assert sum(b - a for a, b, _, p in rng[:-1] if os.path.basename(os.path.dirname(p)) == "C%d" % i) == per_cell[i]
    if ploidy == 2:                                                               # the coin flips of a cell split it about evenly
        for i in range(n_cells):
            e1 = rng[2 * i][1] - rng[2 * i][0]
            assert abs(e1 - per_cell[i] / 2) < 5 * (per_cell[i] / 4) ** 0.5 + 2
    # the reference's layout errors
    (d / "endo" / "stray.fa").write_bytes(b">x\nACGT\n")
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "-n", "10", str(d)])
    assert r.returncode == 1 and b"without any other fasta files" in r.stderr
    os.remove(str(d / "endo" / "stray.fa"))
    os.rename(str(d / "endo" / "C1"), str(d / "endo" / "C7"))
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "-n", "10", str(d)])
    assert r.returncode == 1 and b"Missing directory" in r.stderr and b"endo/C1" in r.stderr
    os.rename(str(d / "endo" / "C7"), str(d / "endo" / "cellB"))
    r = run([os.path.join(BIN, "gargammel-b200"), "--mock", "-n", "10", str(d)])
    assert r.returncode == 1 and b"does not correspond to the pattern CX" in r.stderr
