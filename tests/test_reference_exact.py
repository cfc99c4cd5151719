"""Level-2 parity: wherever the reference's filters are DETERMINISTIC, the oracle must reproduce the reference's own binaries
(oracle/_ref = its unmodified sources + our shim) byte for byte.  CPU only; the CUDA path is held bit-exact to the oracle by
tests/test_gpu_parity.py, so this closes the chain reference -> oracle -> kernels for record grammar, adapter layout, trimming,
reverse complement, -name/-tag/-uniq text and deterministic damage."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth
from tests import util


def _ref(name):
    p = os.path.join(util.ROOT, "oracle", "_ref", "src", name)
    return p if os.path.exists(p) else None


pytestmark = pytest.mark.skipif(_ref("adptSim") is None, reason="oracle/_ref not built")


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr
    return r


@pytest.fixture(scope="module")
def frags(tmp_path_factory):
    """fragments of 16..230 bases (with lower-case stretches): long enough that no adapter padding is drawn at read length 75"""
    d = tmp_path_factory.mktemp("exact")
    rng = np.random.default_rng(11)
    recs = []
    for i in range(400):
        n = int(rng.integers(16, 231))
        s = synth._BASES[rng.integers(0, 4, size=n)].tobytes()
        if i % 7 == 0:
            s = s[:n // 3] + s[n // 3:2 * n // 3].lower() + s[2 * n // 3:]
        recs.append(b">chr%d:%s:%d:%d:%d%s\n%s\n" % (i % 5 + 1, b"+-"[i % 2:i % 2 + 1], i * 13, i * 13 + n, n, b"e1_0" if i % 3 else b"", s))
    data = b"".join(recs)
    (d / "f.fa").write_bytes(data)
    with gzip.open(str(d / "f.fa.gz"), "wb") as f:
        f.write(data)
    return d, data


@pytest.mark.parametrize("opts", [[], ["-name"], ["-tag", "XY"], ["-name", "-tag", "lib7"], ["-f", "ACGTTGCA", "-s", "TTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTGGGGCCCC"]])
def test_adptsim_dialects_byte_exact(oracle_mod, frags, tmp_path, opts):
    """adptSim.cpp:281-322,324-411 with nothing left to chance (len + |adapter| >= read length): all four dialects."""
    d, data = frags
    o = oracle_mod.OracleContext(1)
    adF = opts[opts.index("-f") + 1] if "-f" in opts else api.DEFAULT_ADAPTER_F
    adS = opts[opts.index("-s") + 1] if "-s" in opts else api.DEFAULT_ADAPTER_S
    if "-f" in opts:                      # short forward adapter: keep only fragments that still need no padding
        keep = [(h, s) for h, s in util.records(data) if len(s) + len(adF) >= 75]
        data = b"".join(h + b"\n" + s + b"\n" for h, s in keep)
        (tmp_path / "f.fa").write_bytes(data)
        src = str(tmp_path / "f.fa")
    else:
        src = str(d / "f.fa.gz")          # the driver hands adptSim a .gz (gargammel.pl:1847)
    o.set_adapters(adF, adS, 75)
    o.set_adpt_naming("-name" in opts, opts[opts.index("-tag") + 1] if "-tag" in opts else "")
    assert run([_ref("adptSim"), "-l", "75"] + opts + [src]).stdout == o.run_adpt(data, api.EMIT_STDOUT4)[0]
    run([_ref("adptSim"), "-l", "75"] + opts + ["-arts", str(tmp_path / "s.fa"), src])
    assert (tmp_path / "s.fa").read_bytes() == o.run_adpt(data, api.EMIT_ARTS)[0]
    run([_ref("adptSim"), "-l", "75"] + opts + ["-artp", str(tmp_path / "p.fa"), src])
    assert (tmp_path / "p.fa").read_bytes() == o.run_adpt(data, api.EMIT_ARTP)[0]
    run([_ref("adptSim"), "-l", "75"] + opts + ["-fr", str(tmp_path / "fr.gz"), "-rr", str(tmp_path / "rr.gz"), src])
    assert gzip.open(str(tmp_path / "fr.gz")).read() == o.run_adpt(data, api.EMIT_FR)[0]
    assert gzip.open(str(tmp_path / "rr.gz")).read() == o.run_adpt(data, api.EMIT_RR)[0]


def _const_matrix(rows, ct=0.0, ga=0.0):
    m = np.zeros((rows, 12))
    m[:, 5] = ct        # C->T: from C (1) * 3 + (T=3 -> 2)
    m[:, 6] = ga        # G->A: from G (2) * 3 + 0
    return m


@pytest.mark.parametrize("ct,ga", [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0)])
@pytest.mark.parametrize("name", [False, True])
def test_deamsim_deterministic_matrix_byte_exact(oracle_mod, frags, tmp_path, ct, ga, name):
    """-matfile with rates 0 or 1 (deamSim.cpp:1794-1930,2026-2041): upper-casing, which bases change, and the _DEAM: position lists
    (+i+1 in the 5' half, -(len-i) in the 3' half, clamped rows beyond the table) are all deterministic."""
    d, data = frags
    s5, s3 = _const_matrix(30, ct, ga), _const_matrix(25, ct, ga)      # shorter than many fragments: the clamp is exercised
    synth.write_matrix_dat(str(tmp_path / "m-"), s5, s3)
    o = oracle_mod.OracleContext(5)
    o.set_matrix(0, s5, s3)
    o.set_deam_naming(name)
    got = run([_ref("deamSim"), "--seed", "3", "-matfile", str(tmp_path / "m-")] + (["-name"] if name else []) + [str(d / "f.fa.gz")]).stdout
    assert got == o.run_deam(data, 0)[0]
    if ct == 1.0:
        assert b"C" not in b"".join(s for _, s in util.records(got))


def test_deamsim_last_and_briggs_deterministic(oracle_mod, frags, tmp_path):
    d, data = frags
    s5, s3 = _const_matrix(12, 1.0, 1.0), _const_matrix(9, 1.0, 1.0)
    synth.write_matrix_dat(str(tmp_path / "m-"), s5, s3)
    o = oracle_mod.OracleContext(5)
    o.set_matrix(0, s5, s3, clamp_last=False)                              # -last: positions beyond the table are left alone
    o.set_deam_naming(True)
    assert run([_ref("deamSim"), "-last", "-name", "-matfile", str(tmp_path / "m-"), str(d / "f.fa")]).stdout == o.run_deam(data, 0)[0]
    # Briggs with an (almost surely) unbounded overhang and s = 1: every C -> T (deamSim.cpp:1483-1494); with l = 1, v = 0, d = 1, s = 0:
    # no overhang, no nick, every C -> T through the double-stranded branch (:1586-1592)
    for pars in ("0,0.000000001,0,1", "0,1,1,0"):
        v, l, dd, s = map(float, pars.split(","))
        o.set_briggs(0, v, l, dd, s)
        got = run([_ref("deamSim"), "-name", "-damage", pars, str(d / "f.fa")]).stdout
        assert got == o.run_deam(data, 0)[0] and b"C" not in b"".join(s_ for _, s_ in util.records(got))


def test_fragsim_forced_window_byte_exact(oracle_mod, tmp_path):
    """A contig exactly as long as the fragment leaves one admissible window: the defline grammar (fragSim.cpp:1489-1507), -tag without
    separator, -uniq counters (:119-135) and the multi-line FASTA gather (:305-329) become deterministic with --norev."""
    rng = np.random.default_rng(2)
    seq = synth.random_contig(rng, 83, 0.0)
    fa, fai = synth.fasta_bytes([("only", seq)], 17)
    (tmp_path / "g.fa").write_bytes(fa)
    synth.write_fai(str(tmp_path / "g.fa.fai"), fai)
    o = oracle_mod.OracleContext(9)
    gid = o.upload_genome(fa, fai)
    o.set_sizes(api.SIZE_FIXED, fixed_len=83); o.set_norev(True)
    o.set_plan([api.Range(0, 60, gid, -1, "e1_0")])
    got = run([_ref("fragSim"), "--seed", "1", "--norev", "-n", "25", "-l", "83", "-tag", "e1_0", str(tmp_path / "g.fa")]).stdout
    assert got == o.run_fused(0, 25, api.EMIT_FRAG)[0] == (b">only:+:0:83:83e1_0\n" + bytes(seq).upper() + b"\n") * 25
    got = run([_ref("fragSim"), "--seed", "1", "--norev", "-uniq", "-n", "5", "-l", "83", "-tag", "t", str(tmp_path / "g.fa")]).stdout
    assert got == b"".join(b">only:+:0:83:83_%d_t\n" % (k + 1) + bytes(seq).upper() + b"\n" for k in range(5))
    # both strands: two possible records, each must be one of the oracle's two
    o.set_norev(False)
    ours = set(util.records(o.run_fused(0, 60, api.EMIT_FRAG)[0]))
    ref = set(util.records(run([_ref("fragSim"), "--seed", "4", "-n", "60", "-l", "83", "-tag", "e1_0", str(tmp_path / "g.fa")]).stdout))
    assert ours == ref and len(ref) == 2


# ---------------------------------------------------------------------------------------------------------------------------
# Level 3 on the CPU: the stochastic damage models of the oracle against the reference's binaries (two independent random
# streams, so agreement is statistical: every per-position rate within 5 sigma of the pooled binomial error + 2e-4).
def _rates(out, seqs):
    n, L = seqs.shape
    got = np.frombuffer(b"".join(out.split(b"\n")[1::2]), dtype=np.uint8).reshape(n, L)
    isC, isG = seqs == ord("C"), seqs == ord("G")
    other = (got != seqs) & ~(isC & (got == ord("T"))) & ~(isG & (got == ord("A")))
    return ((got == ord("T")) & isC).sum(0) / isC.sum(0), ((got == ord("A")) & isG).sum(0) / isG.sum(0), other.sum(), isC.sum(0).min()


@pytest.mark.parametrize("model", ["mat_single", "mat_double", "mapdamage_single", "mapdamage_double", "briggs"])
def test_damage_models_statistical_vs_reference_binary(oracle_mod, tmp_path, model):
    import shutil
    from oracle import parsers
    rng = np.random.default_rng(17)
    n, L = 150_000, 36
    seqs = synth._BASES[rng.integers(0, 4, size=(n, L))]
    data = b"".join(b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n))
    (tmp_path / "in.fa").write_bytes(data)
    exe = str(tmp_path / "deamSim")
    shutil.copy(_ref("deamSim"), exe)                                   # -mat reads <dir of the executable>/matrices/ (deamSim.cpp:1105)
    os.makedirs(str(tmp_path / "matrices"))
    o = oracle_mod.OracleContext(23)
    if model.startswith("mat_"):
        kind = model.split("_")[1]
        s5, s3 = synth.golden_matrix(kind + "-")
        synth.write_matrix_dat(str(tmp_path / "matrices" / (kind + "-")), s5, s3)
        o.set_singledouble(0, api.DMG_SINGLE if kind == "single" else api.DMG_DOUBLE, s5, s3)
        args = ["-mat", kind]
    elif model.startswith("mapdamage_"):
        kind = model.split("_")[1]
        md = os.path.join(util.ROOT, "tests", "golden", "mapdamage_sample.txt")
        s5, s3 = parsers.load_mapdamage(md)
        o.set_singledouble(0, api.DMG_SINGLE if kind == "single" else api.DMG_DOUBLE, np.array(s5), np.array(s3))
        args = ["-mapdamage", md, kind]
    else:
        o.set_briggs(0, 0.03, 0.4, 0.01, 0.3)
        args = ["-damage", "0.03,0.4,0.01,0.3"]
    ref = run([exe, "--seed", "8"] + args + [str(tmp_path / "in.fa")]).stdout
    ours = o.run_deam(data, 0)[0]
    a_ct, a_ga, a_other, nmin = _rates(ours, seqs)
    b_ct, b_ga, b_other, _ = _rates(ref, seqs)
    assert a_other == 0 and b_other == 0                                 # these models only ever make C->T and G->A
    for a, b in ((a_ct, b_ct), (a_ga, b_ga)):
        sd = np.sqrt(np.maximum((a + b) / 2 * (1 - (a + b) / 2), 0) * 2.0 / nmin)
        assert (np.abs(a - b) < 5 * sd + 2e-4).all(), (model, a, b)
    assert max(a_ct.max(), a_ga.max()) > 0.02                             # the model does something
