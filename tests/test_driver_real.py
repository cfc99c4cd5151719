"""The reference's own driver, run FOR REAL (no --mock) over the drop-in binaries (GPU): gargammel.pl:1450-1889 composes
`fragSim ... | gzip >> P.e.fa.gz`, `deamSim ... P.e.fa.gz | gzip > P_d.fa.gz` (+ `>>` for bact / cont), `adptSim ... -artp P_a.fa P_d.fa.gz`,
art_illumina and the final gzips.  oracle/_ref/gargammel.pl is the unmodified script installed by oracle/Makefile (git-ignored, it travels
to the GPU box with the other reference artefacts); art_illumina is absent from this image, tests/art_stub.py stands in for it."""
import gzip
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from tests import synth
from tests import util
from tests.conftest import ROOT

REF_PL = os.path.join(ROOT, "oracle", "_ref", "gargammel.pl")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_PL) or shutil.which("perl") is None, reason="oracle/_ref/gargammel.pl or perl not available")

DEFLINE = re.compile(rb"^>([A-Za-z0-9_.]+):([+-]):(\d+):(\d+):(\d+)(e1_0|e2_0|b\d+|c\d+)$")


def make_layout(t, bins):
    """<t>/gg/gargammel.pl next to src/{fragSim,deamSim,adptSim} (gargammel.pl:232-248 resolves them from its own path)."""
    gg = t / "gg"
    (gg / "src").mkdir(parents=True)
    (gg / "art_src_MountRainier").mkdir()
    shutil.copy(REF_PL, gg / "gargammel.pl")
    for b in ("fragSim", "deamSim", "adptSim"):
        os.symlink(os.path.join(bins, b), gg / "src" / b)
    art = gg / "art_src_MountRainier" / "art_illumina"
    art.write_text("#!/bin/sh\nexec python3 %s \"$@\"\n" % os.path.join(ROOT, "tests", "art_stub.py")); art.chmod(0o755)
    return gg


def make_inputs(t):
    d = t / "in"
    seqs = {}                                              # (file tag, contig) -> bases
    for s_i, (sub, names, tags) in enumerate((("endo", ["endo.1.fa", "endo.2.fa"], ["e1_0", "e2_0"]), ("cont", ["cont.1.fa"], ["c1"]), ("bact", ["b1.fa", "b2.fa"], ["b1", "b2"]))):
        (d / sub).mkdir(parents=True)
        for k, nm in enumerate(names):
            # the two endogenous haplotypes must carry the same chromosome names (gargammel.pl checks the .fai files)
            fa, fai = synth.make_genome(10 * s_i + k + 1, 2, 60000, prefix=sub[0] + ("" if sub == "endo" else str(k)), n_frac=0.01, n_run=(20, 200))
            (d / sub / nm).write_bytes(fa); synth.write_fai(str(d / sub / (nm + ".fai")), fai)
            for cname, bases in util.contig_seqs(fa, fai).items():
                seqs.setdefault((sub[0], cname), []).append(bases)        # (which endogenous file gets tag e1_0 is readdir order, gargammel.pl:1105-1130)
    (d / "bact" / "list").write_text("b1.fa\t0.4\nb2.fa\t0.6\n")
    synth.write_sizefreq(str(t / "sf.size"))
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(t / "m-"), s5, s3)
    return d, seqs, (s5, s3)


def fasta_records(path):
    data = gzip.open(path, "rb").read()            # multi-member gzip / BGZF (the driver appends with >>)
    return util.records(data)


@pytest.mark.gpu
def test_reference_driver_runs_over_the_dropin_binaries(tmp_path):
    run_and_check(tmp_path, os.path.join(ROOT, "bin"))


def test_reference_driver_over_its_own_binaries(tmp_path):
    """The same run and the same checks over the reference's own filters (oracle/_ref/src, CPU): pins what the checks above expect."""
    run_and_check(tmp_path, os.path.join(ROOT, "oracle", "_ref", "src"))


def run_and_check(t, bins):
    gg = make_layout(t, bins)
    d, seqs, (s5, s3) = make_inputs(t)
    n = 6000
    (t / "out").mkdir()
    prefix = str(t / "out" / "sim")
    r = subprocess.run(["perl", str(gg / "gargammel.pl"), "--comp", "0.3,0.1,0.6", "-n", str(n), "-f", str(t / "sf.size"), "-matfile", str(t / "m-"),
                        "-rl", "75", "-o", prefix, str(d) + "/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Program finished successfully" in r.stderr
    # ---- the file set of gargammel.pl:1871-1889
    for suffix in (".e.fa.gz", ".b.fa.gz", ".c.fa.gz", "_d.fa.gz", "_a.fa.gz", "_s1.fq.gz", "_s2.fq.gz"):
        assert os.path.getsize(prefix + suffix) > 0, suffix
        assert subprocess.run(["gzip", "-t", prefix + suffix]).returncode == 0, suffix + " is not a valid gzip stream"
    e, b, c, dd = (fasta_records(prefix + s) for s in (".e.fa.gz", ".b.fa.gz", ".c.fa.gz", "_d.fa.gz"))
    # ---- counts = the apportioning of gargammel.pl:1256-1283 (int(n * comp); the two endogenous files share nE by coin flips)
    assert (len(e), len(c), len(b)) == (int(n * 0.6), int(n * 0.1), int(n * 0.3))
    assert len(dd) == len(e) + len(b) + len(c)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    tags = {"e": set(), "b": set(), "c": set()}
    for recs, kind in ((e, "e"), (b, "b"), (c, "c")):
        for hdr, seq in recs:
            m = DEFLINE.match(hdr)
            assert m, hdr
            name, strand, start, end, ln, tag = m.groups()
            start, end, ln = int(start), int(end), int(ln)
            assert tag[:1].decode() == kind and end - start == ln == len(seq)
            tags[kind].add(tag)
            refs = [g[start:end] for g in seqs[(kind, name.decode())]]
            assert any(seq == (ref if strand == b"+" else ref.translate(comp)[::-1]) for ref in refs), hdr   # the fragment is what a genome of its source holds there
    assert tags["e"] == {b"e1_0", b"e2_0"} and tags["b"] == {b"b1", b"b2"} and tags["c"] == {b"c1"}
    # ---- _d.fa.gz = deaminated endo, deaminated bact, untouched cont, in this order (gargammel.pl:640-648,1698-1830)
    ne, nb = len(e), len(b)
    assert [h for h, _ in dd] == [h for h, _ in e] + [h for h, _ in b] + [h for h, _ in c]
    assert dd[ne + nb:] == c                                                                  # the contaminant is recompressed untouched
    changed = sum(1 for (_, s0), (_, s1) in zip(e + b, dd[:ne + nb]) if s0 != s1)
    assert 0.25 * (ne + nb) < changed < 0.9 * (ne + nb)
    # last base: C -> T at the rate of the matrix' first 3' row (0.341), first base at 0.080 (deamSim.cpp:1805-1829,1873-1892)
    last = [(s0[-1:], s1[-1:]) for (_, s0), (_, s1) in zip(e + b, dd[:ne + nb])]
    nC = sum(1 for a, _ in last if a == b"C"); nT = sum(1 for a, z in last if a == b"C" and z == b"T")
    assert abs(nT / nC - s3[0][5]) < 5 * np.sqrt(0.341 * 0.659 / nC)
    first = [(s0[:1], s1[:1]) for (_, s0), (_, s1) in zip(e + b, dd[:ne + nb])]
    nC = sum(1 for a, _ in first if a == b"C"); nT = sum(1 for a, z in first if a == b"C" and z == b"T")
    assert abs(nT / nC - s5[0][5]) < 5 * np.sqrt(0.08 * 0.92 / nC)
    # ---- adptSim -artp amplicons: forward read ++ revcomp(reverse read), 2 x 75 (adptSim.cpp:396-411)
    a = fasta_records(prefix + "_a.fa.gz")
    assert [h for h, _ in a] == [h for h, _ in dd] and all(len(s) == 150 for _, s in a)
    for (_, frag), (_, amp) in list(zip(dd, a))[:500]:
        k = min(len(frag), 75)
        assert amp[:k] == frag[:k] and amp[150 - k:] == frag[len(frag) - k:]
    # ---- the ART stand-in saw every amplicon
    assert gzip.open(prefix + "_s1.fq.gz", "rb").read().count(b"\n") == 4 * len(a)


# ---------------------------------------------------------------------------------------------------------------------------
# gargammel.pl --methyl: fragSim --case for every source, deamSim -matfilenonmeth/-matfilemeth for every source
def lower_some(fa, fai, seed, frac=0.35):
    b = np.frombuffer(fa, dtype=np.uint8).copy()
    rng = np.random.default_rng(seed)
    seqb = np.zeros(len(b), dtype=bool)
    for e in fai:
        seqb[e.offset:e.offset + e.len + e.len // e.line_blen + 1] = True
    low = (rng.random(len(b)) < frac) & seqb & (b >= 65) & (b <= 90)
    return np.where(low, b | 0x20, b).astype(np.uint8).tobytes()


def raw_contigs(fa, fai):
    return {e.name: fa[e.offset:e.offset + e.len + e.len // e.line_blen + 1].replace(b"\n", b"")[:e.len] for e in fai}


@pytest.mark.gpu
def test_reference_driver_methyl_over_the_dropin_binaries(tmp_path):
    run_methyl(tmp_path, os.path.join(ROOT, "bin"))


def test_reference_driver_methyl_over_its_own_binaries(tmp_path):
    run_methyl(tmp_path, os.path.join(ROOT, "oracle", "_ref", "src"))


def run_methyl(t, bins):
    gg = make_layout(t, bins)
    d = t / "in"
    raw = {}
    for sub, nm, seed in (("endo", "endo.1.fa", 5), ("cont", "cont.1.fa", 6)):
        (d / sub).mkdir(parents=True)
        fa, fai = synth.make_genome(seed, 2, 50000, prefix=sub[0], n_frac=0.01, n_run=(20, 200))
        fa = lower_some(fa, fai, seed)
        (d / sub / nm).write_bytes(fa); synth.write_fai(str(d / sub / (nm + ".fai")), fai)
        raw[sub[0]] = raw_contigs(fa, fai)
    (d / "bact").mkdir()
    no5, no3, wi5, wi3 = util.meth_tables()
    synth.write_matrix_dat(str(t / "no-"), no5, no3); synth.write_matrix_dat(str(t / "wi-"), wi5, wi3)
    (t / "out").mkdir()
    prefix = str(t / "out" / "sim")
    n = 8000
    r = subprocess.run(["perl", str(gg / "gargammel.pl"), "--comp", "0,0.25,0.75", "-n", str(n), "-l", "48", "--methyl", "-matfilenonmeth", str(t / "no-"),
                        "-matfilemeth", str(t / "wi-"), "-rl", "75", "-se", "-o", prefix, str(d) + "/"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Program finished successfully" in r.stderr
    e, c, dd = (fasta_records(prefix + s) for s in (".e.fa.gz", ".c.fa.gz", "_d.fa.gz"))
    assert (len(e), len(c)) == (int(n * 0.75), int(n * 0.25)) and len(dd) == n
    comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    for recs, kind in ((e, "e"), (c, "c")):
        for hdr, seq in recs:
            m = re.match(rb"^>([A-Za-z0-9_.]+):([+-]):(\d+):(\d+):(\d+)", hdr)
            ref = raw[kind][m.group(1).decode()][int(m.group(3)):int(m.group(4))]
            assert seq == (ref if m.group(2) == b"+" else ref.translate(comp)[::-1]), hdr     # fragSim --case: the bytes of the file, case included
    assert any(ch in b"".join(s for _, s in e) for ch in b"acgt")
    # every source went through deamSim with the two tables (gargammel.pl:1698-1811): all bases upper-case afterwards, and a methylated
    # cytosine at the first base becomes T far more often than an unmethylated one
    src = e + c
    assert [h for h, _ in dd] == [h for h, _ in src]
    assert not any(ch in b"".join(s for _, s in dd) for ch in b"acgtn")
    for base, rate in ((b"c", wi5[0][5]), (b"C", no5[0][5])):
        k = [(s0[:1], s1[:1]) for (_, s0), (_, s1) in zip(src, dd) if s0[:1] == base]
        nT = sum(1 for _, z in k if z == b"T")
        assert len(k) > 150 and abs(nT / len(k) - rate) < 5 * np.sqrt(rate * (1 - rate) / len(k)) + 1e-3, (base, nT, len(k), rate)
    a = fasta_records(prefix + "_a.fa.gz")
    assert [h for h, _ in a] == [h for h, _ in dd] and all(len(s) == 75 for _, s in a)
