"""Synthetic inputs of the shapes BASELINE.json names (random genomes with N-runs, .fai, reference-format
tables rebuilt from tests/golden/ref_tables.json).  Pure numpy; test and benchmark infrastructure (tests/, bench.py,
__graft_entry__.smoke()), not part of the product package."""
import json
import os

import numpy as np

from gargammel_b200.api import FaiEntry

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tables.json")
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_contig(rng, length, n_frac=0.005, n_run=(100, 10000), lower_frac=0.0):
    """iid uniform bases with about n_frac of positions inside runs of N (SURVEY 8d C2 inputs)."""
    seq = _BASES[rng.integers(0, 4, size=length, dtype=np.uint8)]
    if n_frac > 0 and length > 0:
        covered = 0
        target = int(length * n_frac)
        while covered < target:
            run = int(rng.integers(n_run[0], n_run[1] + 1))
            run = min(run, max(1, target - covered), length)
            start = int(rng.integers(0, max(1, length - run + 1)))
            seq[start:start + run] = ord("N")
            covered += run
    if lower_frac > 0:
        m = rng.random(length) < lower_frac
        seq = np.where(m, seq | 0x20, seq).astype(np.uint8)
    return seq


def fasta_bytes(contigs, width=60):
    """contigs: list of (name, uint8 array).  Returns (bytes, [FaiEntry]) laid out like `samtools faidx` expects."""
    parts, fai, off = [], [], 0
    for name, seq in contigs:
        hdr = (">" + name + "\n").encode()
        parts.append(hdr)
        off += len(hdr)
        n = len(seq)
        fai.append(FaiEntry(name, n, off, width, width + 1))
        full = n // width
        body = np.empty(n + full + (1 if n % width else 0), dtype=np.uint8)
        if full:
            blk = body[:full * (width + 1)].reshape(full, width + 1)
            blk[:, :width] = np.asarray(seq[:full * width]).reshape(full, width)
            blk[:, width] = 10
        if n % width:
            body[full * (width + 1):-1] = seq[full * width:]
            body[-1] = 10
        parts.append(body.tobytes())
        off += len(body)
    return b"".join(parts), fai


def make_genome(seed, n_contigs, total_len, prefix="chr", width=60, n_frac=0.005, n_run=(100, 10000), lower_frac=0.0):
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.integers(1, max(2, total_len), size=max(0, n_contigs - 1))) if n_contigs > 1 else np.array([], dtype=np.int64)
    edges = np.concatenate(([0], cuts, [total_len])).astype(np.int64)
    contigs = []
    for i in range(n_contigs):
        ln = int(edges[i + 1] - edges[i])
        if ln <= 0:
            ln = 1
        contigs.append(("%s%d" % (prefix, i + 1), random_contig(rng, ln, n_frac, n_run, lower_frac)))
    return fasta_bytes(contigs, width)


def write_fai(path, fai):
    with open(path, "w") as f:
        for e in fai:
            f.write("%s\t%d\t%d\t%d\t%d\n" % (e.name, e.len, e.offset, e.line_blen, e.line_len))


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def golden_matrix(name="single-"):
    m = golden()["matrices"][name]
    return np.array(m["sub5"], dtype=np.float64), np.array(m["sub3"], dtype=np.float64)


def golden_sizefreq():
    """(lens, cum) with cum accumulated exactly like fragSim.cpp:765 (running fp64 sum of the parsed text)."""
    rows = golden()["sizefreq"]
    lens = [int(r[0]) for r in rows]
    cum, c = [], 0.0
    for r in rows:
        c += float(r[1])
        cum.append(c)
    return lens, cum


def golden_sizelist():
    out = []
    for k, v in golden()["sizedist_hist"]:
        out.extend([int(k)] * int(v))
    return out


_SUBST = ["A->C", "A->G", "A->T", "C->A", "C->G", "C->T", "G->A", "G->C", "G->T", "T->A", "T->C", "T->G"]


def write_matrix_dat(prefix, sub5, sub3):
    """Write <prefix>5.dat / <prefix>3.dat in the reference's format (src/matrices/*.dat): header, then
    'pos<TAB>v [lo..hi]' x12; the 3' file lists positions -R+1..0 (i.e. reversed w.r.t. sub3)."""
    def w(path, rows, labels):
        with open(path, "w") as f:
            f.write("\t" + "\t".join(_SUBST) + "\n")
            for lab, r in zip(labels, rows):
                f.write(str(lab) + "\t" + "\t".join("%.17g [%.3e..%.3e]" % (v, v * 0.9, v * 1.1) for v in r) + "\n")
    s5, s3 = list(map(list, sub5)), list(map(list, sub3))
    w(prefix + "5.dat", s5, range(len(s5)))
    w(prefix + "3.dat", s3[::-1], range(-(len(s3) - 1), 1))


def write_sizefreq(path):
    with open(path, "w") as f:
        for a, b in golden()["sizefreq"]:
            f.write("%d\t%s\n" % (a, b))


def write_dnacomp(path, seed=3, max_pos=8):
    """A mapDamage-style dnacomp.txt (Chr End Std Pos A C G T Total) with position-specific base counts."""
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        f.write("# table produced by synth.write_dnacomp\n# comment\nChr\tEnd\tStd\tPos\tA\tC\tG\tT\tTotal\n")
        for end in ("3p", "5p"):
            for std in ("+", "-"):
                for pos in list(range(-max_pos, 0)) + list(range(1, max_pos + 1)):
                    w = np.array([0.3, 0.2, 0.2, 0.3]) if abs(pos) > 2 else rng.dirichlet([3, 3, 3, 3])
                    c = rng.multinomial(500000, w)
                    f.write("21\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (end, std, pos, c[0], c[1], c[2], c[3], c.sum()))


def golden_dnacomp():
    """The reference's src/dnacomp.txt parsed with --dist 2: [5p+, 5p-, 3p+, 3p-] x [4 positions][4 bases]."""
    return [np.array(t, dtype=np.float64) for t in golden()["dnacomp_dist2"]]
