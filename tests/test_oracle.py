"""CPU tests of the oracle itself: Philox known answers, golden tables, deterministic transforms,
structural properties of its output.  (No GPU.)"""
import collections
import os

import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth
from tests import util


def test_philox_known_answers(oracle_mod):
    # Random123 kat_vectors, philox4x32-10
    assert oracle_mod.philox_raw([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_mod.philox_raw([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_mod.philox_raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # counter layout: (id_lo, id_hi, block, stream), key = (seed_lo, seed_hi)
    assert oracle_mod.philox(0x299f31d0a4093822, 0x85a308d3243f6a88, 0x13198a2e, 0x03707344) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_golden_fixture_values(golden):
    """values that pin the matrix parser (column order, tokenisation, 3' reversal)."""
    s = golden["matrices"]["single-"]
    d = golden["matrices"]["double-"]
    assert len(s["sub5"]) == 81 and len(s["sub3"]) == 81
    assert s["sub5"][0][5] == 8.004e-2 and s["sub5"][0][6] == 1.236e-2 and s["sub5"][80][5] == 8.569e-3
    assert s["sub3"][0][5] == 0.341 and s["sub3"][0][6] == 3.039e-2 and s["sub3"][80][5] == 6.633e-3
    assert d["sub5"][0][5] == 2.649e-2 and d["sub3"][0][6] == 2.472e-2
    for name, mx in (("single-", (0.0898, 0.3617)), ("double-", (0.0379, 0.0363))):
        for key, m in zip(("sub5", "sub3"), mx):
            got = max(sum(r[3 * b:3 * b + 3]) for r in golden["matrices"][name][key] for b in range(4))
            assert abs(got - m) < 1e-4
    lens, cum = synth.golden_sizefreq()
    assert len(lens) == 157 and lens[0] == 35 and lens[-1] == 191 and abs(cum[-1] - 1.0) < 1e-6
    assert abs(sum(int(a) * float(b) for a, b in golden["sizefreq"]) - 72.4478) < 1e-3
    sl = synth.golden_sizelist()
    assert len(sl) == 50655 and min(sl) == 35 and max(sl) == 191 and int(np.median(sl)) == 64
    assert golden["mapdamage_LaBrana"]["rows"] == 70


def _oracle_ctx(oracle_mod, seed=42, **kw):
    ctx = oracle_mod.OracleContext(seed)
    ids = util.configure(ctx, util.tiny_genomes(), **kw)
    return ctx, ids


@pytest.mark.parametrize("emit", [api.EMIT_FRAG, api.EMIT_DEAM, api.EMIT_STDOUT4, api.EMIT_ARTS, api.EMIT_ARTP, api.EMIT_FR, api.EMIT_RR])
def test_oracle_record_grammar(oracle_mod, emit):
    ctx, ids = _oracle_ctx(oracle_mod)
    plan = util.plan3(ids)
    ctx.set_plan(plan)
    out, info = ctx.run_fused(0, 1200, emit)
    recs = util.records(out)
    per = 2 if emit == api.EMIT_STDOUT4 else 1
    assert len(recs) == 1200 * per and info.records == 1200 and info.bytes == len(out)
    genomes = util.tiny_genomes()
    for k, (hdr, seq) in enumerate(recs):
        fid = k // per
        rg = [r for r in plan if r.first_id <= fid < r.first_id + r.count][0]
        assert hdr.startswith(b">")
        name, strand, start, end, rest = hdr[1:].split(b":")
        suffix = b"_r" if (emit == api.EMIT_RR or (emit == api.EMIT_STDOUT4 and k % 2 == 1)) else b""
        assert rest.endswith(rg.tag.encode() + suffix)                       # tag glued on without separator (fragSim.cpp:1505)
        length = int(rest[:len(rest) - len(rg.tag) - len(suffix)])
        assert int(end) - int(start) == length and strand in (b"+", b"-") and 35 <= length <= 191
        fai = {e.name: e for e in genomes[ids.index(rg.genome)][1]}
        assert int(end) <= fai[name.decode()].len
        assert set(seq) <= set(b"ACGT")
        want = length if emit in (api.EMIT_FRAG, api.EMIT_DEAM) else (150 if emit == api.EMIT_ARTP else 75)
        assert len(seq) == want


def test_oracle_fragments_match_genome(oracle_mod):
    """Undamaged fragments are exact substrings (or reverse complements) of the upper-cased genome, never contain N."""
    ctx, ids = _oracle_ctx(oracle_mod)
    ctx.set_plan(util.plan3(ids))
    out, _ = ctx.run_fused(0, 1200, api.EMIT_FRAG)
    genomes = util.tiny_genomes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    seqs = {}
    for fa, fai in genomes:
        for e in fai:
            raw = fa[e.offset:e.offset + e.len + e.len // e.line_blen + 1].replace(b"\n", b"")[:e.len]
            seqs[e.name] = raw.upper()
    strands = collections.Counter()
    for hdr, seq in util.records(out):
        name, strand, start, end, _ = hdr[1:].split(b":")
        ref = seqs[name.decode()][int(start):int(end)]
        assert b"N" not in ref
        strands[strand] += 1
        assert seq == (ref if strand == b"+" else ref.translate(comp)[::-1])
    assert 450 < strands[b"+"] < 750        # randomBool (fragSim.cpp:1466)


def test_oracle_is_pure_function_of_id(oracle_mod):
    """Counter-based draws: any id range reproduces the same bytes as the corresponding slice of a full run."""
    ctx, ids = _oracle_ctx(oracle_mod, damage="briggs")
    ctx.set_plan(util.plan3(ids))
    full, _ = ctx.run_fused(0, 1200, api.EMIT_ARTP)
    a, _ = ctx.run_fused(0, 333, api.EMIT_ARTP)
    b, _ = ctx.run_fused(333, 867, api.EMIT_ARTP)
    assert a + b == full
    ctx.set_seed(43)
    other, _ = ctx.run_fused(0, 1200, api.EMIT_ARTP)
    assert other != full


@pytest.mark.parametrize("damage", ["matrix", "briggs", "briggs_ss", "single", "double", "none"])
def test_oracle_fused_equals_stage_composition(oracle_mod, damage):
    """fragSim | deamSim | adptSim over files == the fused run (same seed, ids = record index)."""
    ctx, ids = _oracle_ctx(oracle_mod, damage=damage)
    ctx.set_plan([api.Range(0, 700, ids[1], 0 if damage != "none" else -1, "e")])
    frag, _ = ctx.run_fused(0, 700, api.EMIT_FRAG)
    deam, _ = ctx.run_fused(0, 700, api.EMIT_DEAM)
    d2, _ = ctx.run_deam(frag, 0 if damage != "none" else -1, 0)
    assert d2 == deam
    for emit in (api.EMIT_STDOUT4, api.EMIT_ARTS, api.EMIT_ARTP, api.EMIT_FR, api.EMIT_RR):
        fused, _ = ctx.run_fused(0, 700, emit)
        a2, _ = ctx.run_adpt(deam, emit, 0)
        assert a2 == fused


def test_oracle_adapter_layout(oracle_mod):
    """adptSim.cpp:298-322,404: trimming, adapter append, random padding, -artp = fwd ++ revcomp(rev)."""
    ctx = oracle_mod.OracleContext(5)
    F, S = "ACGTTGCA" * 2, "TTGGCCAA"
    ctx.set_adapters(F, S, 30)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    frag = b"GATTACAGATTACACCGT"                       # 18 < 30: adapter (16) fits for F, S (8) needs 4 random bases
    recs = b">x:+:0:18:18\n" + frag + b"\n>y\n" + b"ACGTN" * 8 + b"\n"
    out4, _ = ctx.run_adpt(recs, api.EMIT_STDOUT4)
    r = util.records(out4)
    assert r[0][0] == b">x:+:0:18:18" and r[1][0] == b">x:+:0:18:18_r"
    assert r[0][1] == (frag + F.encode())[:30]
    rc = frag.translate(comp)[::-1]
    assert r[1][1][:26] == rc + S.encode() and len(r[1][1]) == 30 and set(r[1][1][26:]) <= set(b"ACGT")
    long_f, long_r = r[2][1], r[3][1]
    assert long_f == (b"ACGTN" * 8)[:30] and long_r == (b"ACGTN" * 8).translate(bytes.maketrans(b"ACGTN", b"TGCAN"))[::-1][:30]
    artp, _ = ctx.run_adpt(recs, api.EMIT_ARTP)
    ra = util.records(artp)
    assert ra[0][1] == r[0][1] + r[1][1].translate(comp)[::-1]
    arts, _ = ctx.run_adpt(recs, api.EMIT_ARTS)
    assert util.records(arts)[0][1] == r[0][1]
    fr, _ = ctx.run_adpt(recs, api.EMIT_FR)
    rr, _ = ctx.run_adpt(recs, api.EMIT_RR)
    assert util.records(fr) == [r[0], r[2]] and util.records(rr) == [r[1], r[3]]


def test_oracle_matrix_rates_analytic(oracle_mod):
    """C1: -matfile single- on fixed 50-bp fragments: per-position C->T equals the matrix entry
    (5' row i for i<25, 3' row 49-i for i>=25; deamSim.cpp:1805-1829,1873-1892) within +-0.002 at this n."""
    rng = np.random.default_rng(1)
    n, L = 200000, 50
    seqs = _BASES[rng.integers(0, 4, size=(n, L))]
    recs = b"".join(b">chrS:+:%d:%d:50\n%s\n" % (i * 50, i * 50 + 50, seqs[i].tobytes()) for i in range(n))
    ctx = oracle_mod.OracleContext(42)
    s5, s3 = synth.golden_matrix("single-")
    ctx.set_matrix(0, s5, s3)
    out, _ = ctx.run_deam(recs, 0)
    got = np.frombuffer(b"".join(s for _, s in util.records(out)), dtype=np.uint8).reshape(n, L)
    isC = seqs == ord("C")
    rate = ((got == ord("T")) & isC).sum(0) / isC.sum(0)
    want = np.array([s5[i][5] if i < 25 else s3[49 - i][5] for i in range(L)])
    assert np.abs(rate - want).max() < 0.004          # 3 sigma at n/4 = 50k is ~0.006 at p=0.34, ~0.001 elsewhere
    assert abs(rate[0] - 0.08004) < 0.004 and abs(rate[49] - 0.341) < 0.007


_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.mark.parametrize("L", [50, 7, 171])
def test_oracle_matrix_event_draws_follow_the_table(oracle_mod, L):
    """Draw specification v2 draws -matfile damage as events (terminal rows per base, the rest by geometric skips under a majorant
    and an acceptance draw).  The law must stay the reference's per-base categorical (deamSim.cpp:1794-1930): EVERY one of the 12
    substitution rates at EVERY position equals its matrix entry (5 sigma), rows beyond the table clamp to the last one, and
    substitutions at different positions stay independent."""
    rng = np.random.default_rng(L)
    n = 240000 if L <= 50 else 60000
    seqs = _BASES[rng.integers(0, 4, size=(n, L))]
    recs = b"".join(b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n))
    ctx = oracle_mod.OracleContext(1234 + L)
    s5, s3 = synth.golden_matrix("single-")
    ctx.set_matrix(0, s5, s3)
    out, _ = ctx.run_deam(recs, 0)
    got = np.frombuffer(b"".join(s for _, s in util.records(out)), dtype=np.uint8).reshape(n, L)
    worst = 0.0
    for i in range(L):
        row = s5[min(i, len(s5) - 1)] if i < L // 2 else s3[min(L - 1 - i, len(s3) - 1)]
        for b in range(4):
            was = seqs[:, i] == _BASES[b]
            nb = was.sum()
            for a in range(4):
                if a == b:
                    continue
                want = row[b * 3 + (a if a < b else a - 1)]
                obs = ((got[:, i] == _BASES[a]) & was).sum() / nb
                sd = np.sqrt(max(want * (1 - want), 1e-9) / nb)
                worst = max(worst, abs(obs - want) / sd)
                assert abs(obs - want) < 5.2 * sd + 1e-5, (i, b, a, obs, want)
    # independence: changed-at-i and changed-at-j are uncorrelated (adjacent interior positions and the two ends)
    ch = got != seqs
    for i, j in ((L // 2 - 1, L // 2), (0, L - 1), (min(9, L - 2), min(10, L - 1))):
        pi, pj, pij = ch[:, i].mean(), ch[:, j].mean(), (ch[:, i] & ch[:, j]).mean()
        assert abs(pij - pi * pj) < 5 * np.sqrt(max(pi * pj, 1e-9) / n) + 1e-6, (i, j, pi, pj, pij)


def test_oracle_sizefreq_distribution(oracle_mod):
    """-f sizefreq: the empirical length pmf follows the file (KS p > 0.01 as in north_star)."""
    from scipy import stats
    ctx, ids = _oracle_ctx(oracle_mod)
    ctx.set_plan([api.Range(0, 20000, ids[2], -1, "")])
    out, _ = ctx.run_fused(0, 20000, api.EMIT_FRAG)
    L = np.array([len(s) for _, s in util.records(out)])
    lens, cum = synth.golden_sizefreq()
    cdf = lambda x: np.interp(x, lens, cum)            # noqa: E731
    # discrete KS via chi-square on the pmf (lengths conditioned on acceptance ~ the pmf itself for short fragments)
    pmf = np.diff(np.concatenate(([0.0], cum)))
    obs = np.array([(L == l).sum() for l in lens])
    keep = pmf * len(L) > 5
    chi = ((obs[keep] - pmf[keep] * len(L)) ** 2 / (pmf[keep] * len(L))).sum()
    assert stats.chi2.sf(chi, keep.sum() - 1) > 0.001
    assert abs(L.mean() - 72.45) < 1.0


def _check_circ_records(out, seqs, circ_name, d=0):
    """--circ (fragSim.cpp:1407-1426): a window that runs off the circular contig is contig[idx-d : len-1] ++ contig[0:...]
    (the LAST base is skipped: temp1 has end-(coord-d) bases), cut to [d, d+length); deflines keep idx un-wrapped."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    wrapped = 0
    for hdr, seq in util.records(out):
        name, strand, start, end, ln = hdr[1:].split(b":")
        start, end, ln = int(start), int(end), int(ln)
        contig = seqs[name.decode()]
        assert end - start == ln and len(seq) == ln
        if end + d > len(contig):
            assert name.decode() == circ_name, "run-off outside the circular contig"
            n1 = len(contig) - 1 - (start - d)            # may be -1 (coord == end+1): temp1 = "", temp2 gets length+1 bases
            temp = (contig[start - d:len(contig) - 1] if n1 > 0 else b"") + contig[0:ln + 2 * d - n1]
            wrapped += 1
        else:
            temp = contig[start - d:end + d]
        if strand == b"-":
            temp = temp.translate(comp)[::-1]
        assert seq == temp[d:d + ln], hdr
    return wrapped


@pytest.mark.parametrize("circ", [1, 0, 3])
def test_oracle_circular_contig(oracle_mod, circ):
    fa, fai = util.circ_genome()
    seqs = util.contig_seqs(fa, fai)
    o = oracle_mod.OracleContext(3)
    gid = o.upload_genome(fa, fai, circ=circ)
    o.set_sizes(api.SIZE_FIXED, fixed_len=40)
    o.set_plan([api.Range(0, 4000, gid, -1, "")])
    out, _ = o.run_fused(0, 4000, api.EMIT_FRAG)
    wrapped = _check_circ_records(out, seqs, fai[circ].name)
    expect = 4000 * 40.0 / sum(e.len for e in fai)          # P(start within the last ~40 coordinates of the circular contig)
    assert 0.6 * expect < wrapped < 1.5 * expect
    o.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(util.ROOT, "oracle", "_ref", "src", "fragSim")), reason="oracle/_ref not built")
def test_reference_binary_circular_quirk(tmp_path):
    """Pins the oracle's reading of --circ on the reference's own code: the wrapped window drops the contig's last base."""
    import subprocess
    fa, fai = util.circ_genome()
    (tmp_path / "g.fa").write_bytes(fa)
    synth.write_fai(str(tmp_path / "g.fa.fai"), fai)
    ref = os.path.join(util.ROOT, "oracle", "_ref", "src", "fragSim")
    out = subprocess.run([ref, "--seed", "4", "-n", "3000", "-l", "40", "--circ", "m2", str(tmp_path / "g.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
    assert _check_circ_records(out, util.contig_seqs(fa, fai), "m2") > 30


@pytest.mark.skipif(not os.path.exists(os.path.join(util.ROOT, "oracle", "_ref", "src", "deamSim")), reason="oracle/_ref not built")
def test_oracle_damagess_vs_reference_binary(oracle_mod, tmp_path):
    """-damagess d,s5,s3,l5,l3 (deamSim.cpp:310-340,1337-1468): per-position C->T rate of the oracle against the reference's own
    binary (G->A never happens)."""
    import subprocess
    rng = np.random.default_rng(5)
    n, L = 120_000, 30
    seqs = synth._BASES[rng.integers(0, 4, size=(n, L))]
    recs = b"".join(b">r%d\n%s\n" % (i, seqs[i].tobytes()) for i in range(n))
    (tmp_path / "in.fa").write_bytes(recs)
    pars = (0.02, 0.5, 0.7, 0.25, 0.1)
    o = oracle_mod.OracleContext(9)
    o.set_briggs_ss(0, *pars)
    ours = o.run_deam(recs, 0)[0]
    ref = subprocess.run([os.path.join(util.ROOT, "oracle", "_ref", "src", "deamSim"), "--seed", "3", "-damagess", ",".join(map(str, pars)), str(tmp_path / "in.fa")],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True).stdout
    isC = seqs == ord("C")

    def ct(out):
        got = np.frombuffer(b"".join(out.split(b"\n")[1::2]), dtype=np.uint8).reshape(n, L)
        assert ((got != seqs) <= (isC & (got == ord("T")))).all()          # C->T only
        return ((got == ord("T")) & isC).sum(0) / isC.sum(0)
    a, b = ct(ours), ct(ref)
    sd = np.sqrt(np.maximum(a * (1 - a), 1e-4) * 2.0 / isC.sum(0).min())
    assert (np.abs(a - b) < 5 * sd).all(), (a, b)
    assert a[0] > 0.1 and a[-1] > 0.05 and a[L // 2] < a[-1]
