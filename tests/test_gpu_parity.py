"""GPU parity tests proper: the CUDA path (through the C-ABI) must be BIT-EXACT against the CPU oracle on the
same seeded inputs -- integer/byte work, so no tolerance anywhere."""
import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth
from tests import util

pytestmark = pytest.mark.gpu

ALL_EMITS = [api.EMIT_FRAG, api.EMIT_DEAM, api.EMIT_STDOUT4, api.EMIT_ARTS, api.EMIT_ARTP, api.EMIT_FR, api.EMIT_RR]


@pytest.fixture(scope="module")
def genomes():
    return util.tiny_genomes()


def _pair(oracle_mod, genomes, seed=42, **kw):
    g = api.Context(0, seed)
    o = oracle_mod.OracleContext(seed)
    ids = util.configure(g, genomes, **kw)
    assert util.configure(o, genomes, **kw) == ids
    return g, o, ids


def test_pack_genome_matches_fasta(genomes):
    """k_pack_genome vs the FASTA bytes read through the line structure (fetchSeq, fragSim.cpp:305-329)."""
    with api.Context(0, 1) as g:
        for fa, fai in genomes:
            gid = g.upload_genome(fa, fai)
            for ci, e in enumerate(fai):
                raw = fa[e.offset:e.offset + e.len + e.len // e.line_blen + 1].replace(b"\n", b"")[:e.len].upper()
                want = bytes(c if c in b"ACGT" else ord("N") for c in raw)
                assert g.genome_fetch(gid, ci, 0, e.len) == want
                assert g.genome_fetch(gid, ci, e.len // 3, min(77, e.len - e.len // 3)) == want[e.len // 3:e.len // 3 + 77]


@pytest.mark.parametrize("damage", ["matrix", "briggs", "briggs_ss", "single", "double"])
@pytest.mark.parametrize("sizes", ["freq", "list", "fixed", "lognormal"])
def test_fused_bit_exact(oracle_mod, genomes, damage, sizes):
    g, o, ids = _pair(oracle_mod, genomes, sizes=sizes, damage=damage)
    plan = util.plan3(ids)
    g.set_plan(plan); o.set_plan(plan)
    for emit in ALL_EMITS:
        a, ia = g.run_fused(0, 1200, emit)
        b, ib = o.run_fused(0, 1200, emit)
        util.assert_same(a, b, "emit=%d damage=%s sizes=%s" % (emit, damage, sizes))
        assert (ia.bytes, ia.records, ia.attempts, ia.gather_bytes) == (ib.bytes, ib.records, ib.attempts, ib.gather_bytes)
    g.close(); o.close()


def test_fused_id_ranges_and_seeds(oracle_mod, genomes):
    """Any id sub-range gives the corresponding slice (disjoint Philox counters = multi-GPU sharding); seeds differ."""
    g, o, ids = _pair(oracle_mod, genomes, seed=2024)
    plan = util.plan3(ids, counts=(1000, 700, 900))
    g.set_plan(plan); o.set_plan(plan)
    full, _ = g.run_fused(0, 2600, api.EMIT_ARTP)
    util.assert_same(full, o.run_fused(0, 2600, api.EMIT_ARTP)[0], "full")
    cuts = [0, 1, 255, 256, 257, 999, 1000, 1700, 2599, 2600]
    parts = [g.run_fused(a, b - a, api.EMIT_ARTP)[0] for a, b in zip(cuts[:-1], cuts[1:])]
    assert b"".join(parts) == full
    g.set_seed(2025); o.set_seed(2025)
    other, _ = g.run_fused(0, 2600, api.EMIT_ARTP)
    assert other != full
    util.assert_same(other, o.run_fused(0, 2600, api.EMIT_ARTP)[0], "seed 2025")
    empty, info = g.run_fused(5, 0, api.EMIT_ARTP)
    assert empty == b"" and info.bytes == 0


@pytest.mark.parametrize("read_len,adF,adS", [(30, "ACGTTGCA", "TTGGCC"), (75, api.DEFAULT_ADAPTER_F, api.DEFAULT_ADAPTER_S),
                                              (200, api.DEFAULT_ADAPTER_F, api.DEFAULT_ADAPTER_S), (16, "A", "C"), (1, "G", "T")])
def test_fused_adapters_trim_and_pad(oracle_mod, genomes, read_len, adF, adS):
    """Read shorter than, equal to and longer than fragment+adapter: trimming, adapter append, random padding (adptSim.cpp:298-322)."""
    g, o, ids = _pair(oracle_mod, genomes, read_len=read_len, adapters=(adF, adS), damage="matrix")
    plan = util.plan3(ids)
    g.set_plan(plan); o.set_plan(plan)
    for emit in ALL_EMITS[2:]:
        util.assert_same(g.run_fused(0, 1200, emit)[0], o.run_fused(0, 1200, emit)[0], "RL=%d emit=%d" % (read_len, emit))


@pytest.mark.parametrize("kw", [dict(sizes="fixed", fixed_len=0), dict(sizes="fixed", fixed_len=1), dict(sizes="fixed", fixed_len=16), dict(sizes="fixed", fixed_len=17),
                                dict(sizes="fixed", fixed_len=191), dict(sizes="fixed", fixed_len=700, maxlen=1000), dict(minlen=60, maxlen=90),
                                dict(norev=True), dict(clamp_last=False, sizes="fixed", fixed_len=191), dict(matrix="double-")])
def test_fused_edge_configs(oracle_mod, genomes, kw):
    g, o, ids = _pair(oracle_mod, genomes, **kw)
    plan = util.plan3(ids, counts=(300, 200, 250))
    g.set_plan(plan); o.set_plan(plan)
    for emit in (api.EMIT_FRAG, api.EMIT_DEAM, api.EMIT_ARTP, api.EMIT_STDOUT4):
        util.assert_same(g.run_fused(0, 750, emit)[0], o.run_fused(0, 750, emit)[0], "%r emit=%d" % (kw, emit))


def test_fused_many_contigs_and_long_names(oracle_mod):
    """Draft-assembly style input: hundreds of short contigs, long names, contig ends everywhere (run-off rejection)."""
    rng = np.random.default_rng(11)
    contigs = [("scaffold_%04d_len_with_a_rather_long_name|x" % i, synth.random_contig(rng, int(rng.integers(40, 400)), 0.02, (3, 30))) for i in range(300)]
    fa, fai = synth.fasta_bytes(contigs, width=70)
    g = api.Context(0, 9); o = oracle_mod.OracleContext(9)
    for c in (g, o):
        gid = c.upload_genome(fa, fai)
        c.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
        util.set_damage(c, 0, "briggs")
        c.set_adapters(read_len=75)
        c.set_plan([api.Range(100, 3000, gid, 0, "b1000")])
    for emit in (api.EMIT_FRAG, api.EMIT_ARTP):
        a, ia = g.run_fused(100, 3000, emit); b, ib = o.run_fused(100, 3000, emit)
        util.assert_same(a, b, "many contigs emit=%d" % emit)
        assert ia.attempts == ib.attempts and ia.attempts > 3000


@pytest.mark.parametrize("damage", ["matrix", "briggs", "briggs_ss", "single", "double", "none"])
def test_standalone_stages_bit_exact(oracle_mod, genomes, damage):
    """deamSim / adptSim as separate stages over the record stream the previous stage wrote, and fused == composition."""
    g, o, ids = _pair(oracle_mod, genomes, damage=damage)
    slot = 0 if damage != "none" else -1
    plan = [api.Range(0, 900, ids[1], slot, "e")]
    g.set_plan(plan); o.set_plan(plan)
    frag, _ = g.run_fused(0, 900, api.EMIT_FRAG)
    d_g, info = g.run_deam(frag, slot, 0)
    util.assert_same(d_g, o.run_deam(frag, slot, 0)[0], "deam " + damage)
    assert info.records == 900 and info.in_bytes == len(frag)
    util.assert_same(d_g, g.run_fused(0, 900, api.EMIT_DEAM)[0], "fused vs fragSim|deamSim")
    util.assert_same(g.run_deam(frag, slot, 12345)[0], o.run_deam(frag, slot, 12345)[0], "deam first_id")
    for emit in ALL_EMITS[2:]:
        a_g, _ = g.run_adpt(d_g, emit, 0)
        util.assert_same(a_g, o.run_adpt(d_g, emit, 0)[0], "adpt emit=%d" % emit)
        util.assert_same(a_g, g.run_fused(0, 900, emit)[0], "fused vs composition emit=%d" % emit)


def test_standalone_odd_inputs(oracle_mod):
    """Empty stream, missing final newline, lower case, N and IUPAC bases, 1-base and long sequences, long ids."""
    g = api.Context(0, 77); o = oracle_mod.OracleContext(77)
    for c in (g, o):
        util.set_damage(c, 0, "matrix"); util.set_damage(c, 1, "briggs"); util.set_damage(c, 2, "double")
        c.set_adapters("ACGTACGTAC", "TTTTGGGGCC", 40)
    rng = np.random.default_rng(5)
    recs = [b">only_one_base\nC\n", b">lower:+:1:9:8\nacgtnacg\n", b">iupac\nACGTRYKMNNacgtCCCCGGGG\n",
            b">" + b"x" * 300 + b"\n" + synth._BASES[rng.integers(0, 4, 3000)].tobytes() + b"\n", b">empty_seq\n\n"]
    recs += [b">r%d\n%s\n" % (i, synth._BASES[rng.integers(0, 4, int(rng.integers(1, 120)))].tobytes()) for i in range(500)]
    stream = b"".join(recs)
    for data in (b"", stream, stream[:-1], recs[0], recs[0][:-1]):
        for slot in (0, 1, 2, -1):
            util.assert_same(g.run_deam(data, slot)[0], o.run_deam(data, slot)[0], "deam slot=%d n=%d" % (slot, len(data)))
        for emit in ALL_EMITS[2:]:
            util.assert_same(g.run_adpt(data, emit)[0], o.run_adpt(data, emit)[0], "adpt emit=%d n=%d" % (emit, len(data)))
    with pytest.raises(api.GGError):
        g.run_deam(b">a\nACGT\n>b\n", 0)          # odd number of lines
    with pytest.raises(api.GGError):
        g.run_deam(b"a\nACGT\n", 0)               # no '>'


def test_errors_are_loud(genomes):
    g = api.Context(0, 1)
    with pytest.raises(api.GGError):
        g.run_fused(0, 10, api.EMIT_FRAG)                       # no plan
    fa, fai = synth.fasta_bytes([("allN", np.full(5000, ord("N"), dtype=np.uint8))])
    gid = g.upload_genome(fa, fai)
    g.set_sizes(api.SIZE_FIXED, fixed_len=40)
    g.set_plan([api.Range(0, 10, gid, -1, "")])
    with pytest.raises(api.GGError) as e:
        g.run_fused(0, 10, api.EMIT_FRAG)                       # fragSim.cpp:1361 would spin forever
    assert e.value.code == -6
    with pytest.raises(api.GGError):
        g.set_plan([api.Range(0, 10, 99, -1, "")])
    with pytest.raises(api.GGError):
        g.set_sizes(api.SIZE_FIXED, fixed_len=5000, maxlen=10000)
    with pytest.raises(api.GGError):
        g.set_adapters("AC GT", "ACGT", 75)                     # adapters are printable strings (any letters: the reference appends them as given)
    with pytest.raises(api.GGError):
        g.set_briggs(0, 1.5, 0.4, 0.01, 0.3)


def test_large_run_properties(oracle_mod):
    """A BASELINE-shaped run (C2 at 1/25 scale: 2 Mb sources, 400k fragments, --comp 0,0.08,0.92, sizefreq, single- matrix,
    2x75 artp): sampled id windows are bit-exact vs the oracle; the whole stream obeys the size-independent invariants."""
    srcs = [synth.make_genome(100 + i, 1 + i % 3, 2_000_000, prefix="s%d_" % i) for i in range(3)]
    g = api.Context(0, 42); o = oracle_mod.OracleContext(42)
    n = 400_000
    nC, nE = int(n * 0.08), int(n * 0.92)
    for c in (g, o):
        ids = [c.upload_genome(fa, fai) for fa, fai in srcs]
        c.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
        util.set_damage(c, 0, "matrix")
        c.set_adapters(read_len=75)
        c.set_plan([api.Range(0, nE // 2, ids[0], 0, "e1_0"), api.Range(nE // 2, nE - nE // 2, ids[1], 0, "e2_0"), api.Range(nE, nC, ids[2], -1, "c1")])
    tot = nE + nC
    out, info = g.run_fused(0, tot, api.EMIT_ARTP)
    assert info.records == tot and info.bytes == len(out) and info.launches == 2
    lines = out.split(b"\n")
    assert lines[-1] == b"" and len(lines) == 2 * tot + 1
    seqlen = np.array([len(x) for x in lines[1:-1:2]])
    assert (seqlen == 150).all()
    assert all(h.startswith(b">s") for h in lines[0:-1:2][:5000])
    for first, cnt in ((0, 3000), (nE // 2 - 1500, 3000), (nE - 1500, 3000), (tot - 3000, 3000)):
        want, _ = o.run_fused(first, cnt, api.EMIT_ARTP)
        got = b"\n".join(lines[2 * first:2 * (first + cnt)]) + b"\n"
        util.assert_same(got, want, "window %d" % first)
    # fragSim-only stream: record grammar + composition
    frag, _ = g.run_fused(0, tot, api.EMIT_FRAG)
    hdrs = frag.split(b"\n")[0:-1:2]
    assert sum(h.endswith(b"c1") for h in hdrs) == nC and sum(h.endswith(b"e1_0") for h in hdrs) == nE // 2


def test_fused_short_headers(oracle_mod):
    """One-letter contig names and no tag: headers shorter than 13 bytes take the masked read-modify-write path of the
    emit pass (two sequence lines may share a 16-byte group)."""
    rng = np.random.default_rng(3)
    fa, fai = synth.fasta_bytes([("a", synth.random_contig(rng, 900, 0.02, (3, 20))), ("b", synth.random_contig(rng, 700, 0.0))], width=50)
    for sizes, kw in (("fixed", dict(fixed_len=1)), ("fixed", dict(fixed_len=3)), ("fixed", dict(fixed_len=20)), ("freq", {})):
        g = api.Context(0, 5); o = oracle_mod.OracleContext(5)
        for c in (g, o):
            gid = c.upload_genome(fa, fai)
            if sizes == "fixed":
                c.set_sizes(api.SIZE_FIXED, **kw)
            else:
                c.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
            util.set_damage(c, 0, "matrix")
            c.set_adapters("ACGT", "TTGCA", 9)
            c.set_plan([api.Range(0, 2000, gid, 0, "")])
        for emit in ALL_EMITS:
            util.assert_same(g.run_fused(0, 2000, emit)[0], o.run_fused(0, 2000, emit)[0], "short headers %s %r emit=%d" % (sizes, kw, emit))


def test_run_fused_to_host_chunked(oracle_mod, genomes):
    """gg_run_fused_to_host (chunks through two device slabs, D2H overlapped) == one gg_run_fused call."""
    import ctypes as C
    g, o, ids = _pair(oracle_mod, genomes, damage="matrix")
    plan = util.plan3(ids, counts=(3000, 2500, 2000))
    g.set_plan(plan)
    full, info = g.run_fused(0, 7500, api.EMIT_ARTP)
    cap = len(full) + 100
    buf = C.create_string_buffer(cap)
    for chunk in (0, 7500, 1000, 777, 1):
        if chunk == 1:
            i2 = g.run_fused_to_host(100, 50, api.EMIT_ARTP, C.addressof(buf), cap, 1)
            assert buf.raw[:i2.bytes] == g.run_fused(100, 50, api.EMIT_ARTP)[0]
            continue
        i2 = g.run_fused_to_host(0, 7500, api.EMIT_ARTP, C.addressof(buf), cap, chunk)
        assert i2.bytes == len(full) and buf.raw[:i2.bytes] == full and i2.records == 7500 and i2.attempts == info.attempts
    with pytest.raises(api.GGError):
        g.run_fused_to_host(0, 7500, api.EMIT_ARTP, C.addressof(buf), 1000, 1000)


@pytest.mark.parametrize("dist", [1, 2, 5])
def test_fused_base_composition(oracle_mod, genomes, tmp_path, dist):
    """fragSim --comp/--dist (fragSim.cpp:1609-1759): fp64 acceptance product over the fragment ends and their flanks."""
    from oracle import parsers
    p = str(tmp_path / "dnacomp.txt")
    synth.write_dnacomp(p, seed=dist)
    tabs = [np.array(t) for t in parsers.load_dnacomp(p, dist)]
    g, o, ids = _pair(oracle_mod, genomes, damage="matrix")
    for c in (g, o):
        c.set_composition(1, dist, *tabs)
    plan = [api.Range(0, 700, ids[0], 0, "e1_0", 1), api.Range(700, 500, ids[1], 0, "b1", -1), api.Range(1200, 600, ids[2], -1, "c1", 1)]
    g.set_plan(plan); o.set_plan(plan)
    for emit in (api.EMIT_FRAG, api.EMIT_ARTP):
        a, ia = g.run_fused(0, 1800, emit); b, ib = o.run_fused(0, 1800, emit)
        util.assert_same(a, b, "--comp dist=%d emit=%d" % (dist, emit))
        assert (ia.attempts, ia.gather_bytes) == (ib.attempts, ib.gather_bytes) and ia.attempts > 1.3 * 1800
    g.clear_composition(1); o.clear_composition(1)
    util.assert_same(g.run_fused(0, 1800, api.EMIT_FRAG)[0], o.run_fused(0, 1800, api.EMIT_FRAG)[0], "composition cleared")


def test_adpt_name_and_tag(oracle_mod):
    """adptSim -name (_adpt/_rand flags from the forward read) and -tag (adptSim.cpp:375-393), all dialects."""
    g = api.Context(0, 3); o = oracle_mod.OracleContext(3)
    rng = np.random.default_rng(8)
    recs = b"".join(b">f%d:+:1:2:3\n%s\n" % (i, synth._BASES[rng.integers(0, 4, int(rng.integers(1, 70)))].tobytes()) for i in range(400))
    for c in (g, o):
        c.set_adapters("ACGTACGTACGTAC", "TTGCA", 40)           # L < 26: adapter + random; 26 <= L < 40: adapter only; else neither
    for name, tag in ((True, ""), (False, "XY"), (True, "_lib7")):
        g.set_adpt_naming(name, tag); o.set_adpt_naming(name, tag)
        for emit in ALL_EMITS[2:]:
            a = g.run_adpt(recs, emit)[0]; b = o.run_adpt(recs, emit)[0]
            util.assert_same(a, b, "adpt naming %r %r emit=%d" % (name, tag, emit))
    out = g.run_adpt(b">x\nACGT\n>y\n" + b"A" * 30 + b"\n>z\n" + b"C" * 50 + b"\n", api.EMIT_STDOUT4)[0].split(b"\n")
    assert out[0] == b">x_adpt_rand_lib7" and out[2] == b">x_r_adpt_rand_lib7" and out[4] == b">y_adpt_lib7" and out[8] == b">z_lib7"


def test_standalone_deam_long_records(oracle_mod):
    """Stand-alone deamSim has no bound on the record length, the geometric tables have 4096 entries: an overhang / nick draw that runs off
    its table means "longer than any record", a skip that runs off the table carries on from there (DESIGN.md draw specification)."""
    g = api.Context(0, 77); o = oracle_mod.OracleContext(77)
    rng = np.random.default_rng(5)
    recs = b"".join(b">long%d\n%s\n" % (i, synth._BASES[rng.integers(0, 4, int(n))].tobytes()) for i, n in enumerate([9000, 5000, 12000, 4096, 4097, 40, 8191] * 6))
    flat = np.full((3, 12), 1e-4); flat[:, 5] = 2e-4
    for name, setup in (("briggs tiny l", lambda c: c.set_briggs(0, 0.0001, 0.00005, 0.0002, 0.3)),
                        ("briggs no overhang", lambda c: c.set_briggs(0, 0.00002, 0.9, 0.0001, 0.5)),
                        ("matrix small rates", lambda c: c.set_matrix(0, flat, flat)),
                        ("matrix small rates -last", lambda c: c.set_matrix(0, flat, flat, clamp_last=False))):
        for c in (g, o):
            setup(c)
        for naming in (False, True):
            g.set_deam_naming(naming); o.set_deam_naming(naming)
            a = g.run_deam(recs, 0, 3)[0]; b = o.run_deam(recs, 0, 3)[0]
            util.assert_same(a, b, "long records: %s, -name %r" % (name, naming))
        if "last" not in name:
            assert a != recs


def test_adpt_adapters_pass_through_as_given(oracle_mod, genomes):
    """The reference appends -f/-s as they are (adptSim.cpp:298-315): lower case and index placeholders (N) go through the stand-alone
    adptSim stage untouched; the fused kernel (2-bit sequences) refuses them loudly."""
    g = api.Context(0, 3); o = oracle_mod.OracleContext(3)
    rng = np.random.default_rng(18)
    recs = b"".join(b">f%d\n%s\n" % (i, synth._BASES[rng.integers(0, 4, int(rng.integers(1, 60)))].tobytes()) for i in range(300))
    for c in (g, o):
        c.set_adapters("AGATCGGAAGAGCacacgtNNNNNNNNatctcg", "agatcGGAAGNNNNgtgt", 50)
    for emit in ALL_EMITS[2:]:
        util.assert_same(g.run_adpt(recs, emit)[0], o.run_adpt(recs, emit)[0], "raw adapters emit=%d" % emit)
    assert b"acacgtNNNNNNNNatctcg" in g.run_adpt(recs, api.EMIT_FR)[0]
    ids = [g.upload_genome(fa, fai) for fa, fai in genomes]
    g.set_sizes(api.SIZE_FIXED, fixed_len=30); g.set_plan([api.Range(0, 10, ids[0], -1, "x")])
    with pytest.raises(api.GGError, match="upper-case ACGT"):
        g.run_fused(0, 10, api.EMIT_ARTP)
    g.run_fused(0, 10, api.EMIT_FRAG)                          # the fragSim / deamSim dialects do not use the adapters


@pytest.mark.parametrize("damage", ["matrix", "briggs", "briggs_ss", "single", "double"])
def test_deam_name_tags(oracle_mod, damage):
    """deamSim -name: "_DEAM:" + positions in the reference's push order (deamSim.cpp:1857,1919,1949-1994,2035)."""
    g = api.Context(0, 31); o = oracle_mod.OracleContext(31)
    rng = np.random.default_rng(9)
    recs = b"".join(b">f%d:-:10:20:%d\n%s\n" % (i, i, synth._BASES[rng.integers(0, 4, int(rng.integers(1, 260)))].tobytes()) for i in range(600))
    recs += b">allN\nNNNNNNNN\n>one\nC\n"
    for c in (g, o):
        util.set_damage(c, 0, damage); c.set_deam_naming(True)
    a, info = g.run_deam(recs, 0, 5)
    b, _ = o.run_deam(recs, 0, 5)
    util.assert_same(a, b, "deamSim -name " + damage)
    tagged = [h for h, _ in util.records(a) if b"_DEAM:" in h]
    assert len(tagged) > 100 and info.records == 602
    if damage in ("single", "double"):
        assert any(b",-" in h and h.split(b"_DEAM:")[1].split(b",")[0][:1] != b"-" for h in tagged)      # pass-1 hits listed before pass-2 hits
    for c in (g, o):
        c.set_deam_naming(False)
    util.assert_same(g.run_deam(recs, 0, 5)[0], o.run_deam(recs, 0, 5)[0], "naming off")


def test_fused_gc_bias(oracle_mod, tmp_path):
    """fragSim -gc (fragSim.cpp:1190-1269,1509-1576): probe statistics and survival tests, alone and together with --comp."""
    from oracle import parsers
    rng = np.random.default_rng(21)
    contigs = []
    for i in range(4):                                      # GC content drifts along the contigs so that the bias has something to bite on
        n = 9000 + 1500 * i
        pgc = np.clip(0.25 + 0.5 * np.arange(n) / n + 0.1 * np.sin(np.arange(n) / 300.0), 0.05, 0.95)
        isgc = rng.random(n) < pgc
        seq = np.where(isgc, np.where(rng.random(n) < 0.5, ord("G"), ord("C")), np.where(rng.random(n) < 0.5, ord("A"), ord("T"))).astype(np.uint8)
        seq[2000:2040] = ord("N")
        contigs.append(("gc%d" % i, seq))
    fa, fai = synth.fasta_bytes(contigs, 61)
    p = str(tmp_path / "dnacomp.txt"); synth.write_dnacomp(p, seed=2)
    tabs = [np.array(t) for t in parsers.load_dnacomp(p, 2)]
    g = api.Context(0, 77); o = oracle_mod.OracleContext(77)
    stats = []
    for c in (g, o):
        gid = c.upload_genome(fa, fai)
        c.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
        util.set_damage(c, 0, "matrix"); c.set_adapters(read_len=75)
        c.set_composition(0, 2, *tabs)
        stats.append((c.set_gcbias(0, gid, 0, 6.0), c.set_gcbias(1, gid, 2, -4.0)))
        c.set_plan([api.Range(0, 900, gid, 0, "e", -1, 0), api.Range(900, 800, gid, 0, "b1", 0, 1), api.Range(1700, 500, gid, -1, "c1", -1, -1)])
    assert stats[0] == stats[1]                              # integer probe sums -> identical mean / stdev on both sides
    assert 0.3 < stats[0][0][0] < 0.7 and 0.03 < stats[0][0][1] < 0.3
    for emit in (api.EMIT_FRAG, api.EMIT_ARTP):
        a, ia = g.run_fused(0, 2200, emit); b, ib = o.run_fused(0, 2200, emit)
        util.assert_same(a, b, "-gc emit=%d" % emit)
        assert ia.attempts == ib.attempts and ia.attempts > 2 * 2200
    gcf = lambda recs: np.mean([(s.count(b"G") + s.count(b"C")) / len(s) for _, s in recs])   # noqa: E731
    recs = util.records(g.run_fused(0, 2200, api.EMIT_FRAG)[0])
    assert gcf(recs[:900]) < gcf(recs[1700:]) - 0.02        # positive bias factor penalises GC-rich windows
    g.clear_gcbias(0); o.clear_gcbias(0)
    util.assert_same(g.run_fused(0, 900, api.EMIT_FRAG)[0], o.run_fused(0, 900, api.EMIT_FRAG)[0], "gc cleared")


@pytest.mark.parametrize("circ", [0, 1, 3])
@pytest.mark.parametrize("sizes", ["fixed", "freq"])
def test_fused_circular_contig(oracle_mod, genomes, tmp_path, circ, sizes):
    """fragSim --circ (fragSim.cpp:1407-1426): run-off windows of the circular contig are gathered from the wrap image, incl.
    the reference's quirks (last base skipped, coord == end+1, contig shorter than the fragment, u64 wrap of end-length+1),
    alone and together with --comp (flanks wrap too) and -gc."""
    from oracle import parsers
    fa, fai = util.circ_genome()
    p = str(tmp_path / "dnacomp.txt"); synth.write_dnacomp(p, seed=3)
    tabs = [np.array(t) for t in parsers.load_dnacomp(p, 2)]
    g = api.Context(0, 5); o = oracle_mod.OracleContext(5)
    for c in (g, o):
        other = c.upload_genome(*genomes[0])
        gid = c.upload_genome(fa, fai, circ=circ)
        if sizes == "fixed":
            c.set_sizes(api.SIZE_FIXED, fixed_len=40)
        else:
            c.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
        util.set_damage(c, 0, "matrix"); c.set_adapters(read_len=75)
        c.set_composition(0, 2, *tabs)
        c.set_gcbias(0, gid, 0, 3.0)
        c.set_plan([api.Range(0, 3000, gid, 0, "e1_0"), api.Range(3000, 500, other, 0, "b1"), api.Range(3500, 1500, gid, -1, "c1", 0),
                    api.Range(5000, 1000, gid, 0, "c2", -1, 0)])
    for emit in (api.EMIT_FRAG, api.EMIT_ARTP):
        a, ia = g.run_fused(0, 6000, emit); b, ib = o.run_fused(0, 6000, emit)
        util.assert_same(a, b, "--circ %d sizes=%s emit=%d" % (circ, sizes, emit))
        assert (ia.attempts, ia.gather_bytes) == (ib.attempts, ib.gather_bytes)
    clen = fai[circ].len
    wrapped = sum(1 for h, _ in util.records(g.run_fused(0, 3000, api.EMIT_FRAG)[0]) if int(h.split(b":")[3]) > clen and h.startswith(b">" + fai[circ].name.encode() + b":"))
    assert wrapped > (0 if clen < 100 else 20)
    g.close(); o.close()


def test_fused_circ_end_coordinate_one_digit_longer(oracle_mod):
    """A circular contig just short of a power of ten: the un-wrapped end coordinate start+length (fragSim.cpp:1489) takes one digit more
    than the contig length, in the dialects that have no slack in their record bound (worst-case -l, fragSim / deamSim records)."""
    rng = np.random.default_rng(77)
    fa, fai = synth.fasta_bytes([("m1", synth.random_contig(rng, 9990, 0.0)), ("m2", synth.random_contig(rng, 300, 0.0))], 60)
    g = api.Context(0, 5); o = oracle_mod.OracleContext(5)
    for c in (g, o):
        gid = c.upload_genome(fa, fai, circ=0)
        c.set_sizes(api.SIZE_FIXED, fixed_len=60)
        util.set_damage(c, 0, "briggs")
        c.set_plan([api.Range(0, 40000, gid, 0, "")])
    for emit in (api.EMIT_FRAG, api.EMIT_DEAM):
        a, _ = g.run_fused(0, 40000, emit); b, _ = o.run_fused(0, 40000, emit)
        util.assert_same(a, b, "--circ end digits emit=%d" % emit)
    assert sum(1 for h, _ in util.records(a) if int(h.split(b":")[3]) >= 10000 and int(h.split(b":")[2]) < 10000) > 50


def test_gc_bias_on_a_genome_without_gc_variation(oracle_mod):
    """-gc on a genome whose probe windows all have the same GC content: stdev 0, the survival probability is NaN and "rP > NaN" is false in
    the reference (fragSim.cpp:1524-1529): no fragment is ever killed (instead of every attempt being rejected)."""
    unit = np.frombuffer(b"AC" * 5000, dtype=np.uint8).copy()          # every window of any even length is half C
    fa, fai = synth.fasta_bytes([("flat", unit)], 60)
    g = api.Context(0, 9); o = oracle_mod.OracleContext(9)
    for c in (g, o):
        gid = c.upload_genome(fa, fai)
        c.set_sizes(api.SIZE_FIXED, fixed_len=40)
        mean, sd = c.set_gcbias(0, gid, 0, 2.0)
        assert sd == 0.0 and abs(mean - 0.5) < 1e-12
        c.set_plan([api.Range(0, 2000, gid, -1, "x", -1, 0)])
    a, ia = g.run_fused(0, 2000, api.EMIT_FRAG); b, ib = o.run_fused(0, 2000, api.EMIT_FRAG)
    util.assert_same(a, b, "-gc with sd = 0")
    assert ia.attempts == ib.attempts


def test_fused_mixed_damage_slots(oracle_mod, genomes):
    """All five damage models side by side in one plan (the DMODE = any-mix kernel), incl. -damagess."""
    g, o, ids = _pair(oracle_mod, genomes, damage="briggs_ss")
    for c in (g, o):
        util.set_damage(c, 1, "matrix"); util.set_damage(c, 2, "briggs"); util.set_damage(c, 3, "double")
    plan = [api.Range(0, 500, ids[0], 0, "e1_0"), api.Range(500, 400, ids[1], 1, "b1"), api.Range(900, 400, ids[2], 2, "b2"),
            api.Range(1300, 300, ids[0], 3, "b3"), api.Range(1600, 200, ids[1], -1, "c1")]
    g.set_plan(plan); o.set_plan(plan)
    for emit in (api.EMIT_DEAM, api.EMIT_ARTP, api.EMIT_FR):
        util.assert_same(g.run_fused(0, 1800, emit)[0], o.run_fused(0, 1800, emit)[0], "mixed damage emit=%d" % emit)
    recs, _ = o.run_fused(0, 500, api.EMIT_FRAG)
    dam, _ = g.run_fused(0, 500, api.EMIT_DEAM)
    diff = [(a, b) for (_, s), (_, t) in zip(util.records(recs), util.records(dam)) for a, b in zip(s, t) if a != b]
    assert len(diff) > 200 and set(diff) == {(ord("C"), ord("T"))}       # -damagess: C->T only (deamSim.cpp:1337-1468)
    g.close(); o.close()


def _bgzf_members(z):
    """walk the BGZF members: (offset, BSIZE, ISIZE) from the BC extra field"""
    out, p = [], 0
    while p < len(z):
        assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 10:p + 16] == b"\x06\x00BC\x02\x00", "not a BGZF member at %d" % p
        bsize = int.from_bytes(z[p + 16:p + 18], "little") + 1
        out.append((p, bsize, int.from_bytes(z[p + bsize - 4:p + bsize], "little")))
        p += bsize
    assert p == len(z)
    return out


@pytest.mark.parametrize("n", [1, 700, 2500])
def test_device_bgzf_round_trip(oracle_mod, genomes, n):
    """gg_bgzf_dev: the members inflate (zlib checks every CRC-32 and ISIZE) to exactly the stream the oracle gives; member
    framing is valid BGZF (BC field, BSIZE, <= 64 KiB, 65280 input bytes per member)."""
    import gzip
    import zlib
    g, o, ids = _pair(oracle_mod, genomes)
    plan = util.plan3(ids, counts=(1000, 700, 800))
    g.set_plan(plan); o.set_plan(plan)
    for emit in (api.EMIT_FRAG, api.EMIT_ARTP):
        z, info, zi = g.run_fused_bgzf(0, n, emit)
        want = o.run_fused(0, n, emit)[0]
        assert gzip.decompress(z + api.BGZF_EOF) == want
        mem = _bgzf_members(z)
        assert len(mem) == (len(want) + 65279) // 65280 and all(b <= 65536 for _, b, _ in mem)
        assert [i for _, _, i in mem] == [min(65280, len(want) - 65280 * k) for k in range(len(mem))]
        for k, (off, bsize, isize) in enumerate(mem[:3]):               # each member is a self-contained gzip file
            assert zlib.decompress(z[off:off + bsize], wbits=31) == want[k * 65280:k * 65280 + isize]
        assert zi.in_bytes == len(want) and zi.bytes == len(z) and zi.launches == 4
        if len(want) > 50000:
            assert len(z) < 0.45 * len(want)                            # ~2.6 bits per byte on FASTA of reads
    g.close(); o.close()


def test_device_bgzf_incompressible_and_binary(oracle_mod):
    """Streams the code was not built for: all 256 byte values, and random bytes (members fall back to stored blocks)."""
    import gzip
    rng = np.random.default_rng(3)
    g = api.Context(0, 1)
    for data in (bytes(range(256)) * 600, rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes(), b"A" * 70000, b"ACGT\n" * 13057 + b"x", b"", b"G"):
        z, zi = g.bgzf(data)
        assert gzip.decompress(z + api.BGZF_EOF) == data
        assert all(b <= 65536 for _, b, _ in _bgzf_members(z)) and zi.in_bytes == len(data)
    g.close()


def test_device_bgzf_equals_host_model(oracle_mod, genomes):
    """k_bgzf_lz against ggh::bgzf_lz_model (gg_lz.h): the same parse, the same code, the same members byte for byte -- simulator
    dialects, multi-member streams with a partial last member, runs shorter than a match, streams that fall back to stored blocks."""
    import gzip
    rng = np.random.default_rng(11)
    g, o, ids = _pair(oracle_mod, genomes)
    plan = util.plan3(ids, counts=(1500, 1000, 1200))
    g.set_plan(plan); o.set_plan(plan)
    streams = [o.run_fused(0, 3700, emit)[0] for emit in (api.EMIT_FRAG, api.EMIT_ARTP, api.EMIT_STDOUT4)]
    artp = util.read_like_stream(rng, 9000, "artp")                          # ~1.7 MB: 26 members, every 1st sampled
    streams += [artp, util.read_like_stream(rng, 2500, "frag"), artp[:65280], artp[:65281], artp[:2176 * 3 + 5], artp[:7], artp[:2181],
                bytes(range(256)) * 600, rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes(), b"A" * 70000, b"ACGT\n" * 13057 + b"x",
                b"G", b"ACGTAC" * 2, bytes(70000), artp[:40000] + rng.integers(0, 256, size=70_000, dtype=np.uint8).tobytes() + artp[:50000]]
    for k, data in enumerate(streams):
        z, zi = g.bgzf(data)
        want, _ = util.bgzf_lz_model(data, eof=False)
        assert gzip.decompress(z + api.BGZF_EOF) == data, "stream %d does not inflate" % k
        util.assert_same(z, want, "stream %d: device members vs host model" % k)
        assert zi.in_bytes == len(data) and zi.bytes == len(z)
    for k in range(24):                                                       # random phrase soups: alphabets of 2..256 letters, phrases of 1..300 bytes, odd lengths
        alpha = rng.integers(0, 256, size=int(rng.integers(2, 257)), dtype=np.uint8)
        phrases = [alpha[rng.integers(0, len(alpha), size=int(rng.integers(1, 301)))].tobytes() for _ in range(int(rng.integers(1, 40)))]
        n = int(rng.integers(1, 400_000))
        data = b"".join(phrases[int(i)] for i in rng.integers(0, len(phrases), size=n // 20 + 1))[:n]
        z, _ = g.bgzf(data)
        assert gzip.decompress(z + api.BGZF_EOF) == data, "soup %d does not inflate" % k
        util.assert_same(z, util.bgzf_lz_model(data, eof=False)[0], "soup %d: device members vs host model" % k)
    big = util.read_like_stream(rng, 60_000, "artp")                          # ~11 MB: 170 members, every 2nd sampled, more members than SMs
    z, _ = g.bgzf(big)
    util.assert_same(z, util.bgzf_lz_model(big, eof=False)[0], "170 members")
    g.close(); o.close()


def test_device_bgzf_huffman_only_mode(oracle_mod, genomes, monkeypatch):
    """GG_BGZF_MODE=huff keeps the Huffman-only encoder (k_bgzf_deflate) reachable: same framing, larger members."""
    import gzip
    g, o, ids = _pair(oracle_mod, genomes)
    plan = util.plan3(ids, counts=(1000, 700, 800))
    g.set_plan(plan); o.set_plan(plan)
    want = o.run_fused(0, 2500, api.EMIT_ARTP)[0]
    z_lz = g.run_fused_bgzf(0, 2500, api.EMIT_ARTP)[0]
    monkeypatch.setenv("GG_BGZF_MODE", "huff")
    z_h = g.run_fused_bgzf(0, 2500, api.EMIT_ARTP)[0]
    monkeypatch.delenv("GG_BGZF_MODE")
    assert gzip.decompress(z_h + api.BGZF_EOF) == want and gzip.decompress(z_lz + api.BGZF_EOF) == want
    assert len(z_lz) < 0.85 * len(z_h)
    _bgzf_members(z_h); _bgzf_members(z_lz)
    g.close(); o.close()


def test_fused_bgzf_to_host_chunked(oracle_mod, genomes):
    """gg_run_fused_bgzf_to_host: chunk boundaries and the reused Huffman code leave the inflated stream untouched."""
    import ctypes as C
    import gzip
    g, o, ids = _pair(oracle_mod, genomes)
    plan = util.plan3(ids, counts=(3000, 2000, 2500))
    g.set_plan(plan); o.set_plan(plan)
    want = o.run_fused(0, 7500, api.EMIT_ARTP)[0]
    buf = (C.c_uint8 * (len(want) + 65536))()
    for chunk in (1000, 2048, 7500, 0):
        info = g.run_fused_bgzf_to_host(0, 7500, api.EMIT_ARTP, C.addressof(buf), len(buf), chunk)
        z = bytes(buf[:info.bytes])
        assert gzip.decompress(z + api.BGZF_EOF) == want and info.in_bytes == len(want) and info.records == 7500
        _bgzf_members(z)
    with pytest.raises(api.GGError):
        g.run_fused_bgzf_to_host(0, 7500, api.EMIT_ARTP, C.addressof(buf), 1000, 1000)
    g.close(); o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# methylation mode: fragSim --case keeps the case of the genome, deamSim -matfilenonmeth/-matfilemeth reads it
def _case_genomes():
    """genomes whose bases are lower-case one by one (methylation calls), not in stretches"""
    out = []
    for g, (fa, fai) in enumerate(util.tiny_genomes(seed=19)):
        b = np.frombuffer(fa, dtype=np.uint8).copy()
        rng = np.random.default_rng(100 + g)
        letters = ((b >= 65) & (b <= 90)) | ((b >= 97) & (b <= 122))
        hdr = np.zeros(len(b), dtype=bool)
        for e in fai:                                                    # only the sequence lines change case
            hdr[e.offset:e.offset + e.len + e.len // e.line_blen + 1] = True
        up = np.where(letters & hdr, b & 0xDF, b)
        low = (rng.random(len(b)) < 0.3) & letters & hdr
        out.append((np.where(low, up | 0x20, up).astype(np.uint8).tobytes(), fai))
    return out


@pytest.mark.parametrize("sizes", ["freq", "fixed"])
def test_fused_case_bit_exact(oracle_mod, sizes):
    """fragSim --case (fragSim.cpp:318-326,548-551): records keep the case of the file; reverse-strand records complement within the case.
    The later stages upper-case what they read, so only the fragSim records differ from a run without --case."""
    genomes = _case_genomes()
    g, o, ids = _pair(oracle_mod, genomes, sizes=sizes, damage="matrix", case=True)
    plan = util.plan3(ids)
    g.set_plan(plan); o.set_plan(plan)
    a, ia = g.run_fused(0, 1200, api.EMIT_FRAG)
    b, ib = o.run_fused(0, 1200, api.EMIT_FRAG)
    util.assert_same(a, b, "--case fragSim records")
    assert (ia.bytes, ia.attempts, ia.gather_bytes) == (ib.bytes, ib.attempts, ib.gather_bytes)
    seqs = b"".join(s for _, s in util.records(a))
    assert any(c in seqs for c in b"acgt") and any(c in seqs for c in b"ACGT")
    for emit in ALL_EMITS[1:]:
        util.assert_same(g.run_fused(0, 1200, emit)[0], o.run_fused(0, 1200, emit)[0], "--case emit=%d" % emit)
    g2, o2, ids2 = _pair(oracle_mod, genomes, sizes=sizes, damage="matrix", case=False)
    g2.set_plan(util.plan3(ids2))
    plain = util.records(g2.run_fused(0, 1200, api.EMIT_FRAG)[0])      # same fragments, upper-cased
    assert [(d, s_.upper()) for d, s_ in util.records(a)] == plain
    for c in (g, o, g2, o2): c.close()


def test_fused_case_on_the_circular_contig(oracle_mod):
    """--case with --circ: the wrap image carries the case bits too"""
    fa, fai = util.circ_genome()
    b = np.frombuffer(fa, dtype=np.uint8).copy()
    rng = np.random.default_rng(3)
    seqb = np.zeros(len(b), dtype=bool)
    for e in fai: seqb[e.offset:e.offset + e.len + e.len // e.line_blen + 1] = True
    low = (rng.random(len(b)) < 0.4) & seqb & (b >= 65) & (b <= 90)
    fa = np.where(low, b | 0x20, b).astype(np.uint8).tobytes()
    with api.Context(0, 5) as g:
        o = oracle_mod.OracleContext(5)
        for c in (g, o):
            c.set_case(True)
            gid = c.upload_genome(fa, fai, circ=1)
            c.set_sizes(api.SIZE_FIXED, fixed_len=70, maxlen=1000)
            c.set_plan([api.Range(0, 3000, gid, -1, "mt")])
        util.assert_same(g.run_fused(0, 3000, api.EMIT_FRAG)[0], o.run_fused(0, 3000, api.EMIT_FRAG)[0], "--case --circ")
        o.close()


@pytest.mark.parametrize("clamp", [True, False])
@pytest.mark.parametrize("name", [False, True])
def test_deam_methylation_standalone_bit_exact(oracle_mod, clamp, name):
    """deamSim -matfilenonmeth/-matfilemeth (deamSim.cpp:1630-1787) over fragSim --case records plus hand-made ones (unresolved letters in
    both cases, records shorter than the tables, empty sequence): bit-exact vs the oracle, with -name and -last"""
    genomes = _case_genomes()
    g, o, ids = _pair(oracle_mod, genomes, sizes="freq", damage="meth", case=True, clamp_last=clamp)
    plan = util.plan3(ids)
    g.set_plan(plan); o.set_plan(plan)
    recs = g.run_fused(0, 1200, api.EMIT_FRAG)[0]
    recs += b">odd:+:0:9:9\nacgNnTgca\n>one:+:0:1:1\nc\n>none:+:0:0:0\n\n>x:-:5:25:20\nRYacgtACGTnnNNacgtAC\n"
    for c in (g, o): c.set_deam_naming(name)
    a, ia = g.run_deam(recs, 0)
    b, ib = o.run_deam(recs, 0)
    util.assert_same(a, b, "methylation deamSim clamp=%s name=%s" % (clamp, name))
    assert ia.records == ib.records == 1204
    if clamp:                                                            # every resolved base came out upper-case; unresolved ones kept their case
        body = b"".join(s for _, s in util.records(a))
        assert not any(c in body for c in b"acgt") and b"n" in body
    g.close(); o.close()


def test_fused_methylation_plan_runs_as_stage_chain(oracle_mod):
    """a plan whose every range uses a methylation slot: gg_run_fused chains fragSim(--case) -> deamSim -> adptSim stage kernels on the
    device; every dialect, id sub-ranges and the host egress paths must equal the oracle's composition of the three loops"""
    genomes = _case_genomes()
    g, o, ids = _pair(oracle_mod, genomes, sizes="freq", damage="meth", case=True)
    plan = util.plan3(ids, slots=(0, 0, 0))
    g.set_plan(plan); o.set_plan(plan)
    for emit in ALL_EMITS:
        a, ia = g.run_fused(0, 1200, emit)
        b, ib = o.run_fused(0, 1200, emit)
        util.assert_same(a, b, "methylation chain emit=%d" % emit)
        assert (ia.bytes, ia.records, ia.attempts, ia.gather_bytes) == (ib.bytes, ib.records, ib.attempts, ib.gather_bytes)
    full = g.run_fused(0, 1200, api.EMIT_ARTP)[0]
    assert b"".join(g.run_fused(a0, b0 - a0, api.EMIT_ARTP)[0] for a0, b0 in ((0, 1), (1, 399), (399, 400), (400, 1200))) == full
    # methylated and unmethylated cytosines really deaminate at different rates: C->T at the first base of + strand records
    fr = util.records(g.run_fused(0, 1200, api.EMIT_FRAG)[0]); dm = util.records(g.run_fused(0, 1200, api.EMIT_DEAM)[0])
    lo = [d[1][:1] for f, d in zip(fr, dm) if f[1][:1] == b"c"]; up = [d[1][:1] for f, d in zip(fr, dm) if f[1][:1] == b"C"]
    assert len(lo) > 30 and len(up) > 60 and lo.count(b"T") / len(lo) > 0.17 > 0.1 > up.count(b"T") / len(up)   # tables: 0.29 vs 0.026
    # host egress: chunked copies and device BGZF
    import ctypes as C
    import zlib
    cap = int(g.max_record_bytes(api.EMIT_ARTP)) * 1200 + 65536
    buf = (C.c_uint8 * cap)()
    n = g.run_fused_to_host(0, 1200, api.EMIT_ARTP, C.addressof(buf), cap, chunk=500).bytes
    assert bytes(buf[:n]) == full
    n = g.run_fused_bgzf_to_host(0, 1200, api.EMIT_ARTP, C.addressof(buf), cap, chunk=500).bytes
    z = bytes(buf[:n]) + api.BGZF_EOF; out = b""
    while z:
        d = zlib.decompressobj(31); out += d.decompress(z); z = d.unused_data
    assert out == full
    # mixing a methylation slot with another model in one plan is refused (gargammel.pl:670-677)
    util.set_damage(g, 1, "briggs")
    g.set_plan(util.plan3(ids, slots=(0, 1, 0)))
    with pytest.raises(api.GGError):
        g.run_fused(0, 100, api.EMIT_ARTP)
    g.close(); o.close()
