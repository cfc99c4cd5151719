"""CPU tests of the host layer: the C-ABI library loads and exports every symbol of include/gargammel_b200.h,
the C++ parsers agree with the oracle parsers and the golden tables, apportioning follows gargammel.pl."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth, util
from oracle import parsers
from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    hdr = open(os.path.join(ROOT, "include", "gargammel_b200.h")).read()
    names = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n
    lib.gg_abi_version.restype = C.c_int
    assert lib.gg_abi_version() == 1


def test_no_cuda_device_fails_loudly():
    """Without a GPU the context cannot be created: there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.GGError):
        api.Context(0, 1)


def _load_rates(path, kind):
    lib = api.load_library()
    s5 = np.zeros((4096, 12)); s3 = np.zeros((4096, 12))
    r5, r3 = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    rc = lib.ggh_load_rates(path.encode(), kind, s5.ctypes.data_as(C.POINTER(C.c_double)), 4096, C.byref(r5),
                            s3.ctypes.data_as(C.POINTER(C.c_double)), 4096, C.byref(r3), err, 512)
    if rc:
        raise ValueError(err.value.decode())
    return s5[:r5.value], s3[:r3.value]


def test_matrix_dat_parser(tmp_path, golden):
    for name in ("single-", "double-"):
        s5, s3 = synth.golden_matrix(name)
        synth.write_matrix_dat(str(tmp_path / name), s5, s3)
        g5, g3 = _load_rates(str(tmp_path / name), 0)
        assert np.array_equal(g5, s5) and np.array_equal(g3, s3)       # incl. the 3' reversal (deamSim.cpp:1236)
        o5, o3 = parsers.load_matrix_dat(str(tmp_path / name))
        assert np.array_equal(np.array(o5), g5) and np.array_equal(np.array(o3), g3)
    # sum > 1 is rejected (deamSim.cpp:1148-1158)
    bad = s5.copy(); bad[3, 3:6] = 0.4
    synth.write_matrix_dat(str(tmp_path / "bad-"), bad, s3)
    with pytest.raises(ValueError, match="Problem with line"):
        _load_rates(str(tmp_path / "bad-"), 0)
    with pytest.raises(ValueError, match="Unable to open"):
        _load_rates(str(tmp_path / "missing-"), 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/matrices"), reason="reference tree not mounted")
def test_parsers_on_reference_files(golden):
    for name in ("single-", "double-"):
        g5, g3 = _load_rates("/root/reference/src/matrices/" + name, 0)
        s5, s3 = synth.golden_matrix(name)
        assert np.array_equal(g5, s5) and np.array_equal(g3, s3)
    g5, g3 = _load_rates("/root/reference/examplesMapDamage/results_LaBrana/misincorporation.txt", 2)
    md = golden["mapdamage_LaBrana"]
    assert g5.shape == (70, 12) and np.allclose(g5[0], md["sub5_row0"], rtol=0, atol=0, equal_nan=True)
    assert np.allclose(g3[0], md["sub3_row0"], rtol=0, atol=0, equal_nan=True) and np.allclose(g5[69], md["sub5_row69"], rtol=0, atol=0, equal_nan=True)
    lens, cum = _size_freq("/root/reference/src/sizefreq.size.gz")
    ol, oc = synth.golden_sizefreq()
    assert lens == ol and cum == oc
    assert _size_list("/root/reference/src/sizedist.size.gz")[:32] == golden["sizedist_first32"]


def test_profile_parser():
    """-profile (deamSim.cpp:956-1094): <prefix>5p.prof / <prefix>3p.prof, 12 columns, header skipped, the 3' file NOT reversed.  The fixture
    was written by the reference's own mapDamage2prof (tests/golden/make_prof_fixture.py) from tests/golden/mapdamage_sample.txt."""
    prefix = os.path.join(ROOT, "tests", "golden", "sample_")
    g5, g3 = _load_rates(prefix, 1)
    o5, o3 = parsers.load_profile(prefix)
    assert g5.shape == (70, 12) and g3.shape == (70, 12)
    assert np.array_equal(g5, np.array(o5)) and np.array_equal(g3, np.array(o3))
    # the rates are the mapDamage counts turned into frequencies (mapDamage2prof.cpp:334-367): the same numbers -mapdamage computes itself
    m5, m3 = parsers.load_mapdamage(os.path.join(ROOT, "tests", "golden", "mapdamage_sample.txt"))
    assert np.allclose(g5[:, 5], np.array(m5)[:, 5], rtol=1e-5) and np.allclose(g3[:, 6], np.array(m3)[:, 6], rtol=1e-5)
    assert g5[0, 5] > 0.1 > g5[20, 5]                                   # C->T decays from the 5' end
    with pytest.raises(ValueError, match="Unable to open"):
        _load_rates(prefix + "missing_", 1)


def test_mapdamage_parser(golden):
    p = os.path.join(ROOT, "tests", "golden", "mapdamage_sample.txt")
    g5, g3 = _load_rates(p, 2)
    assert np.array_equal(g5, np.array(golden["mapdamage_sample"]["sub5"])) and np.array_equal(g3, np.array(golden["mapdamage_sample"]["sub3"]))
    o5, o3 = parsers.load_mapdamage(p)
    assert np.array_equal(np.array(o5), g5)
    # header order of mapDamage: G>A C>T A>G T>C A>C A>T C>G C>A T>G T>A G>C G>T -> canonical index (deamSim.cpp:483-536)
    first = open(p).read().split("\n")[4].split("\t")
    assert first[1] == "3p" and first[3] == "1"


def _size_freq(path):
    lib = api.load_library()
    ln = np.zeros(100000, dtype=np.int32); cm = np.zeros(100000)
    n = C.c_int(); err = C.create_string_buffer(512)
    rc = lib.ggh_load_size_freq(path.encode(), ln.ctypes.data_as(C.POINTER(C.c_int32)), cm.ctypes.data_as(C.POINTER(C.c_double)), 100000, C.byref(n), err, 512)
    if rc:
        raise ValueError(err.value.decode())
    return ln[:n.value].tolist(), cm[:n.value].tolist()


def _size_list(path):
    lib = api.load_library()
    ln = np.zeros(100000, dtype=np.int32)
    n = C.c_int(); err = C.create_string_buffer(512)
    rc = lib.ggh_load_size_list(path.encode(), ln.ctypes.data_as(C.POINTER(C.c_int32)), 100000, C.byref(n), err, 512)
    if rc:
        raise ValueError(err.value.decode())
    return ln[:n.value].tolist()


def test_size_parsers(tmp_path):
    p = str(tmp_path / "sf.size")
    synth.write_sizefreq(p)
    lens, cum = _size_freq(p)
    ol, oc = synth.golden_sizefreq()
    assert lens == ol and cum == oc and (lens, cum) == tuple(map(list, parsers.load_size_freq(p)))
    import gzip
    with gzip.open(p + ".gz", "wt") as f:
        f.write(open(p).read())
    assert _size_freq(p + ".gz") == (lens, cum)                         # igzstream reads gz or plain (fragSim.cpp:753)
    with open(p, "w") as f:
        f.write("40\t0.5\n41\t0.3\n")
    with pytest.raises(ValueError, match="does not sum to 1"):
        _size_freq(p)
    with open(p, "w") as f:
        f.write("40\n0\n-3\n77\n")
    assert _size_list(p) == [40, 77]                                   # keep t >= 1 (fragSim.cpp:736)


def test_fai_parse_and_build(tmp_path):
    fa, fai = synth.make_genome(3, 4, 5000, width=50, lower_frac=0.1)
    p = str(tmp_path / "g.fa")
    open(p, "wb").write(fa)
    synth.write_fai(p + ".fai", fai)
    lib = api.load_library()
    for from_fasta, path in ((0, p + ".fai"), (1, p)):
        h = C.c_void_p(); n = C.c_int(); err = C.create_string_buffer(512)
        assert lib.ggh_fai_load(path.encode(), from_fasta, C.byref(h), C.byref(n), err, 512) == 0, err.value
        assert n.value == len(fai)
        for i, e in enumerate(fai):
            name = C.create_string_buffer(256); ln = C.c_int64(); off = C.c_uint64(); bl = C.c_int32(); ll = C.c_int32()
            assert lib.ggh_fai_get(h, i, name, 256, C.byref(ln), C.byref(off), C.byref(bl), C.byref(ll)) == 0
            assert (name.value.decode(), ln.value, off.value) == (e.name, e.len, e.offset)
            if e.len > e.line_blen:
                assert (bl.value, ll.value) == (e.line_blen, e.line_len)
        lib.ggh_fai_free(h)


def _apportion(n, comp, n_endo, cont_len, bact_w, seed):
    lib = api.load_library()
    tot = (C.c_uint64 * 3)(); e = (C.c_uint64 * 2)(); c = (C.c_uint64 * max(1, len(cont_len)))(); b = (C.c_uint64 * max(1, len(bact_w)))()
    err = C.create_string_buffer(512)
    rc = lib.ggh_apportion(C.c_uint64(n), C.c_double(comp[0]), C.c_double(comp[1]), C.c_double(comp[2]), n_endo,
                           (C.c_uint64 * max(1, len(cont_len)))(*cont_len), len(cont_len), (C.c_double * max(1, len(bact_w)))(*bact_w), len(bact_w),
                           C.c_uint64(seed), tot, e, c, b, err, 512)
    if rc:
        raise ValueError(err.value.decode())
    return list(tot), list(e)[:n_endo], list(c)[:len(cont_len)], list(b)[:len(bact_w)]


def test_apportion_follows_gargammel_pl():
    """gargammel.pl:1256-1283 (int(n*comp), diploid coin flips), :1320-1349 (bacterial multinomial), :1421-1431 (contaminant by length)."""
    w = np.random.default_rng(0).lognormal(size=10); w /= w.sum()
    tot, e, c, b = _apportion(10_000_000, (0.9, 0.02, 0.08), 2, [30_000_000, 10_000_000], w.tolist(), 42)
    assert tot == [int(10_000_000 * 0.08), int(10_000_000 * 0.02), int(10_000_000 * 0.9)]
    assert sum(e) == tot[0] and sum(c) == tot[1] and sum(b) == tot[2]
    assert abs(e[0] - tot[0] / 2) < 5 * (tot[0] / 4) ** 0.5                              # Binomial(nE, 1/2)
    assert abs(c[0] / tot[1] - 0.75) < 5 * (0.75 * 0.25 / tot[1]) ** 0.5                 # by genome length
    z = (np.array(b) - tot[2] * w) / np.sqrt(tot[2] * w * (1 - w))
    assert np.abs(z).max() < 5
    # deterministic in the seed, different across seeds
    assert _apportion(10_000_000, (0.9, 0.02, 0.08), 2, [30_000_000, 10_000_000], w.tolist(), 42) == (tot, e, c, b)
    assert _apportion(10_000_000, (0.9, 0.02, 0.08), 2, [30_000_000, 10_000_000], w.tolist(), 43)[1] != e
    with pytest.raises(ValueError, match="sum up to 1"):
        _apportion(1000, (0.5, 0.2, 0.2), 1, [10], [1.0], 1)
    tot, e, c, b = _apportion(1000, (0, 0.08, 0.92), 1, [5000], [], 1)                   # README example composition
    assert tot == [920, 80, 0] and e == [920] and c == [80]


def test_binomial_sampler_moments():
    lib = api.load_library()
    lib.ggh_binomial.restype = C.c_uint64
    for n, p in ((50, 0.1), (1000, 0.3), (10 ** 9, 0.37), (10 ** 6, 1e-7), (10 ** 7, 0.999)):
        x = np.array([lib.ggh_binomial(C.c_uint64(n), C.c_double(p), C.c_uint64(7), C.c_uint64(k)) for k in range(4000)], dtype=np.float64)
        m, s = n * p, (n * p * (1 - p)) ** 0.5
        assert abs(x.mean() - m) < 5 * s / 4000 ** 0.5 + 1e-9
        if s > 0.5:
            assert abs(x.std() / s - 1) < 0.08
        assert x.min() >= 0 and x.max() <= n


def test_lognormal_table_matches_scipy():
    """fragSim --loc/--scale: length = int(lognormal(loc, scale)) (fragSim.cpp:184) as an inverse-CDF table."""
    from scipy import stats
    from tests import util
    for loc, scale in ((4.1, 0.35), (3.7, 0.6), (5.0, 0.2)):
        lens, cum = util.lognormal_table(loc, scale)
        assert lens[:3] == [0, 1, 2] and lens[-1] == 4096 and cum[-1] == 1.0
        want = stats.lognorm.cdf(np.arange(1, 4097), s=scale, scale=np.exp(loc))       # P(int(X) <= l) = F(l+1)
        assert np.abs(np.array(cum[:-1]) - want).max() < 1e-12
        assert all(b >= a for a, b in zip(cum[:-1], cum[1:]))


def test_bgzf_writer_roundtrip():
    """Parallel BGZF (blocked gzip) used for every .gz output: inflates with a plain gzip reader, carries the BC block
    sizes and the 28-byte EOF marker of the SAM spec."""
    import gzip
    lib = api.load_library()
    rng = np.random.default_rng(0)
    for n in (0, 1, 65279, 65280, 65281, 1_000_003):
        data = (b">chr1:+:%d:%d:40e1\n" % (n, n + 40)) * 3 + synth._BASES[rng.integers(0, 4, n)].tobytes()
        out = C.create_string_buffer(len(data) + len(data) // 200 + 70000)
        m = C.c_size_t()
        assert lib.ggh_bgzf_compress(data, C.c_size_t(len(data)), 3, 1, 1, out, C.c_size_t(len(out)), C.byref(m)) == 0
        z = out.raw[:m.value]
        assert gzip.decompress(z) == data
        assert z[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
        # walk the blocks through their BSIZE fields
        p, blocks = 0, 0
        while p < len(z):
            assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 12:p + 14] == b"BC"
            p += (z[p + 16] | (z[p + 17] << 8)) + 1; blocks += 1
        assert p == len(z) and blocks == (len(data) + 65279) // 65280 + 1


def _load_dnacomp(path, dist):
    lib = api.load_library()
    t = [np.zeros((2 * dist, 4)) for _ in range(4)]
    err = C.create_string_buffer(512)
    rc = lib.ggh_load_dnacomp(path.encode(), dist, *[x.ctypes.data_as(C.POINTER(C.c_double)) for x in t], err, 512)
    if rc:
        raise ValueError(err.value.decode())
    return t


def test_dnacomp_parser(tmp_path, golden):
    """fragSim --comp/--dist loader (fragSim.cpp:798-1012) vs the Python restatement and the golden table of src/dnacomp.txt."""
    p = str(tmp_path / "dnacomp.txt")
    synth.write_dnacomp(p)
    for dist in (1, 2, 3):
        got = _load_dnacomp(p, dist)
        want = parsers.load_dnacomp(p, dist)
        for a, b in zip(got, want):
            assert np.array_equal(a, np.array(b))
            assert np.allclose(a.sum(1), 1.0, atol=1e-9)            # frequency - background + 0.25 still sums to 1
    with pytest.raises(ValueError):
        _load_dnacomp(p, 9)                                         # the file only holds +-8
    if os.path.exists("/root/reference/src/dnacomp.txt"):
        got = _load_dnacomp("/root/reference/src/dnacomp.txt", 2)
        for a, b in zip(got, synth.golden_dnacomp()):
            assert np.array_equal(a, b)
    assert len(golden["dnacomp_dist2"]) == 4 and len(golden["dnacomp_dist2"][0]) == 4


def test_deflate_code_for_device_bgzf():
    """ggh::build_deflate_code: the dynamic-Huffman description + per-symbol codes the device encoder uses decode under zlib for
    any byte values (unseen bytes still get a code), also when the histogram forces the 15-bit length limit; CRC shift algebra."""
    import ctypes as C
    import zlib
    lib = api.load_library()
    lib.ggh_crc32_x8n.restype = C.c_uint32; lib.ggh_crc32_mul.restype = C.c_uint32
    rng = np.random.default_rng(0)
    fasta = b"".join(b">chr1:+:%d:%d:75e1_0\n%s\n" % (i * 3, i * 3 + 75, synth._BASES[rng.integers(0, 4, size=150)].tobytes()) for i in range(300))

    def code_for(hist):
        sym = (C.c_uint32 * 257)(); hdr = C.create_string_buffer(512); nb = C.c_int(0)
        assert lib.ggh_deflate_code((C.c_uint64 * 256)(*hist), sym, hdr, 512, C.byref(nb)) == 0
        assert all(1 <= (sym[i] >> 16) <= 15 for i in range(257)) and nb.value <= 1280

        def encode(d):
            bits, n = 0, 0
            for i in range(nb.value):
                bits |= ((hdr.raw[i >> 3] >> (i & 7)) & 1) << n; n += 1
            for b in list(d) + [256]:
                bits |= (sym[b] & 0xFFFF) << n; n += sym[b] >> 16
            return bits.to_bytes((n + 7) // 8, "little"), n
        return sym, encode
    sym, enc = code_for(np.bincount(np.frombuffer(fasta, dtype=np.uint8), minlength=256).tolist())
    raw, nbits = enc(fasta)
    assert zlib.decompress(raw, wbits=-15) == fasta and nbits / len(fasta) < 3.0
    assert max(sym[c] >> 16 for c in b"ACGT") <= 3
    allbytes = bytes(range(256)) * 3
    assert zlib.decompress(enc(allbytes)[0], wbits=-15) == allbytes
    sym2, enc2 = code_for([int(1.7 ** i) if i < 60 else 0 for i in range(256)])        # Fibonacci-like: unlimited depth would be ~60
    assert max(sym2[i] >> 16 for i in range(257)) == 15
    assert zlib.decompress(enc2(allbytes)[0], wbits=-15) == allbytes
    a, b = fasta[:1234], fasta[1234:]
    comb = lib.ggh_crc32_mul(C.c_uint32(zlib.crc32(a)), C.c_uint32(lib.ggh_crc32_x8n(C.c_uint64(len(b))))) ^ zlib.crc32(b)
    assert comb == zlib.crc32(fasta)


def test_bam_writer_and_reader_round_trip(tmp_path):
    """ggh::bam_append_record / bam_read (SAM/BAM specification 4.2) through the C exports: a record stream -> stored-BGZF BAM -> back;
    the file is also parsed by the independent reader in tests/util.py (flag 4, no coordinates, phred 0, 4-bit bases, odd lengths)."""
    import ctypes as C
    from tests import util
    lib = api.load_library()
    recs = [(b"chr1:+:10:15:5e1_0", b"ACGTN"), (b"x", b""), (b"y:-:0:1:1", b"T"), (b"long" * 40, b"ACGTRYKMSWBDHVN" * 9000), (b"n" * 254, b"GG")]
    data = b"".join(b">" + n + b"\n" + s + b"\n" for n, s in recs)
    err = C.create_string_buffer(256)
    path = str(tmp_path / "t.bam").encode()
    assert lib.ggh_bam_from_fasta(data, C.c_size_t(len(data)), path, err, 256) == 0, err.value
    text, got = util.read_bam(path.decode())
    assert "@PG\tID:test" in text and text.startswith("@HD")
    assert [(n, s) for n, _, s, _, _ in got] == recs
    assert all(f == 4 and q == bytes(len(s)) and t == {} for _, f, s, q, t in got)
    out = C.create_string_buffer(len(data) + 16); n = C.c_size_t()
    assert lib.ggh_bam_to_fasta(path, out, C.c_size_t(len(out)), C.byref(n), err, 256) == 0, err.value
    assert out.raw[:n.value] == data
    open(str(tmp_path / "bad.bam"), "wb").write(b"not a bam")
    assert lib.ggh_bam_to_fasta(str(tmp_path / "bad.bam").encode(), out, C.c_size_t(len(out)), C.byref(n), err, 256) != 0


def test_bgzf_lz_host_model():
    """ggh::bgzf_lz_model, the host model that pins the device BGZF encoder with LZ77 matches (gg_lz.h): the members inflate under zlib
    (CRC-32 and ISIZE checked) for read-like streams and for streams the code was not built for; the files are about the size of zlib -1's
    (fragSim records within 2 %, -artp smaller); member framing is BGZF; incompressible members are stored."""
    import gzip
    lib = api.load_library()
    rng = np.random.default_rng(5)

    def zlib_bgzf(data, level):
        out = C.create_string_buffer(len(data) + len(data) // 100 + 70000); m = C.c_size_t()
        assert lib.ggh_bgzf_compress(data, C.c_size_t(len(data)), 2, level, 1, out, C.c_size_t(len(out)), C.byref(m)) == 0
        return m.value
    for dialect in ("frag", "artp"):
        data = util.read_like_stream(rng, 6000, dialect)
        z, hist = util.bgzf_lz_model(data)
        assert gzip.decompress(z) == data
        assert len(z) <= (1.02 if dialect == "frag" else 1.0) * zlib_bgzf(data, 1), (dialect, len(z), zlib_bgzf(data, 1))    # frag: within 2 %; -artp: smaller
        assert sum(hist[257:286]) == sum(hist[286:]) > 0 and hist[256] == min(96, (len(data) + 65279) // 65280)      # one distance per length; sampled members
        p, sizes = 0, []
        while p < len(z):
            assert z[p:p + 4] == b"\x1f\x8b\x08\x04" and z[p + 12:p + 14] == b"BC"
            sizes.append((z[p + 16] | (z[p + 17] << 8)) + 1); p += sizes[-1]
        assert p == len(z) and len(sizes) == (len(data) + 65279) // 65280 + 1 and max(sizes) <= 65536
    artp = util.read_like_stream(rng, 3000, "artp")
    assert len(util.bgzf_lz_model(artp)[0]) < 0.33 * len(artp)                       # the forward / reverse overlap is found: ~2.5 bits per byte (Huffman alone: 3.1)
    for data in (bytes(range(256)) * 600, rng.integers(0, 256, size=200_000, dtype=np.uint8).tobytes(), b"A" * 70000, b"ACGT\n" * 13057 + b"x",
                 b"", b"G", b"ACGTACGTA", b"ACGTAC" * 2, bytes(70000), artp[:65280], artp[:65281], artp[:2176 * 3 + 5]):
        z, _ = util.bgzf_lz_model(data)
        assert gzip.decompress(z) == data
    z, _ = util.bgzf_lz_model(rng.integers(0, 256, size=65280 * 2, dtype=np.uint8).tobytes())
    assert len(z) == 2 * (18 + 5 + 65280 + 8) + 28                                    # random bytes: stored blocks
    assert len(util.bgzf_lz_model(b"A" * 65280)[0]) < 600                             # a run of one letter: distance-1 matches of 258
