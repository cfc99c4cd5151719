"""Shared scaffolding of the parity tests: build the same state in a product Context and an oracle Context."""
import os

import numpy as np

from gargammel_b200 import api
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_genomes(seed=7, n_files=3):
    """A few small multi-contig genomes with N runs, lower-case stretches and awkward line widths."""
    out = []
    for g in range(n_files):
        fa, fai = synth.make_genome(seed * 100 + g, n_contigs=3 + g, total_len=6000 + 2500 * g, prefix="c%dx" % g,
                                    width=[60, 37, 80][g % 3], n_frac=0.03, n_run=(5, 120), lower_frac=0.2)
        out.append((fa, fai))
    return out


def circ_genome(seed=11, lens=(45, 310, 900, 130), width=37):
    """Small N-free contigs (names m1..mK) so that --circ wraps often; returns (fasta, fai)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return synth.fasta_bytes([("m%d" % (i + 1), synth.random_contig(rng, n, 0.0)) for i, n in enumerate(lens)], width)


def contig_seqs(fa, fai):
    return {e.name: fa[e.offset:e.offset + e.len + e.len // e.line_blen + 1].replace(b"\n", b"")[:e.len].upper() for e in fai}


def configure(ctx, genomes, sizes="freq", damage="matrix", read_len=75, minlen=0, maxlen=1000, norev=False,
              adapters=(api.DEFAULT_ADAPTER_F, api.DEFAULT_ADAPTER_S), fixed_len=40, clamp_last=True, matrix="single-", case=False):
    if case:
        ctx.set_case(True)                                   # fragSim --case: before the uploads
    ids = [ctx.upload_genome(fa, fai) for fa, fai in genomes]
    if sizes == "freq":
        lens, cum = synth.golden_sizefreq()
        ctx.set_sizes(api.SIZE_FREQ, lens, cum, minlen=minlen, maxlen=maxlen)
    elif sizes == "lognormal":
        lens, cum = lognormal_table(4.1, 0.35)
        ctx.set_sizes(api.SIZE_FREQ, lens, cum, minlen=minlen, maxlen=min(maxlen, 4095))
    elif sizes == "list":
        ctx.set_sizes(api.SIZE_LIST, synth.golden_sizelist(), minlen=minlen, maxlen=maxlen)
    else:
        ctx.set_sizes(api.SIZE_FIXED, fixed_len=fixed_len, minlen=minlen, maxlen=maxlen)
    ctx.set_norev(norev)
    set_damage(ctx, 0, damage, clamp_last=clamp_last, matrix=matrix)
    ctx.set_adapters(adapters[0], adapters[1], read_len)
    return ids


def lognormal_table(loc, scale, cap=4095):
    """--loc/--scale lengths through the host library's inverse-CDF table (ggh::lognormal_table)."""
    import ctypes as C
    lib = api.load_library()
    ln = np.zeros(cap + 2, dtype=np.int32); cm = np.zeros(cap + 2)
    n = lib.ggh_lognormal_table(C.c_double(loc), C.c_double(scale), C.c_int(cap), ln.ctypes.data_as(C.POINTER(C.c_int32)), cm.ctypes.data_as(C.POINTER(C.c_double)))
    assert n == cap + 2
    return ln.tolist(), cm.tolist()


def set_damage(ctx, slot, damage, clamp_last=True, matrix="single-"):
    s5, s3 = synth.golden_matrix(matrix)
    if damage == "matrix":
        ctx.set_matrix(slot, s5, s3, clamp_last=clamp_last)
    elif damage == "briggs":
        ctx.set_briggs(slot, 0.03, 0.4, 0.01, 0.3)
    elif damage == "briggs_ss":
        ctx.set_briggs_ss(slot, 0.02, 0.5, 0.7, 0.25, 0.03)      # a small l3 makes "inside both overhangs" and "all single-stranded" common
    elif damage == "single":
        ctx.set_singledouble(slot, api.DMG_SINGLE, s5, s3)
    elif damage == "double":
        ctx.set_singledouble(slot, api.DMG_DOUBLE, s5, s3)
    elif damage == "meth":
        no5, no3, wi5, wi3 = meth_tables()
        ctx.set_matrix_meth(slot, no5, no3, wi5, wi3, clamp_last=clamp_last)
    elif damage == "none":
        ctx.clear_damage(slot)
    else:
        raise ValueError(damage)


def meth_tables():
    """-matfilenonmeth / -matfilemeth test tables with different row counts: unmethylated = the double-stranded profile cut to 9/7 rows,
    methylated = the single-stranded profile with its C->T rates raised (methylated cytosines deaminate to T and are not repaired)."""
    d5, d3 = synth.golden_matrix("double-")
    s5, s3 = synth.golden_matrix("single-")
    wi5, wi3 = s5.copy(), s3.copy()
    wi5[:, 5] = np.minimum(0.9, 0.05 + 3.0 * wi5[:, 5]); wi3[:, 5] = np.minimum(0.9, 0.05 + 3.0 * wi3[:, 5])
    return d5[:9].copy(), d3[:7].copy(), wi5, wi3


def plan3(ids, counts=(400, 300, 500), tags=("e1_0", "b1", "c1"), slots=(0, 0, -1)):
    """endo / bact / cont ranges in the reference's output order with its tag grammar (gargammel.pl:1476,1600,1648)."""
    out, first = [], 0
    for g, n, t, s in zip(ids, counts, tags, slots):
        out.append(api.Range(first, n, g, s, t))
        first += n
    return out


def records(data: bytes):
    lines = data.split(b"\n")
    assert lines[-1] == b""
    lines = lines[:-1]
    assert len(lines) % 2 == 0
    return [(lines[i], lines[i + 1]) for i in range(0, len(lines), 2)]


def first_diff(a: bytes, b: bytes):
    n = min(len(a), len(b))
    aa, bb = np.frombuffer(a[:n], dtype=np.uint8), np.frombuffer(b[:n], dtype=np.uint8)
    d = np.nonzero(aa != bb)[0]
    if len(d) == 0:
        return None if len(a) == len(b) else n
    return int(d[0])


def assert_same(a: bytes, b: bytes, what=""):
    i = first_diff(a, b)
    if i is not None:
        lo = max(0, i - 60)
        raise AssertionError("%s: outputs differ at byte %d (lens %d vs %d)\n cuda  : %r\n oracle: %r" % (what, i, len(a), len(b), a[lo:i + 60], b[lo:i + 60]))


def read_bam(path):
    """Minimal reader of an unaligned BAM (SAM/BAM specification 4.2): returns (header text, [(name, flag, seq, qual bytes, {tag: value})]).
    Every BGZF member is checked for the 'BC' extra field and inflated on its own."""
    import struct
    import zlib
    z = open(path, "rb").read()
    assert z[-28:] == api.BGZF_EOF, "BGZF EOF marker missing"
    raw = b""
    while z:
        assert z[:4] == b"\x1f\x8b\x08\x04" and z[12:14] == b"BC", "not a BGZF member"
        bsize = struct.unpack("<H", z[16:18])[0] + 1
        d = zlib.decompressobj(31)
        raw += d.decompress(z[:bsize])
        assert d.eof and d.unused_data == b""
        z = z[bsize:]
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack("<i", raw[4:8])[0]
    text = raw[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
    assert n_ref == 0
    recs = []
    while p < len(raw):
        block = struct.unpack("<i", raw[p:p + 4])[0]; p += 4
        b = raw[p:p + block]; p += block
        refid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack("<iiBBHHHiiii", b[:32])
        assert (refid, pos, nref, npos, tlen, n_cig, mapq) == (-1, -1, -1, -1, 0, 0, 0)
        q = 32
        name = b[q:q + l_name - 1]; assert b[q + l_name - 1] == 0; q += l_name
        packed = b[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        seq = bytes(b"=ACMGRSVTWYHKDBN"[(packed[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        qual = b[q:q + l_seq]; q += l_seq
        tags = {}
        while q < len(b):
            tag, typ = b[q:q + 2].decode(), chr(b[q + 2]); q += 3
            assert typ == "Z"
            e = b.index(b"\0", q); tags[tag] = b[q:e].decode(); q = e + 1
        recs.append((name, flag, seq, qual, tags))
    return text, recs


def bgzf_lz_model(data: bytes, eof=True):
    """Host model of the device BGZF encoder with LZ77 matches (ggh::bgzf_lz_model, parse specification gg_lz.h):
    (members [+ EOF marker], the 316 sampled symbol counts)."""
    import ctypes as C
    lib = api.load_library()
    out = C.create_string_buffer(len(data) + len(data) // 100 + 70000)
    m = C.c_size_t(); hist = (C.c_uint64 * 316)()
    rc = lib.ggh_bgzf_lz_model(data, C.c_size_t(len(data)), 1 if eof else 0, out, C.c_size_t(len(out)), C.byref(m), hist)
    assert rc == 0, rc
    return out.raw[:m.value], list(hist)


def read_like_stream(rng, n_rec, dialect="artp", read_len=75):
    """FASTA records shaped like the simulator's output (random bases: nothing repeats but what the record structure repeats)."""
    from tests import synth
    ad_f, ad_s = api.DEFAULT_ADAPTER_F.encode(), api.DEFAULT_ADAPTER_S.encode()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    out = []
    for i in range(n_rec):
        L = int(rng.integers(30, 140)); st = int(rng.integers(1, 50_000_000))
        frag = synth._BASES[rng.integers(0, 4, L)].tobytes()
        hdr = b">chr%d:%s:%d:%d:%de1_0" % (rng.integers(1, 23), b"+-"[i & 1:(i & 1) + 1], st, st + L, L)
        if dialect == "frag":
            out.append(hdr + b"\n" + frag + b"\n"); continue
        pad = synth._BASES[rng.integers(0, 4, 2 * read_len)].tobytes()
        fwd = (frag + ad_f + pad)[:read_len]
        rev = (frag.translate(comp)[::-1] + ad_s + pad[::-1])[:read_len]
        out.append(hdr + b"\n" + fwd + rev.translate(comp)[::-1] + b"\n")
    return b"".join(out)
