"""Host-side parsers and record readers of the drop-in executables under AddressSanitizer / UBSan (CPU only): mutated size tables,
matrices, profiles, mapDamage tables, dnacomp files, .fai / FASTA, bact lists and BAM files must be parsed or rejected without a memory
error; the whole-file and the streaming FASTA readers must agree on every random record stream."""
import gzip
import os
import shutil
import subprocess

import pytest

from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAT = os.path.join(ROOT, "tests", "native")
SAN = ["-std=c++11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]


@pytest.fixture(scope="module")
def host_objs(tmp_path_factory):
    """gg_host.cpp + the upload stub, compiled once with the sanitizers."""
    d = tmp_path_factory.mktemp("san")
    objs = []
    for src in (os.path.join(ROOT, "gargammel_b200", "csrc", "gg_host.cpp"), os.path.join(NAT, "stub_upload.cpp")):
        o = str(d / (os.path.basename(src) + ".o"))
        r = subprocess.run(["g++"] + SAN + ["-c", "-o", o, src, "-I" + os.path.join(ROOT, "include")], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("sanitizer build not available here: " + r.stderr[-300:])
        objs.append(o)
    return objs


def _build(tmp, name, objs):
    exe = str(tmp / name)
    r = subprocess.run(["g++"] + SAN + ["-o", exe, os.path.join(NAT, name + ".cpp")] + objs + ["-I" + os.path.join(ROOT, "include"), "-lz", "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available here: " + r.stderr[-300:])
    return exe


pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def test_parsers_survive_mutated_inputs(tmp_path, host_objs):
    exe = _build(tmp_path, "fuzz_parsers", host_objs)
    s, w = tmp_path / "seeds", tmp_path / "work"
    s.mkdir(); w.mkdir()
    fa, fai = synth.make_genome(3, 3, 5000, prefix="chr")
    (s / "g.fa").write_bytes(fa); synth.write_fai(str(s / "g.fai"), fai)
    synth.write_sizefreq(str(s / "sf"))
    (s / "sl").write_text("\n".join(str(x) for x in synth.golden_sizelist()[:300]) + "\n")
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(str(s / "m-"), s5, s3); synth.write_matrix_dat(str(w / "m-"), s5, s3)
    synth.write_dnacomp(str(s / "dc"))
    (s / "list").write_text("b1.fa\t0.4\nb2.fa\t0.6\n")
    gold = os.path.join(ROOT, "tests", "golden")
    shutil.copy(os.path.join(gold, "mapdamage_sample.txt"), s / "mis")
    shutil.copy(os.path.join(gold, "sample_5p.prof"), s / "p5"); shutil.copy(os.path.join(gold, "sample_5p.prof"), w / "p_5p.prof")
    shutil.copy(os.path.join(gold, "sample_3p.prof"), w / "p_3p.prof")
    # an unaligned BAM as the library writes it, inflated (gzread passes plain bytes through: the mutations reach the record parser)
    import ctypes as C
    from gargammel_b200 import api
    lib = api.load_library()
    lib.ggh_bam_from_fasta.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_int]
    recs = b"".join(b">r%d\nACGTNACGTTGCA\n" % i for i in range(50))
    err = C.create_string_buffer(256)
    assert lib.ggh_bam_from_fasta(recs, len(recs), str(s / "x.bam").encode(), err, 256) == 0, err.value
    (s / "x.bam.raw").write_bytes(gzip.open(str(s / "x.bam")).read())
    seeds = [("fai", "g.fai"), ("fasta", "g.fa"), ("sizelist", "sl"), ("sizefreq", "sf"), ("dnacomp", "dc"), ("dat5", "m-5.dat"), ("dat3", "m-3.dat"),
             ("prof5", "p5"), ("mapd", "mis"), ("list", "list"), ("bam", "x.bam.raw")]
    args = [exe, str(w), "4000", "11"]
    for name, f in seeds:
        args += [name, str(s / f)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "done:" in r.stdout and "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, (r.stdout[-300:], r.stderr[-3000:])


def test_record_readers_agree_on_random_streams(tmp_path, host_objs):
    exe = _build(tmp_path, "fuzz_records", host_objs)
    r = subprocess.run([exe, "4000", "3", str(tmp_path / "r.fa")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "different 0" in r.stdout and "runtime error" not in r.stderr, (r.stdout[-2000:], r.stderr[-3000:])
