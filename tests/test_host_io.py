"""Host-side I/O of the drop-in executables without a GPU: the parallel plain-file writer (Sink::write_big) and the in-process
`gzip -f` of the fused driver's closing steps (gargammel.pl:1871-1889).  tests/native/io_check.cpp is compiled here with g++."""
import gzip
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gargammel_b200")


@pytest.fixture(scope="module")
def io_check(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(LIBDIR, "libgargammel_b200.so")):
        pytest.skip("needs g++ and the built library")
    exe = str(tmp_path_factory.mktemp("io") / "io_check")
    subprocess.run(["g++", "-std=c++11", "-O2", "-o", exe, os.path.join(ROOT, "tests", "native", "io_check.cpp"), "-L" + LIBDIR, "-lgargammel_b200",
                    "-Wl,-rpath," + LIBDIR, "-lz", "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("min_bytes", [None, "1", "4", "1000000"])
def test_sink_parallel_writer_is_the_concatenation(io_check, tmp_path, min_bytes):
    env = dict(os.environ)
    if min_bytes is not None:
        env["GG_PWRITE_MIN"] = min_bytes            # pieces of at least this many bytes take the pwrite path (default 32 MiB)
    r = subprocess.run([io_check, "sink", str(tmp_path / "out.bin")], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "SAME" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.parametrize("n", [0, 1, 65279, 65280, 65281, 5_000_000])
def test_in_process_gzip_replaces_the_file(io_check, tmp_path, n):
    rec = b"@r1\nACGTTGCATTGACCA\n+\nIIIIHHHHGGGGFFF\n"
    data = (rec * (n // len(rec) + 1))[:n]
    p = tmp_path / "reads.fq"
    p.write_bytes(data)
    r = subprocess.run([io_check, "gzip", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "running cmd gzip -f " + str(p) in r.stderr              # echoed like the script's own commands
    assert not p.exists()
    z = (tmp_path / "reads.fq.gz").read_bytes()
    assert gzip.decompress(z) == data
    assert z.endswith(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))   # the BGZF EOF marker
    if shutil.which("zcat"):
        assert subprocess.run(["zcat", str(tmp_path / "reads.fq.gz")], capture_output=True).stdout == data
