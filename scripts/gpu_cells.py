#!/usr/bin/env python
"""bin/gargammel-b200 on a multi-cell directory (endo/C0..C2, diploid) for real: the records of <prefix>_a.fa follow the plan that --mock
prints (ranges in order, one tag per range), the intermediates hold the same records per source class."""
import collections
import gzip
import os
import pathlib
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_cli  # noqa: E402

t = pathlib.Path(tempfile.mkdtemp())
d = test_cli._cell_dirs(t, 3, 2)
common = ["--comp", "0,0.1,0.9", "-n", "30000", "-l", "40", "--seed", "5", "--noart", "--keep", "-damage", "0.03,0.4,0.01,0.3"]
exe = os.path.join(ROOT, "bin", "gargammel-b200")
m = subprocess.run([exe, "--mock"] + common + ["-o", str(t / "m"), str(d)], capture_output=True, text=True)
plan = [(int(a), int(b), tag) for a, b, tag in re.findall(r"fused range ids \[(\d+),(\d+)\) tag (\S+) file", m.stderr)]
r = subprocess.run([exe] + common + ["-o", str(t / "s"), str(d)], capture_output=True, text=True)
assert r.returncode == 0, r.stderr[-1500:]
heads = [l for l in open(t / "s_a.fa") if l.startswith(">")]
assert len(heads) == plan[-1][1] == 30002, (len(heads), plan[-1])
for a, b, tag in plan:
    got = collections.Counter(re.sub(r"^.*:\d+", "", h.strip()) for h in heads[a:b])          # what follows the length field = the tag
    assert got == {tag: b - a}, (tag, got.most_common(3))
n_e = sum(b - a for a, b, tag in plan if tag.startswith("e"))
assert gzip.open(t / "s.e.fa.gz").read().count(b">") == n_e and gzip.open(t / "s.c.fa.gz").read().count(b">") == 3000
assert gzip.open(t / "s_d.fa.gz").read().count(b">") == 30002
print("multi-cell run ok:", [(tag, b - a) for a, b, tag in plan])
