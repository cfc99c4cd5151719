#!/usr/bin/env python3
"""Regenerates tests/golden/ref_tables.json and tests/golden/mapdamage_sample.txt from the reference's
data files (run in the build container, where /root/reference exists; the GPU box only sees the outputs).

The numbers are the reference's shipped tables (src/matrices/*.dat, src/sizefreq.size.gz, src/sizedist.size.gz,
examplesMapDamage/results_LaBrana/misincorporation.txt) parsed by oracle/parsers.py, stored as data so that
tests and bench.py can rebuild .dat/.size files in the reference's formats without reading /root/reference."""
import collections
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import parsers  # noqa: E402

REF = os.environ.get("GG_REFERENCE", "/root/reference")


def main():
    out = {"source": "grenaud/gargammel data files, parsed by tests/golden/make_golden.py", "matrices": {}}
    for name in ("single-", "double-"):
        s5, s3 = parsers.load_matrix_dat(os.path.join(REF, "src", "matrices", name))
        out["matrices"][name] = {"sub5": s5, "sub3": s3}          # sub3 already reversed: row 0 = terminal base
    lens, cum = parsers.load_size_freq(os.path.join(REF, "src", "sizefreq.size.gz"))
    import gzip
    with gzip.open(os.path.join(REF, "src", "sizefreq.size.gz"), "rt") as f:
        out["sizefreq"] = [[int(a), b] for a, b in (l.rstrip("\n").split("\t") for l in f if l.strip())]   # freq kept as TEXT
    sl = parsers.load_size_list(os.path.join(REF, "src", "sizedist.size.gz"))
    h = collections.Counter(sl)
    out["sizedist_hist"] = sorted([[k, v] for k, v in h.items()])
    out["sizedist_first32"] = sl[:32]
    md = os.path.join(REF, "examplesMapDamage", "results_LaBrana", "misincorporation.txt")
    m5, m3 = parsers.load_mapdamage(md)
    out["mapdamage_LaBrana"] = {"rows": len(m5), "sub5_row0": m5[0], "sub3_row0": m3[0], "sub5_row69": m5[-1], "sub3_row69": m3[-1]}
    # a small sample in the original format: header block + every line of the first two chromosomes
    with open(md) as f:
        lines = f.read().split("\n")
    keep = lines[:4] + [l for l in lines[4:] if l and l.split("\t")[0] in ("1", "2")]
    with open(os.path.join(HERE, "mapdamage_sample.txt"), "w") as f:
        f.write("\n".join(keep) + "\n")
    s5, s3 = parsers.load_mapdamage(os.path.join(HERE, "mapdamage_sample.txt"))
    out["mapdamage_sample"] = {"sub5": s5, "sub3": s3}
    out["dnacomp_dist2"] = parsers.load_dnacomp(os.path.join(REF, "src", "dnacomp.txt"), 2)     # 5p+,5p-,3p+,3p- x [4][4]
    with open(os.path.join(HERE, "ref_tables.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote ref_tables.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
