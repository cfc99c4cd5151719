#!/usr/bin/env python3
"""Golden OUTPUTS of the reference's own filters (oracle/_ref: unmodified reference sources + our shim) on a small fixed input,
for every case where the filters are deterministic.  Run in the build container (needs oracle/_ref); writes ref_exact.json,
which travels with the repo so that `-m gpu` tests compare the CUDA path with the reference's bytes directly.

    python tests/golden/make_golden_ref.py
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "src")


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr
    return r.stdout


def main():
    rng = np.random.default_rng(2026)
    recs = []
    for i in range(64):
        n = int(rng.integers(16, 200))
        s = synth._BASES[rng.integers(0, 4, size=n)].tobytes()
        if i % 5 == 0:
            s = s[:n // 2].lower() + s[n // 2:]
        recs.append(b">chr%d:%s:%d:%d:%d%s\n%s\n" % (i % 3 + 1, b"+-"[i % 2:i % 2 + 1], i * 101, i * 101 + n, n, b"b7" if i % 2 else b"", s))
    data = b"".join(recs)
    out = {"input": data.decode(), "adpt": {}, "deam": {}}
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "f.fa")
        open(src, "wb").write(data)
        for key, opts in (("plain", []), ("name_tag", ["-name", "-tag", "lib7"])):
            a = {"opts": opts}
            a["stdout4"] = run([REF + "/adptSim", "-l", "75"] + opts + [src]).decode()
            run([REF + "/adptSim", "-l", "75"] + opts + ["-arts", d + "/s.fa", src]); a["arts"] = open(d + "/s.fa").read()
            run([REF + "/adptSim", "-l", "75"] + opts + ["-artp", d + "/p.fa", src]); a["artp"] = open(d + "/p.fa").read()
            run([REF + "/adptSim", "-l", "75"] + opts + ["-fr", d + "/fr.gz", "-rr", d + "/rr.gz", src])
            a["fr"] = gzip.open(d + "/fr.gz").read().decode(); a["rr"] = gzip.open(d + "/rr.gz").read().decode()
            out["adpt"][key] = a
        for key, ct, ga in (("identity", 0.0, 0.0), ("all_ct", 1.0, 0.0), ("all_ct_ga", 1.0, 1.0)):
            s5 = np.zeros((30, 12)); s3 = np.zeros((25, 12))
            s5[:, 5] = s3[:, 5] = ct; s5[:, 6] = s3[:, 6] = ga
            synth.write_matrix_dat(d + "/m-", s5, s3)
            out["deam"][key] = {"ct": ct, "ga": ga, "rows5": 30, "rows3": 25,
                                "plain": run([REF + "/deamSim", "-matfile", d + "/m-", src]).decode(),
                                "name": run([REF + "/deamSim", "-name", "-matfile", d + "/m-", src]).decode(),
                                "last_name": run([REF + "/deamSim", "-last", "-name", "-matfile", d + "/m-", src]).decode()}
        out["deam"]["briggs_all_ss"] = {"pars": [0, 1e-9, 0, 1], "name": run([REF + "/deamSim", "-name", "-damage", "0,0.000000001,0,1", src]).decode()}
        out["deam"]["briggs_all_ds"] = {"pars": [0, 1, 1, 0], "name": run([REF + "/deamSim", "-name", "-damage", "0,1,1,0", src]).decode()}
        # ---- methylation mode (deamSim -matfilenonmeth / -matfilemeth, deamSim.cpp:1630-1787): records with per-base case (lower-case =
        # methylated), some unresolved letters in both cases; tables of 0/1 rates with DIFFERENT row counts make the output deterministic
        mrecs = []
        for i in range(48):
            n = int(rng.integers(8, 120))
            b = synth._BASES[rng.integers(0, 4, size=n)].copy()
            low = rng.random(n) < (0.5 if i % 3 else 0.15)
            b[low] += 32
            if i % 7 == 0:
                b[int(rng.integers(0, n))] = ord("N"); b[int(rng.integers(0, n))] = ord("n")
            mrecs.append(b">m%d:+:%d:%d:%d\n%s\n" % (i % 4, i * 7, i * 7 + n, n, b.tobytes()))
        mdata = b"".join(mrecs)
        msrc = os.path.join(d, "m.fa")
        open(msrc, "wb").write(mdata)
        out["meth_input"] = mdata.decode(); out["meth"] = {}
        for key, no, wi in (("wi_ctga", (0.0, 0.0, 30, 25), (1.0, 1.0, 20, 15)), ("no_ct", (1.0, 0.0, 12, 40), (0.0, 0.0, 33, 9))):
            for tag, (ct, ga, r5, r3) in (("no", no), ("wi", wi)):
                s5 = np.zeros((r5, 12)); s3 = np.zeros((r3, 12))
                s5[:, 5] = s3[:, 5] = ct; s5[:, 6] = s3[:, 6] = ga
                synth.write_matrix_dat(d + "/%s-" % tag, s5, s3)
            base = [REF + "/deamSim", "-matfilenonmeth", d + "/no-", "-matfilemeth", d + "/wi-"]
            out["meth"][key] = {"no": list(no), "wi": list(wi),
                                "plain": run(base + [msrc]).decode(), "name": run(base + ["-name", msrc]).decode(),
                                "last_name": run(base + ["-last", "-name", msrc]).decode()}
        # ---- fragSim --case (fragSim.cpp:318-326,548-551) on a contig exactly as long as the fragment (one admissible window): the
        # record keeps the case of the file on the + strand and complements within the case on the - strand
        seq = synth._BASES[rng.integers(0, 4, size=61)].copy(); seq[rng.random(61) < 0.4] += 32
        fa, fai = synth.fasta_bytes([("cs", seq)], 23)
        open(d + "/c.fa", "wb").write(fa); synth.write_fai(d + "/c.fa.fai", fai)
        got = set()
        for sd in range(1, 9):
            got |= set(run([REF + "/fragSim", "--seed", str(sd), "--case", "-n", "8", "-l", "61", "-tag", "e", d + "/c.fa"]).decode().split("\n>"))
        out["frag_case"] = {"fasta": fa.decode(), "fai": [[e.name, e.len, e.offset, e.line_blen, e.line_len] for e in fai],
                            "records": sorted(">" + r.lstrip(">").rstrip("\n") + "\n" for r in got)}
    json.dump(out, open(os.path.join(HERE, "ref_exact.json"), "w"), indent=0)
    print("wrote ref_exact.json", os.path.getsize(os.path.join(HERE, "ref_exact.json")), "bytes")


if __name__ == "__main__":
    main()
