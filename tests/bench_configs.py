#!/usr/bin/env python3
"""Device-timed throughput of the other BASELINE.json configurations on one B200 (bench.py covers C2, the headline).

  C1  deamSim -matfile single- on 1 M fixed 50-bp fragments (stand-alone stage over a FASTA record stream)
  C3' Briggs -damage 0.03,0.4,0.01,0.3 on a synthetic diploid genome (2 x G bp, default G = 500 Mb; BASELINE: 3.1 Gb), -f sizefreq
  C4' microbial-heavy: K synthetic bacterial genomes (2-6 Mb, 1-50 contigs, log-normal list weights), --comp 0.9,0.02,0.08
  C5  fragSim only: -l 40 / -f sizefreq / -s sizedist on a 50 Mb genome (EMIT_FRAG), and the fused -artp variants

Each line: fragments/s from CUDA-event time, algorithmic GB/s and the fraction of the measured HBM copy bandwidth.
Sampled id windows of every run are checked bit-exact against the CPU oracle: that makes this a test harness, so it lives under tests/
(run it as `python tests/bench_configs.py`; it is not collected by pytest)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ -> repo root
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from gargammel_b200 import api
from tests import synth  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def genome(seed, n_contigs, total, prefix, n_frac=0.005):
    rng = np.random.default_rng(seed)
    edges = np.linspace(0, total, n_contigs + 1).astype(np.int64)
    return synth.fasta_bytes([("%s%d" % (prefix, i + 1), synth.random_contig(rng, int(edges[i + 1] - edges[i]), n_frac)) for i in range(n_contigs)], 60)


def timed(ctx, first, count, emit, reps=3):
    ctx.run_fused(first, count, emit, fetch=False)
    best, info = 1e30, None
    for _ in range(reps):
        ctx.flush_l2()
        _, info = ctx.run_fused(first, count, emit, fetch=False)
        best = min(best, info.device_ms)
    return best, info


def report(name, frags, ms, alg_bytes, extra=""):
    gbs = alg_bytes / (ms / 1e3) / 1e9
    line = {"config": name, "fragments": frags, "ms": round(ms, 3), "fragments_per_s": frags / (ms / 1e3), "algorithmic_GBps": gbs,
            "frac_of_measured_hbm": gbs / peak(), "note": extra}
    print(json.dumps(line), flush=True)


def check_window(ctx, orc, first, count, emit, what):
    a, _ = ctx.run_fused(first, count, emit)
    b, _ = orc.run_fused(first, count, emit)
    assert a == b, what + ": CUDA output differs from the oracle"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-bp", type=int, default=500_000_000)
    ap.add_argument("--c3-frags", type=int, default=20_000_000)
    ap.add_argument("--c4-genomes", type=int, default=200)
    ap.add_argument("--c4-frags", type=int, default=20_000_000)
    ap.add_argument("--c5-frags", type=int, default=20_000_000)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    from oracle import oracle
    lens, cum = synth.golden_sizefreq()
    s5, s3 = synth.golden_matrix("single-")

    if "c1" not in args.skip:
        # ---- C1: stand-alone deamSim over a record stream resident on the device
        rng = np.random.default_rng(1)
        n, L = 1_000_000, 50
        seqs = synth._BASES[rng.integers(0, 4, size=(n, L))]
        rec_list = [b">chrS:+:%d:%d:50\n%s\n" % (i * 50, i * 50 + 50, seqs[i].tobytes()) for i in range(n)]
        recs = b"".join(rec_list)
        head = b"".join(rec_list[:30000])                   # whole records: a record's damage depends on its length
        ctx = api.Context(0, 42)
        ctx.set_matrix(0, s5, s3)
        ctx.run_deam(recs, 0)
        t = []
        for _ in range(5):
            ctx.flush_l2()
            out, info = ctx.run_deam(recs, 0)
            t.append(info.device_ms)
        o = oracle.OracleContext(42); o.set_matrix(0, s5, s3)
        assert out[:len(head)] == o.run_deam(head, 0)[0]
        report("C1 deamSim -matfile single-, 1M x 50bp (record index + damage kernels; the H2D of the record stream is enqueued before the first event)", n, min(t), 2.0 * len(recs),
               "bytes = record stream read + written; %d kernels" % info.launches)
        ctx.close()

    if "c5" not in args.skip:
        # ---- C5: fragSim-only and fused variants on a 50 Mb genome
        fa, fai = genome(7, 5, 50_000_000, "chr")
        ctx = api.Context(0, 42); o = oracle.OracleContext(42)
        gid = ctx.upload_genome(fa, fai); o.upload_genome(fa, fai)
        n = args.c5_frags
        for c in (ctx, o):
            c.set_matrix(0, s5, s3); c.set_adapters(read_len=75)
            c.set_plan([api.Range(0, n, gid, 0, "e")])
        for name, setter in (("-l 40", lambda c: c.set_sizes(api.SIZE_FIXED, fixed_len=40)),
                             ("-f sizefreq", lambda c: c.set_sizes(api.SIZE_FREQ, lens, cum)),
                             ("-s sizedist", lambda c: c.set_sizes(api.SIZE_LIST, synth.golden_sizelist()))):
            setter(ctx); setter(o)
            for emit, en in ((api.EMIT_FRAG, "fragSim only"), (api.EMIT_ARTP, "fused -matfile -artp 2x75")):
                ms, info = timed(ctx, 0, n, emit)
                check_window(ctx, o, n // 2, 2000, emit, "C5 " + name)
                report("C5 %s, %s, 50 Mb genome" % (name, en), n, ms, info.bytes + info.gather_bytes, "attempts/fragment %.4f" % (info.attempts / n))
        ctx.close()

    if "c3" not in args.skip:
        # ---- C3': Briggs on a large diploid genome (24 contigs per haplotype, ~7.6 % N)
        t0 = time.time()
        h1 = genome(31, 24, args.c3_bp, "chr", n_frac=0.076)
        h2 = genome(32, 24, args.c3_bp, "chr", n_frac=0.076)
        ctx = api.Context(0, 42); o = oracle.OracleContext(42)
        n = args.c3_frags
        for c in (ctx, o):
            g1 = c.upload_genome(*h1); g2 = c.upload_genome(*h2)
            c.set_sizes(api.SIZE_FREQ, lens, cum); c.set_briggs(0, 0.03, 0.4, 0.01, 0.3); c.set_adapters(read_len=75)
            c.set_plan([api.Range(0, n // 2, g1, 0, "e1_0"), api.Range(n // 2, n - n // 2, g2, 0, "e2_0")])
        for emit, en in ((api.EMIT_DEAM, "fragSim|deamSim"), (api.EMIT_ARTP, "fused -artp 2x75")):
            ms, info = timed(ctx, 0, n, emit)
            check_window(ctx, o, n // 2 - 1000, 2000, emit, "C3")
            report("C3' Briggs 0.03,0.4,0.01,0.3, diploid 2 x %d Mb, %s" % (args.c3_bp // 1_000_000, en), n, ms, info.bytes + info.gather_bytes,
                   "attempts/fragment %.4f; setup %.0f s" % (info.attempts / n, time.time() - t0))
        ctx.close()

    if "c4" not in args.skip:
        # ---- C4': many bacterial genomes with abundance weights
        rng = np.random.default_rng(4)
        K = args.c4_genomes
        w = rng.lognormal(size=K); w /= w.sum()
        ctx = api.Context(0, 42); o = oracle.OracleContext(42)
        endo = genome(41, 5, 50_000_000, "chrE"); cont = genome(42, 5, 50_000_000, "chrC")
        bact = [genome(1000 + k, int(rng.integers(1, 51)), int(rng.integers(2_000_000, 6_000_001)), "b%d_" % k, n_frac=0.001) for k in range(K)]
        n = args.c4_frags
        import ctypes as C
        lib = api.load_library()
        tot = (C.c_uint64 * 3)(); e = (C.c_uint64 * 1)(); cc = (C.c_uint64 * 1)(); b = (C.c_uint64 * K)(); err = C.create_string_buffer(256)
        assert lib.ggh_apportion(C.c_uint64(n), C.c_double(0.9), C.c_double(0.02), C.c_double(0.08), 1, (C.c_uint64 * 1)(50_000_000), 1,
                                 (C.c_double * K)(*w.tolist()), K, C.c_uint64(42), tot, e, cc, b, err, 256) == 0
        for c in (ctx, o):
            plan, first = [], 0
            ge = c.upload_genome(*endo)
            plan.append(api.Range(first, int(e[0]), ge, 0, "e")); first += int(e[0])
            for k in range(K):
                if c is o and k >= 3:
                    gb = 0          # the oracle only replays the first ranges (its genomes are host copies); windows below stay inside them
                else:
                    gb = c.upload_genome(*bact[k])
                if b[k]:
                    plan.append(api.Range(first, int(b[k]), gb, 0, "b%d" % (k + 1))); first += int(b[k])
            gc_ = c.upload_genome(*cont)
            plan.append(api.Range(first, int(cc[0]), gc_, -1, "c1")); first += int(cc[0])
            c.set_sizes(api.SIZE_FREQ, lens, cum); c.set_matrix(0, s5, s3); c.set_adapters(read_len=75)
            c.set_plan(plan)
        ms, info = timed(ctx, 0, first, api.EMIT_ARTP)
        check_window(ctx, o, int(e[0]) - 1000, 2000, api.EMIT_ARTP, "C4")      # spans the endo -> b1 boundary
        report("C4' %d bacterial genomes (2-6 Mb, 1-50 contigs), --comp 0.9,0.02,0.08, fused -matfile -artp" % K, first, ms, info.bytes + info.gather_bytes,
               "%d plan ranges; attempts/fragment %.4f" % (len(plan), info.attempts / first))
        ctx.close()


if __name__ == "__main__":
    main()
