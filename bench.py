#!/usr/bin/env python3
"""bench.py -- benchmark of the gargammel hot path on B200.

Metric (BASELINE.json): simulated fragments/sec, device-timed, plus % of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2]     our arm
  python bench.py --impl reference ...                                  the reference's own CPU filters (oracle/_ref: its
                                                                        sources + our shim) on the box's host cores

--config selects one of BASELINE.json's configurations at its NAMED size (C2 = the headline and the default):
  C1  deamSim -matfile single- on 1 M fixed 50-bp fragments (stand-alone stage over a FASTA record stream)
  C2  fragSim->deamSim->adptSim on a synthetic 50 Mb endo(x2, diploid)/cont/bact directory, --comp 0,0.08,0.92, -f sizefreq,
      -matfile single-, 2x75 (-artp), 10 M fragments per step and GPU
  C3  Briggs -damage 0.03,0.4,0.01,0.3 on a synthetic diploid 2 x 3.1 Gb genome, -f sizefreq, 100 M fragments per step and GPU
  C4  1000 synthetic bacterial genomes (2-6 Mb, 1-50 contigs, log-normal `list` weights), --comp 0.9,0.02,0.08, a 1 B-fragment job
      over 8 GPUs = 125 M fragments per step and GPU
  C5  1 B fragments per step over the GPUs of the run (strong scaling), -l 40 or --c5-sizes sizefreq|sizedist, fused -artp (or
      --c5-emit frag = fragSim only); the output goes through a device slab ring (device-timed) and is drained to the host
      with and without device BGZF (e2e)
One "step" = one pass of the hot path over the batch, in chunks of at most --chunk fragments (a slab ring in HBM).  N>1: one process
per GPU (torchrun), rank r takes the r-th contiguous slice of the fragment ids; genomes replicated, no data-path collective.
--scale f shrinks the batch (and, for C3/C4, the genomes) for quick runs; the line then says so.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 42
READ_LEN = 75
AD_F = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG"
AD_S = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- synthetic inputs of the named shapes
def _genome(seed, n_contigs, total, prefix, n_frac=0.005, snp_of=None):
    """iid bases, n_frac of the positions in N runs, 60-column lines; snp_of: copy of that contig list with 0.1 % SNPs."""
    from tests import synth
    rng = np.random.default_rng(seed)
    if snp_of is not None:
        contigs = []
        for name, seq in snp_of:
            s = seq.copy()
            pos = rng.integers(0, len(s), size=max(1, len(s) // 1000))
            ok = s[pos] != ord("N")
            s[pos[ok]] = synth._BASES[rng.integers(0, 4, size=int(ok.sum()))]
            contigs.append((name, s))
        return contigs
    edges = np.linspace(0, total, n_contigs + 1).astype(np.int64)
    return [("%s%d" % (prefix, i + 1), synth.random_contig(rng, int(edges[i + 1] - edges[i]), n_frac)) for i in range(n_contigs)]


def make_job(name, args, world):
    """-> dict describing the workload of one configuration: files (kind e/b/c, bytes, fai, weight), --comp, size model, damage, emit,
    fragments per step and GPU."""
    from tests import synth
    sc = args.scale
    J = {"name": name, "read_len": READ_LEN, "emit": "artp", "sizes": "sizefreq", "damage": "matrix", "scaling": "weak", "files": [], "comp": (0.0, 0.0, 1.0)}
    fa = lambda contigs: synth.fasta_bytes(contigs, 60)          # noqa: E731
    if name == "C2":
        bp = 50_000_000
        e1 = _genome(20251019, 5, bp, "chrE")
        J.update(comp=(0.0, 0.08, 0.92), n=int(10_000_000 * sc), genome_bp=bp,
                 workload="C2: fused fragSim->deamSim->adptSim, synthetic 50 Mb endo(x2)/cont/bact, --comp 0,0.08,0.92, -f sizefreq, -matfile single-, 2x75 -artp")
        J["files"] = [("e", "endo.1.fa") + fa(e1), ("e", "endo.2.fa") + fa(_genome(7, 0, 0, "", snp_of=e1)), ("c", "cont.1.fa") + fa(_genome(11, 5, bp, "chrC"))]
    elif name == "C3":
        bp = int(3_100_000_000 * min(1.0, sc * 4 if sc < 1 else 1.0))
        h1 = _genome(31, 24, bp, "chr", n_frac=0.076)
        J.update(comp=(0.0, 0.0, 1.0), n=int(100_000_000 * sc), genome_bp=bp, damage="briggs",
                 workload="C3: Briggs -damage 0.03,0.4,0.01,0.3, synthetic diploid 2 x %.2f Gb (24 contigs, 7.6 %% N), -f sizefreq, fused 2x75 -artp" % (bp / 1e9))
        J["files"] = [("e", "endo.1.fa") + fa(h1), ("e", "endo.2.fa") + fa(_genome(32, 0, 0, "", snp_of=h1))]
    elif name == "C4":
        K = max(3, int(1000 * min(1.0, sc * 4 if sc < 1 else 1.0)))
        rng = np.random.default_rng(4)
        w = rng.lognormal(size=K); w /= w.sum()
        J.update(comp=(0.9, 0.02, 0.08), n=int(125_000_000 * sc), genome_bp=50_000_000,
                 workload="C4: %d synthetic bacterial genomes (2-6 Mb, 1-50 contigs, log-normal list weights) + 50 Mb endo/cont, --comp 0.9,0.02,0.08, -f sizefreq, -matfile single-, 2x75 -artp; 1 B-fragment job over 8 GPUs" % K)
        J["files"] = [("e", "endo.1.fa") + fa(_genome(41, 5, 50_000_000, "chrE"))]
        for k in range(K):
            J["files"].append(("b", "b%d.fa" % (k + 1)) + fa(_genome(1000 + k, int(rng.integers(1, 51)), int(rng.integers(2_000_000, 6_000_001)), "b%d_" % k, n_frac=0.001)) + (float(w[k]),))
        J["files"].append(("c", "cont.1.fa") + fa(_genome(42, 5, 50_000_000, "chrC")))
    elif name == "C5":
        J.update(comp=(0.0, 0.0, 1.0), n=int(1_000_000_000 * sc) // world, genome_bp=50_000_000, scaling="strong", sizes=args.c5_sizes, emit=args.c5_emit,
                 workload="C5: 1 B fragments per step over the GPUs of the run, %s, %s, 50 Mb genome" % (
                     {"l40": "-l 40", "sizefreq": "-f sizefreq", "sizedist": "-s sizedist"}[args.c5_sizes], "fused -matfile single- 2x75 -artp" if args.c5_emit == "artp" else "fragSim only"))
        J["files"] = [("e", "endo.1.fa") + fa(_genome(7, 5, 50_000_000, "chr"))]
    else:
        raise SystemExit("unknown --config " + name)
    if sc != 1.0:
        J["workload"] += " [--scale %g: reduced size]" % sc
    if args.frags:
        J["n"] = args.frags
    return J


def job_plan(api, J, ids, n_total, seed=SEED):
    """gargammel.pl:1256-1283,1320-1349,1421-1431 through the host library; output order e, b, c (gargammel.pl:1698-1830)."""
    import ctypes as C
    lib = api.load_library()
    kinds = [f[0] for f in J["files"]]
    nE, nB, nC = kinds.count("e"), kinds.count("b"), kinds.count("c")
    tot = (C.c_uint64 * 3)(); e = (C.c_uint64 * max(nE, 1))(); c = (C.c_uint64 * max(nC, 1))(); b = (C.c_uint64 * max(nB, 1))()
    err = C.create_string_buffer(512)
    clen = [sum(x.len for x in f[3]) for f in J["files"] if f[0] == "c"]
    bw = [f[4] for f in J["files"] if f[0] == "b"]
    rc = lib.ggh_apportion(C.c_uint64(n_total), C.c_double(J["comp"][0]), C.c_double(J["comp"][1]), C.c_double(J["comp"][2]), max(nE, 1),
                           (C.c_uint64 * max(nC, 1))(*(clen or [1])), nC, (C.c_double * max(nB, 1))(*(bw or [1.0])), nB, C.c_uint64(seed), tot, e, c, b, err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    plan, first, counts = [], 0, []
    ie = ib = ic = 0
    # the driver's damage options deaminate endogenous and bacterial fragments only; the contaminant is passed through (gargammel.pl:640-648)
    for i, f in enumerate(J["files"]):
        if f[0] == "e":
            cnt, slot, tag = int(e[ie]), 0, ("e%d_0" % (ie + 1)) if nE == 2 else "e"; ie += 1
        elif f[0] == "b":
            cnt, slot, tag = int(b[ib]), 0, "b%d" % (ib + 1); ib += 1
        else:
            cnt, slot, tag = int(c[ic]), -1, "c%d" % (ic + 1); ic += 1
        counts.append(cnt)
        if cnt:
            plan.append(api.Range(first, cnt, ids[i], slot, tag)); first += cnt
    return plan, first, counts


def setup_tables(api, ctx, J):
    from tests import synth
    if J["sizes"] == "sizefreq":
        lens, cum = synth.golden_sizefreq()
        ctx.set_sizes(api.SIZE_FREQ, lens, cum, minlen=0, maxlen=1000)       # gargammel.pl passes -m 0 -M 1000
    elif J["sizes"] == "sizedist":
        ctx.set_sizes(api.SIZE_LIST, synth.golden_sizelist(), minlen=0, maxlen=1000)
    else:
        ctx.set_sizes(api.SIZE_FIXED, fixed_len=40, minlen=0, maxlen=1000)
    if J["damage"] == "briggs":
        ctx.set_briggs(0, 0.03, 0.4, 0.01, 0.3)
    else:
        s5, s3 = synth.golden_matrix("single-")
        ctx.set_matrix(0, s5, s3, clamp_last=True)
    ctx.set_adapters(api.DEFAULT_ADAPTER_F, api.DEFAULT_ADAPTER_S, J["read_len"])


# --------------------------------------------------------------------------- clocks sampler
class Clocks:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (nvidia-smi -lms cannot go below ~100 ms, longer than a whole timed region here); nvidia-smi as fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.stop_flag, self.nvml = [], None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), mhz, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            sm = [r[1] for r in rows]
            return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": int(self.max_mhz), "reasons": sorted({x for r in rows for x in r[2]}),
                    "samples": len(rows), "how": "NVML polled every 2 ms during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = [int(r[0]) for r in rows if r[0].isdigit()]
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(rows), "how": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------- host placement
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(local):
    """Run this rank (and allocate its pinned buffers: first touch) on the CPUs of the NUMA node its GPU hangs off, so that the D2H stream
    does not cross the socket interconnect.  Returns a description for the bench line; harmless no-op when sysfs does not tell."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id"):
            busid = "%04x:%02x:%02x.0" % (int(getattr(pr, "pci_domain_id", 0)), int(pr.pci_bus_id), int(pr.pci_device_id))
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True).stdout.strip().lower()
            busid = out[4:] if len(out) > 12 else out                    # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % busid).read())
        if node < 0:
            return {"numa_node": None, "note": "single NUMA domain / not reported"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as ex:
        return {"numa_node": None, "note": "not bound: %r" % (ex,)}


# --------------------------------------------------------------------------- reference arm (CPU)
def reference_step(workdir, J, counts, n_frag, cores, gzip_hops=False):
    """One bounded sample of the job: `cores` independent shards, each the reference composition of gargammel.pl:1457-1848 without the gzip
    hops: fragSim per source file (>> P.{e,b}.fa / P.c.fa), deamSim on endogenous+bacterial, contaminant appended untouched, adptSim -artp.
    gzip_hops=True inserts the driver's own `| gzip >` / gz-input hops (gargammel.pl:1510,1729,1847).  The per-file counts are the job's
    apportioned counts scaled to the sample.  Returns (fragments, seconds, kind)."""
    from oracle import oracle
    kind = "reference" if oracle.ref_bin("fragSim") else "port"
    total = max(1, sum(counts))
    per = max(1, n_frag // cores)
    share = [int(per * c / total) for c in counts]
    files = [(f[0], f[1], s) for f, s in zip(J["files"], share) if s > 0]
    tags = {}
    def tag_of(kind_, idx, n_of_kind):
        return ("e%d_0" % (idx + 1) if n_of_kind == 2 else "e") if kind_ == "e" else "%s%d" % (kind_, idx + 1)
    nk = {k: sum(1 for f in J["files"] if f[0] == k) for k in "ebc"}
    idx = {"e": 0, "b": 0, "c": 0}
    for f in J["files"]:
        tags[f[1]] = tag_of(f[0], idx[f[0]], nk[f[0]]); idx[f[0]] += 1
    sizes = {"sizefreq": "-f sf.size", "sizedist": "-s sl.size", "l40": "-l 40"}[J["sizes"]]
    dmg = "-damage 0.03,0.4,0.01,0.3" if J["damage"] == "briggs" else "-matfile m-"
    done = sum(s for _, _, s in files) * cores
    if kind == "reference":
        frag, deam, adpt = (oracle.ref_bin(x) for x in ("fragSim", "deamSim", "adptSim"))
        z = " | gzip" if gzip_hops else ""
        x = ".gz" if gzip_hops else ""
        parts = ["set -e; cd %s; S=$1; : > P$S.d.fa%s; : > P$S.c.fa%s" % (workdir, x, x)]
        for k_, name, s in files:
            dst = "P$S.c.fa" if k_ == "c" else "P$S.d.fa"
            parts.append("%s -tag %s -n %d -m 0 -M 1000 %s %s 2>/dev/null%s >> %s%s" % (frag, tags[name], s, sizes, name, z, dst, x))
        if J["emit"] == "artp":
            parts.append("%s %s P$S.d.fa%s 2>/dev/null%s > P${S}_d.fa%s" % (deam, dmg, x, z, x))
            if any(k_ == "c" for k_, _, _ in files):
                parts.append(("gzip -dc P$S.c.fa.gz | gzip >> P${S}_d.fa.gz" if gzip_hops else "cat P$S.c.fa >> P${S}_d.fa"))
            parts.append("%s -f %s -s %s -l %d -artp P${S}_a.fa P${S}_d.fa%s 2>/dev/null" % (adpt, AD_F, AD_S, J["read_len"], x))
        parts.append("rm -f P$S.d.fa%s P$S.c.fa%s P${S}_d.fa%s P${S}_a.fa" % (x, x, x))
        with open(os.path.join(workdir, "shard.sh"), "w") as fh:           # (a file: a thousand fragSim lines do not fit an argument)
            fh.write("\n".join(parts) + "\n")
        t0 = time.perf_counter()
        procs = [subprocess.Popen(["bash", os.path.join(workdir, "shard.sh"), str(i)]) for i in range(cores)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference shard failed: %r" % rcs)
        return done, dt, kind
    # port: the single-threaded oracle, one process per core (only when the reference could not be compiled here)
    code = ("import sys; sys.path.insert(0, %r); import bench; bench.oracle_shard(%r, %r, int(sys.argv[1]), %r)" % (ROOT, workdir, json.dumps({k: J[k] for k in ("sizes", "damage", "emit", "read_len")}), files))
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", code, str(i)]) for i in range(cores)]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("oracle shard failed: %r" % rcs)
    return done, dt, kind


def oracle_shard(workdir, jj, shard, files):
    from gargammel_b200 import api
    from oracle import oracle
    J = json.loads(jj)
    ctx = oracle.OracleContext(SEED + shard)
    plan, first = [], 0
    for k_, name, s in files:
        fa = open(os.path.join(workdir, name), "rb").read()
        fai = [api.FaiEntry(a, int(b), int(c), int(d), int(e)) for a, b, c, d, e in (l.split("\t") for l in open(os.path.join(workdir, name + ".fai")).read().splitlines())]
        plan.append(api.Range(first, s, ctx.upload_genome(fa, fai), -1 if k_ == "c" else 0, k_)); first += s
    setup_tables(api, ctx, J)
    ctx.set_plan(plan)
    ctx.run_fused(0, first, api.EMIT_ARTP if J["emit"] == "artp" else api.EMIT_FRAG)


def prepare_reference_dir(J):
    from tests import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    w = tempfile.mkdtemp(prefix="gg_ref_", dir=base)
    for f in J["files"]:
        with open(os.path.join(w, f[1]), "wb") as fh:
            fh.write(f[2])
        synth.write_fai(os.path.join(w, f[1] + ".fai"), f[3])
    synth.write_sizefreq(os.path.join(w, "sf.size"))
    with open(os.path.join(w, "sl.size"), "w") as fh:
        fh.write("\n".join(str(x) for x in synth.golden_sizelist()) + "\n")
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(os.path.join(w, "m-"), s5, s3)
    return w


def c1_records(n, L=50):
    from tests import synth
    rng = np.random.default_rng(1)
    seqs = synth._BASES[rng.integers(0, 4, size=(n, L))]
    return [b">chrS:+:%d:%d:50\n%s\n" % (i * 50, i * 50 + 50, seqs[i].tobytes()) for i in range(n)]


def c1_reference(n, cores):
    """C1 on the CPU: `deamSim -matfile single-` of the reference over the same records, one shard per core."""
    from oracle import oracle
    from tests import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    w = tempfile.mkdtemp(prefix="gg_ref_", dir=base)
    try:
        s5, s3 = synth.golden_matrix("single-")
        synth.write_matrix_dat(os.path.join(w, "m-"), s5, s3)
        recs = c1_records(n)
        per = n // cores
        for i in range(cores):
            with open(os.path.join(w, "f%d.fa" % i), "wb") as fh:
                fh.write(b"".join(recs[i * per:(i + 1) * per]))
        deam = oracle.ref_bin("deamSim")
        if deam is None:
            return None
        t0 = time.perf_counter()
        procs = [subprocess.Popen(["bash", "-c", "%s -matfile %s/m- %s/f%d.fa > /dev/null 2>&1" % (deam, w, w, i)]) for i in range(cores)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference deamSim failed: %r" % rcs)
        return per * cores, dt
    finally:
        shutil.rmtree(w, ignore_errors=True)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    cores = host_cores()
    if args.config == "C1":
        n = 1_000_000
        for _ in range(args.warmup):
            c1_reference(max(cores, n // 50), cores)
        frags, t = 0, 0.0
        for _ in range(args.steps):
            f, dt = c1_reference(n, cores); frags += f; t += dt
        kind, workload = "reference", "C1: deamSim -matfile single- on 1 M fixed 50-bp fragments"
        sample = "%d records per step, `deamSim -matfile` of the reference (its sources + our shim), one shard per core on %d cores" % (n, cores)
    else:
        from gargammel_b200 import api
        J = make_job(args.config, args, max(1, args.gpus))
        _, _, counts = job_plan(api, J, list(range(len(J["files"]))), J["n"] * max(1, args.gpus))
        w = prepare_reference_dir(J)
        try:
            per_core = 150000          # ~1 s per core per step: 15-25 core-seconds of CPU work per step
            n = per_core * cores
            for _ in range(args.warmup):
                reference_step(w, J, counts, n, cores)
            t, frags, kind = 0.0, 0, "port"
            for _ in range(args.steps):
                f, dt, kind = reference_step(w, J, counts, n, cores)
                frags += f; t += dt
        finally:
            shutil.rmtree(w, ignore_errors=True)
        workload = J["workload"]
        sample = "%d fragments per step (%d per core x %d cores), %s composition without gzip hops, files on %s" % (
            frags // max(1, args.steps), per_core, cores, "fragSim;deamSim;adptSim" if kind == "reference" else "oracle port", os.path.dirname(w))
    v = frags / t
    line = {"impl": "reference", "metric": "simulated fragments/sec", "value": v, "unit": "fragments/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "fragments_per_step": frags // max(1, args.steps)},
            "cpu_baseline": {"value": v, "unit": "fragments/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- the reference's driver, for real (C2, N=1)
def via_gargammel_pl(J, n, bins, label, fused=False):
    """_via_gargammel_pl, never raising: a leg that cannot run (no space on tmpfs, ...) reports its error instead of a number."""
    try:
        return _via_gargammel_pl(J, n, bins, label, fused)
    except Exception as ex:
        return {"value": None, "error": repr(ex)[-300:]}


def _via_gargammel_pl(J, n, bins, label, fused=False):
    """perl gargammel.pl (the unmodified script, oracle/_ref) over `bins` = {fragSim,deamSim,adptSim} on the job's directory, n fragments:
    what a user of the reference actually runs (process launches, gzip hops and, for our binaries, one CUDA context per process included).
    art_illumina is replaced by a no-op stub writing empty FASTQ files.
    fused=True: the same job, same directory, same options and the same art stub through bin/gargammel-b200 (one process, the
    intermediates compressed on the device) -- what a user who switches drivers runs instead of the script."""
    pl = os.path.join(ROOT, "oracle", "_ref", "gargammel.pl")
    if fused:
        if not os.path.exists(os.path.join(bins, "gargammel-b200")):
            return None
    elif not os.path.exists(pl) or shutil.which("perl") is None or not all(os.path.exists(os.path.join(bins, b)) for b in ("fragSim", "deamSim", "adptSim")):
        return None
    from tests import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    t = tempfile.mkdtemp(prefix="gg_pl_", dir=base)
    try:
        gg = os.path.join(t, "gg"); os.makedirs(os.path.join(gg, "src")); os.makedirs(os.path.join(gg, "art_src_MountRainier"))
        if not fused:
            shutil.copy(pl, os.path.join(gg, "gargammel.pl"))
            for b in ("fragSim", "deamSim", "adptSim"):
                os.symlink(os.path.join(bins, b), os.path.join(gg, "src", b))
        art = os.path.join(gg, "art_src_MountRainier", "art_illumina")
        with open(art, "w") as fh:
            fh.write("#!/bin/sh\nwhile [ $# -gt 0 ]; do if [ \"$1\" = \"-o\" ]; then O=$2; fi; shift; done\n: > ${O}1.fq; : > ${O}2.fq\n")
        os.chmod(art, 0o755)
        d = os.path.join(t, "in")
        for sub in ("endo", "cont", "bact"):
            os.makedirs(os.path.join(d, sub))
        for f in J["files"]:
            sub = {"e": "endo", "c": "cont", "b": "bact"}[f[0]]
            with open(os.path.join(d, sub, f[1]), "wb") as fh:
                fh.write(f[2])
            synth.write_fai(os.path.join(d, sub, f[1] + ".fai"), f[3])
        synth.write_sizefreq(os.path.join(t, "sf.size"))
        s5, s3 = synth.golden_matrix("single-")
        synth.write_matrix_dat(os.path.join(t, "m-"), s5, s3)
        os.makedirs(os.path.join(t, "out"))
        opts = ["--comp", "%g,%g,%g" % J["comp"], "-n", str(n), "-f", os.path.join(t, "sf.size"), "-matfile", os.path.join(t, "m-"),
                "-rl", str(J["read_len"]), "-o", os.path.join(t, "out", "sim")]
        cmd = ([os.path.join(bins, "gargammel-b200"), "--keep", "--art", art] if fused else ["perl", os.path.join(gg, "gargammel.pl")]) + opts + [d + "/"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            return {"value": None, "error": r.stderr[-300:]}
        out_files = sorted(os.path.basename(f) for f in os.listdir(os.path.join(t, "out")))
        if fused:
            return {"value": n / dt, "unit": "fragments/s", "seconds": dt, "fragments": n, "processes": 1, "files": out_files, "binaries": label}
        nproc = sum(1 for l in r.stderr.splitlines() if l.startswith("running cmd ") and "/src/" in l)
        return {"value": n / dt, "unit": "fragments/s", "seconds": dt, "fragments": n, "filter_processes": nproc, "files": out_files, "binaries": label}
    finally:
        shutil.rmtree(t, ignore_errors=True)


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the batch (and the C3/C4 genomes) by this factor; the line says so")
    ap.add_argument("--frags", type=int, default=0, help="fragments per step per GPU (overrides the configuration's)")
    ap.add_argument("--chunk", type=int, default=1 << 24, help="fragments per kernel launch of the device-timed path (slab ring in HBM)")
    ap.add_argument("--c5-sizes", default="l40", choices=["l40", "sizefreq", "sizedist"])
    ap.add_argument("--c5-emit", default="artp", choices=["artp", "frag"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-perl", action="store_true", help="skip cpu_baseline.via_gargammel_pl")
    ap.add_argument("--perl-10x", action="store_true", help="also time gargammel.pl over bin/ at ten times the fragments (about a minute and a half)")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 20, help="fragments per chunk of the plain e2e path")
    ap.add_argument("--bgzf-chunk", type=int, default=1 << 19, help="fragments per chunk of the device-BGZF e2e pipeline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from gargammel_b200 import api, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    placement = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "C1":
        return run_c1(args, rank, world, local, barrier, placement)

    t_setup = time.perf_counter()
    J = make_job(args.config, args, world)
    n = J["n"]
    emit = api.EMIT_ARTP if J["emit"] == "artp" else api.EMIT_FRAG
    ctx = api.Context(local, SEED)
    pinned_src = []
    for f in J["files"]:
        t = torch.empty(len(f[2]), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = np.frombuffer(f[2], dtype=np.uint8)
        pinned_src.append((t, f[3]))
    ids = [ctx.upload_genome(t, fai) for t, fai in pinned_src]
    setup_tables(api, ctx, J)
    plan, n_total, counts = job_plan(api, J, ids, n * world)
    ctx.set_plan(plan)
    first, last = shard.rank_slice(n_total, rank, world)           # rank r takes the r-th contiguous slice of the fragment ids
    count = last - first
    log("[bench] %s rank %d/%d setup %.1fs; %d files, %d plan ranges; ids [%d,%d)" % (J["name"], rank, world, time.perf_counter() - t_setup, len(J["files"]), len(plan), first, last))

    def run_job(fetch_none=True):
        """one step: the rank's ids in chunks of --chunk through the device slab; returns (device ms, last info, bytes, gather bytes, attempts, launches)"""
        ms = 0.0; nb = ng = na = nl = 0; info = None
        for a in range(first, last, args.chunk):
            _, info = ctx.run_fused(a, min(args.chunk, last - a), emit, fetch=False)
            ms += info.device_ms; nb += info.bytes; ng += info.gather_bytes; na += info.attempts; nl += info.launches
        return ms, info, nb, ng, na, nl

    # ---- device-timed: inputs (genomes, tables) resident in HBM; L2 flushed between steps (untimed)
    for _ in range(args.warmup if count <= (1 << 26) else 1):
        ctx.flush_l2()
        run_job()
    barrier()
    clocks = Clocks(local)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        ms, info, out_bytes, gather_bytes, attempts, launches = run_job()
        dev_ms += ms
    barrier()
    t1 = time.perf_counter()
    clk = clocks.stop(t0, t1)
    dev_ms_max = shard.max_over_ranks(dev_ms, device="cuda")
    frags_all = float(shard.reduce_counts([count * args.steps], device="cuda")[0])      # the final count reduction (no data-path collective)
    value = frags_all / (dev_ms_max / 1000.0)

    # ---- end to end through the C-ABI with HOST buffers: every step uploads the FASTA files (pinned host -> HBM, pack), runs the fused
    # kernel and drains the FASTA byte stream to pinned host memory (a ring of `sup` fragments: a 1 B-fragment stream does not fit host RAM)
    e2e = None
    if not args.no_e2e:
        rec_max = int(ctx.max_record_bytes(emit))
        sup = min(count, 8 << 20)                                   # fragments per gg_run_fused_to_host call
        cap = rec_max * sup + 4096
        host_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        h2d = sum(int(t.numel()) for t, _ in pinned_src)
        e_steps = max(1, min(args.steps, 5)) if count <= (1 << 25) else 1

        def upload_inputs():
            ctx.genome_clear()
            gids = [ctx.upload_genome(t, fai) for t, fai in pinned_src]
            assert gids == ids                                      # ids restart after genome_clear: the plan made at setup is still the plan
            ctx.set_plan(plan)

        def e2e_step():
            upload_inputs()
            nb = 0
            for a in range(first, last, sup):
                nb += ctx.run_fused_to_host(a, min(sup, last - a), emit, host_out.data_ptr(), cap, args.e2e_chunk).bytes
            return nb
        e2e_step()
        barrier()
        te0 = time.perf_counter()
        for _ in range(e_steps):
            d2h = e2e_step()
        barrier()
        te = shard.max_over_ranks(time.perf_counter() - te0, device="cuda")
        e2e = {"value": count * world * e_steps / te, "unit": "fragments/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(d2h), "steps": e_steps, "chunk_fragments": args.e2e_chunk, "pinned_buffers": placement,
               "what": "per step: gg_genome_upload of every FASTA file from pinned host + gg_run_fused_to_host (fused kernel per chunk, D2H of chunk k overlapping kernel k+1) into a pinned ring; PCIe-bound: the D2H of the FASTA stream"}
        # ---- the ceiling of that path: the same bytes moved by plain cudaMemcpyAsync (pinned host <-> device), all ranks at once
        try:
            dev_buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
            pieces = [min(cap - 4096, d2h - o) for o in range(0, d2h, cap - 4096)]

            def probe():
                for t, _ in pinned_src:
                    dev_buf[:min(cap, t.numel())].copy_(t[:min(cap, t.numel())], non_blocking=True)     # the step's H2D bytes (first `cap` of each file: files larger than the ring are a second-order term)
                for p in pieces:
                    host_out[:p].copy_(dev_buf[:p], non_blocking=True)
                torch.cuda.synchronize()
            probe()
            barrier()
            tp0 = time.perf_counter()
            for _ in range(e_steps):
                probe()
            barrier()
            tp = shard.max_over_ranks(time.perf_counter() - tp0, device="cuda")
            moved = float(d2h + min(h2d, cap * len(pinned_src)))
            e2e["roofline"] = {"achieved_GBps": moved * world * e_steps / te / 1e9, "pcie_ceiling_GBps": moved * world * e_steps / tp / 1e9,
                               "frac": tp / te, "what": "ceiling = the step's H2D + D2H bytes by plain cudaMemcpyAsync from/to the same pinned buffers, %d rank(s) concurrently" % world}
            del dev_buf
        except Exception as ex:
            e2e["roofline"] = {"error": repr(ex)}
        # ---- the same with host BGZF of the byte stream on all host threads (north_star: with and without gzip/BGZF output); C2-sized jobs only
        if count <= (1 << 24):
            try:
                import ctypes as C
                lib = api.load_library()
                zcap = int(d2h) + int(d2h) // 200 + (1 << 20)
                zbuf = torch.empty(zcap, dtype=torch.uint8)
                zn = C.c_size_t(0)
                barrier()
                tz0 = time.perf_counter()
                upload_inputs()
                rcz, zbytes, pbytes = 0, 0, 0
                for a in range(first, last, sup):                           # every piece of the ring is compressed before the next one overwrites it
                    nbp = ctx.run_fused_to_host(a, min(sup, last - a), emit, host_out.data_ptr(), cap, args.e2e_chunk).bytes
                    rcz |= lib.ggh_bgzf_compress(C.c_void_p(host_out.data_ptr()), C.c_size_t(int(nbp)), 0, 1, 1, C.c_void_p(zbuf.data_ptr()), C.c_size_t(zcap), C.byref(zn))
                    zbytes += int(zn.value); pbytes += int(nbp)
                barrier()
                tz = shard.max_over_ranks(time.perf_counter() - tz0, device="cuda")
                if rcz == 0:
                    e2e["with_bgzf"] = {"value": count * world / tz, "unit": "fragments/s", "host_threads": host_cores(), "level": 1, "compressed_bytes_per_step": zbytes,
                                        "bits_per_byte": 8.0 * zbytes / max(pbytes, 1), "steps": 1}
            except Exception as ex:
                e2e["with_bgzf"] = {"value": None, "error": repr(ex)}
        # ---- ... and with the BGZF members produced ON THE DEVICE (gg_bgzf_dev), so that only the compressed members cross PCIe.
        # One piece of one step is inflated on the host afterwards and compared with the plain stream.  Both device encoders are measured:
        # the default one with LZ77 matches (k_bgzf_lz: files of zlib -1's size) and the Huffman-only one (GG_BGZF_MODE=huff: k_bgzf_deflate,
        # a fifth of the kernel time, files 30-40 % larger).
        zhost = None
        for key, mode in (("with_device_bgzf", None), ("with_device_bgzf_huffman", "huff")):
            try:
                if mode:
                    os.environ["GG_BGZF_MODE"] = mode
                zcap = cap
                if zhost is None:
                    zhost = torch.empty(zcap, dtype=torch.uint8, pin_memory=True)

                zms = [0.0]

                def e2e_step_z():
                    upload_inputs()
                    zb = ib = 0; zi = None; zms[0] = 0.0
                    for a in range(first, last, sup):
                        zi = ctx.run_fused_bgzf_to_host(a, min(sup, last - a), emit, zhost.data_ptr(), zcap, args.bgzf_chunk)
                        zb += zi.bytes; ib += zi.in_bytes; zms[0] += zi.device_ms
                    return zb, ib, zi
                e2e_step_z()
                barrier()
                td0 = time.perf_counter()
                for _ in range(e_steps):
                    zb, ib, zi = e2e_step_z()
                barrier()
                td = shard.max_over_ranks(time.perf_counter() - td0, device="cuda")
                e2e[key] = {"value": count * world * e_steps / td, "unit": "fragments/s", "steps": e_steps, "d2h_bytes_per_step": int(zb),
                            "uncompressed_bytes_per_step": int(ib), "bits_per_byte": 8.0 * zb / max(ib, 1), "chunk_fragments": args.bgzf_chunk,
                            "kernel_ms_per_step": zms[0], "wall_ms_per_step": 1000.0 * td / e_steps,
                            "encoder": "k_bgzf_deflate (Huffman only, GG_BGZF_MODE=huff)" if mode else "k_bgzf_lz (LZ77 matches + Huffman; the default)",
                            "what": "per step: genome uploads + gg_run_fused_bgzf_to_host (fused kernel + device BGZF per chunk, D2H of the members overlapped with the next chunk)"}
                if rank == 0:
                    import zlib
                    a_last = first + ((count - 1) // sup) * sup                 # the piece both buffers hold at this point
                    ctx.run_fused_to_host(a_last, min(sup, last - a_last), emit, host_out.data_ptr(), cap, args.e2e_chunk)
                    zbv = zhost[:int(zi.bytes)].numpy(); plain = host_out.numpy()
                    off, pos, okz = 0, 0, True
                    while off < len(zbv) and pos < (32 << 20):                  # the first 32 MB of members (zlib verifies CRC-32 and ISIZE of each)
                        bsize = int(zbv[off + 16]) + (int(zbv[off + 17]) << 8) + 1
                        part = zlib.decompress(zbv[off:off + bsize].tobytes(), wbits=31)
                        okz = okz and part == plain[pos:pos + len(part)].tobytes()
                        off += bsize; pos += len(part)
                    e2e[key]["verified_bytes"] = pos
                    assert okz, "device BGZF does not inflate to the plain stream"
            except Exception as ex:
                e2e[key] = {"value": None, "error": repr(ex)}
            finally:
                os.environ.pop("GG_BGZF_MODE", None)
        head = bytes(host_out[:64].numpy().tobytes())
        assert head.startswith(b">"), head

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg_bytes = float(out_bytes + gather_bytes)                     # per step: emitted FASTA bytes + 2-bit/mask gather bytes
    k_ms = dev_ms / args.steps
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
        if int(tj.get("fragments_per_launch", 0)) == count and J["name"] == "C2":
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "simulated fragments/sec", "value": value, "unit": "fragments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": J["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": J["workload"], "fragments_per_step_per_gpu": count, "genome_bp_per_file": J["genome_bp"], "files": len(J["files"]), "plan_ranges": len(plan),
                   "read_len": J["read_len"], "seed": SEED, "chunk_fragments": args.chunk,
                   "l2": "flushed between steps (256 MiB write, untimed); output per step (%.2f GB) exceeds L2" % (out_bytes / 1e9),
                   "timing": "CUDA events on the launching stream around each launch (memsets + k_fused + k_publish = the byte count reaching the host through mapped memory), summed over the step's chunks and the steps, max over ranks",
                   "attempts_per_fragment": attempts / max(1, count), "bytes_per_fragment": alg_bytes / max(1, count)},
        "wall_ms_per_step": 1000.0 * (t1 - t0) / args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "k_fused", "algorithmic_bytes_per_launch": alg_bytes / max(1, (count + args.chunk - 1) // args.chunk),
                     "frac_of_8TBps_spec": achieved / 8000.0},
        "gpu_launches": args.steps * launches,
        "clocks": clk,
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = host_cores()
            w = prepare_reference_dir(J)
            try:
                per_core = 150000
                reference_step(w, J, counts, 2000 * cores, cores)
                f, dt, kind = reference_step(w, J, counts, per_core * cores, cores)
                line["cpu_baseline"] = {"value": f / dt, "unit": "fragments/s", "cores": cores, "kind": kind,
                                        "sample": "%d fragments (%d per core x %d cores) of the same workload, %s, no gzip hops, files on tmpfs" % (
                                            f, per_core, cores, "reference sources + our shim (oracle/_ref): fragSim;deamSim;adptSim per shard" if kind == "reference" else "oracle port")}
                if kind == "reference" and shutil.which("gzip"):           # the same with the driver's gzip hops, as gargammel.pl runs it
                    fz, dtz, _ = reference_step(w, J, counts, 50000 * cores, cores, gzip_hops=True)
                    line["cpu_baseline"]["with_gzip_hops"] = {"value": fz / dtz, "unit": "fragments/s", "sample": "%d fragments (50000 per core x %d cores), `| gzip >` after fragSim and deamSim, gz inputs to deamSim/adptSim (gargammel.pl:1510,1729,1847)" % (fz, cores)}
            finally:
                shutil.rmtree(w, ignore_errors=True)
            if J["name"] == "C2" and not args.no_perl:                      # the reference's driver for real, over its own filters and over ours
                ctx.close()
                line["cpu_baseline"]["via_gargammel_pl"] = {
                    "reference_binaries": via_gargammel_pl(J, 200000, os.path.join(ROOT, "oracle", "_ref", "src"), "oracle/_ref/src (reference sources + our shim), single-threaded filters"),
                    "dropin_binaries": via_gargammel_pl(J, 200000, os.path.join(ROOT, "bin"), "bin/ (this repository): one CUDA context (about a second) per filter process"),
                    "fused_driver": via_gargammel_pl(J, 200000, os.path.join(ROOT, "bin"), "bin/gargammel-b200 (this repository): one process, one CUDA context, .gz intermediates made on the device", fused=True),
                    "fused_driver_10x": via_gargammel_pl(J, 2000000, os.path.join(ROOT, "bin"), "bin/gargammel-b200, ten times the fragments", fused=True),
                    "fused_driver_50x": via_gargammel_pl(J, 10000000, os.path.join(ROOT, "bin"), "bin/gargammel-b200, fifty times the fragments", fused=True),
                    "what": "perl gargammel.pl --comp 0,0.08,0.92 -n 200000 -f sizefreq -matfile single- -rl 75 (unmodified script, art_illumina stubbed, the script's own gzip hops included); "
                            "fused_driver*: the same directory, options and art stub through bin/gargammel-b200 --keep, which replaces the script's middle with one process "
                            "(the script's file set: intermediates .{e,b,c}.fa.gz and _d.fa.gz, _a.fa.gz, _s1/_s2.fq.gz; wall clock from exec to exit, genome loading "
                            "and all output files on tmpfs included)"}
                if args.perl_10x:       # 94 s on the round-2 box (21 k fragments/s: the script's gzip -6 hops and Perl loops, not the filters, set the pace)
                    line["cpu_baseline"]["via_gargammel_pl"]["dropin_binaries_10x"] = via_gargammel_pl(J, 2000000, os.path.join(ROOT, "bin"), "bin/ (this repository), ten times the fragments: the per-process cost is fixed")
        except Exception as ex:      # the baseline is a reported figure; never let it kill the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "fragments/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (ex,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c1(args, rank, world, local, barrier, placement):
    """C1: the stand-alone deamSim stage over a FASTA record stream (BASELINE configs[0]); every rank processes the same-shaped stream."""
    import torch
    import torch.distributed as dist
    from gargammel_b200 import api, shard
    from tests import synth
    n = args.frags or int(1_000_000 * args.scale)
    recs = b"".join(c1_records(n))
    s5, s3 = synth.golden_matrix("single-")
    ctx = api.Context(local, SEED)
    ctx.set_matrix(0, s5, s3)
    for _ in range(args.warmup):
        ctx.flush_l2(); out, info = ctx.run_deam(recs, 0)
    barrier()
    clocks = Clocks(local)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        out, info = ctx.run_deam(recs, 0)
        dev_ms += info.device_ms
    barrier()
    t1 = time.perf_counter()
    clk = clocks.stop(t0, t1)
    dev_ms_max = shard.max_over_ranks(dev_ms, device="cuda")
    value = n * world * args.steps / (dev_ms_max / 1000.0)
    # e2e = the C-ABI calls a host program makes -- gg_run_deam (record stream in pinned host memory -> device -> stage kernels) +
    # gg_out_fetch (-> pinned host memory) -- timed by the wall clock around them; the Python wrapper above (bytes objects, a fresh ctypes
    # buffer per call) is left out of it
    import ctypes as C
    lib = api.load_library()
    h_in = torch.empty(len(recs), dtype=torch.uint8, pin_memory=True)
    h_in.numpy()[:] = np.frombuffer(recs, dtype=np.uint8)
    h_out = torch.empty(len(recs) + 4096, dtype=torch.uint8, pin_memory=True)
    e_steps = max(args.steps, 5)

    def e2e_step():
        o = api._Out(); got = C.c_size_t(0)
        ctx._ck(lib.gg_run_deam(ctx._h, C.c_void_p(h_in.data_ptr()), C.c_size_t(len(recs)), C.c_int(0), C.c_uint64(0), C.byref(o)))
        ctx._ck(lib.gg_out_fetch(ctx._h, C.byref(o), C.c_void_p(h_out.data_ptr()), C.c_size_t(h_out.numel()), C.byref(got)))
        return int(got.value)
    e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(e_steps):
        got = e2e_step()
    barrier()
    te = shard.max_over_ranks(time.perf_counter() - te0, device="cuda") * args.steps / e_steps
    assert h_out[:got].numpy().tobytes() == out, "C1 e2e: the C-ABI path and the wrapper disagree"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    from oracle import oracle
    o = oracle.OracleContext(SEED); o.set_matrix(0, s5, s3)
    head = b"".join(c1_records(20000))
    assert out[:len(head)] == o.run_deam(head, 0)[0], "C1: CUDA output differs from the oracle"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = 2.0 * len(recs)
    achieved = alg / (dev_ms / args.steps / 1000.0) / 1e9
    line = {"metric": "simulated fragments/sec", "value": value, "unit": "fragments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C1: deamSim -matfile single- on %d fixed 50-bp fragments (stand-alone stage over a FASTA record stream)" % n, "records_per_step_per_gpu": n,
                       "l2": "flushed between steps", "timing": "CUDA events around the stage's kernels (the H2D of the record stream is enqueued before the first event)"},
            "wall_ms_per_step": 1000.0 * (t1 - t0) / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "kernel": "stand-alone deamSim kernels",
                         "algorithmic_bytes_per_launch": alg, "what": "record stream read + written"},
            "gpu_launches": args.steps * info.launches, "clocks": clk,
            "e2e": {"value": n * world * args.steps / te, "unit": "fragments/s", "h2d_bytes_per_step": len(recs), "d2h_bytes_per_step": len(out), "pinned_buffers": placement,
                    "what": "gg_run_deam + gg_out_fetch through the C-ABI: record stream in pinned host memory in, pinned host memory out (wall clock)"}}
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        r = c1_reference(n, cores)
        if r:
            line["cpu_baseline"] = {"value": r[0] / r[1], "unit": "fragments/s", "cores": cores, "kind": "reference", "sample": "the whole configuration: %d records, `deamSim -matfile` of the reference, one shard per core" % r[0]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
