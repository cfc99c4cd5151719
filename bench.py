#!/usr/bin/env python3
"""bench.py -- headline benchmark of the gargammel hot path on B200.

Metric (BASELINE.json): simulated fragments/sec, device-timed, plus % of the HBM roofline.
Workload at N=1 = BASELINE.json configs[1] (C2): full fragSim->deamSim->adptSim on a synthetic
endo(x2, diploid)/cont/bact directory of 50 Mb files, --comp 0,0.08,0.92, -f sizefreq, -matfile single-,
2x75 bp paired (-artp), -n 10,000,000 fragments per step, seed 42.  One "step" = one pass of the fused
kernel over that batch.  N>1: one process per GPU (torchrun), rank r takes fragment ids
[r*n, (r+1)*n) of an N*n-fragment job (weak scaling, genome replicated, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]         our arm
  python bench.py --impl reference ...                        the reference's own CPU filters (oracle/_ref: its
                                                              sources + our shim) on the box's host cores

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

FRAGS_PER_STEP = 10_000_000
GENOME_BP = 50_000_000
COMP = (0.0, 0.08, 0.92)            # --comp B,C,E
READ_LEN = 75
SEED = 42
WORKLOAD = "C2: fused fragSim->deamSim->adptSim, synthetic 50 Mb endo(x2)/cont/bact, --comp 0,0.08,0.92, -f sizefreq, -matfile single-, 2x75 -artp"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- synthetic C2 directory
def make_sources(genome_bp=GENOME_BP):
    """endo.1, endo.2 (= endo.1 with 0.1% SNPs), cont.1; iid bases, 0.5% of positions in N runs, 60-col lines."""
    from tests import synth
    rng = np.random.default_rng(20251019)
    n_contigs = 5
    edges = np.linspace(0, genome_bp, n_contigs + 1).astype(np.int64)
    endo1 = [("chrE%d" % (i + 1), synth.random_contig(rng, int(edges[i + 1] - edges[i]))) for i in range(n_contigs)]
    endo2 = []
    for name, seq in endo1:
        s = seq.copy()
        k = max(1, len(s) // 1000)
        pos = rng.integers(0, len(s), size=k)
        ok = s[pos] != ord("N")
        s[pos[ok]] = synth._BASES[rng.integers(0, 4, size=int(ok.sum()))]
        endo2.append((name, s))
    cont = [("chrC%d" % (i + 1), synth.random_contig(rng, int(edges[i + 1] - edges[i]))) for i in range(n_contigs)]
    out = []
    for contigs in (endo1, endo2, cont):
        out.append(synth.fasta_bytes(contigs, 60))
    return out      # [(fasta_bytes, fai)] for endo.1, endo.2, cont.1


def make_plan(api, n_total, ids, seed=SEED):
    """gargammel.pl:1256-1283,1421-1431 through the host library; output order e1,e2,(b),c (gargammel.pl:1698-1830)."""
    import ctypes as C
    lib = api.load_library()
    tot = (C.c_uint64 * 3)(); e = (C.c_uint64 * 2)(); c = (C.c_uint64 * 1)(); b = (C.c_uint64 * 1)()
    err = C.create_string_buffer(512)
    rc = lib.ggh_apportion(C.c_uint64(n_total), C.c_double(COMP[0]), C.c_double(COMP[1]), C.c_double(COMP[2]), 2,
                           (C.c_uint64 * 1)(GENOME_BP), 1, (C.c_double * 1)(1.0), 0, C.c_uint64(seed), tot, e, c, b, err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    plan, first = [], 0
    # -matfile deaminates endogenous (and bacterial) fragments only; the contaminant is passed through (gargammel.pl:640-648)
    for cnt, g, slot, tag in ((e[0], ids[0], 0, "e1_0"), (e[1], ids[1], 0, "e2_0"), (c[0], ids[2], -1, "c1")):
        if cnt:
            plan.append(api.Range(first, int(cnt), g, slot, tag))
            first += int(cnt)
    return plan, first, (int(e[0]), int(e[1]), int(c[0]))


def setup_tables(api, ctx):
    from tests import synth
    lens, cum = synth.golden_sizefreq()
    ctx.set_sizes(api.SIZE_FREQ, lens, cum, minlen=0, maxlen=1000)       # gargammel.pl passes -m 0 -M 1000
    s5, s3 = synth.golden_matrix("single-")
    ctx.set_matrix(0, s5, s3, clamp_last=True)
    ctx.set_adapters(api.DEFAULT_ADAPTER_F, api.DEFAULT_ADAPTER_S, READ_LEN)


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


# --------------------------------------------------------------------------- reference arm (CPU)
def reference_step(workdir, n_frag, cores, gzip_hops=False):
    """One bounded sample: `cores` independent shards, each the reference composition of gargammel.pl:1457-1848 without
    the gzip hops:  fragSim(e1) ; fragSim(e2) >> P.e.fa ; fragSim(c1) > P.c.fa ; deamSim -matfile P.e.fa > P_d.fa ;
    cat P.c.fa >> P_d.fa ; adptSim -artp P_a.fa P_d.fa.   gzip_hops=True inserts the driver's own `| gzip >` / gz-input hops
    (gargammel.pl:1510,1729,1847).  Returns (fragments, seconds, kind)."""
    from oracle import oracle
    kind = "reference" if oracle.ref_bin("fragSim") else "port"
    per = max(1, n_frag // cores)
    nE = int(per * COMP[2]); nC = int(per * COMP[1])
    nE1 = nE // 2; nE2 = nE - nE1
    if kind == "reference":
        frag, deam, adpt = (oracle.ref_bin(x) for x in ("fragSim", "deamSim", "adptSim"))
        sh = ("set -e; cd {w}; S=$1; "
              "{f} -tag e1_0 -n {a} -m 0 -M 1000 -f sf.size endo.1.fa > P$S.e.fa 2>/dev/null; "
              "{f} -tag e2_0 -n {b} -m 0 -M 1000 -f sf.size endo.2.fa >> P$S.e.fa 2>/dev/null; "
              "{f} -tag c1 -n {c} -m 0 -M 1000 -f sf.size cont.1.fa > P$S.c.fa 2>/dev/null; "
              "{d} -matfile m- P$S.e.fa > P${{S}}_d.fa 2>/dev/null; cat P$S.c.fa >> P${{S}}_d.fa; "
              "{p} -f {af} -s {as_} -l {rl} -artp P${{S}}_a.fa P${{S}}_d.fa 2>/dev/null; rm -f P$S.e.fa P$S.c.fa P${{S}}_d.fa P${{S}}_a.fa").format(
            w=workdir, f=frag, d=deam, p=adpt, a=nE1, b=nE2, c=nC, rl=READ_LEN,
            af="AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG", as_="AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT")
        if gzip_hops:
            sh = ("set -e; cd {w}; S=$1; "
                  "( {f} -tag e1_0 -n {a} -m 0 -M 1000 -f sf.size endo.1.fa 2>/dev/null; {f} -tag e2_0 -n {b} -m 0 -M 1000 -f sf.size endo.2.fa 2>/dev/null ) | gzip > P$S.e.fa.gz; "
                  "{f} -tag c1 -n {c} -m 0 -M 1000 -f sf.size cont.1.fa 2>/dev/null | gzip > P$S.c.fa.gz; "
                  "{d} -matfile m- P$S.e.fa.gz 2>/dev/null | gzip > P${{S}}_d.fa.gz; gzip -dc P$S.c.fa.gz | gzip >> P${{S}}_d.fa.gz; "
                  "{p} -f {af} -s {as_} -l {rl} -artp P${{S}}_a.fa P${{S}}_d.fa.gz 2>/dev/null; rm -f P$S.e.fa.gz P$S.c.fa.gz P${{S}}_d.fa.gz P${{S}}_a.fa").format(
                w=workdir, f=frag, d=deam, p=adpt, a=nE1, b=nE2, c=nC, rl=READ_LEN,
                af="AGATCGGAAGAGCACACGTCTGAACTCCAGTCACCGATTCGATCTCGTATGCCGTCTTCTGCTTG", as_="AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATTT")
        t0 = time.perf_counter()
        procs = [subprocess.Popen(["bash", "-c", sh, "shard", str(i)]) for i in range(cores)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference shard failed: %r" % rcs)
        return (nE + nC) * cores, dt, kind
    # port: the single-threaded oracle, one process per core
    code = ("import sys; sys.path.insert(0, %r); import bench; bench.oracle_shard(%r, int(sys.argv[1]), %d, %d, %d)" % (ROOT, workdir, nE1, nE2, nC))
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", code, str(i)]) for i in range(cores)]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("oracle shard failed: %r" % rcs)
    return (nE + nC) * cores, dt, kind


def oracle_shard(workdir, shard, nE1, nE2, nC):
    from gargammel_b200 import api
    from tests import synth
    from oracle import oracle
    ctx = oracle.OracleContext(SEED + shard)
    ids = []
    for name in ("endo.1.fa", "endo.2.fa", "cont.1.fa"):
        fa = open(os.path.join(workdir, name), "rb").read()
        fai = [api.FaiEntry(a, int(b), int(c), int(d), int(e)) for a, b, c, d, e in (l.split("\t") for l in open(os.path.join(workdir, name + ".fai")).read().splitlines())]
        ids.append(ctx.upload_genome(fa, fai))
    setup_tables(api, ctx)
    plan = [api.Range(0, nE1, ids[0], 0, "e1_0"), api.Range(nE1, nE2, ids[1], 0, "e2_0"), api.Range(nE1 + nE2, nC, ids[2], -1, "c1")]
    ctx.set_plan(plan)
    ctx.run_fused(0, nE1 + nE2 + nC, api.EMIT_ARTP)


def prepare_reference_dir(sources):
    from tests import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    w = tempfile.mkdtemp(prefix="gg_ref_", dir=base)
    for (fa, fai), name in zip(sources, ("endo.1.fa", "endo.2.fa", "cont.1.fa")):
        with open(os.path.join(w, name), "wb") as f:
            f.write(fa)
        synth.write_fai(os.path.join(w, name + ".fai"), fai)
    synth.write_sizefreq(os.path.join(w, "sf.size"))
    s5, s3 = synth.golden_matrix("single-")
    synth.write_matrix_dat(os.path.join(w, "m-"), s5, s3)
    return w


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    cores = host_cores()
    sources = make_sources()
    w = prepare_reference_dir(sources)
    try:
        per_core = 150000          # ~1 s per core per step: 15-25 core-seconds of CPU work per step
        n = per_core * cores
        for _ in range(args.warmup):
            reference_step(w, n, cores)
        t, frags, kind = 0.0, 0, "port"
        for _ in range(args.steps):
            f, dt, kind = reference_step(w, n, cores)
            frags += f; t += dt
    finally:
        shutil.rmtree(w, ignore_errors=True)
    v = frags / t
    sample = "%d fragments per step (%d per core x %d cores), %s composition without gzip hops, files on %s" % (
        frags // max(1, args.steps), per_core, cores, "fragSim;deamSim;adptSim" if kind == "reference" else "oracle port", os.path.dirname(w))
    line = {"impl": "reference", "metric": "simulated fragments/sec", "value": v, "unit": "fragments/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fragments_per_step": frags // max(1, args.steps)},
            "cpu_baseline": {"value": v, "unit": "fragments/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frags", type=int, default=FRAGS_PER_STEP, help="fragments per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 20, help="fragments per chunk of the plain e2e path (0 = one gg_run_fused + one gg_out_fetch)")
    ap.add_argument("--bgzf-chunk", type=int, default=1 << 19, help="fragments per chunk of the device-BGZF e2e pipeline (0 = one piece)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from gargammel_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.frags
    t_setup = time.perf_counter()
    sources = make_sources()
    ctx = api.Context(local, SEED)
    pinned_src = []
    for fa, fai in sources:
        t = torch.empty(len(fa), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = np.frombuffer(fa, dtype=np.uint8)
        pinned_src.append((t, fai))
    ids = [ctx.upload_genome(t, fai) for t, fai in pinned_src]
    setup_tables(api, ctx)
    plan, n_total, split = make_plan(api, n * world, ids)
    ctx.set_plan(plan)
    first = rank * n
    count = min(n, n_total - first)
    log("[bench] rank %d/%d setup %.1fs; plan e1=%d e2=%d c1=%d; ids [%d,%d)" % (rank, world, time.perf_counter() - t_setup, split[0], split[1], split[2], first, first + count))

    # ---- device-timed: inputs (genomes, tables) resident in HBM; L2 flushed between steps (untimed)
    infos = []
    for _ in range(args.warmup):
        ctx.flush_l2()
        ctx.run_fused(first, count, api.EMIT_ARTP, fetch=False)
    barrier()
    clocks = Clocks(local)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()
        _, info = ctx.run_fused(first, count, api.EMIT_ARTP, fetch=False)
        infos.append(info)
        dev_ms += info.device_ms
    barrier()
    t1 = time.perf_counter()
    clk = clocks.stop(t0, t1)
    info = infos[-1]
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    frags = torch.tensor([float(count * args.steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(frags, op=dist.ReduceOp.SUM)       # the final count reduction (no data-path collective)
    dev_ms_max = float(tmax.item()); frags_all = float(frags.item())
    value = frags_all / (dev_ms_max / 1000.0)

    # ---- end to end through the C-ABI with HOST buffers: every step uploads the FASTA files (pinned host -> HBM, pack),
    # runs the fused kernel and copies the FASTA byte stream back to pinned host memory
    e2e = None
    if not args.no_e2e:
        cap = int(ctx.max_record_bytes(api.EMIT_ARTP)) * count + 4096
        host_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        h2d = sum(int(t.numel()) for t, _ in pinned_src)
        e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            ctx.genome_clear()
            gids = [ctx.upload_genome(t, fai) for t, fai in pinned_src]
            p2, _, _ = make_plan(api, n * world, gids)
            ctx.set_plan(p2)
            if args.e2e_chunk > 0:
                return ctx.run_fused_to_host(first, count, api.EMIT_ARTP, host_out.data_ptr(), cap, args.e2e_chunk)
            return ctx.run_fused_into(first, count, api.EMIT_ARTP, host_out.data_ptr(), cap)
        e2e_step()
        barrier()
        te0 = time.perf_counter()
        for _ in range(e_steps):
            ei = e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - te0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": count * world * e_steps / float(te.item()), "unit": "fragments/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(ei.bytes), "steps": e_steps,
               "chunk_fragments": args.e2e_chunk,
               "what": "per step: gg_genome_upload x3 from pinned host + gg_run_fused_to_host (fused kernel per chunk, D2H of chunk k overlapping kernel k+1) to pinned host; PCIe-bound: the D2H of the FASTA stream"}
        # the same with host BGZF of the byte stream on all host threads (north_star: with and without gzip/BGZF output)
        try:
            import ctypes as C
            lib = api.load_library()
            zcap = int(ei.bytes) + int(ei.bytes) // 200 + (1 << 20)
            zbuf = torch.empty(zcap, dtype=torch.uint8)
            zn = C.c_size_t(0)
            barrier()
            tz0 = time.perf_counter()
            ei = e2e_step()
            rcz = lib.ggh_bgzf_compress(C.c_void_p(host_out.data_ptr()), C.c_size_t(int(ei.bytes)), 0, 1, 1, C.c_void_p(zbuf.data_ptr()), C.c_size_t(zcap), C.byref(zn))
            barrier()
            tz = torch.tensor([time.perf_counter() - tz0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tz, op=dist.ReduceOp.MAX)
            if rcz == 0:
                e2e["with_bgzf"] = {"value": count * world / float(tz.item()), "unit": "fragments/s", "host_threads": host_cores(), "level": 1,
                                    "compressed_bytes_per_step": int(zn.value), "steps": 1}
        except Exception as ex:
            e2e["with_bgzf"] = {"value": None, "error": repr(ex)}
        # ... and with the BGZF members produced ON THE DEVICE (gg_bgzf_dev: Huffman-only deflate + CRC-32 kernels), so that only
        # the compressed members cross PCIe.  One step is inflated on the host afterwards and compared with the plain stream.
        try:
            def e2e_step_z():
                ctx.genome_clear()
                gids = [ctx.upload_genome(t, fai) for t, fai in pinned_src]
                p2, _, _ = make_plan(api, n * world, gids)
                ctx.set_plan(p2)
                if args.bgzf_chunk > 0:
                    zi_ = ctx.run_fused_bgzf_to_host(first, count, api.EMIT_ARTP, zhost.data_ptr(), zcap, args.bgzf_chunk)
                    return zi_, zi_
                return ctx.run_fused_bgzf(first, count, api.EMIT_ARTP, zhost.data_ptr(), zcap)
            zhost = torch.empty(zcap, dtype=torch.uint8, pin_memory=True)
            ctx.run_fused_bgzf(first, count, api.EMIT_ARTP, zhost.data_ptr(), zcap)             # (first call allocates the slabs)
            _, z0 = ctx.run_fused_bgzf(first, count, api.EMIT_ARTP, zhost.data_ptr(), zcap)     # untimed: the compression alone, in one piece
            e2e_step_z()
            barrier()
            td0 = time.perf_counter()
            for _ in range(e_steps):
                ri, zi = e2e_step_z()
            barrier()
            td = torch.tensor([time.perf_counter() - td0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            e2e["with_device_bgzf"] = {"value": count * world * e_steps / float(td.item()), "unit": "fragments/s", "steps": e_steps,
                                       "d2h_bytes_per_step": int(zi.bytes), "uncompressed_bytes_per_step": int(zi.in_bytes),
                                       "bgzf_device_ms": z0.device_ms, "bgzf_GBps_in": z0.in_bytes / (z0.device_ms / 1e3) / 1e9,
                                       "chunk_fragments": args.bgzf_chunk,
                                       "what": "per step: gg_genome_upload x3 + gg_run_fused_bgzf_to_host (fused kernel + device BGZF per chunk, D2H of the members overlapped with the next chunk)"}
            if rank == 0:
                import zlib
                zb = zhost[:int(zi.bytes)].numpy()
                plain = host_out[:int(ei.bytes)].numpy()
                off, pos, okz = 0, 0, True
                while off < len(zb) and pos < (32 << 20):                   # the first 32 MB of members (zlib verifies CRC-32 and ISIZE of each)
                    bsize = int(zb[off + 16]) + (int(zb[off + 17]) << 8) + 1
                    part = zlib.decompress(zb[off:off + bsize].tobytes(), wbits=31)
                    okz = okz and part == plain[pos:pos + len(part)].tobytes()
                    off += bsize; pos += len(part)
                e2e["with_device_bgzf"]["verified_bytes"] = pos
                assert okz, "device BGZF does not inflate to the plain stream"
        except Exception as ex:
            e2e["with_device_bgzf"] = {"value": None, "error": repr(ex)}
        # cheap end-to-end sanity on what reached the host
        head = bytes(host_out[:64].numpy().tobytes())
        assert head.startswith(b">chr"), head

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
    alg_bytes = float(info.bytes + info.gather_bytes)               # per launch: emitted FASTA bytes + 2-bit/mask gather bytes
    k_ms = dev_ms / args.steps
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
        if int(tj.get("fragments_per_launch", 0)) == count:
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "simulated fragments/sec", "value": value, "unit": "fragments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "fragments_per_step_per_gpu": count, "genome_bp_per_file": GENOME_BP, "read_len": READ_LEN, "seed": SEED,
                   "l2": "flushed between steps (256 MiB write, untimed); output per step (%.2f GB) exceeds L2" % (info.bytes / 1e9),
                   "timing": "CUDA events on the launching stream around each step (memsets + k_fused + k_publish = the byte count reaching the host through mapped memory), summed over steps, max over ranks",
                   "attempts_per_fragment": info.attempts / max(1, info.records), "bytes_per_fragment": alg_bytes / max(1, info.records)},
        "wall_ms_per_step": 1000.0 * (t1 - t0) / args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "k_fused", "algorithmic_bytes_per_launch": alg_bytes,
                     "frac_of_8TBps_spec": achieved / 8000.0},
        "gpu_launches": args.steps * info.launches,
        "clocks": clk,
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = host_cores()
            w = prepare_reference_dir(sources)
            try:
                per_core = 150000
                reference_step(w, 2000 * cores, cores)
                f, dt, kind = reference_step(w, per_core * cores, cores)
            finally:
                shutil.rmtree(w, ignore_errors=True)
            line["cpu_baseline"] = {"value": f / dt, "unit": "fragments/s", "cores": cores, "kind": kind,
                                    "sample": "%d fragments (%d per core x %d cores) of the same workload, %s, no gzip hops, files on tmpfs" % (
                                        f, per_core, cores, "reference sources + our shim (oracle/_ref): fragSim;deamSim;adptSim per shard" if kind == "reference" else "oracle port")}
            if kind == "reference" and shutil.which("gzip"):               # the same with the driver's gzip hops, as gargammel.pl runs it
                w = prepare_reference_dir(sources)
                try:
                    fz, dtz, _ = reference_step(w, 50000 * cores, cores, gzip_hops=True)
                    line["cpu_baseline"]["with_gzip_hops"] = {"value": fz / dtz, "unit": "fragments/s", "sample": "%d fragments (50000 per core x %d cores), `| gzip >` after fragSim and deamSim, gz inputs to deamSim/adptSim (gargammel.pl:1510,1729,1847)" % (fz, cores)}
                finally:
                    shutil.rmtree(w, ignore_errors=True)
        except Exception as ex:      # the baseline is a reported figure; never let it kill the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "fragments/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (ex,)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
