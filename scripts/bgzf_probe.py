#!/usr/bin/env python3
"""Device BGZF on the headline stream: time of gg_bgzf_dev (events around histogram + host code build + deflate kernel)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from gargammel_b200 import api
from tests import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(7)
edges = np.linspace(0, 50_000_000, 6).astype(np.int64)
fa, fai = synth.fasta_bytes([("chr%d" % (i + 1), synth.random_contig(rng, int(edges[i + 1] - edges[i]), 0.005)) for i in range(5)], 60)
ctx = api.Context(0, 42)
gid = ctx.upload_genome(fa, fai)
ctx.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq())
ctx.set_matrix(0, *synth.golden_matrix("single-")); ctx.set_adapters(read_len=75)
ctx.set_plan([api.Range(0, n, gid, 0, "e1_0")])
import ctypes as C
lib = api.load_library()
for rep in range(4):
    o, z = api._Out(), api._Out()
    ctx._ck(lib.gg_run_fused(ctx._h, C.c_uint64(0), C.c_uint64(n), C.c_int(api.EMIT_ARTP), C.byref(o)))
    ctx._ck(lib.gg_bgzf_dev(ctx._h, C.byref(o), C.byref(z)))
    print("fused %.3f ms  bgzf %.3f ms  in %d out %d  %.1f GB/s in" % (o.device_ms, z.device_ms, z.in_bytes, z.bytes, z.in_bytes / z.device_ms / 1e6), flush=True)
