import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from gargammel_b200 import api
src = bench.make_sources()
ctx = api.Context(0, 42)
pinned = []
for fa, fai in src:
    t = torch.empty(len(fa), dtype=torch.uint8, pin_memory=True); t.numpy()[:] = np.frombuffer(fa, dtype=np.uint8); pinned.append((t, fai))
bench.setup_tables(api, ctx)
n = 10_000_000
cap = None
for it in range(4):
    t0 = time.perf_counter(); ctx.genome_clear(); t1 = time.perf_counter()
    ids = [ctx.upload_genome(t, fai) for t, fai in pinned]; t2 = time.perf_counter()
    plan, _, _ = bench.make_plan(api, n, ids); ctx.set_plan(plan); t3 = time.perf_counter()
    if cap is None:
        cap = int(ctx.max_record_bytes(api.EMIT_ARTP)) * n + 4096
        host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    _, info = ctx.run_fused(0, n, api.EMIT_ARTP, fetch=False); t4 = time.perf_counter()
    info2 = ctx.run_fused_into(0, n, api.EMIT_ARTP, host.data_ptr(), cap); t5 = time.perf_counter()
    print("clear %.1f ms  upload %.1f ms  plan %.1f ms  run %.1f ms (dev %.2f)  run+fetch %.1f ms  bytes %.2f GB" % (
        1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), info.device_ms, 1e3*(t5-t4), info2.bytes/1e9), flush=True)
