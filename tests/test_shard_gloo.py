"""world_size-2 test of the N>1 host logic on CPU (gloo): each rank computes its contiguous id slice (through the
oracle -- there is no GPU here), rank outputs concatenated in rank order equal the single-process stream, and the count
reduction matches."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gargammel_b200 import api, shard
from tests import util


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    ctx = oracle.OracleContext(99)
    ids = util.configure(ctx, util.tiny_genomes(), damage="briggs")
    ctx.set_plan(util.plan3(ids, counts=(500, 400, 300)))
    lo, hi = shard.rank_slice(total, rank, world)
    out, info = ctx.run_fused(lo, hi - lo, api.EMIT_ARTP)
    open(os.path.join(outdir, "shard%d" % rank), "wb").write(out)
    tot = shard.reduce_counts([info.records, info.bytes, info.attempts])
    tmax = shard.max_over_ranks(float(rank + 1))
    dist.barrier()
    if rank == 0:
        open(os.path.join(outdir, "totals"), "w").write("%d %d %d %g" % (tot[0], tot[1], tot[2], tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_stream(tmp_path, oracle_mod):
    total, world = 1200, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ctx = oracle_mod.OracleContext(99)
    ids = util.configure(ctx, util.tiny_genomes(), damage="briggs")
    ctx.set_plan(util.plan3(ids, counts=(500, 400, 300)))
    full, info = ctx.run_fused(0, total, api.EMIT_ARTP)
    got = b"".join(open(str(tmp_path / ("shard%d" % r)), "rb").read() for r in range(world))
    assert got == full
    rec, nbytes, attempts, tmax = open(str(tmp_path / "totals")).read().split()
    assert (int(rec), int(nbytes), int(attempts)) == (info.records, info.bytes, info.attempts) and float(tmax) == 2.0


def test_rank_slices_partition():
    for total in (0, 1, 7, 10_000_000, 10 ** 9 + 7):
        for world in (1, 2, 4, 8):
            s = [shard.rank_slice(total, r, world) for r in range(world)]
            assert s[0][0] == 0 and s[-1][1] == total and all(a[1] == b[0] for a, b in zip(s[:-1], s[1:]))
            assert max(h - l for l, h in s) - min(h - l for l, h in s) <= 1
