"""Randomised parity sweep: seeded random combinations of genomes, size models, damage models, adapters, read lengths, emit dialects,
strands, --comp, -gc, --circ and id windows; the CUDA path must stay bit-exact against the oracle on every one of them."""
import numpy as np
import pytest

from gargammel_b200 import api
from tests import synth
from tests import util

pytestmark = pytest.mark.gpu

EMITS = [api.EMIT_FRAG, api.EMIT_DEAM, api.EMIT_STDOUT4, api.EMIT_ARTS, api.EMIT_ARTP, api.EMIT_FR, api.EMIT_RR]
DAMAGES = ["matrix", "briggs", "briggs_ss", "single", "double", "none"]


def _random_case(rng, tmp_path):
    n_genomes = int(rng.integers(1, 4))
    genomes = []
    for g in range(n_genomes):
        n_contigs = int(rng.integers(1, 7))
        total = int(rng.integers(400, 20000))
        genomes.append(synth.make_genome(int(rng.integers(1, 10**6)), n_contigs, total, prefix="g%dc" % g, width=int(rng.choice([13, 37, 60, 80, 200])),
                                         n_frac=float(rng.choice([0.0, 0.01, 0.05])), n_run=(3, 60), lower_frac=float(rng.choice([0.0, 0.3]))))
    case = {"genomes": genomes, "circ": [int(rng.integers(-1, len(fai))) if rng.random() < 0.3 else -1 for _, fai in genomes]}
    case["sizes"] = str(rng.choice(["freq", "list", "fixed", "lognormal"]))
    case["fixed_len"] = int(rng.integers(1, 120))
    case["minlen"] = int(rng.choice([0, 0, 30])); case["maxlen"] = int(rng.choice([1000, 1000, 90]))
    case["norev"] = bool(rng.random() < 0.2)
    case["read_len"] = int(rng.choice([30, 75, 100, 150]))
    ads = [api.DEFAULT_ADAPTER_F, "ACGT", "T" * 90, "GATTACAGATTACA"]
    case["adapters"] = (str(rng.choice(ads)), str(rng.choice(ads + [api.DEFAULT_ADAPTER_S])))
    case["damage"] = [str(rng.choice(DAMAGES)) for _ in range(3)]
    case["comp_dist"] = int(rng.choice([0, 0, 1, 3]))
    case["gc"] = bool(rng.random() < 0.2)
    case["case"] = bool(rng.random() < 0.25)                      # fragSim --case; half of those runs use the methylation tables for every range
    case["meth"] = case["case"] and bool(rng.random() < 0.5)
    case["seed"] = int(rng.integers(0, 2**62))
    case["counts"] = [int(rng.integers(1, 400)) for _ in range(n_genomes + 1)]
    return case


def _setup(ctx, case, tabs):
    ctx.set_case(case["case"])
    ids = [ctx.upload_genome(fa, fai, circ=c) for (fa, fai), c in zip(case["genomes"], case["circ"])]
    if case["sizes"] == "freq":
        ctx.set_sizes(api.SIZE_FREQ, *synth.golden_sizefreq(), minlen=case["minlen"], maxlen=case["maxlen"])
    elif case["sizes"] == "lognormal":
        ctx.set_sizes(api.SIZE_FREQ, *util.lognormal_table(3.8, 0.4), minlen=case["minlen"], maxlen=case["maxlen"])
    elif case["sizes"] == "list":
        ctx.set_sizes(api.SIZE_LIST, synth.golden_sizelist(), minlen=case["minlen"], maxlen=case["maxlen"])
    else:
        ctx.set_sizes(api.SIZE_FIXED, fixed_len=case["fixed_len"], minlen=0, maxlen=1000)
    ctx.set_norev(case["norev"])
    for slot, dm in enumerate(case["damage"]):
        util.set_damage(ctx, slot, "meth" if (case["meth"] and slot == 0) else dm, clamp_last=(slot != 1))
    ctx.set_adapters(case["adapters"][0], case["adapters"][1], case["read_len"])
    if tabs is not None:
        ctx.set_composition(0, case["comp_dist"], *tabs)
    if case["gc"]:
        ctx.set_gcbias(0, ids[0], case["comp_dist"] if tabs is not None else 0, 4.0)
    plan, first = [], 0
    for k, cnt in enumerate(case["counts"]):
        g = ids[k % len(ids)]
        slot = [0, 1, 2, -1][k % 4] if case["damage"][k % 3] != "none" else -1
        if case["meth"]:
            slot = 0
        plan.append(api.Range(first, cnt, g, slot, ["e1_0", "b12", "c1", ""][k % 4], 0 if (tabs is not None and k % 2 == 0) else -1,
                              0 if (case["gc"] and g == ids[0] and k % 2 == 0) else -1))
        first += cnt
    ctx.set_plan(plan)
    return first


@pytest.mark.parametrize("block", range(int(__import__("os").environ.get("GG_FUZZ_BLOCKS", "6"))))
def test_random_configurations_bit_exact(oracle_mod, tmp_path, block):
    from oracle import parsers
    rng = np.random.default_rng(9000 + block)
    done = 0
    for it in range(10):
        case = _random_case(rng, tmp_path)
        tabs = None
        if case["comp_dist"]:
            p = str(tmp_path / ("dnacomp_%d.txt" % it)); synth.write_dnacomp(p, seed=it)
            tabs = [np.array(t) for t in parsers.load_dnacomp(p, case["comp_dist"])]
        g = api.Context(0, case["seed"]); o = oracle_mod.OracleContext(case["seed"])
        try:
            try:
                total = _setup(o, case, tabs)
            except api.GGError:
                continue                                            # an inadmissible combination (e.g. -gc on a genome without a resolved window)
            assert _setup(g, case, tabs) == total
            for emit in rng.choice(EMITS, size=3, replace=False):
                lo = int(rng.integers(0, total)); cnt = int(rng.integers(1, total - lo + 1))
                try:
                    want, iw = o.run_fused(lo, cnt, int(emit))
                except api.GGError as e:
                    with pytest.raises(api.GGError):                # e.g. no admissible window: both sides must give up
                        g.run_fused(lo, cnt, int(emit))
                    continue
                got, ig = g.run_fused(lo, cnt, int(emit))
                util.assert_same(got, want, "block %d case %d emit %d: %r" % (block, it, int(emit), {k: v for k, v in case.items() if k != "genomes"}))
                assert (ig.attempts, ig.gather_bytes) == (iw.attempts, iw.gather_bytes)
                done += 1
        finally:
            g.close(); o.close()
    assert done >= 10


@pytest.mark.parametrize("block", range(int(__import__("os").environ.get("GG_FUZZ_BLOCKS", "4"))))
def test_random_record_streams_bit_exact(oracle_mod, block):
    """The stand-alone deamSim / adptSim kernels over random FASTA record streams: sequence lengths 0..400, lower case, N and other
    IUPAC letters, odd headers, every damage model with and without -name / -last, every dialect with and without -name / -tag."""
    rng = np.random.default_rng(7000 + block)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTacgtNnRYKM", dtype=np.uint8)
    for it in range(8):
        recs = []
        for i in range(int(rng.integers(1, 300))):
            n = int(rng.choice([0, 1, 2, 15, 16, 17, int(rng.integers(0, 400))]))
            hdr = b">" + bytes(rng.choice(np.frombuffer(b"abcXYZ019:+-_|.", dtype=np.uint8), size=int(rng.integers(1, 40))))
            recs.append(hdr + b"\n" + alphabet[rng.integers(0, len(alphabet), size=n)].tobytes() + b"\n")
        data = b"".join(recs)
        seed = int(rng.integers(0, 2**62)); first_id = int(rng.integers(0, 2**40))
        g = api.Context(0, seed); o = oracle_mod.OracleContext(seed)
        dm = str(rng.choice(["matrix", "briggs", "briggs_ss", "single", "double", "meth"]))
        name, clamp = bool(rng.random() < 0.5), bool(rng.random() < 0.7)
        rl = int(rng.choice([20, 75, 130])); tag = str(rng.choice(["", "x", "lib_42"])); addname = bool(rng.random() < 0.5)
        for c in (g, o):
            util.set_damage(c, 0, dm, clamp_last=clamp)
            c.set_deam_naming(name)
            c.set_adapters(read_len=rl); c.set_adpt_naming(addname, tag)
        util.assert_same(g.run_deam(data, 0, first_id)[0], o.run_deam(data, 0, first_id)[0], "deam %s name=%s clamp=%s block %d/%d" % (dm, name, clamp, block, it))
        for emit in (api.EMIT_STDOUT4, api.EMIT_ARTS, api.EMIT_ARTP, api.EMIT_FR, api.EMIT_RR):
            util.assert_same(g.run_adpt(data, emit, first_id)[0], o.run_adpt(data, emit, first_id)[0], "adpt emit=%d rl=%d tag=%r name=%s block %d/%d" % (emit, rl, tag, addname, block, it))
        g.close(); o.close()
