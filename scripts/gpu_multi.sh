#!/bin/bash
# 2-GPU check of the torchrun path (weak scaling, rank slices) -- run with: gpurun --gpus 2 -- bash scripts/gpu_multi.sh
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/multi2_bench.json 2> gpurun_out/multi2_bench.err
cat gpurun_out/multi2_bench.json; tail -3 gpurun_out/multi2_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/multi2_ref.json 2> gpurun_out/multi2_ref.err; cat gpurun_out/multi2_ref.json
# fused driver on 2 GPUs == 1 GPU byte for byte
python - <<'PY'
import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
from tests import synth
d = tempfile.mkdtemp()
for sub, names in (("endo", ["endo.1.fa"]), ("cont", ["cont.1.fa"]), ("bact", [])):
    os.makedirs(os.path.join(d, "in", sub))
    for k, nm in enumerate(names):
        fa, fai = synth.make_genome(5 + k, 3, 500000, prefix=sub[0])
        open(os.path.join(d, "in", sub, nm), "wb").write(fa); synth.write_fai(os.path.join(d, "in", sub, nm + ".fai"), fai)
synth.write_sizefreq(os.path.join(d, "sf.size"))
outs = []
for g in (1, 2):
    r = subprocess.run(["bin/gargammel-b200", "--comp", "0,0.1,0.9", "-n", "300000", "-f", os.path.join(d, "sf.size"), "-damage", "0.03,0.4,0.01,0.3", "--seed", "9", "--gpus", str(g), "--keep", "-o", os.path.join(d, "s%d" % g), os.path.join(d, "in")], stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr
    import gzip
    outs.append([open(os.path.join(d, "s%d_a.fa" % g), "rb").read()] + [gzip.open(os.path.join(d, "s%d%s" % (g, sfx))).read() for sfx in ("_d.fa.gz", ".e.fa.gz", ".b.fa.gz", ".c.fa.gz")])
    assert not [f for f in os.listdir(d) if ".part" in f], "part files left behind"
print("driver 1 GPU vs 2 GPUs identical (_a.fa, _d, .e, .b, .c):", [a == b for a, b in zip(outs[0], outs[1])], [len(a) for a in outs[0]])
PY
