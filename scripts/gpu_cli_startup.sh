#!/bin/bash
# Where does a drop-in filter process spend its time?  (per-process cost matters when gargammel.pl launches one process per source file)
mkdir -p gpurun_out; T=$(mktemp -d); cd $GRAFT_REPO_ROOT
python - <<PY
import sys; sys.path.insert(0, ".")
import numpy as np
from tests import synth
rng = np.random.default_rng(1)
fa, fai = synth.fasta_bytes([("chr%d" % i, synth.random_contig(rng, 10_000_000)) for i in range(5)], 60)
open("$T/g.fa", "wb").write(fa); synth.write_fai("$T/g.fa.fai", fai)
s5, s3 = synth.golden_matrix("single-"); synth.write_matrix_dat("$T/m-", s5, s3)
PY
tm() { local w="$1"; shift; local a=$(date +%s%N); "$@"; local b=$(date +%s%N); echo "$w: $(( (b - a) / 1000000 )) ms wall"; }
nvidia-smi --query-gpu=persistence_mode --format=csv
for n in 1000 100000 1000000; do
  tm "fragSim -n $n" bin/fragSim -n $n -l 60 $T/g.fa > $T/f$n.fa 2> $T/err; tail -1 $T/err
done
tm "deamSim 1000" bin/deamSim -matfile $T/m- $T/f1000.fa > /dev/null 2> $T/err; tail -1 $T/err
tm "deamSim 1000000" bin/deamSim -matfile $T/m- $T/f1000000.fa > /dev/null 2> $T/err; tail -1 $T/err
tm "adptSim 1000000" bin/adptSim -artp $T/a.fa $T/f1000000.fa 2> $T/err; tail -1 $T/err
tm "python ctx create+destroy" python -c "
import time; t=time.time()
from gargammel_b200 import api
t1=time.time(); c=api.Context(0,1); t2=time.time(); c.close(); print('import %.2f ctx %.2f'%(t1-t,t2-t1))"
