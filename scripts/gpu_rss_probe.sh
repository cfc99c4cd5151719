cd $GRAFT_REPO_ROOT
python - <<PY
import numpy as np, sys
sys.path.insert(0,".")
from tests import synth
n, L = 10_000_000, 60
rec = np.empty((n, 1 + 8 + 1 + L + 1), dtype=np.uint8)
rec[:, 0] = ord(">"); rec[:, 9] = 10; rec[:, -1] = 10
idx = np.arange(n, dtype=np.int64)
for d in range(8): rec[:, 8 - d] = 48 + (idx // 10 ** d) % 10
rec[:, 10:10 + L] = synth._BASES[np.random.default_rng(4).integers(0, 4, size=(n, L), dtype=np.uint8)]
rec.tofile("/tmp/big.fa")
s5, s3 = synth.golden_matrix("single-"); synth.write_matrix_dat("/tmp/m-", s5, s3)
PY
GG_PRINT_RSS=1 bin/deamSim -matfile /tmp/m- /tmp/big.fa > /tmp/big.out
