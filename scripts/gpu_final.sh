#!/bin/bash
# Last GPU call of the round (a few minutes of budget): the parity suite at HEAD, the driver wall clocks (gargammel.pl over the reference
# filters / over bin/ against bin/gargammel-b200 on the same job), C5 with -s sizedist, then the default bench as far as the budget goes.
# Every step writes its own file under gpurun_out/ as soon as it ends.
TAG=${1:-r02_final}
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -x -q -m gpu ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
( time timeout 200 python scripts/driver_walltime.py --big ) > gpurun_out/${TAG}_driver_walltime.json 2> gpurun_out/${TAG}_driver_walltime.err; cat gpurun_out/${TAG}_driver_walltime.json; tail -4 gpurun_out/${TAG}_driver_walltime.err
( time timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --config C5 --c5-sizes sizedist --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c5_sizedist.json 2> gpurun_out/${TAG}_c5_sizedist.err; cut -c1-300 gpurun_out/${TAG}_c5_sizedist.json
