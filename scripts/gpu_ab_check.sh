#!/bin/bash
# A/B timing like gpu_ab.sh, preceded by the smoke check (bit-exact vs the oracle) of every variant library
for spec in "$@"; do
  lib=${spec%%:*}
  if [ "$lib" == "cur" ]; then python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; else GG_LIB=$PWD/$lib python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; fi
done
bash scripts/gpu_ab.sh "$@"
