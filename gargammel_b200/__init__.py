"""gargammel_b200 -- B200-native implementation of gargammel's per-fragment pipeline
(fragSim -> deamSim -> adptSim).  The product is the CUDA library behind the C-ABI in
include/gargammel_b200.h and the drop-in binaries in bin/; this package is the thin ctypes
harness used by the tests and bench.py.  It never falls back to a CPU path: importing
`gargammel_b200.api` without the built library raises."""
from .api import (Context, GGError, FaiEntry, Range, RunInfo, EMIT_FRAG, EMIT_DEAM, EMIT_STDOUT4, EMIT_ARTS,
                  EMIT_ARTP, EMIT_FR, EMIT_RR, SIZE_FIXED, SIZE_LIST, SIZE_FREQ, DMG_NONE, DMG_MATRIX,
                  DMG_BRIGGS, DMG_SINGLE, DMG_DOUBLE, DEFAULT_ADAPTER_F, DEFAULT_ADAPTER_S, lib_path)
__all__ = [n for n in dir() if not n.startswith("_")]
