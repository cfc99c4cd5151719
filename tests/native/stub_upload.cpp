// gg_host.cpp without the CUDA library: the one C-ABI entry it calls, never reached by the fuzzers
#include "../../include/gargammel_b200.h"
extern "C" int gg_genome_upload_circ(gg_ctx*, const uint8_t*, size_t, const gg_fai_entry*, int, int, int*) { return -1; }
