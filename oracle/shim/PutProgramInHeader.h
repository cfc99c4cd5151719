/* PutProgramInHeader.h -- SHIM written for this repo.  BAM I/O is out of scope: stubs only. */
#ifndef SHIM_PPIH_H
#define SHIM_PPIH_H
#include <string>
#include "api/BamAux.h"
using namespace BamTools;
inline void putProgramInHeader(SamHeader*, const std::string&, const std::string&, const std::string&, const std::string&) {}
#endif
