/* api/BamReader.h -- SHIM, see api/BamAux.h */
#include "BamAux.h"
