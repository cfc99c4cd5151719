/* api/BamAux.h -- SHIM: inert BamTools stand-ins so the unmodified reference sources compile.  Every BAM
 * open fails, so the reference's own "Could not open ... BAM" error paths fire if -b/-bs/-bp is used. */
#ifndef SHIM_BAMAUX_H
#define SHIM_BAMAUX_H
#include <string>
#include <vector>
#include <stdint.h>
namespace BamTools {
struct SamProgram {};
struct SamHeader {};
struct RefData {};
typedef std::vector<RefData> RefVector;
struct BamAlignment {
    std::string Name, QueryBases, Qualities; uint32_t AlignmentFlag;
    BamAlignment() : AlignmentFlag(0) {}
    void SetIsMapped(bool) {}
    void SetIsPaired(bool) {}
    template <typename T> bool AddTag(const std::string&, const std::string&, const T&) { return true; }
};
class BamReader {
public:
    bool Open(const std::string&) { return false; }
    SamHeader GetHeader() const { return SamHeader(); }
    RefVector GetReferenceData() const { return RefVector(); }
    bool GetNextAlignment(BamAlignment&) { return false; }
    void Close() {}
};
class BamWriter {
public:
    enum CompressionMode { Compressed, Uncompressed };
    void SetCompressionMode(CompressionMode) {}
    bool Open(const std::string&, const SamHeader&, const RefVector&) { return false; }
    void SaveAlignment(const BamAlignment&) {}
    void Close() {}
};
}
#endif
