// gg_host.h -- host-side (C++11) file-format layer and fragment apportioning of the gargammel hot path.
// Parsers follow the reference's loaders line by line (cited at each function); no CUDA here.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/gargammel_b200.h"
#include "gg_lz.h"

namespace ggh {

struct FaiEntry { std::string name; int64_t len; uint64_t offset; int32_t line_blen, line_len; };

// whole file into memory; transparently inflates gzip (zlib gzread)
bool read_file(const std::string& path, std::vector<uint8_t>& out, std::string& err);
// split on a delimiter keeping empty tokens (libgab allTokens semantics as used by the reference)
std::vector<std::string> all_tokens(const std::string& s, char delim);

// samtools .fai, 5 tab-separated columns (fragSim.cpp:231-269)
bool parse_fai(const std::string& path, std::vector<FaiEntry>& out, std::string& err);
// build a .fai from FASTA bytes (the driver runs `samtools faidx` when it is missing, gargammel.pl:900-903)
bool build_fai(const std::vector<uint8_t>& fasta, std::vector<FaiEntry>& out, std::string& err);

// -s file: one length per line, keep t >= 1 (fragSim.cpp:720-745)
bool load_size_list(const std::string& path, std::vector<int32_t>& len, std::string& err);
// -f file: "length<TAB>freq", running cumulative sum, require 0.99 <= sum <= 1.01 (fragSim.cpp:751-784)
bool load_size_freq(const std::string& path, std::vector<int32_t>& len, std::vector<double>& cum, std::string& err);

// --loc/--scale: length = int(lognormal(loc, scale)) (fragSim.cpp:184,671) as an inverse-CDF table over lengths 0..cap:
// cum[l] = P(length <= l) = Phi((ln(l+1) - loc) / scale); a last row `cap+1` with cum 1 holds the tail (rejected by -M).
void lognormal_table(double loc, double scale, int cap, std::vector<int32_t>& len, std::vector<double>& cum);

// --comp file (mapDamage dnacomp.txt) with --dist d: four [2d][4] tables of "frequency - background + 0.25" (fragSim.cpp:798-1012)
struct CompTables { int dist; std::vector<double> s5p, s5m, s3p, s3m; };
bool load_dnacomp(const std::string& path, int dist, CompTables& out, std::string& err);

typedef std::vector<double> Rates;   // rows x 12, idx = from*3 + (to<from ? to : to-1), bases A,C,G,T
// -matfile prefix: <prefix>5.dat / <prefix>3.dat; first column dropped, "value [lo..hi]" -> value; per-base sums <= 1;
// the 3' file is reversed so that row 0 is the terminal base (deamSim.cpp:1098-1241)
bool load_matrix_dat(const std::string& prefix, Rates& sub5, Rates& sub3, std::string& err);
// -profile prefix: <prefix>5p.prof / <prefix>3p.prof, 12 columns from field 0, 3' NOT reversed (deamSim.cpp:956-1094)
bool load_profile(const std::string& prefix, Rates& sub5, Rates& sub3, std::string& err);
// -mapdamage misincorporation.txt: header-driven column remap, counts summed over chromosomes/strands (deamSim.cpp:453-668)
bool load_mapdamage(const std::string& path, Rates& sub5, Rates& sub3, std::string& err);

// bact/list: "name<TAB>weight" (select_random_speciesftp.py:66-70; gargammel.pl:1040-1105)
bool load_bact_list(const std::string& path, std::vector<std::string>& names, std::vector<double>& w, std::string& err);

// ---- apportioning (gargammel.pl:1256-1283,1320-1349,1421-1431) ----
struct ApportionIn {
    uint64_t n_fragments;                  // -n
    double compB, compC, compE;            // --comp B,C,E
    int n_endo;                            // 1 (haploid) or 2 (diploid)
    std::vector<uint64_t> cont_len;        // contaminant file lengths (bp)
    std::vector<double> bact_w;            // bacterial weights (sum ~ 1)
    uint64_t seed;
    // -c coverage path (gargammel.pl:1255-1258): the endogenous count is fixed by the coverage and kept as it is; only the contaminant and
    // bacterial counts derive from n_fragments = int(nE / compE), as int(n * comp / (compE + compC + compB))
    bool has_nE; uint64_t nE;
    // multi-cell mode (endo/C0 .. endo/CN, gargammel.pl:1017-1087,1363-1403): number of cell directories, 0 = the plain layout
    int n_cells;
    ApportionIn() : n_fragments(0), compB(0), compC(0), compE(1), n_endo(1), seed(0), has_nE(false), nE(0), n_cells(0) {}
};
struct ApportionOut {
    uint64_t nE, nC, nB;                   // int(n*comp)  (gargammel.pl:1261-1263)
    std::vector<uint64_t> endo, cont, bact;
    std::vector<uint64_t> cell, cell_e1, cell_e2;   // multi-cell mode: fragments per cell and from each cell's first / second file
};
bool apportion(const ApportionIn& in, ApportionOut& out, std::string& err);
// exact Binomial(n,p) draw from a Philox-fed uniform stream (inversion for small n*p, BTRS otherwise)
uint64_t binomial_draw(uint64_t n, double p, uint64_t seed, uint64_t stream);

// ---- parallel BGZF (blocked gzip, readable by zcat/gzstream/htslib) of an output stream ----
// Splits `in` into 65280-byte blocks, deflates them on `threads` host threads (0 = all), concatenates in order.
bool bgzf_compress(const uint8_t* in, size_t n, int threads, int level, std::vector<uint8_t>& out, bool eof_marker, std::string& err);

// ---- Huffman-only deflate code for the device BGZF encoder (k_bgzf_deflate) ----
// One dynamic-Huffman code (RFC 1951 3.2.7) for a whole output stream, built from a byte histogram: every literal 0..255 and the
// end-of-block symbol get a code of at most 15 bits (unseen bytes are floored to a count of 1, so any input is encodable); no
// distance codes (HDIST = 0 with one zero-length code: "all literals").  sym[s] = bit-reversed code | len << 16 (ready to be OR-ed
// into an LSB-first bit stream); header = BFINAL=1, BTYPE=10 and the code description, LSB-first, header_nbits bits.
struct DeflateCode { uint32_t sym[257]; std::vector<uint8_t> header; int header_nbits; };
bool build_deflate_code(const uint64_t hist[256], DeflateCode& out, std::string& err);
// ---- LZ77 + Huffman deflate code for the device BGZF encoder with matches (k_bgzf_lz; parse specification in gg_lz.h) ----
// sym[0..286): literal/length codes, sym[286..316): distance codes (bit-reversed code | len << 16), every symbol encodable; header =
// BFINAL=1, BTYPE=10, HLIT=29, HDIST=29 and the code description.  hist = symbol counts of the sampled members (316).
struct LzCode { uint32_t sym[gglz::N_SYM]; std::vector<uint8_t> header; int header_nbits; };
bool build_lz_code(const uint64_t hist[gglz::N_SYM], LzCode& out, std::string& err);
// host model of the device encoder: the same parse, the same code, the same members byte for byte (tests pin the kernel to it)
bool bgzf_lz_model(const uint8_t* in, size_t n, std::vector<uint8_t>& out, bool eof_marker, std::string& err, uint64_t* hist_out = nullptr);
// CRC-32 (IEEE, reflected) helpers for the device encoder: slicing-by-4 tables and x^(8n) mod P in the reflected domain
void crc32_tables(uint32_t t[4][256]);
uint32_t crc32_x8n(uint64_t n_bytes);                       // multiply a CRC by this (crc32_mul) to append n_bytes zero-effect bytes
uint32_t crc32_mul(uint32_t a, uint32_t b);

// ---- unaligned BAM in/out (the reference goes through bamtools: fragSim.cpp:1305-1325,1589-1601, deamSim.cpp:1260-1286,
// 2016-2024, adptSim.cpp:217-243,328-373).  Records are serialised as the SAM/BAM specification lays them out (little-endian; 4-bit
// bases "=ACMGRSVTWYHKDBN"; qualities as raw phred, 0xFF = absent); the container is BGZF: members of <= 65280 bytes made by the device
// encoder (gg_bgzf_host), by zlib on host threads, or stored (-u).
struct BamRecord {
    std::string name, seq;                  // seq in ASCII (upper-case IUPAC; "" = '*')
    std::vector<uint8_t> qual;              // raw phred, one per base; empty = 0xFF for every base
    std::vector<uint8_t> aux;               // tag bytes as they sit in the file
    std::vector<uint32_t> cigar;
    int32_t refid, pos, next_refid, next_pos, tlen; uint16_t flag, bin; uint8_t mapq;
    BamRecord() : refid(-1), pos(-1), next_refid(-1), next_pos(-1), tlen(0), flag(4), bin(4680), mapq(0) {}     // bamtools' fresh BamAlignment; bin = reg2bin(-1, 0)
    void add_tag_z(const char tag[2], const std::string& v) { aux.push_back((uint8_t)tag[0]); aux.push_back((uint8_t)tag[1]); aux.push_back('Z'); aux.insert(aux.end(), v.begin(), v.end()); aux.push_back(0); }
};
struct BamHeader { std::string text; std::vector<std::pair<std::string, uint32_t> > refs; };
// "BAM\1" + text + references
void bam_header_bytes(const BamHeader& h, std::vector<uint8_t>& out);
void bam_append_record(const BamRecord& r, std::vector<uint8_t>& out);
// the whole file: BGZF members inflated, header and records parsed
bool bam_read(const std::string& path, BamHeader& h, std::vector<BamRecord>& recs, std::string& err);
// "@PG\tID:..\tPN:..\tCL:..\tVN:.." appended to a header text (putProgramInHeader; an ID already present gets ".1", ".2", ...)
void bam_add_program(BamHeader& h, const std::string& id, const std::string& cmdline);

// ---- one genome file ready for upload ----
struct GenomeFile { std::string path; std::vector<uint8_t> bytes; std::vector<FaiEntry> fai; uint64_t total_len; };
bool load_genome(const std::string& fasta_path, GenomeFile& g, std::string& err);
int  upload_genome(gg_ctx* ctx, const GenomeFile& g, int* id_out, int circ_chr = -1);   // circ_chr: fai index of the --circ contig

} // namespace ggh

// C exports for the Python harness (ctypes); caller-provided buffers, 0 on success, message in err
extern "C" {
int ggh_load_size_freq(const char* path, int32_t* len, double* cum, int cap, int* n, char* err, int errcap);
int ggh_load_size_list(const char* path, int32_t* len, int cap, int* n, char* err, int errcap);
/* kind: 0 = -matfile (.dat), 1 = -profile (.prof), 2 = -mapdamage (path is the file) */
int ggh_load_rates(const char* path, int kind, double* sub5, int cap5, int* rows5, double* sub3, int cap3, int* rows3, char* err, int errcap);
int ggh_fai_load(const char* path, int from_fasta, void** handle, int* n, char* err, int errcap);
int ggh_fai_get(void* handle, int i, char* name, int namecap, int64_t* len, uint64_t* offset, int32_t* blen, int32_t* llen);
void ggh_fai_free(void* handle);
int ggh_apportion(uint64_t n, double compB, double compC, double compE, int n_endo, const uint64_t* cont_len, int n_cont,
                  const double* bact_w, int n_bact, uint64_t seed, uint64_t* totals3, uint64_t* endo, uint64_t* cont, uint64_t* bact,
                  char* err, int errcap);
uint64_t ggh_binomial(uint64_t n, double p, uint64_t seed, uint64_t stream);
/* returns 0 and the compressed size in *out_n; out needs n + n/255 + 64 KiB */
int ggh_bgzf_compress(const uint8_t* in, size_t n, int threads, int level, int eof_marker, uint8_t* out, size_t cap, size_t* out_n);
int ggh_bgzf_lz_model(const uint8_t* in, size_t n, int eof_marker, uint8_t* out, size_t cap, size_t* out_n, uint64_t* hist);
int ggh_load_dnacomp(const char* path, int dist, double* s5p, double* s5m, double* s3p, double* s3m, char* err, int errcap);   /* each [2*dist][4] */
int ggh_lognormal_table(double loc, double scale, int cap, int32_t* len, double* cum);   /* cap+2 rows */
/* BAM round trip for the tests: a 2-line FASTA record stream -> unaligned BAM file (flag 4, phred 0, stored BGZF members), and back */
int ggh_bam_from_fasta(const uint8_t* recs, size_t n, const char* path, char* err, int errcap);
int ggh_bam_to_fasta(const char* path, uint8_t* out, size_t cap, size_t* out_n, char* err, int errcap);
}
