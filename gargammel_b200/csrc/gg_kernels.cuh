// gg_kernels.cuh -- parameter blocks shared by the host API (gg_api.cu) and the kernels (gg_kernels.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "gg_device.cuh"
#include "gg_lz.h"

namespace ggk {

struct GenomeDev {           // one uploaded FASTA file, packed (DESIGN.md "Data layout in HBM")
    const uint32_t* seq2;    // 2-bit codes, 16 bases per word
    const uint32_t* nmask;   // 1 bit per base, 1 = not A/C/G/T, 32 bases per word
    uint64_t G;              // sum of contig lengths = size of the sampling space (fragSim.cpp:1163)
    int first_contig;        // index of this genome's first contig in the (name-sorted) contig table
    int n_contigs;
    // fragSim --circ (fragSim.cpp:1407-1426): windows that run off the circular contig are gathered from a "wrap image"
    // = contig[wrap_s0 .. len-1) ++ contig[0 ..) packed behind the last contig (the reference drops the contig's last base there)
    int circ_contig;         // absolute index in the contig table, -1 = none
    uint32_t wrap_s0;
    uint64_t wrap_gbase;
};
struct ContigTab {           // struct-of-arrays over all contigs of all genomes
    const uint64_t* cum_end; // inclusive running end within its genome, name-sorted order (fragSim.cpp:1160-1163)
    const uint64_t* gbase;   // first base of the contig in seq2/nmask base space (64-aligned)
    const uint32_t* len;
    const uint32_t* name_off;
    const uint16_t* name_len;
    const char*     names;
};
struct RangeDev {            // one `fragSim -tag T -n N file` of the plan
    uint64_t first_id, count;
    int genome, slot, tag_len, comp;
    int gc, pad0;
    char tag[24];
};
struct CompDev {             // fragSim --comp tables of one slot (fp64: the acceptance product must round like the reference)
    int dist, pad;
    const double* tab;       // [4 tables: 5p+,5p-,3p+,3p-][2*dist][4]
    double maxP, maxM;       // fragSim.cpp:1030-1047
};
struct AdptNaming {           // adptSim -name / -tag (adptSim.cpp:375-393)
    int add_in_name, tag_len; char tag[24];
};
struct GcDev {               // fragSim -gc: survive iff u31 < thr[W*(W+1)/2 + gcCount], W = window length (fragment + flanks)
    const uint32_t* thr; int wmax, on;
};
struct Adapters {
    const uint32_t* adF2;    // forward adapter, 2-bit packed (+slack)
    const uint32_t* adS2;    // second adapter
    const uint32_t* rcS2;    // reverse complement of the second adapter
    int lenF, lenS, read_len, pad;
};
struct SmemLayout {          // byte offsets into dynamic shared memory, computed by the host
    int thr, thr_bytes;      // CTA-shared copy of the matrix tables of the used damage slots (DMODE 1)
    int thr_slot[4];         // uint4 index of each slot's acceptance rows inside that copy ...
    int term_slot[4];        // ... and of its terminal rows
    int geo_slot[4], geo_n;  // byte offset (in that copy) of each slot's skip tables: geo_n thresholds + 512 guide bytes, once (matrix) or twice (Briggs: s, then d); -1 = none
    int warp0, warp_bytes;   // per-warp regions start at warp0 + warp * warp_bytes; the offsets below are relative to a region
    int stage, stage_bytes;  // output staging (readable slack before and after)
    int bh;                                        // per-fragment Briggs header draws (overhangs, nick), u64 each
    int dwords, lwords, cmap, gmap, total;
    // CTA-shared copy of the plan / genome / contig tables when they are small (tab < 0: read them from global memory)
    int tab, tab_ranges, tab_genomes, tab_cum, tab_gbase, tab_name8, tab_len, tab_noff, tab_nlen, tab_names, tab_bytes;
    int n_contigs, names_bytes;
    // ... plus the FREQ length tables (n_sizes_sm entries, 0 = not copied) and the three packed adapters (ad_words[k] words each incl.
    // guard and slack, 0 = not copied), byte offsets relative to tab
    int tab_sthr, tab_slen, n_sizes_sm, tab_ad[3], ad_words[3];
};
struct FusedParams {
    // work
    uint64_t first, count;   // fragment ids [first, first+count)
    int F;                   // fragments per warp tile (32 unless a record needs more shared memory)
    int n_tiles;
    int emit;                // gg_emit_mode
    int norev;
    int fast_boundary;       // every header >= 13 bytes: whole-group stores, headers written afterwards
    int dmode;               // 1: every damage slot used by the plan is a matrix (specialised kernel), 2: every slot is Briggs, 0: any mix
    // constants of the adptSim dialects (every fragment has the same sequence-line length there)
    int S_ad;                // bases per sequence line: read length, or twice that for -artp
    int LWL, lw_stride;      // words per packed line image (guard + data + slack), words per fragment (lines x LWL)
    uint32_t gpl_ad, gpf_ad; // 16-byte staging groups per line / per fragment
    uint32_t gmagic;         // ceil(2^32 / gpf_ad): group index -> fragment by one multiply
    int sfx0;                // 2 when the only line carries "_r" (-rr), else 0
    int E;                   // fragments per emission half (the staging buffer holds E records)
    uint32_t lut;            // 0x54474341: the bytes 'A','C','G','T'
    ggd::Key key;
    ggd::KeySched ks;        // the ten Philox round keys of `key`
    // plan + genomes
    const RangeDev* ranges; int n_ranges;
    const GenomeDev* genomes; int n_genomes;
    ContigTab ctab;
    // fragment lengths
    int size_mode, fixed_len, minlen, maxlen, n_sizes;
    const uint32_t* size_thr;   // FREQ: u31 < size_thr[i]  <=>  p <= cum[i]
    const int32_t*  size_len;
    // damage + adapters
    ggd::DamageDev dmg[4];
    CompDev comp[4];
    GcDev gc[4];
    Adapters ad;
    // output
    uint8_t* out; uint64_t out_cap;
    unsigned long long* tile_state;   // decoupled look-back: [63:62] status, [61:0] bytes
    unsigned int* ticket;
    unsigned long long* stats;        // [0] total bytes [1] attempts [2] gather bytes [3] error flags
    // fragSim --case (null otherwise): per fragment, the output offset of its first base, its genome position and its META word
    unsigned long long* case_off; unsigned long long* case_gpos; uint32_t* case_meta;
    SmemLayout sm;
};

#ifndef GG_FUSED_MAXT
#define GG_FUSED_MAXT 1024     // launch bound of k_fused: 32 warps, 64 registers
#endif
enum { ST_TOTAL = 0, ST_ATTEMPTS = 1, ST_GATHER = 2, ST_ERR = 3, ST_N = 8 };
enum { ERRF_GIVEUP = 1, ERRF_OVERFLOW = 2, ERRF_PARSE = 4, ERRF_FALLBACK = 8 };


// ---- launchers (gg_kernels.cu) ----
// ---- device BGZF (Huffman-only deflate) ----
constexpr int BGZF_BLOCK = 65280;        // input bytes per BGZF member (the same 0xff00 htslib uses)
constexpr int BGZF_RUN = 68;             // bytes per thread: 17 words, an odd word stride (conflict-free shared-memory reads)
constexpr int BGZF_THREADS = 960;        // 960 x 68 = 65280
constexpr int BGZF_HDR_WORDS = 40;       // room for the dynamic-code description (<= 1280 bits)
// table layout (uint32): [0,257) symbol codes | [260,1284) CRC slicing-by-4 | [1284,2244) per-thread shift polynomials x^(8*68*(959-t))
// | [2244,2261) x^(8*2^k) | [2264, 2264+BGZF_HDR_WORDS) header bits
constexpr int BGZF_TAB_SYM = 0, BGZF_TAB_CRC = 260, BGZF_TAB_SHIFT = 1284, BGZF_TAB_X8 = 2244, BGZF_TAB_HDR = 2264, BGZF_TAB_WORDS = BGZF_TAB_HDR + BGZF_HDR_WORDS;
struct BgzfParams {
    const uint8_t* in; uint64_t n;       // plain stream (16-byte aligned)
    uint8_t* out; uint64_t out_cap;      // BGZF members, back to back (16-byte aligned)
    const uint32_t* tab; int hdr_nbits;
    uint32_t n_blocks;
    unsigned long long* state;           // look-back state per block
    unsigned int* ticket; unsigned long long* total;
};
// ---- device BGZF with LZ77 matches (k_bgzf_lz; parse specification in gg_lz.h): the default encoder ----
// table layout (uint32): [0,316) literal/length + distance codes | [320,1344) CRC slicing-by-4 | [1344,2304) per-thread shift polynomials
// | [2304,2321) x^(8*2^k) | [2324, 2324+BGZL_HDR_WORDS) header bits
constexpr int BGZL_HDR_WORDS = 64;       // room for the description of both codes (<= 2048 bits)
constexpr int BGZL_TAB_SYM = 0, BGZL_TAB_CRC = 320, BGZL_TAB_SHIFT = 1344, BGZL_TAB_X8 = 2304, BGZL_TAB_HDR = 2324, BGZL_TAB_WORDS = BGZL_TAB_HDR + BGZL_HDR_WORDS;
struct BgzfLzParams {
    const uint8_t* in; uint64_t n;       // plain stream (16-byte aligned)
    uint8_t* out; uint64_t out_cap;      // BGZF members, back to back (16-byte aligned)
    const uint32_t* tab; int hdr_nbits;
    uint32_t n_blocks;                   // members to visit (counting pass: the sampled ones, member index = visit * stride)
    uint32_t stride;
    unsigned long long* state;           // look-back state per block
    unsigned int* ticket; unsigned long long* total;
    unsigned long long* hist;            // counting pass: 316 symbol counts
};
cudaError_t launch_bgzf_lz(const BgzfLzParams& P, bool count_only, int grid, cudaStream_t st);
cudaError_t launch_byte_hist(const uint8_t* in, uint64_t n, uint64_t group_stride, unsigned long long* hist, cudaStream_t st);
cudaError_t bgzf_prepare(size_t* smem_bytes);
cudaError_t launch_bgzf_deflate(const BgzfParams& P, int grid, cudaStream_t st);

cudaError_t launch_pack_genome(const uint8_t* fasta, uint64_t n, const uint64_t* c_off, const uint32_t* c_len,
                               const int32_t* c_blen, const int32_t* c_llen, const uint64_t* c_gbase,
                               int n_contigs, int wrap_desc, uint32_t wrap_keff, uint32_t wrap_srclen, uint64_t n_units,
                               uint32_t* seq2, uint32_t* nmask, uint32_t* cmask, cudaStream_t st);
cudaError_t launch_case_apply(uint64_t n_rec, const unsigned long long* off, const unsigned long long* gpos, const uint32_t* meta,
                              const uint32_t* const* cmasks, uint8_t* out, cudaStream_t st);
cudaError_t launch_gc_probe(const GenomeDev& G, const uint64_t* cum_end, const uint64_t* gbase, const uint32_t* clen, int dist, ggd::Key key,
                            unsigned long long* sums /* [0]=sum gc, [1]=sum gc^2, [2]=give-ups */, cudaStream_t st);
int fused_desc_bytes();      // bytes of the fixed-layout descriptors at the head of a warp's shared-memory region
cudaError_t launch_fused(const FusedParams& P, int emitc, int grid, int threads, cudaStream_t st);
cudaError_t fused_prepare(int fast, int dmode, int emitc, int threads, int smem_bytes, int* blocks_per_sm, int* regs);

// stand-alone record-stream kernels
struct RecStream {
    const uint8_t* in; uint64_t n;        // input bytes
    const uint64_t* nl; uint64_t n_nl;    // positions of '\n' (an implicit one at n if the stream lacks it)
    uint64_t n_rec;
};
cudaError_t launch_count_newlines(const uint8_t* in, uint64_t n, unsigned long long* block_counts, int n_blocks, cudaStream_t st);
cudaError_t launch_fill_newlines(const uint8_t* in, uint64_t n, const unsigned long long* block_base, uint64_t* nl, int n_blocks, cudaStream_t st);
cudaError_t launch_exclusive_scan(unsigned long long* data, uint64_t n, unsigned long long* scratch, unsigned long long* total, cudaStream_t st);
cudaError_t launch_deam_records(const RecStream& R, uint8_t* out, const ggd::DamageDev& D, int has_damage, ggd::Key key,
                                uint64_t first_id, unsigned long long* stats, cudaStream_t st);
// single-pass stand-alone deamSim (no record index, no host round trip); ERRF_FALLBACK in stats[ST_ERR] = take the general path instead
cudaError_t launch_deam_stream(const uint8_t* in, uint64_t n, uint8_t* out, const ggd::DamageDev& D, int has_damage, ggd::Key key, uint64_t first_id,
                               unsigned long long* tile_state, unsigned int* ticket, unsigned long long* stats /* [ST_TOTAL] = lines */, uint32_t n_tiles, int sm_count, cudaStream_t st);
int deam_stream_tile_bytes();
cudaError_t launch_deam_name_sizes(const RecStream& R, const ggd::DamageDev& D, ggd::Key key, uint64_t first_id, unsigned long long* sizes, unsigned long long* stats, cudaStream_t st);
cudaError_t launch_deam_name_records(const RecStream& R, const unsigned long long* out_off, uint8_t* out, const ggd::DamageDev& D, ggd::Key key, uint64_t first_id, cudaStream_t st);
cudaError_t launch_adpt_sizes(const RecStream& R, int emit, Adapters ad, AdptNaming nm, unsigned long long* sizes, unsigned long long* stats, cudaStream_t st);
cudaError_t launch_adpt_records(const RecStream& R, const unsigned long long* out_off, uint8_t* out, int emit, Adapters ad,
                                const uint8_t* adF, const uint8_t* adS, AdptNaming nm, ggd::Key key, uint64_t first_id, cudaStream_t st);
cudaError_t launch_l2_flush(uint4* buf, uint64_t n16, cudaStream_t st);
cudaError_t launch_publish(const unsigned long long* src, unsigned long long* dst_mapped, int n, cudaStream_t st);
constexpr int NL_BLOCK_BYTES = 1 << 16;

} // namespace ggk
