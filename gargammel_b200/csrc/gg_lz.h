// gg_lz.h -- the parse specification of the device BGZF encoder with LZ77 matches (k_bgzf_lz), shared by the kernel and by the host
// model that pins it (ggh::bgzf_lz_model, gg_host.cpp).  RFC 1951 symbol arithmetic and the constants of the parse.
//
// A BGZF member holds up to 65280 input bytes = 30 runs of 2176 bytes; one warp parses one run, in steps of 32 consecutive positions:
//   * every position p of the step with p + 6 <= run end hashes its next 6 bytes (lz_hash) and reads the candidate T[h] of the warp's
//     512-entry table -- the most recent EARLIER-step position with that hash; then the step's positions are inserted (of equal hashes
//     the highest position stays).  Before its first step a warp inserts the 160 positions in front of its run (no look-ups), so
//     matches may reach back into the previous run.
//   * a candidate is a match start iff distance <= 32768 and its 6 bytes equal the position's 6 bytes.
//   * greedy, left to right: the first match start that is not covered by an earlier match is taken and extended byte by byte up to
//     min(258, run end - p); everything it covers is skipped; uncovered positions without a match are literals.
//   * a run whose codes need more than REGION bytes (6.6 bits per input byte) makes its member a stored block.
// (measured on simulated reads, bits per byte, this parse / zlib -1: fragSim records 3.21 / 3.22, -arts 3.04 / 3.09, -artp 2.20 / 2.33,
// four-line adptSim 2.62 / 2.76; with 8 hashed bytes and 256 entries: 3.34, 3.18, 2.29, 2.66)
// The result depends on nothing but the member's bytes: the device output is reproducible and equals the host model's byte for byte.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define GG_LZ_HD __host__ __device__ __forceinline__
#else
#define GG_LZ_HD static inline
#endif

namespace gglz {
constexpr int RUN = 2176;            // bytes per warp run (30 runs per 65280-byte member)
constexpr int WARM = 160;            // positions in front of a run inserted into its table before the first step
constexpr int MIN_MATCH = 6;         // = the hashed bytes
constexpr int MAX_MATCH = 258;
constexpr int TABLE = 512;           // entries per warp (u16 positions within the member; 0xFFFF = empty)
constexpr int REGION = 1792;         // bytes of shared memory that stage the codes of one run
constexpr int N_LITLEN = 286, N_DIST = 30, N_SYM = N_LITLEN + N_DIST;
constexpr int SAMPLE_MEMBERS = 96;   // the code is built from the symbol counts of every stride-th member, stride = ceil(members / 96)

GG_LZ_HD uint32_t lz_hash(uint32_t lo /* bytes 0..3 */, uint32_t hi16 /* bytes 4,5 */) { return (lo * 0x9E3779B1u + hi16 * 0x85EBCA77u) >> 23; }
GG_LZ_HD int ilog2(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }      // (the kernel uses clz)

// length 3..258 -> literal/length symbol, number of extra bits, extra value (RFC 1951 3.2.5)
GG_LZ_HD void len_symbol(int len, int lg /* floor(log2(len - 3)), any value when len - 3 < 8 */, int* sym, int* nx, uint32_t* xv) {
    const uint32_t y = (uint32_t)(len - 3);
    if (y < 8u) { *sym = 257 + (int)y; *nx = 0; *xv = 0; }
    else if (len == 258) { *sym = 285; *nx = 0; *xv = 0; }
    else { *nx = lg - 2; *sym = 257 + 4 * (lg - 1) + (int)((y >> (lg - 2)) & 3u); *xv = y & ((1u << (lg - 2)) - 1u); }
}
// distance 1..32768 -> distance symbol, extra bits, extra value
GG_LZ_HD void dist_symbol(int dist, int lg /* floor(log2(dist - 1)), any value when dist - 1 < 4 */, int* sym, int* nx, uint32_t* xv) {
    const uint32_t x = (uint32_t)(dist - 1);
    if (x < 4u) { *sym = (int)x; *nx = 0; *xv = 0; }
    else { *nx = lg - 1; *sym = 2 * lg + (int)((x >> (lg - 1)) & 1u); *xv = x & ((1u << (lg - 1)) - 1u); }
}
} // namespace gglz
