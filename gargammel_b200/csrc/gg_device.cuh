// gg_device.cuh -- device-side building blocks shared by the sm_100a kernels:
// Philox4x32-10 counter RNG, 2-bit genome access, per-base damage models.
// Draw layout and threshold semantics: DESIGN.md "Draw specification".
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace ggd {

// stream ids (Philox counter word 3)
enum { ST_FRAG = 0, ST_DEAM = 1, ST_ADPT_F = 2, ST_DEAM2 = 3, ST_ADPT_R = 4, ST_FRAG2 = 5, ST_GCPROBE = 6, ST_FRAG3 = 7 };

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
struct Key { uint32_t k0, k1; };
__device__ __forceinline__ uint4 rng_block(const Key& k, uint64_t id, uint32_t blk, uint32_t stream) {
    return philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), blk, stream, k.k0, k.k1);
}
__device__ __forceinline__ uint32_t pick(const uint4& v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}
__device__ __forceinline__ uint32_t rng_word(const Key& k, uint64_t id, uint32_t stream, uint32_t w) {
    return pick(rng_block(k, id, w >> 2, stream), (int)(w & 3));
}

// ---------------------------------------------------------------- 2-bit sequence helpers (A=0 C=1 G=2 T=3; 16 bases per u32, base k at bits 2k)
// 16 bases starting at base index `pos` of a packed array (reads words pos/16 and pos/16+1)
__device__ __forceinline__ uint32_t fetch16(const uint32_t* __restrict__ w, int64_t pos) {
    const int64_t wi = pos >> 4;
    const uint32_t sh = ((uint32_t)pos & 15u) * 2u;
    return __funnelshift_r(__ldg(w + wi), __ldg(w + wi + 1), sh);
}
__device__ __forceinline__ uint32_t fetch16_smem(const uint32_t* w, int pos) {
    const int wi = pos >> 4;
    const uint32_t sh = ((uint32_t)pos & 15u) * 2u;
    return __funnelshift_r(w[wi], w[wi + 1], sh);
}
// reverse the order of the sixteen 2-bit fields and complement each (A<->T, C<->G is 3-v == ~v)
__device__ __forceinline__ uint32_t revcomp16(uint32_t x) {
    uint32_t y = __brev(x);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
}
// number of G/C among W bases starting at base index p0 (G = 10b, C = 01b: exactly one of the two bits set)
__device__ __forceinline__ int gc_count(const uint32_t* __restrict__ seq2, uint64_t p0, int W) {
    int c = 0; uint64_t p = p0; const uint64_t end = p0 + (uint64_t)W;
#pragma unroll 1
    while (p < end) {
        const uint32_t w = __ldg(seq2 + (p >> 4));
        const int o = (int)(p & 15), n = (int)min((uint64_t)(16 - o), end - p);
        uint32_t x = ((w ^ (w >> 1)) & 0x55555555u) >> (2 * o);
        if (n < 16) x &= (1u << (2 * n)) - 1u;
        c += __popc(x); p += (uint64_t)n;
    }
    return c;
}
// mask with the low n 2-bit fields set, n in [0,16]
// (PTX shl.b32 clamps shift amounts above 31, so n == 16 gives (1<<32)-1 == 0xFFFFFFFF without a compare)
__device__ __forceinline__ uint32_t lowfields(int n) {
    uint32_t r; asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(2 * n)); return r - 1u;
}

// ---------------------------------------------------------------- damage model tables (device view)
struct DamageDev {
    int mode;            // gg_damage_mode
    int rows5, rows3;    // matrix / single-double rows
    int clamp_last;      // matrix: clamp to last row (default) or skip (-last)
    // MATRIX: thr[(r*4+b)] = {K1,K2,K3,K4}: new base = first a with u31 < K[a+1] (K4 == 2^31 always, checked on the host).
    // Layout: rows5 real 5' rows, one sentinel row (= last row when clamping, identity when -last), then the 3' block REVERSED:
    // its sentinel, then distance rows3-1 .. 0, so that the row of position i on the 3' side is max(i + rows5+rows3+2-L, rows5+1).
    const uint4* thr;
    // SINGLE/DOUBLE: sd5[r] (C->T 5'), sd3[r] (C->T single / G->A double)
    const uint32_t* sd5; const uint32_t* sd3;
    // BRIGGS: Ks,Kd strict thresholds (u31 < K); geometric tables (u31 < cdf_thr[k] first k) of n_geo entries
    uint32_t Ks, Kd; const uint32_t* geoL; const uint32_t* geoV; int n_geo;
    // BRIGGS_SS (-damagess d,s5,s3,l5,l3): Ks = 5' overhang rate, K3 = 3' overhang rate, geoL = 5' lengths, geoL3 = 3' lengths
    // (plain Briggs: K3 == Ks, geoL3 == geoL)
    uint32_t K3; const uint32_t* geoL3;
    // 1 / ln(1 - p) of the three geometric tables: starting guess of the inverse-CDF search (the table decides)
    float inv_logL, inv_logL3, inv_logV;
};

// first k with u < thr[k], thr ascending; returns n if none
__device__ __forceinline__ int first_below(const uint32_t* __restrict__ thr, int n, uint32_t u) {
    int lo = 0, hi = n;            // invariant: answer in [lo,hi]
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < __ldg(thr + mid)) hi = mid; else lo = mid + 1; }
    return lo;
}

// same result, for draws that are usually tiny (geometric tables): probe the first entries, then gallop, then bisect
__device__ __forceinline__ int first_below_gallop(const uint32_t* __restrict__ thr, int n, uint32_t u) {
#pragma unroll 1
    for (int k = 0; k < 4 && k < n; k++) if (u < __ldg(thr + k)) return k;
    int lo = 4, hi = 8;
#pragma unroll 1
    while (hi < n && u >= __ldg(thr + hi)) { lo = hi + 1; hi <<= 1; }
    hi = min(hi, n);
    if (lo > hi) return hi;
#pragma unroll 1
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < __ldg(thr + mid)) hi = mid; else lo = mid + 1; }
    return lo;
}

// -matfile, one base (deamSim.cpp:1794-1930): i = position, L = fragment length, b = code, x = raw Philox word
__device__ __forceinline__ uint32_t dmg_matrix_base(const DamageDev& D, int i, int L, uint32_t b, uint32_t x) {
    const int r = (i < (L >> 1)) ? min(i, D.rows5) : max(i + (D.rows5 + D.rows3 + 2 - L), D.rows5 + 1);   // sentinel rows absorb the clamp / -last cases
    const uint4 T = __ldg(D.thr + (r * 4 + (int)b));
    const uint32_t u = x >> 1;
    return (uint32_t)(u >= T.x) + (uint32_t)(u >= T.y) + (uint32_t)(u >= T.z);
}
// same, returning the new base already placed in its 2-bit field: (new base) * unit with unit = 1 << 2j
__device__ __forceinline__ uint32_t dmg_matrix_base_shifted(const DamageDev& D, int i, int L, uint32_t b, uint32_t x, uint32_t unit) {
    const int r = (i < (L >> 1)) ? min(i, D.rows5) : max(i + (D.rows5 + D.rows3 + 2 - L), D.rows5 + 1);
    const uint4 T = __ldg(D.thr + (r * 4 + (int)b));
    const uint32_t u = x >> 1;
    uint32_t v = 0;
    if (u >= T.x) v += unit;
    if (u >= T.y) v += unit;
    if (u >= T.z) v += unit;
    return v;
}
// Inverse CDF of a geometric table thr[k] = thr_le(1 - q^(k+1)): first k with u < thr[k] (n if none).  The closed form
// k = ceil(ln(1-p) / ln q) - 1 gives a starting point in one step; the integer table then decides, so the result is exactly
// first_below's (two independent loads instead of a dependent chain of 4-12).
__device__ __forceinline__ int first_below_geo(const uint32_t* __restrict__ thr, int n, uint32_t u, float inv_logq) {
    const float omp = (float)(2147483648u - u) * 4.656612873e-10f;      // 1 - p, exact for the small values where it matters
    float x = ceilf(__logf(omp) * inv_logq) - 1.0f;
    x = fminf(fmaxf(x, 0.0f), (float)n);                                // NaN (q = 0 or 1) -> 0
    int k = (int)x;
    const uint32_t below = k > 0 ? __ldg(thr + k - 1) : 0u, here = k < n ? __ldg(thr + k) : 0xFFFFFFFFu;
    if (u >= below && u < here) return k;
#pragma unroll 1
    while (k > 0 && u < __ldg(thr + k - 1)) k--;
#pragma unroll 1
    while (k < n && u >= __ldg(thr + k)) k++;
    return k;
}
// Briggs per-fragment header draws (deamSim.cpp:1478-1483,1503-1507)
struct BriggsHdr { int o5, o3, nick; };
__device__ __forceinline__ BriggsHdr dmg_briggs_hdr(const DamageDev& D, const uint4& h) {
    BriggsHdr B;
    B.o5 = first_below_geo(D.geoL, D.n_geo, h.x >> 1, D.inv_logL);
    B.o3 = first_below_geo(D.geoL3, D.n_geo, h.y >> 1, D.inv_logL3);
    B.nick = first_below_geo(D.geoV, D.n_geo, h.z >> 1, D.inv_logV);
    return B;
}
// -damage, one base (deamSim.cpp:1483-1594), branch-free:
//   C->T at s  if all single-stranded, or inside the 5' overhang            (:1486-1494, :1513-1523, :1557-1567)
//   G->A at s  inside the 3' overhang                                       (:1527-1537, :1571-1581)
//   G->A at d  in the double-stranded part at/after the nick, C->T at d before it   (:1542-1548, :1586-1592)
//   -damagess (mode 5, deamSim.cpp:1337-1468): C->T only; 5' overhang at s5, 3' overhang at s3, double-stranded part at d; a base
//   inside both overhangs takes the nearer end (i < int(len*0.5) -> 5', :1351-1371).  SS = false compiles that model out.
template <bool SS>
__device__ __forceinline__ uint32_t dmg_briggs_base(const DamageDev& D, const BriggsHdr& B, int i, int L, uint32_t b, uint32_t x) {
    const uint32_t u = x >> 1;
    const bool in5 = i + 1 <= B.o5, in3 = L - i <= B.o3;
    if (SS && D.mode == 5) {
        const bool s5 = (in5 && in3) ? (i < (L >> 1)) : in5;
        const uint32_t K = s5 ? D.Ks : in3 ? D.K3 : D.Kd;
        return (b == 1u && u < K) ? 3u : b;
    }
    const bool ss5 = (B.o5 + B.o3 >= L) || in5;
    const bool ss3 = !ss5 && in3;
    const bool ct = ss5 || (!ss3 && B.nick > i);
    const uint32_t K = (ss5 || ss3) ? D.Ks : D.Kd;
    const uint32_t from = ct ? 1u : 2u, to = ct ? 3u : 0u;
    return (b == from && u < K) ? to : b;
}
// -mat single|double / -mapdamage, one base, both passes (deamSim.cpp:1941-2002); x1 pass-1 word, x2 pass-2 word
__device__ __forceinline__ uint32_t dmg_sd_base(const DamageDev& D, int i, int L, uint32_t b, uint32_t x1, uint32_t x2) {
    const int r5 = min(i, D.rows5 - 1), r3 = min(L - 1 - i, D.rows3 - 1);
    if (b == 1u && (x1 >> 1) < __ldg(D.sd5 + r5)) b = 3u;                  // pass 1: C->T by 5' row
    if (D.mode == 3 /*SINGLE*/) { if (b == 1u && (x2 >> 1) < __ldg(D.sd3 + r3)) b = 3u; }
    else                        { if (b == 2u && (x2 >> 1) < __ldg(D.sd3 + r3)) b = 0u; }
    return b;
}

// ASCII <-> code
__device__ __forceinline__ int ascii2code(uint8_t c) {   // upper-case input; -1 = unresolved
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}
__device__ __forceinline__ uint8_t code2ascii(uint32_t b) { return (uint8_t)(0x54474341u >> (8 * b)); }
__device__ __forceinline__ uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
__device__ __forceinline__ uint8_t complement_ascii(uint8_t c) {
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
}

static __constant__ uint32_t c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
// decimal digits
// (coordinates are < 2^33: contig lengths are u32 and fragment lengths <= 4095)
__device__ __forceinline__ int ndigits(uint64_t v) {
    if (v <= 0xFFFFFFFFull) {                      // floor(log10) from the bit length, corrected by one table compare
        const uint32_t w = (uint32_t)v;
        const uint32_t t = ((32u - (uint32_t)__clz((int)(w | 1u))) * 1233u) >> 12;      // 0..9
        const uint32_t p10 = c_pow10[t];
        return (int)(t + 1u - (w < p10 ? 1u : 0u)) + (w == 0u ? 1 : 0);
    }
    return v >= 10000000000ull ? 11 : 10;          // 2^32 .. 2^33: 10 digits below 1e10
}

} // namespace ggd
