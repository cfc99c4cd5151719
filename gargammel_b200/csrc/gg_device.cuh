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
// The same generator with its ten round keys precomputed on the host.  As a field of a __grid_constant__ kernel parameter the keys are
// constant-bank operands of the XORs: the two key additions per round (20 of ~60 instructions per block) disappear.
struct KeySched { uint32_t a[10], b[10]; };
__host__ __device__ inline KeySched make_key_sched(uint32_t k0, uint32_t k1) {
    KeySched K;
    for (int r = 0; r < 10; r++) { K.a[r] = k0 + (uint32_t)r * 0x9E3779B9u; K.b[r] = k1 + (uint32_t)r * 0xBB67AE85u; }
    return K;
}
__device__ __forceinline__ uint4 rng_block(const KeySched& K, uint64_t id, uint32_t blk, uint32_t stream) {
    uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = blk, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.a[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.b[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t pick(const uint4& v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}
template <class KeyT>
__device__ __forceinline__ uint32_t rng_word(const KeyT& k, uint64_t id, uint32_t stream, uint32_t w) {
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
    int rows5, rows3;    // matrix / single-double rows (METH: the larger row count of the two tables of an end)
    int clamp_last;      // matrix: clamp to last row (default) or skip (-last)
    // MATRIX (draw specification v2: events).  Thresholds are uint4 per (row, base):
    //   term[(t)*4 + b], t < n5t: categorical of the 5' terminal position t; term[(n5t + t)*4 + b], t < n3t: 3' terminal position L-1-t.
    //     {K1,K2,K3,K4}: new base = number of K <= u31 (K4 == 2^31 always, checked on the host); rows past a -last table are identities.
    //   acc[(r)*4 + b]: a candidate of the interior takes the k-th other base (ascending) iff K_k <= u31 < K_{k+1} ... i.e. the number
    //     of K <= u31 indexes {a1,a2,a3,unchanged}.  Row layout: rows5 real 5' rows, one sentinel row (= last row when clamping,
    //     "unchanged" with -last), then the 3' block REVERSED: its sentinel, then distance rows3-1 .. 0, so that the row of position i
    //     on the 3' side is max(i + rows5+rows3+2-L, rows5+1).
    //   geoM: geometric skips under the majorant M (inverse CDF, like geoL/geoV).
    const uint4* term; const uint4* acc; const uint32_t* geoM; int n5t, n3t;
    // SINGLE/DOUBLE: sd5[r] (C->T 5'), sd3[r] (C->T single / G->A double)
    const uint32_t* sd5; const uint32_t* sd3;
    // BRIGGS: geometric tables (u31 < cdf_thr[k] first k) of n_geo entries: overhang lengths, nick position, and the skips of the
    // event draws over the single-stranded (rate s) and double-stranded (rate d) positions
    const uint32_t* geoL; const uint32_t* geoV; const uint32_t* geoS; const uint32_t* geoD; int n_geo;
    const uint8_t* guideA; const uint8_t* guideB;      // 512-entry guides of geoM / geoS (A) and geoD (B)
    // BRIGGS_SS (-damagess d,s5,s3,l5,l3; still drawn per base): Ks = 5' overhang rate, K3 = 3' overhang rate, Kd = middle, geoL = 5' lengths,
    // geoL3 = 3' lengths (plain Briggs: geoL3 == geoL)
    uint32_t Ks, Kd, K3; const uint32_t* geoL3;
    // 1 / ln(1 - p) of the geometric tables: starting guess of the inverse-CDF search (the table decides)
    float inv_logL, inv_logL3, inv_logV;
};
#define GG_GEO_MISS (1 << 29)     // an overhang / nick draw that ran off its table: beyond any record

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
// The same inverse CDF for the hot draws (the skips of the event draws): guide[u31 >> 22] = first k whose threshold lies above the cell's
// lower edge, then a short upward scan; one byte load + (usually) one threshold load.  Entries of 255 mean "255 or later".  The scan
// stops at lim <= n (the positions that are left): a result of lim means "lim or more".
// SM: thr / guide are shared-memory copies (plain loads; a skip is two DEPENDENT loads and a fragment draws several skips in a row, so with
// the tables in global memory -- L1 is a few KB next to a full shared-memory carve-out -- the event pass waits on L2 most of the time)
template <bool SM = false>
__device__ __forceinline__ int first_below_guided(const uint32_t* __restrict__ thr, const uint8_t* __restrict__ guide, int lim, uint32_t u) {
    int k = SM ? (int)guide[u >> 22] : (int)__ldg(guide + (u >> 22));
#pragma unroll 1
    while (k < lim && u >= (SM ? thr[k] : __ldg(thr + k))) k++;
    return min(k, lim);
}
// Briggs per-fragment header draws (deamSim.cpp:1478-1483,1503-1507)
struct BriggsHdr { int o5, o3, nick; };
__device__ __forceinline__ BriggsHdr dmg_briggs_hdr(const DamageDev& D, const uint4& h) {
    BriggsHdr B;
    B.o5 = first_below_geo(D.geoL, D.n_geo, h.x >> 1, D.inv_logL);
    B.o3 = first_below_geo(D.geoL3, D.n_geo, h.y >> 1, D.inv_logL3);
    B.nick = first_below_geo(D.geoV, D.n_geo, h.z >> 1, D.inv_logV);
    if (B.o5 == D.n_geo) B.o5 = GG_GEO_MISS;                            // off the table: longer than any record
    if (B.o3 == D.n_geo) B.o3 = GG_GEO_MISS;
    if (B.nick == D.n_geo) B.nick = GG_GEO_MISS;
    return B;
}
// -damagess, one base (deamSim.cpp:1337-1468; still drawn per base, word 4+i): C->T only; 5' overhang at s5, 3' overhang at s3,
// double-stranded part at d; a base inside both overhangs takes the nearer end (i < int(len*0.5) -> 5', :1351-1371).
__device__ __forceinline__ uint32_t dmg_briggs_ss_base(const DamageDev& D, const BriggsHdr& B, int i, int L, uint32_t b, uint32_t x) {
    const uint32_t u = x >> 1;
    const bool in5 = i + 1 <= B.o5, in3 = L - i <= B.o3;
    const bool s5 = (in5 && in3) ? (i < (L >> 1)) : in5;
    const uint32_t K = s5 ? D.Ks : in3 ? D.K3 : D.Kd;
    return (b == 1u && u < K) ? 3u : b;
}
// ---------------------------------------------------------------- damage as events (draw specification v2)
// The reference draws one uniform per base (deamSim.cpp:1847,1488-1587).  Per-base draws are independent, so the same law comes
// from thinning: geometric skips under a majorant find the few candidate positions of a fragment, and a candidate with base b
// becomes a with probability rate(b->a)/majorant.  A fragment then costs a handful of Philox words instead of one per base.
// SeqT gives access to the bases of one fragment / record:  code(i) -> 0..3 or -1 (unresolved), put(i, old, new).
// number of thresholds <= u among T.x <= T.y <= T.z (ascending by construction: cumulative sums): a two-level select instead of three compare-and-adds
__device__ __forceinline__ uint32_t count_le(uint32_t u, const uint4& T) { return u >= T.y ? (u >= T.z ? 3u : 2u) : (u >= T.x ? 1u : 0u); }

// matrix row (acc layout) of position i of a fragment of length L
__device__ __forceinline__ int matrix_row(const DamageDev& D, int i, int L) {
    return (i < (L >> 1)) ? min(i, D.rows5) : max(i + (D.rows5 + D.rows3 + 2 - L), D.rows5 + 1);
}
// PH: bit 0 = 5' terminal positions, bit 1 = interior, bit 2 = 3' terminal positions (deamSim -name walks them in this order);
// TSM: term/acc are shared-memory copies (plain loads) instead of global tables
// GSM: geoM / guideA point at shared-memory copies of the skip table (its first entries: a skip never exceeds the record) and its guide
template <int PH, bool TSM, bool GSM, class SeqT, class KeyT>
__device__ __forceinline__ void matrix_events(const KeyT& key, const DamageDev& D, const uint4* term, const uint4* acc, const uint32_t* geoM, const uint8_t* guideA,
                                              SeqT& S, int L, uint64_t id) {
    const int half = L >> 1;                                    // positions i < half are on the 5' side (deamSim.cpp:1805)
    const int n5 = min(D.n5t, half), n3 = min(D.n3t, L - half);
    const uint32_t b3 = (uint32_t)(D.n5t + 3) >> 2, bg = b3 + ((uint32_t)(D.n3t + 3) >> 2);
    if (PH == 7) {
        // both ends in one loop (the order of the positions does not matter when nobody lists the events): step s < n5 is the 5' terminal
        // position s, step n5 + t the 3' terminal position L-1-t; a block is drawn when the step enters it.  One copy of the generator and
        // of the categorical draw instead of eight: the kernel's hot code has to fit the 32 KB instruction cache.
        uint4 x = make_uint4(0, 0, 0, 0); uint32_t have = 0xFFFFFFFFu;
#pragma unroll 1
        for (int s = 0; s < n5 + n3; s++) {
            const bool five = s < n5;
            const int t = five ? s : s - n5;
            const uint32_t blk = (five ? 0u : b3) + ((uint32_t)t >> 2);
            if (blk != have) { x = rng_block(key, id, blk, ST_DEAM); have = blk; }
            const int i = five ? t : L - 1 - t;
            const int b = S.code(i);
            if (b >= 0) { const int ti = ((five ? 0 : D.n5t) + t) * 4 + b; const uint4 T = TSM ? term[ti] : __ldg(term + ti); S.put(i, (uint32_t)b, count_le(pick(x, t & 3) >> 1, T)); }
        }
    }
    if (PH != 7 && (PH & 1)) {
#pragma unroll 1
        for (int t0 = 0; t0 < n5; t0 += 4) {
            const uint4 x = rng_block(key, id, (uint32_t)t0 >> 2, ST_DEAM);
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int t = t0 + jj;
                if (t < n5) {
                    const int b = S.code(t);
                    if (b >= 0) { const uint4 T = TSM ? term[t * 4 + b] : __ldg(term + t * 4 + b); S.put(t, (uint32_t)b, count_le(pick(x, jj) >> 1, T)); }
                }
            }
        }
    }
    if (PH & 2) {
        int pos = n5; const int end = L - n3;
#pragma unroll 1
        for (uint32_t n = 0; pos < end; n += 2) {
            const uint4 x = rng_block(key, id, bg + (n >> 1), ST_DEAM);
#pragma unroll
            for (int sub = 0; sub < 2; sub++) {
                if (pos < end) {
                    const int lim = min(D.n_geo, end - pos);    // a skip of lim or more ends the walk (or, off a table shorter than the record, goes on from there)
                    const int j = first_below_guided<GSM>(geoM, guideA, lim, (sub ? x.z : x.x) >> 1);
                    pos += j;
                    if (j < lim) {
                        const int b = S.code(pos);
                        if (b >= 0) {
                            const int r = matrix_row(D, pos, L);
                            const uint4 A = TSM ? acc[r * 4 + b] : __ldg(acc + r * 4 + b);
                            const uint32_t cnt = count_le((sub ? x.w : x.y) >> 1, A);            // 0,1,2: the cnt-th other base (ascending); 3: unchanged
                            if (cnt < 3u) S.put(pos, (uint32_t)b, cnt + (cnt >= (uint32_t)b ? 1u : 0u));
                        }
                        pos++;
                    }
                }
            }
        }
    }
    if (PH != 7 && (PH & 4)) {
#pragma unroll 1
        for (int t0 = ((n3 - 1) >> 2) << 2; t0 >= 0; t0 -= 4) {     // t descending = position ascending
            const uint4 x = rng_block(key, id, b3 + ((uint32_t)t0 >> 2), ST_DEAM);
#pragma unroll
            for (int jj = 3; jj >= 0; jj--) {
                const int t = t0 + jj;
                if (t < n3) {
                    const int i = L - 1 - t;
                    const int b = S.code(i);
                    if (b >= 0) { const uint4 T = TSM ? term[(D.n5t + t) * 4 + b] : __ldg(term + (D.n5t + t) * 4 + b); S.put(i, (uint32_t)b, count_le(pick(x, jj) >> 1, T)); }
                }
            }
        }
    }
}

// Briggs (deamSim.cpp:1477-1613) as events.  Block 0 = header (overhangs, nick).  Block 1+n: words 0,2 = geometric skips (rate s) over
// the single-stranded positions (5' overhang then 3' overhang; everything when the overhangs meet), words 1,3 = skips (rate d) over the
// double-stranded middle.  A candidate changes iff its base is the one its region deaminates: C->T in the 5' overhang / before the
// nick, G->A in the 3' overhang / from the nick on; both are "flip the high bit of the code".
// PH: bit 0 = 5' overhang (or everything, when all single-stranded), bit 1 = middle, bit 2 = 3' overhang.
// GSM: geoS / guideA / geoD / guideB point at shared-memory copies (see matrix_events)
template <int PH, bool GSM, class SeqT, class KeyT>
__device__ __forceinline__ void briggs_events(const KeyT& key, const DamageDev& D, const BriggsHdr& B, const uint32_t* geoS, const uint8_t* guideA,
                                              const uint32_t* geoD, const uint8_t* guideB, SeqT& S, int L, uint64_t id) {
    const bool allss = B.o5 + B.o3 >= L;                            // deamSim.cpp:1483
    const int nS = allss ? L : B.o5 + B.o3, qend = allss ? L : L - B.o3;
    int v = (PH & 5) ? 0 : nS, q = (PH & 2) ? (allss ? L : B.o5) : qend;
#pragma unroll 1
    for (uint32_t n = 0; v < nS || q < qend; n++) {
        const uint4 x = rng_block(key, id, 1u + n, ST_DEAM);
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            if ((PH & 5) && v < nS) {
                const int lim = min(D.n_geo, nS - v);
                const int j = first_below_guided<GSM>(geoS, guideA, lim, (sub ? x.z : x.x) >> 1);
                v += j;
                if (j < lim) {
                    const bool five = allss || v < B.o5;
                    const int i = five ? v : L - B.o3 + (v - B.o5);
                    if (five ? (PH & 1) : (PH & 4)) { if (S.code(i) == (five ? 1 : 2)) S.put(i, five ? 1u : 2u, five ? 3u : 0u); }
                    v++;
                }
            }
            if ((PH & 2) && q < qend) {
                const int lim = min(D.n_geo, qend - q);
                const int j = first_below_guided<GSM>(geoD, guideB, lim, (sub ? x.w : x.y) >> 1);
                q += j;
                if (j < lim) {
                    const bool ct = B.nick > q;                     // placedNick = nick <= i (deamSim.cpp:1510,1542,1586)
                    if (S.code(q) == (ct ? 1 : 2)) S.put(q, ct ? 1u : 2u, ct ? 3u : 0u);
                    q++;
                }
            }
        }
    }
}

// 2-bit packed fragment in shared memory (16 bases per word)
struct SeqPacked {
    uint32_t* w;
    __device__ __forceinline__ int code(int i) const { return (int)((w[i >> 4] >> (2 * (i & 15))) & 3u); }
    __device__ __forceinline__ void put(int i, uint32_t oldb, uint32_t newb) { w[i >> 4] ^= (oldb ^ newb) << (2 * (i & 15)); }
};

// -mat single|double / -mapdamage, one base, both passes (deamSim.cpp:1941-2002); x1 pass-1 word, x2 pass-2 word
__device__ __forceinline__ uint32_t dmg_sd_base(const DamageDev& D, int i, int L, uint32_t b, uint32_t x1, uint32_t x2) {
    const int r5 = min(i, D.rows5 - 1), r3 = min(L - 1 - i, D.rows3 - 1);
    if (b == 1u && (x1 >> 1) < __ldg(D.sd5 + r5)) b = 3u;                  // pass 1: C->T by 5' row
    if (D.mode == 3 /*SINGLE*/) { if (b == 1u && (x2 >> 1) < __ldg(D.sd3 + r3)) b = 3u; }
    else                        { if (b == 2u && (x2 >> 1) < __ldg(D.sd3 + r3)) b = 0u; }
    return b;
}

// -matfilemeth / -matfilenonmeth, one base (deamSim.cpp:1630-1787; drawn per base, word i): bm = base code | 4 for a methylated
// (lower-case) base.  term[(row)*8 + bm], rows = rows5 5' rows then rows3 3' rows (by distance from the 3' end), holds the reference's
// categorical thresholds; w == 0 marks a position without a row in that table (-last): the byte stays as it is (returns -1).
__device__ __forceinline__ int dmg_meth_base(const DamageDev& D, int i, int L, int bm, uint32_t x) {
    const bool five = i < (L >> 1);                                        // deamSim.cpp:1661
    int dist = five ? i : L - 1 - i;
    const int R = five ? D.rows5 : D.rows3;
    if (dist >= R) { if (!D.clamp_last) return -1; dist = R - 1; }
    const uint4 T = __ldg(D.term + ((five ? 0 : D.rows5) + dist) * 8 + bm);
    if (T.w == 0u) return -1;
    return (int)count_le(x >> 1, T);
}

// ASCII <-> code
__device__ __forceinline__ int ascii2code(uint8_t c) {   // upper-case input; -1 = unresolved
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}
__device__ __forceinline__ uint8_t code2ascii(uint32_t b) { return (uint8_t)(0x54474341u >> (8 * b)); }
__device__ __forceinline__ uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
__device__ __forceinline__ uint8_t complement_ascii(uint8_t c) {    // lower-case letters complement within their case (raw adapters given in lower case)
    const uint8_t u = c & 0xDF, lc = c & 0x20;
    const uint8_t r = u == 'A' ? 'T' : u == 'C' ? 'G' : u == 'G' ? 'C' : u == 'T' ? 'A' : 0;
    return (r && ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) ? (uint8_t)(r | lc) : 'N';
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
