// gg_kernels.cu -- hand-written sm_100a kernels of the gargammel hot path.
//
//   k_pack_genome   FASTA bytes -> 2-bit codes + non-ACGT mask (replaces mmap + per-char fetchSeq, fragSim.cpp:305-329)
//   k_fused         fragSim -> deamSim -> adptSim, ONE kernel of persistent, autonomous warps.  Each warp takes tiles of 32
//                   fragments: lane-per-fragment rejection sampling (fragSim.cpp:1361-1507), lane-per-16-base-chunk damage
//                   (deamSim.cpp:1477-1613,1794-2002), lane-per-fragment 2-bit line image + lane-per-16-byte-group ASCII
//                   (adptSim.cpp:281-411), records staged in the warp's shared memory, global byte offsets from a decoupled
//                   look-back over warp tiles, and 16-byte vector stores of the final FASTA byte stream.
//   k_deam_records / k_adpt_records   the stand-alone stages over a 2-line FASTA record stream (for the drop-in CLIs)
//
// HBM-bound integer/byte work: no tensor cores.  Layout and byte accounting: DESIGN.md.
#include "gg_kernels.cuh"
#ifndef GG_SPIN_SLEEP
#define GG_SPIN_SLEEP 0
#endif
#ifndef GG_GATHER_U
#define GG_GATHER_U 2          // 16-base chunks gathered per lane and trip (4 loads in flight; measured: 4 and 6 per trip lose 4 % / 3 % -- the unrolled code pushes the hot path past the 32 KB instruction cache)
#endif
#ifndef GG_COPY_UNROLL
#define GG_COPY_UNROLL 1       // copy-out with two 16-byte copies per trip (measured: -2.3 %; four per trip: -1.4 %)
#endif

namespace ggk {
using namespace ggd;

__device__ __forceinline__ unsigned long long warp_sum_u64_early(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ======================================================================================= pack
// one thread per 32 bases of the packed base space (2 seq2 words + 1 mask word).  Descriptor `wrap_desc` (if >= 0) is the
// --circ wrap image of a contig of `wrap_srclen` bases: image base j is contig base (srclen-1-keff+j) for j < keff and
// (j-keff) behind it (fragSim.cpp:1413-1422: temp1 stops one base short of the contig end, temp2 restarts at 0).
__global__ void __launch_bounds__(256) k_pack_genome(const uint8_t* __restrict__ fasta, uint64_t n,
        const uint64_t* __restrict__ c_off, const uint32_t* __restrict__ c_len,
        const int32_t* __restrict__ c_blen, const int32_t* __restrict__ c_llen,
        const uint64_t* __restrict__ c_gbase, int n_contigs, int wrap_desc, uint32_t wrap_keff, uint32_t wrap_srclen, uint64_t n_units,
        uint32_t* __restrict__ seq2, uint32_t* __restrict__ nmask, uint32_t* __restrict__ cmask /* fragSim --case: 1 = lower-case; may be null */) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    uint32_t lc = 0;
    const uint64_t base0 = u * 32ull;
    int lo = 0, hi = n_contigs;                       // last contig with gbase <= base0
    while (lo < hi) { int mid = (lo + hi) >> 1; if (c_gbase[mid] <= base0) lo = mid + 1; else hi = mid; }
    const int c = lo - 1;
    uint32_t w0 = 0, w1 = 0, m = 0xFFFFFFFFu;
    if (c >= 0) {
        const uint64_t local = base0 - c_gbase[c];
        const uint64_t len = c_len[c];
        if (local < len) {
            const uint64_t blen = (uint64_t)c_blen[c], llen = (uint64_t)c_llen[c];
            m = 0;
            if (c == wrap_desc) {
                for (int t = 0; t < 32; t++) {
                    uint32_t code = 0, bad = 1;
                    const uint64_t j = local + t;
                    if (j < len) {
                        const uint64_t src = j < wrap_keff ? (uint64_t)wrap_srclen - 1 - wrap_keff + j : j - wrap_keff;
                        if (src < wrap_srclen) {
                            const uint64_t pos = c_off[c] + src / blen * llen + src % blen;
                            const uint8_t raw = pos < n ? fasta[pos] : 0;
                            const int cd = ascii2code(upper(raw));
                            if (cd >= 0) { code = (uint32_t)cd; bad = 0; lc |= (raw >= 'a' ? 1u : 0u) << t; }
                        }
                    }
                    if (t < 16) w0 |= code << (2 * t); else w1 |= code << (2 * (t - 16));
                    m |= bad << t;
                }
            } else {
                uint64_t line = local / blen, col = local % blen;
                uint64_t pos = c_off[c] + line * llen + col;     // fetchSeq, fragSim.cpp:312-314
                for (int t = 0; t < 32; t++) {
                    uint32_t code = 0, bad = 1;
                    if (local + t < len) {
                        const uint8_t raw = pos < n ? fasta[pos] : 0;
                        const int cd = ascii2code(upper(raw));                // toupper, fragSim.cpp:315
                        if (cd >= 0) { code = (uint32_t)cd; bad = 0; lc |= (raw >= 'a' ? 1u : 0u) << t; }
                        pos++; col++;
                        if (col == blen) { col = 0; pos += llen - blen; }
                    }
                    if (t < 16) w0 |= code << (2 * t); else w1 |= code << (2 * (t - 16));
                    m |= bad << t;
                }
            }
        }
    }
    seq2[2 * u] = w0; seq2[2 * u + 1] = w1; nmask[u] = m;
    if (cmask) cmask[u] = lc;
}

cudaError_t launch_pack_genome(const uint8_t* fasta, uint64_t n, const uint64_t* c_off, const uint32_t* c_len,
                               const int32_t* c_blen, const int32_t* c_llen, const uint64_t* c_gbase,
                               int n_contigs, int wrap_desc, uint32_t wrap_keff, uint32_t wrap_srclen, uint64_t n_units,
                               uint32_t* seq2, uint32_t* nmask, uint32_t* cmask, cudaStream_t st) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t blocks = (n_units + 255) / 256;
    k_pack_genome<<<(unsigned)blocks, 256, 0, st>>>(fasta, n, c_off, c_len, c_blen, c_llen, c_gbase, n_contigs, wrap_desc, wrap_keff, wrap_srclen, n_units, seq2, nmask, cmask);
    return cudaGetLastError();
}

// ======================================================================================= GC probes (fragSim -gc)
// GC content of 1,000,000 random 50-bp windows (fragSim.cpp:1190-1269): thread per probe, integer sums so that the
// statistics do not depend on the reduction order.
__global__ void __launch_bounds__(256) k_gc_probe(const GenomeDev G, const uint64_t* __restrict__ cum_end, const uint64_t* __restrict__ gbase,
                                                  const uint32_t* __restrict__ clen, int d, Key key, unsigned long long* __restrict__ sums) {
    const uint64_t pidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long s1 = 0, s2 = 0, bad = 0;
    if (pidx < 1000000ull) {
        const int length = 50, W = length + 2 * d;
        const uint64_t* cum = cum_end + G.first_contig;
        bool done = false;
        for (uint32_t attempt = 0; attempt < 65536u && !done; attempt++) {
            const uint4 x = rng_block(key, pidx, attempt, ST_GCPROBE);
            const uint64_t coord = __umul64hi(((uint64_t)x.w << 32) | x.z, G.G);
            if (coord == 0) continue;
            int a = 0, b = G.n_contigs - 1;                            // contig whose span holds coord
            while (a < b) { int mid = (a + b) >> 1; if (coord <= __ldg(cum + mid)) b = mid; else a = mid + 1; }
            const int ci = G.first_contig + a;
            const uint64_t len = __ldg(clen + ci);
            const uint64_t idx = (coord - 1) - (__ldg(cum + a) - len);
            if (idx < (uint64_t)d || idx + (uint64_t)length + (uint64_t)d > len) continue;      // start+d <= coord <= end-length-d+1 (fragSim.cpp:1217-1219)
            const uint64_t s0 = __ldg(gbase + ci) + idx - (uint64_t)d, e = s0 + (uint64_t)W - 1;
            uint32_t any = 0;
            for (uint64_t w = s0 >> 5; w <= (e >> 5); w++) {
                uint32_t m = __ldg(G.nmask + w);
                if (w == (s0 >> 5)) m &= 0xFFFFFFFFu << (s0 & 31);
                if (w == (e >> 5)) m &= 0xFFFFFFFFu >> (31 - (e & 31));
                any |= m;
            }
            if (any) continue;                                         // isResolvedDNAstring (fragSim.cpp:1241-1242)
            const unsigned long long g = (unsigned long long)gc_count(G.seq2, s0, W);
            s1 = g; s2 = g * g; done = true;
        }
        if (!done) bad = 1;
    }
    s1 = warp_sum_u64_early(s1); s2 = warp_sum_u64_early(s2); bad = warp_sum_u64_early(bad);
    if ((threadIdx.x & 31) == 0) { atomicAdd(sums + 0, s1); atomicAdd(sums + 1, s2); if (bad) atomicAdd(sums + 2, bad); }
}
cudaError_t launch_gc_probe(const GenomeDev& G, const uint64_t* cum_end, const uint64_t* gbase, const uint32_t* clen, int dist, Key key,
                            unsigned long long* sums, cudaStream_t st) {
    k_gc_probe<<<(1000000 + 255) / 256, 256, 0, st>>>(G, cum_end, gbase, clen, dist, key, sums);
    return cudaGetLastError();
}

// ======================================================================================= fused
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_global_v4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive warp scan
__device__ __forceinline__ unsigned long long warp_scan_u64(unsigned long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}

// packed scan element: bytes [63:40] | chunks [39:20] | groups [19:0]
#define PK(bytes, chunks, groups) (((unsigned long long)(bytes) << 40) | ((unsigned long long)(chunks) << 20) | (unsigned long long)(groups))
#define PK_BYTES(v)  ((uint32_t)((v) >> 40))
#define PK_CHUNKS(v) ((uint32_t)(((v) >> 20) & 0xFFFFFu))
#define PK_GROUPS(v) ((uint32_t)((v) & 0xFFFFFu))

// meta word: L [11:0] | plus [12] | slot+1 [15:13] | genome [31:16]
#define META(L, plus, slot, gen) ((uint32_t)(L) | ((uint32_t)(plus) << 12) | ((uint32_t)((slot) + 1) << 13) | ((uint32_t)(gen) << 16))
#define META_L(m)    ((int)((m) & 0xFFFu))
#define META_PLUS(m) (((m) >> 12) & 1u)
#define META_SLOT(m) ((int)(((m) >> 13) & 7u) - 1)
#define META_GEN(m)  ((int)((m) >> 16))

__device__ __forceinline__ int emit_lines(int emit) { return emit == 2 /*STDOUT4*/ ? 2 : 1; }
__device__ __forceinline__ int emit_seqlen(int emit, int L, int RL) {
    return (emit <= 1) ? L : (emit == 4 /*ARTP*/ ? 2 * RL : RL);
}
// "_r" suffix on line `line`?   RR: line 0; STDOUT4: line 1  (adptSim.cpp:286)
__device__ __forceinline__ int emit_suffix(int emit, int line) { return (emit == 6 || (emit == 2 && line == 1)) ? 2 : 0; }

// positions p..p+15 of a packed source, p >= -16.  Every source has one readable guard word in front of it, so a
// window that starts before the source simply carries garbage in fields that the caller's run mask removes.
__device__ __forceinline__ uint32_t fetch_run_smem(const uint32_t* src, int p) { return fetch16_smem(src - 1, p + 16); }
__device__ __forceinline__ uint32_t fetch_run_glob(const uint32_t* __restrict__ src, int p) { return fetch16(src - 1, (int64_t)(p + 16)); }
// group-relative mask of [lo,hi) clipped to [0,16)
__device__ __forceinline__ uint32_t run_mask(int lo, int hi) {
    lo = min(max(lo, 0), 16); hi = min(max(hi, 0), 16);
    return lowfields(hi) & ~lowfields(lo);          // empty when hi <= lo
}

// ---- line image: the 2-bit packed sequence line(s) of one record, assembled once per fragment by its owner thread
// (adptSim.cpp:288-322,396-411 restated as a concatenation of runs over 2-bit sources; D = damaged fragment in smem)
struct BitAppender {                  // pending bits: lo holds bits [0,fill) of the current word (fill < 32 between puts), hi one put's spill
    uint32_t* out; uint32_t lo, hi; int fill;
    __device__ __forceinline__ void put(uint32_t x, int n) {        // n fields (1..16); fields of x above n are zero
        hi = __funnelshift_l(x, 0u, (uint32_t)fill); lo |= x << fill; fill += 2 * n;
        if (fill >= 32) { *out++ = lo; lo = hi; fill -= 32; }
    }
    __device__ __forceinline__ void flush() { if (fill > 0) *out++ = lo; }
};
// bases [start, start+len) of a shared-memory source appended as one unaligned bit copy: bit b of the image's current word is source bit
// s + b, s = 2 start - fill (>= -31: the guard word), so every output word is ONE funnel shift of two consecutive source words with the
// same shift amount -- one load, one shift and one store per 16 bases.  Reads at most two words past the last source word it needs.
__device__ __forceinline__ void put_run_smem(BitAppender& A, const uint32_t* src, int start, int len) {
    if (len <= 0) return;
    const int s = 2 * start - A.fill;
    const uint32_t* sp = src + (s >> 5);
    const uint32_t sh = (uint32_t)s & 31u;
    const int bits = A.fill + 2 * len;              // bits [fill, bits) of the word stream that starts at the current word are new
    uint32_t prev = sp[0], next = sp[1];
    uint32_t x = A.lo | (__funnelshift_r(prev, next, sh) & ~((1u << A.fill) - 1u));
    const int nw = bits >> 5;
#pragma unroll 1
    for (int k = 0; k < nw; k++) { *A.out++ = x; prev = next; next = sp[k + 2]; x = __funnelshift_r(prev, next, sh); }
    A.fill = bits & 31;
    A.lo = x & ((1u << A.fill) - 1u);
}
__device__ __forceinline__ void put_run_glob(BitAppender& A, const uint32_t* __restrict__ src, int start, int len) {
#pragma unroll 1
    for (int o = 0; o < len; o += 16) { const int n = min(16, len - o); A.put(fetch_run_glob(src, start + o) & lowfields(n), n); }
}
// first `len` bases of revcomp(D[0:L])
__device__ __forceinline__ void put_rc_run(BitAppender& A, const uint32_t* D, int L, int len) {
#pragma unroll 1
    for (int o = 0; o < len; o += 16) { const int n = min(16, len - o); A.put(revcomp16(fetch_run_smem(D, L - 16 - o)) & lowfields(n), n); }
}
__device__ __forceinline__ uint32_t pad_base(const FusedParams& P, uint64_t id, uint32_t stream, int q) {   // randomDNASeq base q
    return (rng_word(P.key, id, stream, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u;
}
// forward read S1 = (frag + adapterF)[0:RL] + random tail      adptSim.cpp:298-310
// adapter k (0 forward, 1 second, 2 revcomp of the second): from the shared-memory copy when there is one
__device__ __forceinline__ void put_run_ad(BitAppender& A, const FusedParams& P, const uint32_t* adsm, int k, int start, int len) {
    if (adsm) put_run_smem(A, adsm + (P.sm.tab_ad[k] >> 2) + 1, start, len);
    else put_run_glob(A, k == 0 ? P.ad.adF2 : k == 1 ? P.ad.adS2 : P.ad.rcS2, start, len);
}
__device__ __forceinline__ void put_forward_read(BitAppender& A, const FusedParams& P, const uint32_t* adsm, const uint32_t* D, int L, uint64_t id) {
    const int RL = P.ad.read_len, nD = min(L, RL), nA = min(max(RL - L, 0), P.ad.lenF);
    if (A.fill == 0) {                                   // the read opens its line: whole words of D are copied as they are
        const int full = nD >> 4;
#pragma unroll 1
        for (int j = 0; j < full; j++) *A.out++ = D[j];
        if (nD & 15) A.put(D[full] & lowfields(nD & 15), nD & 15);
    } else put_run_smem(A, D, 0, nD);
    put_run_ad(A, P, adsm, 0, 0, nA);
#pragma unroll 1
    for (int q = 0; q < RL - nD - nA; q++) A.put(pad_base(P, id, ST_ADPT_F, q), 1);
}
// reverse read SR = (revcomp(frag) + adapterS)[0:RL] + random tail      adptSim.cpp:289,313-322
__device__ __forceinline__ void put_reverse_read(BitAppender& A, const FusedParams& P, const uint32_t* adsm, const uint32_t* D, int L, uint64_t id) {
    const int RL = P.ad.read_len, nD = min(L, RL), nA = min(max(RL - L, 0), P.ad.lenS);
    put_rc_run(A, D, L, nD);
    put_run_ad(A, P, adsm, 1, 0, nA);
#pragma unroll 1
    for (int q = 0; q < RL - nD - nA; q++) A.put(pad_base(P, id, ST_ADPT_R, q), 1);
}
// revcomp(SR): complemented random tail reversed, suffix of revcomp(adapterS), the last min(L,RL) fragment bases   adptSim.cpp:404
__device__ __forceinline__ void put_reverse_read_rc(BitAppender& A, const FusedParams& P, const uint32_t* adsm, const uint32_t* D, int L, uint64_t id) {
    const int RL = P.ad.read_len, nD = min(L, RL), mAd = RL - nD, nS = min(mAd, P.ad.lenS), nP = mAd - nS;
#pragma unroll 1
    for (int t = 0; t < nP; t++) A.put(3u - pad_base(P, id, ST_ADPT_R, nP - 1 - t), 1);
    put_run_ad(A, P, adsm, 2, P.ad.lenS - nS, nS);
    put_run_smem(A, D, L - nD, nD);
}

// nd decimal digits of v at p; d100 = shared table of the 100 two-digit strings ("00".."99", low byte = tens)
__device__ __forceinline__ uint8_t* put_decimal(uint8_t* p, uint64_t v, int nd, const uint16_t* d100) {
    int k = nd - 1;
#pragma unroll 1
    for (; v > 0xFFFFFFFFull; k--) { const uint64_t q = v / 10ull; p[k] = (uint8_t)('0' + (uint32_t)(v - q * 10ull)); v = q; }
    uint32_t w = (uint32_t)v;
#pragma unroll 1
    for (; k >= 1; k -= 2) { const uint32_t q = w / 100u; const uint32_t two = d100[w - q * 100u]; p[k - 1] = (uint8_t)two; p[k] = (uint8_t)(two >> 8); w = q; }
    if (k == 0) p[0] = (uint8_t)('0' + w);
    return p + nd;
}

// One packed sequence line: kind 0 = forward read, 1 = reverse read, 2 = forward read ++ revcomp(reverse read) (-artp).
// Deliberately not inlined: it is called from two places and the kernel's code size matters (asynchronous warps).
__device__ __noinline__ void build_line_image(const FusedParams& P, const uint32_t* adsm, const uint32_t* D, int L, uint64_t id, int kind, uint32_t* out) {
    BitAppender A; A.out = out; A.lo = 0; A.hi = 0; A.fill = 0;
    if (kind == 1) put_reverse_read(A, P, adsm, D, L, id);
    else {
        put_forward_read(A, P, adsm, D, L, id);
        if (kind == 2) put_reverse_read_rc(A, P, adsm, D, L, id);
    }
    A.flush();
}

// fragSim --comp acceptance (fragSim.cpp:1609-1759): product of 4*dist position-specific base frequencies over the
// fragment ends and their flanks, in the reference's multiplication order and in IEEE fp64 (no contraction), divided by
// the maximum attainable product; accept iff uniform < product.  temp[t] = fragment-orientation window incl. flanks.
__device__ __noinline__ bool comp_accept(const FusedParams& P, const CompDev& C, const uint32_t* __restrict__ seq2, uint64_t gpos, int L,
                                         uint32_t plus, uint64_t id, uint32_t attempt) {
    const int d = C.dist, W = L + 2 * d;
    const uint64_t w0 = gpos - (uint64_t)d;
    auto temp = [&](int t) -> uint32_t {
        const uint64_t p = plus ? w0 + (uint64_t)t : w0 + (uint64_t)(W - 1 - t);
        const uint32_t c = (__ldg(seq2 + (p >> 4)) >> (2 * (p & 15))) & 3u;
        return plus ? c : 3u - c;
    };
    const double* s5 = C.tab + (plus ? 0 : 1) * (2 * d * 4);
    const double* s3 = C.tab + (plus ? 2 : 3) * (2 * d * 4);
    double acc = 1.0;
#pragma unroll 1
    for (int i = 0; i < d; i++) acc = __dmul_rn(acc, __ldg(s5 + i * 4 + temp(i)));                    // 5p before frag
#pragma unroll 1
    for (int i = 0; i < d; i++) acc = __dmul_rn(acc, __ldg(s5 + (i + d) * 4 + temp(d + i)));          // 5p in frag
#pragma unroll 1
    for (int i = 0; i < d; i++) acc = __dmul_rn(acc, __ldg(s3 + i * 4 + temp(L + i)));                // 3p in frag: frag[L-d+i] = temp[L+i]
#pragma unroll 1
    for (int i = 0; i < d; i++) acc = __dmul_rn(acc, __ldg(s3 + (i + d) * 4 + temp(L + d + i)));      // 3p after frag
    acc = __ddiv_rn(acc, plus ? C.maxP : C.maxM);
    const uint4 y = rng_block(P.key, id, attempt, ST_FRAG2);
    const double p = __dmul_rn((double)(y.x >> 1) + 0.5, 1.0 / 2147483648.0);
    return p < acc;
}

// Per-base damage of one 16-base chunk for the models that are still drawn per base (-damagess, -mat single|double / -mapdamage;
// -matfile and -damage are drawn as events, see matrix_events / briggs_events).  Fields 0..15 of W = fragment positions 16k..16k+15, n of
// them inside the fragment; fields beyond the fragment end are masked off by the caller.
__device__ __noinline__ uint32_t damage_chunk_perbase(const FusedParams& P, const DamageDev& D, unsigned long long bh, uint64_t id, int k, int fl, int n, uint32_t W) {
    const int i0 = 16 * k;
    const int nq = (n + 3) >> 2;
    uint32_t Wn = 0;
    if (D.mode == 5) {                                   // -damagess: header block 0 (drawn once per fragment in pass A), per-base word 4+i
        BriggsHdr B; B.o5 = (int)(uint32_t)(bh & 0xFFFFFFFFu); B.o3 = (int)(uint32_t)(bh >> 32); B.nick = 0;
#pragma unroll 1
        for (int q = 0; q < nq; q++) {
            const uint4 x = rng_block(P.key, id, (uint32_t)(4 * k + q + 1), ST_DEAM);
            const uint32_t Wq = W >> (8 * q);
            uint32_t nb = 0;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) nb |= dmg_briggs_ss_base(D, B, i0 + 4 * q + jj, fl, (Wq >> (2 * jj)) & 3u, pick(x, jj)) << (2 * jj);
            Wn |= nb << (8 * q);
        }
    } else {                                             // single / double: two passes, two streams
#pragma unroll 1
        for (int q = 0; q < nq; q++) {
            const uint4 x1 = rng_block(P.key, id, (uint32_t)(4 * k + q), ST_DEAM);
            const uint4 x2 = rng_block(P.key, id, (uint32_t)(4 * k + q), ST_DEAM2);
            const uint32_t Wq = W >> (8 * q);
            uint32_t nb = 0;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) nb |= dmg_sd_base(D, min(i0 + 4 * q + jj, fl - 1), fl, (Wq >> (2 * jj)) & 3u, pick(x1, jj), pick(x2, jj)) << (2 * jj);
            Wn |= nb << (8 * q);
        }
    }
    return Wn;
}

// 16 bases (2 bits each) -> 16 ASCII bytes with byte permutes: a nibble of the selector picks one byte of "ACGT" (both operands hold
// the table, so bit 2 of a nibble is don't-care and only bit 3 has to be cleared); even and odd bases are looked up separately and
// interleaved.  13 ALU instructions, no shared-memory table.
__device__ __forceinline__ uint4 codes16_to_ascii(uint32_t x, uint32_t lut /* 0x54474341 = 'A','C','G','T', in a register */) {
    const uint32_t e = x & 0x77777777u, o = (x >> 2) & 0x77777777u;
    const uint32_t e0 = __byte_perm(lut, lut, e), o0 = __byte_perm(lut, lut, o);               // bases 0,2,4,6 / 1,3,5,7
    const uint32_t e1 = __byte_perm(lut, lut, e >> 16), o1 = __byte_perm(lut, lut, o >> 16);   // bases 8,10,12,14 / 9,11,13,15
    uint4 a;
    a.x = __byte_perm(e0, o0, 0x5140u); a.y = __byte_perm(e0, o0, 0x7362u);
    a.z = __byte_perm(e1, o1, 0x5140u); a.w = __byte_perm(e1, o1, 0x7362u);
    return a;
}
// four ASCII digits of h < 10000, thousands in byte 0 (print order in little-endian memory): h/100 and h%100 side by side in 16-bit
// lanes, then each lane's /10 and %10 side by side in bytes
__device__ __forceinline__ uint32_t dec4_ascii(uint32_t h) {
    const uint32_t q = (h * 5243u) >> 19;                        // h / 100, exact below 43699
    const uint32_t x = q | ((h - q * 100u) << 16);
    const uint32_t t = ((x * 103u) >> 10) & 0x000F000Fu;         // lane / 10, exact below 179; no carry between the lanes
    return (t | ((x - t * 10u) << 8)) + 0x30303030u;
}
// nd decimal digits of a 32-bit value, two at a time from the right
__device__ __forceinline__ void put_dec32(uint8_t* p, uint32_t w, int nd, const uint16_t* d100) {
    int k = nd - 2;
#pragma unroll 1
    for (; k >= 0; k -= 2) { const uint32_t q = w / 100u; const uint32_t two = d100[w - q * 100u]; p[k] = (uint8_t)two; p[k + 1] = (uint8_t)(two >> 8); w = q; }
    if (k == -1) p[0] = (uint8_t)('0' + w);
}

// ---- staged tile -> global memory with the bulk-copy engine (one instruction for the 16-byte aligned body of a tile)
#ifndef GG_BULK_COPY
#define GG_BULK_COPY 0      /* measured: plain 16-byte stores 1.633 ms, bulk copy of the last half 1.660 ms, of both halves 1.645 ms per 10 M fragments */
#endif
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per-warp tile descriptors (a tile has at most 32 fragments); gg_api.cu reserves sizeof(WarpDesc) at the head of a warp's region
struct WarpDesc {
    uint64_t gpos[32];
    uint32_t idx[32], meta[32], so[32], ci[32];
    uint16_t hdr[32], nd[32], cstart[32], rg[32], gstart[36];
};
int fused_desc_bytes() { return (int)sizeof(WarpDesc); }

// One warp owns one tile of up to 32 fragments from sampling to the global store; warps never wait for each other
// (no CTA barrier inside the loop).  Byte offsets come from a decoupled look-back over per-warp-tile states.  Sampling, gather,
// damage events and line images work on the whole tile (a lane per fragment / per 16-base chunk); the records are then rendered and
// stored in two halves of up to 16 fragments, so that the staging buffer -- the largest piece of a warp's shared memory -- holds
// half a tile and 32 warps fit on an SM.
//
// FASTB: every header is >= 13 bytes, so a partial 16-byte group of one sequence line can never reach another line's
// sequence bytes: groups are stored whole and the headers/newlines are written afterwards (no read-modify-write).
// DMODE: 1 = every damage slot of the plan is a -matfile matrix (tables in shared memory), 2 = every slot is Briggs, 0 = any mix.
// EMITC: 0 = fragSim / deamSim records (sequence line = the fragment), 1 = the one-line adptSim dialects (-arts, -artp, -fr, -rr),
//        2 = adptSim's default four-line output.
template <bool FASTB, int DMODE, int EMITC>
__global__ void __launch_bounds__(GG_FUSED_MAXT, 1) k_fused(const __grid_constant__ FusedParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint16_t s_d100[100];                // "00".."99" for the decimal fields of the deflines
    __shared__ uint16_t s_guide[1024];              // FREQ lengths: first candidate row for the top 10 bits of the uniform

    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    asm volatile("" : "+r"(lane));                  // opaque: kept in a register instead of being re-derived from %tid at every use
    const int emit = P.emit;
    constexpr int NLINES = EMITC == 2 ? 2 : 1;
    const int sfx0 = EMITC == 1 ? P.sfx0 : 0;                              // "_r" on the only line of -rr ...
    constexpr int sfx1 = 2;                                                 // ... and on the second line of the default output (adptSim.cpp:286)

    for (int v = tid; v < 100; v += blockDim.x) s_d100[v] = (uint16_t)(('0' + v / 10) | (('0' + v % 10) << 8));
    const uint4* const s_thr = (const uint4*)smem;      // the CTA-shared matrix tables open the dynamic shared memory (SmemLayout::thr == 0)
    if (DMODE == 1)                                   // matrix tables of the used slots (acceptance rows, then terminal rows): global -> shared, once per CTA
        for (int slot = 0; slot < 4; slot++)
            if (P.dmg[slot].mode == 1 && P.sm.thr_slot[slot] >= 0) {
                const int na = 4 * (P.dmg[slot].rows5 + P.dmg[slot].rows3 + 2), nt = 4 * (P.dmg[slot].n5t + P.dmg[slot].n3t);
                for (int q = tid; q < na; q += blockDim.x) ((uint4*)(smem + P.sm.thr))[P.sm.thr_slot[slot] + q] = __ldg(P.dmg[slot].acc + q);
                for (int q = tid; q < nt; q += blockDim.x) ((uint4*)(smem + P.sm.thr))[P.sm.term_slot[slot] + q] = __ldg(P.dmg[slot].term + q);
            }
    if (DMODE != 0)                                   // skip tables of the event draws (their first geo_n entries) and their guides
        for (int slot = 0; slot < 4; slot++)
            if (P.sm.geo_slot[slot] >= 0) {
                uint32_t* g = (uint32_t*)(smem + P.sm.geo_slot[slot]);
                const int n = P.sm.geo_n;
                for (int q = tid; q < n; q += blockDim.x) g[q] = __ldg(P.dmg[slot].geoM + q);                       // (geoM and geoS are the same table)
                for (int q = tid; q < 128; q += blockDim.x) g[n + q] = __ldg((const uint32_t*)P.dmg[slot].guideA + q);
                if (DMODE == 2) {
                    for (int q = tid; q < n; q += blockDim.x) g[n + 128 + q] = __ldg(P.dmg[slot].geoD + q);
                    for (int q = tid; q < 128; q += blockDim.x) g[2 * n + 128 + q] = __ldg((const uint32_t*)P.dmg[slot].guideB + q);
                }
            }
    if (P.sm.tab >= 0) {                              // small plan / genome / contig tables: global -> shared, once per CTA
        uint8_t* tb = smem + P.sm.tab;
        for (int q = tid; q < P.n_ranges * (int)(sizeof(RangeDev) / 4); q += blockDim.x) ((uint32_t*)(tb + P.sm.tab_ranges))[q] = __ldg((const uint32_t*)P.ranges + q);
        for (int q = tid; q < P.n_genomes * (int)(sizeof(GenomeDev) / 4); q += blockDim.x) ((uint32_t*)(tb + P.sm.tab_genomes))[q] = __ldg((const uint32_t*)P.genomes + q);
        for (int q = tid; q < P.sm.n_contigs; q += blockDim.x) {
            ((uint64_t*)(tb + P.sm.tab_cum))[q] = __ldg(P.ctab.cum_end + q); ((uint64_t*)(tb + P.sm.tab_gbase))[q] = __ldg(P.ctab.gbase + q);
            ((uint32_t*)(tb + P.sm.tab_len))[q] = __ldg(P.ctab.len + q); ((uint32_t*)(tb + P.sm.tab_noff))[q] = __ldg(P.ctab.name_off + q);
            ((uint16_t*)(tb + P.sm.tab_nlen))[q] = __ldg(P.ctab.name_len + q);
        }
        for (int q = tid; q < P.sm.names_bytes; q += blockDim.x) ((char*)(tb + P.sm.tab_names))[q] = __ldg(P.ctab.names + q);
        for (int q = tid; q < P.sm.n_contigs; q += blockDim.x) {          // the first 8 bytes of every contig name as one word (header pass)
            uint64_t w8 = 0; const int nl = min((int)__ldg(P.ctab.name_len + q), 8); const char* nm = P.ctab.names + __ldg(P.ctab.name_off + q);
            for (int k = 0; k < nl; k++) w8 |= (uint64_t)(uint8_t)__ldg(nm + k) << (8 * k);
            ((uint64_t*)(tb + P.sm.tab_name8))[q] = w8;
        }
        for (int q = tid; q < P.sm.n_sizes_sm; q += blockDim.x) {
            ((uint32_t*)(tb + P.sm.tab_sthr))[q] = __ldg(P.size_thr + q); ((int32_t*)(tb + P.sm.tab_slen))[q] = __ldg(P.size_len + q);
        }
        for (int k = 0; k < 3; k++) {
            const uint32_t* src = (k == 0 ? P.ad.adF2 : k == 1 ? P.ad.adS2 : P.ad.rcS2) - 1;      // guard word first
            for (int q = tid; q < P.sm.ad_words[k]; q += blockDim.x) ((uint32_t*)(tb + P.sm.tab_ad[k]))[q] = __ldg(src + q);
        }
    }
    if (P.size_mode == 2)
        for (int q = tid; q < 1024; q += blockDim.x) s_guide[q] = (uint16_t)first_below(P.size_thr, P.n_sizes, (uint32_t)q << 21);
    __syncthreads();                                  // the only CTA barrier

    // plan / genome / contig tables: the shared-memory copy when there is one (generic pointers, warp-uniform)
    const bool tsm = P.sm.tab >= 0;
    const uint8_t* const tb = smem + (tsm ? P.sm.tab : 0);
    const RangeDev*  const t_ranges  = tsm ? (const RangeDev*)(tb + P.sm.tab_ranges) : P.ranges;
    const GenomeDev* const t_genomes = tsm ? (const GenomeDev*)(tb + P.sm.tab_genomes) : P.genomes;
    const uint64_t*  const t_cum     = tsm ? (const uint64_t*)(tb + P.sm.tab_cum) : P.ctab.cum_end;
    const uint64_t*  const t_gbase   = tsm ? (const uint64_t*)(tb + P.sm.tab_gbase) : P.ctab.gbase;
    const uint32_t*  const t_len     = tsm ? (const uint32_t*)(tb + P.sm.tab_len) : P.ctab.len;
    const uint32_t*  const t_noff    = tsm ? (const uint32_t*)(tb + P.sm.tab_noff) : P.ctab.name_off;
    const uint16_t*  const t_nlen    = tsm ? (const uint16_t*)(tb + P.sm.tab_nlen) : P.ctab.name_len;
    const char*      const t_names   = tsm ? (const char*)(tb + P.sm.tab_names) : P.ctab.names;
    const uint64_t*  const t_name8   = tsm ? (const uint64_t*)(tb + P.sm.tab_name8) : nullptr;

    const bool ssm = tsm && P.sm.n_sizes_sm > 0;
    const uint32_t* const t_sthr = ssm ? (const uint32_t*)(tb + P.sm.tab_sthr) : P.size_thr;
    const int32_t*  const t_slen = ssm ? (const int32_t*)(tb + P.sm.tab_slen) : P.size_len;
    const uint32_t* const adsm = (tsm && P.sm.ad_words[0] > 0) ? (const uint32_t*)tb : nullptr;
#define TLD(p) (*(p))                           /* generic load: the shared copy or global memory (a per-use select of __ldg costs 0.7 %) */
    // this warp's private shared memory: the fixed-layout descriptors first (one base register + immediate offsets), then the
    // regions whose size depends on the run
    uint32_t wofs = (uint32_t)P.sm.warp0 + (uint32_t)warp * (uint32_t)P.sm.warp_bytes;
    asm volatile("" : "+r"(wofs));                  // opaque, like lane: one register, every region is an immediate or constant-bank offset from it
    uint8_t*  const wb       = smem + wofs;
    WarpDesc* const wd       = (WarpDesc*)wb;
    uint64_t* const s_gpos   = wd->gpos;
    uint32_t* const s_idx    = wd->idx;        // start coordinate within the contig (contigs are shorter than 2^32)
    uint32_t* const s_meta   = wd->meta;
    uint32_t* const s_so     = wd->so;         // offset within the tile's bytes of the first sequence byte of line 0
    uint32_t* const s_ci     = wd->ci;
    uint16_t* const s_hdr    = wd->hdr;
    uint16_t* const s_nd     = wd->nd;         // digit counts of start | end << 4 | length << 8
    uint16_t* const s_cstart = wd->cstart;
    uint16_t* const s_gstart = wd->gstart;
    uint16_t* const s_rg     = wd->rg;
    constexpr int W0 = ((int)sizeof(WarpDesc) + 15) / 16 * 16 + 16;     // fixed offset of the damaged-fragment words; the adptSim dialects stage there too (gg_api.cu make_layout)
    uint8_t*  const s_stage  = EMITC != 0 ? wb + W0 : wb + P.sm.stage;
    unsigned long long* const s_bh = (unsigned long long*)(wb + P.sm.bh);
    uint32_t* const s_dwords = (uint32_t*)(wb + W0);
    uint32_t* const s_lwords = (uint32_t*)(wb + P.sm.lwords);   // line images (adptSim dialects only)
    uint8_t*  const s_cmap   = wb + P.sm.cmap;

    uint32_t my_attempts = 0, my_gather = 0, my_err = 0;
    const uint32_t lut = P.lut;                     // 'A','C','G','T' for the byte permutes of the ASCII pass (a kernel parameter: as an immediate it was re-materialised four times per group)

    // Two tiles are in flight per warp: a tile is SAMPLED (pass A) and its byte count published as soon as its ticket is taken, then the
    // previous tile is rendered and stored, then the new tile is gathered, damaged and given its byte offset.  A ticket is therefore
    // never held by a warp that is busy with something else (the tiles behind it would spin in their look-back), and a tile's
    // look-back comes a whole emission after its predecessors published: it practically never waits.  Pass A leaves its results in
    // registers (the previous tile's descriptors are still needed by the emission).
    int tile = 0;
    if (lane == 0) tile = (int)atomicAdd(P.ticket, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    bool have_prev = false;
    int p_nf = 0; uint32_t p_T = 0, p_roff = 0; unsigned long long p_gofs = 0; bool p_fits = false;   // the tile waiting for its emission
    while (true) {
        const bool have_tile = tile < P.n_tiles;
        // (measured: warps that start their tiles together in groups of 2 / 4 / 8 / 32 to share instruction-cache lines lose 2.5 / 3 / 5 / 23 %)
        if (!have_tile && !have_prev) break;
        const uint64_t tile_first = (uint64_t)tile * (uint64_t)P.F;
        const int nf = have_tile ? (int)min((uint64_t)P.F, P.count - tile_first) : 0;
        uint32_t nx_meta = 0, nx_idx = 0, nx_ci = 0, nx_rghdr = 0, nx_nd = 0;

        // ------------------------------------------------------------ pass A: sample (lane per fragment)
        unsigned long long pk = 0;
        if (lane < nf) {
            uint64_t idx = 0; int L = 0, ci = 0, hdr = 0, tag_len = 0; uint32_t plus = 1;
            const uint64_t id = P.first + tile_first + (uint64_t)lane;
            int lo = 0, hi = P.n_ranges - 1;                    // last range with first_id <= id
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (t_ranges[mid].first_id <= id) lo = mid; else hi = mid - 1; }
            const RangeDev* rg = t_ranges + lo;
            const GenomeDev G = t_genomes[rg->genome];
            const uint64_t* cum = t_cum + G.first_contig;
            const int d = rg->comp >= 0 ? P.comp[rg->comp].dist : 0;       // distFromEnd: 0 without --comp (fragSim.cpp:1019-1021)
            uint64_t gpos = 0; bool ok = false;
            for (uint32_t attempt = 0; attempt < 65536u; attempt++) {
                const uint4 x = rng_block(P.ks, id, attempt, ST_FRAG);
                my_attempts++;
                // selectFragmentLength, fragSim.cpp:137-187
                if (P.size_mode == 0) L = P.fixed_len;
                else if (P.size_mode == 1) L = __ldg(P.size_len + __umulhi(x.x, (uint32_t)P.n_sizes));
                else {
                    const uint32_t u = x.x >> 1;
                    int k = s_guide[u >> 21];                                       // first i with p <= cProb[i] (fragSim.cpp:162-171)
                    while (k < P.n_sizes && u >= t_sthr[k]) k++;
                    if (k == P.n_sizes) k = (int)(((uint64_t)(x.y >> 1) * (uint64_t)P.n_sizes) >> 31);   // fragSim.cpp:176
                    L = t_slen[k];
                }
                if (L < d || L < P.minlen || L > P.maxlen) continue;                 // fragSim.cpp:1375-1381
                const uint64_t r64 = ((uint64_t)x.w << 32) | x.z;
                const uint64_t coord = __umul64hi(r64, G.G);                           // randGenomicCoord, fragSim.cpp:80-93
                if (coord == 0) continue;                                              // never matches a chromosome (starts are 1-based)
                const uint64_t cp = coord - 1;
                // first contig with start+d <= coord <= end-d+1 (fragSim.cpp:1397-1403); for d == 0 that is the first with cp <= cum_end
                const uint64_t key = d ? coord : cp;
                int a = 0, b = G.n_contigs - 1;
                while (a < b) { int mid = (a + b) >> 1; if (key <= TLD(cum + mid)) b = mid; else a = mid + 1; }
                ci = G.first_contig + a;
                const uint64_t clen = TLD(t_len + ci);
                idx = cp - (TLD(cum + a) - clen);                                    // coord - startIndexChr, fragSim.cpp:1404
                if (idx < (uint64_t)d) continue;                                       // outside [start+d, end-d+1]
                gpos = TLD(t_gbase + ci) + idx;
                int xtra = 0;
                if (idx + (uint64_t)L + (uint64_t)d > clen) {                          // run-off past the contig end: always rejected (fragSim.cpp:1472) ...
                    const GenomeDev* Gc = t_genomes + rg->genome;                      // (rare branch: fields re-read here, not kept live)
                    if (ci != Gc->circ_contig || idx + (uint64_t)d > clen || cum[a] + 1 < (uint64_t)(L + d)) continue;   // coord <= end-d+1 (:1400); u64 wrap of end-length-d+1 (:1410)
                    uint64_t s = idx - (uint64_t)d;                                    // ... except on the --circ contig (fragSim.cpp:1412-1426)
                    if (s >= clen) { s = clen - 1; xtra = 1; }                         // coord == end+1 (d == 0): temp1 = "" and temp2 has length+1 bases
                    gpos = Gc->wrap_gbase + (s - (uint64_t)Gc->wrap_s0) + (uint64_t)d;
                }
                if (L + 2 * d > 0) {                                                   // isResolvedDNAstring over the window incl. the flanks, fragSim.cpp:333-339,1472
                    const uint64_t s0 = gpos - (uint64_t)d, e = gpos + (uint64_t)L + (uint64_t)d - 1 + (uint64_t)xtra;
                    const uint64_t w0 = s0 >> 5;                                       // first mask word of the window
                    const int nw = (int)((e >> 5) - w0);                               // index of its last word
                    const uint32_t first = 0xFFFFFFFFu << (s0 & 31), last = 0xFFFFFFFFu >> (31 - (e & 31));
                    // two aligned 16-byte loads cover the first five words of any window (a 32-bit load per word costs one L1 wavefront per
                    // lane and word: the lanes look at unrelated places); words w0 .. w0+4 are elements k .. k+4 of the eight
                    const uint4* q4 = (const uint4*)G.nmask + (w0 >> 2);
                    const uint4 Q0 = __ldg(q4), Q1 = __ldg(q4 + 1);
                    const bool k1 = (w0 & 1u) != 0, k2 = (w0 & 2u) != 0;
                    const uint32_t t0 = k1 ? Q0.y : Q0.x, t1 = k1 ? Q0.z : Q0.y, t2 = k1 ? Q0.w : Q0.z, t3 = k1 ? Q1.x : Q0.w,
                                   t4 = k1 ? Q1.y : Q1.x, t5 = k1 ? Q1.z : Q1.y, t6 = k1 ? Q1.w : Q1.z;
                    const uint32_t m0 = k2 ? t2 : t0, m1 = k2 ? t3 : t1, m2 = k2 ? t4 : t2, m3 = k2 ? t5 : t3, m4 = k2 ? t6 : t4;
                    const uint32_t* mw = G.nmask + w0;
                    const uint32_t mn = nw == 0 ? m0 : nw == 1 ? m1 : nw == 2 ? m2 : nw == 3 ? m3 : nw == 4 ? m4 : __ldg(mw + nw);
                    uint32_t any = nw == 0 ? (m0 & first & last) : ((m0 & first) | (mn & last));
                    if (nw > 1) any |= m1;
                    if (nw > 2) any |= m2;
                    if (nw > 3) any |= m3;
                    if (nw > 4) any |= m4;
#pragma unroll 1
                    for (int w = 5; w < nw; w++) any |= __ldg(mw + w);                 // windows beyond 160 bases
                    if (any) continue;
                }
                plus = P.norev ? 1u : (x.y & 1u);                                      // randomBool, fragSim.cpp:1462-1467
                if (rg->gc >= 0) {                                                     // -gc survival (fragSim.cpp:1509-1534, repeated :1548-1576 without --comp)
                    const int Wn = L + 2 * d + xtra;
                    const uint32_t thr = __ldg(P.gc[rg->gc].thr + (Wn * (Wn + 1) / 2 + gc_count(G.seq2, gpos - (uint64_t)d, Wn)));
                    const uint4 y = rng_block(P.key, id, attempt, ST_FRAG3);
                    if ((y.x >> 1) >= thr) continue;
                    if (!d && (y.y >> 1) >= thr) continue;
                }
                if (xtra && !plus) gpos++;                                             // reverseComplement of the length+1 window, then substr(0,length) (fragSim.cpp:1484-1498)
                if (d && !comp_accept(P, P.comp[rg->comp], G.seq2, gpos, L, plus, id, attempt)) continue;   // --comp, fragSim.cpp:1609-1759
                ok = true; break;
            }
            if (!ok) { my_err |= ERRF_GIVEUP; L = 0; idx = 0; }
            else {      // the gather of pass C is a dependent random access: start pulling its lines into L2 now (+6 % on a 6.2 Gb genome)
                const uint32_t* sp = G.seq2 + (gpos >> 4);
                asm volatile("prefetch.global.L2 [%0];" :: "l"(sp));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(sp + ((L + 15) >> 4)));
            }
            my_gather += (uint32_t)((L + 2 * d + 3) / 4 + (L + 2 * d + 7) / 8);
            tag_len = rg->tag_len;
            const int nlen = t_nlen[ci];
            // digit counts: the end coordinate has as many digits as the start unless it crosses a power of ten (a few, for starts below L); L <= 4095
            const int n1 = ndigits(idx);
            int n2 = n1;
#pragma unroll 1
            while (n2 < 10 && idx + (uint64_t)L >= (uint64_t)c_pow10[n2]) n2++;
            const int n3 = 1 + (L >= 10 ? 1 : 0) + (L >= 100 ? 1 : 0) + (L >= 1000 ? 1 : 0);
            hdr = 1 + nlen + 3 + n1 + 1 + n2 + 1 + n3 + tag_len;
            const int S = EMITC ? P.S_ad : L;
            uint32_t bytes = 0;
            for (int ln = 0; ln < NLINES; ln++) bytes += hdr + (ln ? sfx1 : sfx0) + 1 + S + 1;
            const uint32_t chunks = (L + 15) >> 4;
            const uint32_t groups = EMITC ? P.gpf_ad : (uint32_t)((S + 15) >> 4) + 1u;
            pk = PK(bytes, chunks, groups);
            s_gpos[lane] = gpos;                                 // (dead since the previous tile's gather: written in place)
            nx_meta = META(L, plus, rg->slot, rg->genome);
            nx_nd = (uint32_t)(n1 | (n2 << 4) | (n3 << 8));
            nx_idx = (uint32_t)idx; nx_ci = (uint32_t)ci; nx_rghdr = (uint32_t)lo | ((uint32_t)hdr << 16);   // parked for the header pass
            if (DMODE == 0 && emit != 0 && rg->slot >= 0 && P.dmg[rg->slot].mode == 5) {   // -damagess (drawn per base in pass C): the two overhangs, once per fragment (deamSim.cpp:1342-1343)
                const BriggsHdr B = dmg_briggs_hdr(P.dmg[rg->slot], rng_block(P.key, id, 0u, ST_DEAM));
                s_bh[lane] = (unsigned long long)(uint32_t)B.o5 | ((unsigned long long)(uint32_t)B.o3 << 32);
            }
        }
        // warp exclusive scan of pk
        const unsigned long long incl = warp_scan_u64(pk, lane);
        const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned long long excl = incl - pk;
        const uint32_t T = PK_BYTES(tot), n_chunks = PK_CHUNKS(tot);
        const uint32_t roff = PK_BYTES(excl);                    // byte offset of this lane's record within the tile
        if (have_tile && lane == 0)                              // publish this tile's byte count for the look-back
            st_volatile_u64(P.tile_state + tile, ((tile == 0 ? 2ull : 1ull) << 62) | (unsigned long long)T);

        // ------------------------------------------------------------ the previous tile: rendered and stored while this tile's byte count is already visible
#ifndef GG_STAGE_TEST
#define GG_STAGE_TEST 0          /* timing experiments only: 1 = no emission, no look-back (the front half of the kernel alone) */
#endif
        if (have_prev && GG_STAGE_TEST != 1) {
            // ------------------------------------------------------------ emission in halves of up to P.E fragments
            for (int f0 = 0; f0 < p_nf; f0 += P.E) {
                const int f1 = min(p_nf, f0 + P.E);
                const uint32_t b0 = __shfl_sync(0xffffffffu, p_roff, f0);                              // bytes [b0, b1) of the tile
                const uint32_t b1x = __shfl_sync(0xffffffffu, p_roff, f1 & 31);
                const uint32_t b1 = f1 < p_nf ? b1x : p_T;
                const uint32_t TB = b1 - b0;
                const int pad = (int)((uintptr_t)(P.out + p_gofs + b0) & 15);                           // the half is staged with its global misalignment
                const int pofs = pad - (int)b0;                                                     // staging offset = pofs + offset within the tile
                const bool ok = p_fits && TB <= (uint32_t)P.sm.stage_bytes;                           // host sizing guard: never write past staging
                if (!ok) my_err |= ERRF_OVERFLOW;
    #if GG_BULK_COPY == 2
                if (f0 > 0) { if (lane == 0) bulk_wait_read(); __syncwarp(); }                      // the first half's bulk copy has read the staging buffer
    #endif

                // ---- headers + newlines: two lanes per fragment running the same code on different fields.  Lane 2k: '>' name ":s:" start ':';
                // lane 2k+1: end ':' length tag ["_r"] '\n' and the newline behind the sequence   (fragSim.cpp:1489,1498,1505-1507,1603)
                auto write_headers = [&]() {
                    const int fh = f0 + (lane >> 1), part = lane & 1;
                    if (fh < f1 && ok) {
                        const uint32_t meta = s_meta[fh];
                        const int L = META_L(meta);
                        const int hdr = (int)s_hdr[fh];
                        const uint32_t nd = s_nd[fh];
                        const int n3 = (int)(nd >> 8);
                        const int S = EMITC ? P.S_ad : L;
                        uint8_t* rec = s_stage + pofs + (int)s_so[fh] - hdr - sfx0 - 1;             // first byte of the record
                        const char* csrc; int cn, nv, coff, voff; uint64_t v;                       // bytes to copy (name / tag) and the coordinate to print, with their offsets in the header
                        uint64_t cw = 0; bool wordcopy;                                             // names / tags of up to 8 bytes travel in a register pair
                        if (part == 0) {
                            const int ci = (int)s_ci[fh];
                            csrc = t_names + t_noff[ci]; cn = t_nlen[ci]; coff = 1;
                            v = s_idx[fh]; nv = (int)(nd & 15u); voff = 1 + cn + 3;
                            wordcopy = t_name8 != nullptr && cn <= 8;
                            if (wordcopy) cw = t_name8[ci];
                        } else {
                            const RangeDev* rg = t_ranges + s_rg[fh];
                            csrc = rg->tag; cn = rg->tag_len; coff = hdr - cn;
                            v = (uint64_t)s_idx[fh] + (uint64_t)L; nv = (int)((nd >> 4) & 15u); voff = hdr - cn - n3 - 1 - nv;
                            wordcopy = cn <= 8;
                            if (wordcopy) cw = *(const uint64_t*)rg->tag;
                        }
                        // Fast path: the coordinate is converted in registers (eight ASCII digits with leading zeros) and stored as eight
                        // unconditional bytes -- right-aligned by lane 2k (the surplus lands on ":s:" / the name / '>', which this lane writes
                        // afterwards), left-aligned by lane 2k+1 (the surplus lands on ':' length tag ["_r"] '\n', which that lane writes
                        // afterwards).  The surplus must stay inside the lane's own bytes; otherwise the lane takes the exact byte loops.
                        // Everything below is the same instruction stream for both lanes (selects on an opaque copy of `part`: specialised
                        // per part by the compiler it would run twice).
                        int pb = part; asm volatile("" : "+r"(pb));
                        const int pmask = -pb;
                        const bool fastp = wordcopy && v <= 0xFFFFFFFFull && nv + cn + (pmask & (n3 + sfx0 - 2)) >= 4;
                        if (fastp) {
                            const uint32_t v32 = (uint32_t)v;
                            const uint32_t top = v32 / 100000000u, r8 = v32 - top * 100000000u;
                            const uint32_t h4 = r8 / 10000u, l4 = r8 - h4 * 10000u;
                            const int lead = max(0, 8 - nv) & pmask;                                // lane 2k+1 drops the leading zeros
                            const uint64_t X = (((uint64_t)dec4_ascii(l4) << 32) | (uint64_t)dec4_ascii(h4)) >> (8 * lead);
                            const uint32_t X0 = (uint32_t)X, X1 = (uint32_t)(X >> 32);
                            const int qoff = voff + nv - 8 + lead;
                            // a field of up to four bytes that ends where the next one starts: ":s:" in front of the start coordinate / the length in front of the tag
                            const uint32_t F4 = pb ? dec4_ascii((uint32_t)L) : (0x3A003A00u | ((META_PLUS(meta) ? (uint32_t)'+' : (uint32_t)'-') << 16));
                            const int eoff = pb ? coff : voff, nfld = pb ? n3 : 3;
                            const uint8_t ch = pb ? (uint8_t)'\n' : (uint8_t)'>';                   // '>' in front / the newlines behind the header and behind the sequence
    #pragma unroll
                            for (int ln = 0; ln < NLINES; ln++) {
                                const int sfx = ln ? sfx1 : sfx0;
                                uint8_t* q = rec + qoff;
                                q[0] = (uint8_t)X0; q[1] = (uint8_t)(X0 >> 8); q[2] = (uint8_t)(X0 >> 16); q[3] = (uint8_t)(X0 >> 24);
                                q[4] = (uint8_t)X1; q[5] = (uint8_t)(X1 >> 8); q[6] = (uint8_t)(X1 >> 16); q[7] = (uint8_t)(X1 >> 24);
                                uint8_t* p = rec + voff;
                                if (nv > 8) {                                                       // coordinates of 1e8 and more: one or two digits in front
                                    const uint32_t two = s_d100[top];
                                    p[nv - 9] = (uint8_t)(two >> 8);
                                    if (nv > 9) p[0] = (uint8_t)two;
                                }
                                p[nv] = ':';
                                uint8_t* tp = rec + coff;
    #pragma unroll
                                for (int k = 0; k < 8; k++) if (k < cn) tp[k] = (uint8_t)(cw >> (8 * k));
                                uint8_t* e = rec + eoff;
                                e[-1] = (uint8_t)(F4 >> 24);
                                if (nfld > 1) e[-2] = (uint8_t)(F4 >> 16);
                                if (nfld > 2) e[-3] = (uint8_t)(F4 >> 8);
                                if (nfld > 3) e[-4] = (uint8_t)F4;
                                if (sfx && pb) { rec[hdr] = '_'; rec[hdr + 1] = 'r'; }
                                rec[pmask & (hdr + sfx)] = ch; rec[pmask & (hdr + sfx + S + 1)] = ch;
                                rec += hdr + sfx + 1 + S + 1;
                            }
                        } else {
    #pragma unroll
                        for (int ln = 0; ln < NLINES; ln++) {
                            const int sfx = ln ? sfx1 : sfx0;
                            if (wordcopy) {
    #pragma unroll
                                for (int k = 0; k < 8; k++) if (k < cn) rec[coff + k] = (uint8_t)(cw >> (8 * k));
                            } else {
    #pragma unroll 1
                                for (int k = 0; k < cn; k++) rec[coff + k] = (uint8_t)csrc[k];
                            }
                            uint8_t* p = rec + voff;
                            if (v <= 0xFFFFFFFFull) put_dec32(p, (uint32_t)v, nv, s_d100); else put_decimal(p, v, nv, s_d100);
                            p[nv] = ':';
                            if (part == 0) { rec[0] = '>'; p[-3] = ':'; p[-2] = META_PLUS(meta) ? '+' : '-'; p[-1] = ':'; }
                            else {
                                put_dec32(p + nv + 1, (uint32_t)L, n3, s_d100);
                                uint8_t* q = rec + hdr;
                                if (sfx) { q[0] = '_'; q[1] = 'r'; q += 2; }
                                q[0] = '\n'; q[S + 1] = '\n';
                            }
                            rec += hdr + sfx + 1 + S + 1;
                        }
                        }
                    }
                };
                if (!FASTB) { write_headers(); __syncwarp(); }

                // ---- ASCII: two lanes per fragment; a lane walks half of the staging groups of a sequence line (or one of the two lines of the
                // four-line dialect) with one running source word: one shared-memory load, one funnel shift, the expansion and one 16-byte
                // store per group
                {
                    const int fd = f0 + (lane >> 1), h = lane & 1;
                    if (fd < f1 && ok) {
                        int S = P.S_ad;
                        if (EMITC == 0) S = META_L(s_meta[fd]);
                        int ls = pofs + (int)s_so[fd];                                              // staging offset of the first sequence byte of the line
                        const uint32_t* src = EMITC == 0 ? s_dwords + s_cstart[fd] : s_lwords + fd * P.lw_stride + 1;
                        if (EMITC == 2 && h) { ls += (int)s_hdr[fd] + sfx1 + S + 2; src += P.LWL; }   // the reverse read's line
                        const int sigma = ls & 15;
                        const int ngr = (sigma + S + 15) >> 4;                                      // staging groups that hold bases of the line
                        const int half = (ngr + 1) >> 1;
                        const int g_lo = (EMITC == 2 || !h) ? 0 : half, g_hi = (EMITC == 2 || h) ? ngr : half;
                        uint8_t* dst = s_stage + (ls - sigma) + 16 * g_lo;                          // 16-byte aligned
                        // group g holds bases [16 g - sigma, 16 g - sigma + 16): source words (16 (g+1) - sigma) >> 4 - 1 and the next one (a guard
                        // word sits in front of every source)
                        const uint32_t* sp = src + (((16 * (g_lo + 1) - sigma) >> 4) - 1);
                        const uint32_t sh = ((uint32_t)(16 - sigma) & 15u) * 2u;
                        uint32_t prev = *sp;
    #pragma unroll 1
                        for (int g = g_lo; g < g_hi; g++) {
                            const uint32_t next = *++sp;
                            const uint4 a = codes16_to_ascii(__funnelshift_r(prev, next, sh), lut);
                            prev = next;
                            const int j0 = 16 * g - sigma;
                            if (FASTB || (j0 >= 0 && j0 + 16 <= S)) {
                                *(uint4*)dst = a;                           // FASTB: bytes outside the line are rewritten by the header pass
                            } else {
                                // boundary group: merge under a byte mask.  Headers were written before; two sequence lines are always
                                // >= 12 bytes apart (newline + a header), so no other lane touches these 4-byte words in this pass.
                                const int t0 = max(j0, 0) - j0, t1 = min(j0 + 16, S) - j0;          // valid bytes [t0,t1) of the group
                                const uint32_t V = ((1u << t1) - 1u) & ~((1u << t0) - 1u);
                                uint32_t* d4 = (uint32_t*)dst;
                                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    #pragma unroll
                                for (int w = 0; w < 4; w++) {
                                    const uint32_t nib = (V >> (4 * w)) & 0xFu;
                                    const uint32_t m = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;     // 4 bits -> 4 byte masks
                                    if (nib == 0xFu) d4[w] = aw[w];
                                    else if (nib) d4[w] = (d4[w] & ~m) | (aw[w] & m);
                                }
                            }
                            dst += 16;
                        }
                    }
                }
                __syncwarp();
                if (FASTB) { write_headers(); __syncwarp(); }

                // ---- copy-out: staging -> global
                if (ok) {
                    uint8_t* const gal = P.out + p_gofs + b0 - pad;            // 16-byte aligned; staging byte s <-> gal[s], valid s in [pad, pad+TB)
                    const int end = pad + (int)TB;
                    const int c_lo = (pad + 15) >> 4, c_hi = end >> 4;       // chunks [c_lo, c_hi) are complete
    #if GG_BULK_COPY
                    if (GG_BULK_COPY == 2 || f1 == p_nf) {                     // the tile's last half goes out through the bulk-copy engine: its staging is not needed before the next tile's emission
                        fence_async_smem();                                  // this lane's staging writes -> visible to the bulk-copy engine
                        __syncwarp();
                        if (lane == 0 && c_hi > c_lo) bulk_store(gal + 16 * c_lo, s_stage + 16 * c_lo, (uint32_t)(16 * (c_hi - c_lo)));
                    } else
    #endif
    #if GG_COPY_UNROLL == 4
    #pragma unroll 1
                    for (int c = c_lo + lane; c < c_hi; c += 128) {  // four independent 16-byte copies per trip (a half tile is one or two trips)
                        const bool p1 = c + 32 < c_hi, p2 = c + 64 < c_hi, p3 = c + 96 < c_hi;
                        const uint8_t* sp = s_stage + 16 * c; uint8_t* gp = gal + 16 * c;
                        uint4 v0 = *(const uint4*)sp, v1 = v0, v2 = v0, v3 = v0;
                        if (p1) v1 = *(const uint4*)(sp + 512);
                        if (p2) v2 = *(const uint4*)(sp + 1024);
                        if (p3) v3 = *(const uint4*)(sp + 1536);
                        st_global_v4(gp, v0);
                        if (p1) st_global_v4(gp + 512, v1);
                        if (p2) st_global_v4(gp + 1024, v2);
                        if (p3) st_global_v4(gp + 1536, v3);
                    }
    #elif GG_COPY_UNROLL
                    {
                        int c = c_lo + lane;
    #pragma unroll 1
                        for (; c + 32 < c_hi; c += 64) {             // two independent 16-byte copies per trip
                            const uint4 v0 = *(const uint4*)(s_stage + 16 * c), v1 = *(const uint4*)(s_stage + 16 * c + 512);
                            st_global_v4(gal + 16 * c, v0); st_global_v4(gal + 16 * c + 512, v1);
                        }
                        if (c < c_hi) st_global_v4(gal + 16 * c, *(const uint4*)(s_stage + 16 * c));
                    }
    #else
                    for (int c = c_lo + lane; c < c_hi; c += 32) st_global_v4(gal + 16 * c, *(const uint4*)(s_stage + 16 * c));
    #endif
                    // the partial head [pad, 16 c_lo) and tail [16 max(c_hi, c_lo), end): at most 15 bytes each, one byte per lane
                    const int t = lane < 16 ? pad + lane : 16 * max(c_hi, c_lo) + (lane - 16);
                    if (lane < 16 ? t < min(16 * c_lo, end) : t < end) gal[t] = s_stage[t];
                }
                __syncwarp();
            }
        }
        have_prev = false;
        if (!have_tile) break;
        __syncwarp();
        if (lane < nf) {                                         // the new tile's descriptors (the previous tile's are dead now)
            s_meta[lane] = nx_meta; s_hdr[lane] = (uint16_t)(nx_rghdr >> 16); s_nd[lane] = (uint16_t)nx_nd;
            s_idx[lane] = nx_idx; s_ci[lane] = nx_ci; s_rg[lane] = (uint16_t)nx_rghdr;
        }

        // ------------------------------------------------------------ pass B1: work maps (owner lane)
        if (lane < nf) {
            const uint32_t c0 = PK_CHUNKS(excl);
            s_so[lane] = roff + (nx_rghdr >> 16) + (uint32_t)sfx0 + 1u; s_cstart[lane] = (uint16_t)c0;
            const uint32_t nch = PK_CHUNKS(pk);
#pragma unroll 1
            for (uint32_t k = 0; k < nch; k++) s_cmap[c0 + k] = (uint8_t)lane;
        }
        // first look-back probe, issued now and consumed after the damage pass (its latency overlaps pass C)
        unsigned long long lb = 2ull << 62;
        if (tile > 0 && tile - 1 - lane >= 0) lb = ld_volatile_u64(P.tile_state + (tile - 1 - lane));
        __syncwarp();

        // ------------------------------------------------------------ pass C: gather (lane per 16-base chunk)
        // GG_GATHER_U chunks per lane and trip: all loads of a trip are issued before any of them is used (the gather is a dependent L2
        // access: a trip costs one exposed L2 latency, so a tile should need as few trips as the registers allow)
        for (uint32_t e0 = lane; e0 < n_chunks; e0 += 32u * GG_GATHER_U) {
            uint32_t wlo[GG_GATHER_U], whi[GG_GATHER_U], wmt[GG_GATHER_U];     // wmt: shift [4:0] | fields [10:5] | reverse strand [11]
#pragma unroll
            for (int u = 0; u < GG_GATHER_U; u++) {
                const uint32_t e = e0 + 32u * (uint32_t)u;
                wlo[u] = whi[u] = wmt[u] = 0u;
                if (e < n_chunks) {
                    const int f = s_cmap[e];
                    const int k = (int)e - (int)s_cstart[f];
                    const uint32_t meta = s_meta[f];
                    const int fl = META_L(meta);
                    const uint32_t* seq2 = t_genomes[META_GEN(meta)].seq2;
                    // forward strand: bases gp+16k ..; reverse strand: the 16 bases that end at gp+fl-16k, reverse-complemented (fragSim.cpp:1484).
                    // Chunks are 16 bases = one word apart: the word index is that of the fragment's first / last window plus or minus k and
                    // the shift is the same for all of them (32-bit arithmetic: a packed genome has fewer than 2^32 words)
                    const uint64_t base = META_PLUS(meta) ? s_gpos[f] : s_gpos[f] + (uint64_t)fl;
                    const uint32_t w0 = (uint32_t)(base >> 4);
                    const uint32_t* wp = seq2 + (META_PLUS(meta) ? w0 + (uint32_t)k : w0 - (uint32_t)k - 1u);
                    wlo[u] = __ldg(wp); whi[u] = __ldg(wp + 1);
                    wmt[u] = (((uint32_t)base & 15u) * 2u) | ((uint32_t)min(16, fl - 16 * k) << 5) | (META_PLUS(meta) ? 0u : 0x800u);
                }
            }
#pragma unroll
            for (int u = 0; u < GG_GATHER_U; u++) {
                const uint32_t e = e0 + 32u * (uint32_t)u;
                if (e < n_chunks) {
                    uint32_t W = __funnelshift_r(wlo[u], whi[u], wmt[u] & 31u);
                    if (wmt[u] & 0x800u) W = revcomp16(W);
                    const int wn = (int)((wmt[u] >> 5) & 63u);
                    if (DMODE == 0 && emit != 0) {                   // the models still drawn per base
                        const int f = s_cmap[e];
                        const uint32_t meta = s_meta[f];
                        const int slot = META_SLOT(meta);
                        if (slot >= 0 && P.dmg[slot].mode >= 3)
                            W = damage_chunk_perbase(P, P.dmg[slot], s_bh[f], P.first + tile_first + (uint64_t)f, (int)e - (int)s_cstart[f], META_L(meta), wn, W);
                    }
                    s_dwords[e] = W & lowfields(wn);
                }
            }
        }
        __syncwarp();

        // ------------------------------------------------------------ pass E: damage events (lane per fragment; deamSim.cpp:1477-1613,1794-1930)
        if (emit != 0 && lane < nf) {
            const uint32_t meta = s_meta[lane];
            const int slot = META_SLOT(meta);
            if (slot >= 0) {
                const int mode = DMODE != 0 ? DMODE : P.dmg[slot].mode;
                const int L = META_L(meta);
                const uint64_t id = P.first + tile_first + (uint64_t)lane;
                SeqPacked Sq; Sq.w = s_dwords + s_cstart[lane];
                // DMODE 1 / 2: the slot's skip tables sit in shared memory (the host falls back to DMODE 0 when they do not fit)
                const uint32_t* const gA = (const uint32_t*)(smem + (DMODE != 0 ? P.sm.geo_slot[slot] : 0)); const uint8_t* const uA = (const uint8_t*)(gA + P.sm.geo_n);
                if (mode == 1) {
                    if (DMODE == 1) matrix_events<7, true, true>(P.ks, P.dmg[slot], s_thr + P.sm.term_slot[slot], s_thr + P.sm.thr_slot[slot], gA, uA, Sq, L, id);
                    else matrix_events<7, false, false>(P.ks, P.dmg[slot], P.dmg[slot].term, P.dmg[slot].acc, P.dmg[slot].geoM, P.dmg[slot].guideA, Sq, L, id);
                } else if (mode == 2) {
                    const BriggsHdr B = dmg_briggs_hdr(P.dmg[slot], rng_block(P.ks, id, 0u, ST_DEAM));
                    if (DMODE == 2) {
                        const uint32_t* const gB = (const uint32_t*)(uA + 512); const uint8_t* const uB = (const uint8_t*)(gB + P.sm.geo_n);
                        briggs_events<7, true>(P.ks, P.dmg[slot], B, gA, uA, gB, uB, Sq, L, id);
                    } else briggs_events<7, false>(P.ks, P.dmg[slot], B, P.dmg[slot].geoS, P.dmg[slot].guideA, P.dmg[slot].geoD, P.dmg[slot].guideB, Sq, L, id);
                }
            }
        }
        __syncwarp();

        // ------------------------------------------------------------ pass C2: line images (owner lane; adptSim dialects only)
        if (EMITC != 0 && lane < nf) {
            const int L = META_L(s_meta[lane]);
            const uint64_t id = P.first + tile_first + (uint64_t)lane;
            const uint32_t* D = s_dwords + s_cstart[lane];
            uint32_t* lw = s_lwords + lane * P.lw_stride;
            // -rr: reverse read; -artp: forward ++ revcomp(reverse); otherwise the forward read (first line)
            build_line_image(P, adsm, D, L, id, emit == 6 ? 1 : emit == 4 ? 2 : 0, lw + 1);
            if (EMITC == 2) build_line_image(P, adsm, D, L, id, 1, lw + P.LWL + 1);     // default 4-line output: second line = reverse read
        }

        // the next ticket: asked for now, used at the end of the iteration (the atomic's latency hides behind the look-back; sampling starts
        // as soon as the look-back is done, so the ticket is still not held for long)
        int next_tile = 0;
        if (lane == 0) next_tile = (int)atomicAdd(P.ticket, 1u);

        // ------------------------------------------------------------ look-back: global byte offset of this tile
        unsigned long long gofs = 0;
        if (tile > 0 && GG_STAGE_TEST != 1) {
            int look = tile - 1;
            while (true) {
                const int i = look - lane;
                while (__any_sync(0xffffffffu, (lb >> 62) == 0)) {
#if GG_SPIN_SLEEP
                    __nanosleep(GG_SPIN_SLEEP);                  // a predecessor has not published yet: leave the issue slots to the warps that work
#endif
                    if ((lb >> 62) == 0) lb = ld_volatile_u64(P.tile_state + i);
                }
                const uint32_t inc = __ballot_sync(0xffffffffu, (lb >> 62) == 2ull);
                const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                gofs += warp_sum_u64(lane <= first_inc ? (lb & 0x3FFFFFFFFFFFFFFFull) : 0ull);
                if (inc) break;
                look -= 32;
                lb = (look - lane >= 0) ? ld_volatile_u64(P.tile_state + (look - lane)) : (2ull << 62);
            }
            if (lane == 0) st_volatile_u64(P.tile_state + tile, (2ull << 62) | (gofs + (unsigned long long)T));
        }
        if (lane == 0 && tile == P.n_tiles - 1) P.stats[ST_TOTAL] = gofs + (unsigned long long)T;
        const bool fits = gofs + T <= P.out_cap;                         // host sizing guard: never write past the output
        if (!fits) my_err |= ERRF_OVERFLOW;
        if (EMITC == 0 && P.case_off != nullptr && lane < nf) {          // fragSim --case: where the record's bases are and where they came from (k_case_apply)
            const uint64_t f = tile_first + (uint64_t)lane;
            P.case_off[f] = gofs + (unsigned long long)s_so[lane]; P.case_gpos[f] = s_gpos[lane]; P.case_meta[f] = s_meta[lane];
        }
        // this tile now waits for its emission; the next ticket (its sampling starts at once, at the top of the loop)
        have_prev = true; p_nf = nf; p_T = T; p_roff = roff; p_gofs = gofs; p_fits = fits;
        tile = __shfl_sync(0xffffffffu, next_tile, 0);
    }
#if GG_BULK_COPY
    if (lane == 0) bulk_wait_all();
#endif

    // per-warp statistics
    unsigned long long a64 = warp_sum_u64((unsigned long long)my_attempts), g64 = warp_sum_u64((unsigned long long)my_gather);
    uint32_t e = my_err;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e |= __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) {
        if (a64) atomicAdd(P.stats + ST_ATTEMPTS, a64);
        if (g64) atomicAdd(P.stats + ST_GATHER, g64);
        if (e) atomicOr(P.stats + ST_ERR, (unsigned long long)e);
    }
}

typedef void (*fused_fn)(const FusedParams);
static fused_fn fused_variant(int fast, int dmode, int emitc) {
#define GG_V(F, D) (emitc == 0 ? k_fused<F, D, 0> : emitc == 1 ? k_fused<F, D, 1> : k_fused<F, D, 2>)
    if (fast) return dmode == 1 ? GG_V(true, 1) : dmode == 2 ? GG_V(true, 2) : GG_V(true, 0);
    return GG_V(false, 0);                           // short headers (rare): only the general damage variant is built
#undef GG_V
}
cudaError_t fused_prepare(int fast, int dmode, int emitc, int threads, int smem_bytes, int* blocks_per_sm, int* regs) {
    fused_fn f = fused_variant(fast, dmode, emitc);
    cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes a; e = cudaFuncGetAttributes(&a, f);
    if (e != cudaSuccess) return e;
    *regs = a.numRegs;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, f, threads, smem_bytes);
}
cudaError_t launch_fused(const FusedParams& P, int emitc, int grid, int threads, cudaStream_t st) {
    fused_variant(P.fast_boundary, P.dmode, emitc)<<<grid, threads, P.sm.total, st>>>(P);
    return cudaGetLastError();
}

// fragSim --case (fragSim.cpp:318-326,548-551): the fused kernel writes upper-case bases; this pass lower-cases the bases whose genome
// position carries the case bit (reverse-strand fragments read the mask backwards: complementing keeps the case).  Warp per record.
__global__ void __launch_bounds__(256) k_case_apply(uint64_t n_rec, const unsigned long long* __restrict__ off, const unsigned long long* __restrict__ gpos,
                                                    const uint32_t* __restrict__ meta, const uint32_t* const* __restrict__ cmasks, uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rec; r += nw) {
        const uint32_t m = meta[r];
        const int L = META_L(m);
        const uint32_t* cm = cmasks[META_GEN(m)];
        if (cm == nullptr) continue;
        const uint64_t g = gpos[r], o = off[r];
        for (int i = lane; i < L; i += 32) {
            const uint64_t p = META_PLUS(m) ? g + (uint64_t)i : g + (uint64_t)(L - 1 - i);
            if ((__ldg(cm + (p >> 5)) >> (p & 31)) & 1u) out[o + (uint64_t)i] |= 0x20;
        }
    }
}
cudaError_t launch_case_apply(uint64_t n_rec, const unsigned long long* off, const unsigned long long* gpos, const uint32_t* meta,
                              const uint32_t* const* cmasks, uint8_t* out, cudaStream_t st) {
    if (n_rec == 0) return cudaSuccess;
    const uint64_t want = (n_rec + 7) / 8;
    k_case_apply<<<(unsigned)(want < 148ull * 16 ? want : 148ull * 16), 256, 0, st>>>(n_rec, off, gpos, meta, cmasks, out);
    return cudaGetLastError();
}

// ======================================================================================= record-stream helpers
__device__ __forceinline__ int count_nl16(uint4 v) {
    const uint32_t nl = 0x0A0A0A0Au;
    return __popc(__vcmpeq4(v.x, nl) & 0x01010101u) + __popc(__vcmpeq4(v.y, nl) & 0x01010101u) +
           __popc(__vcmpeq4(v.z, nl) & 0x01010101u) + __popc(__vcmpeq4(v.w, nl) & 0x01010101u);
}
__device__ __forceinline__ uint4 load16_guard(const uint8_t* in, uint64_t pos, uint64_t n) {
    if (pos + 16 <= n) return *(const uint4*)(in + pos);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int t = 0; t < 16; t++) if (pos + t < n) w[t >> 2] |= (uint32_t)in[pos + t] << (8 * (t & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void __launch_bounds__(256) k_count_newlines(const uint8_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ block_counts) {
    const uint64_t b0 = (uint64_t)blockIdx.x * NL_BLOCK_BYTES;
    int cnt = 0;
    for (uint64_t p = b0 + 16ull * threadIdx.x; p < b0 + NL_BLOCK_BYTES && p < n; p += 16ull * 256) cnt += count_nl16(load16_guard(in, p, n));
    __shared__ int s[8];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s[w]; block_counts[blockIdx.x] = (unsigned long long)t; }
}
__global__ void __launch_bounds__(256) k_fill_newlines(const uint8_t* __restrict__ in, uint64_t n, const unsigned long long* __restrict__ block_base, uint64_t* __restrict__ nl) {
    const uint64_t b0 = (uint64_t)blockIdx.x * NL_BLOCK_BYTES;
    __shared__ int s[8];
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_base = block_base[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t step = b0; step < b0 + NL_BLOCK_BYTES && step < n; step += 16ull * 256) {
        const uint64_t p = step + 16ull * threadIdx.x;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (p < n) v = load16_guard(in, p, n);
        const int c = count_nl16(v);
        int inc = c;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s[warp] = inc;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 8; w++) { if (w < warp) woff += s[w]; tot += s[w]; }
        unsigned long long o = s_base + (unsigned long long)(woff + inc - c);
        if (c) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            for (int t = 0; t < 16; t++) if (((w4[t >> 2] >> (8 * (t & 3))) & 0xFFu) == 0x0Au) nl[o++] = p + t;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += (unsigned long long)tot;
        __syncthreads();
    }
}
cudaError_t launch_count_newlines(const uint8_t* in, uint64_t n, unsigned long long* block_counts, int n_blocks, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    k_count_newlines<<<n_blocks, 256, 0, st>>>(in, n, block_counts); return cudaGetLastError();
}
cudaError_t launch_fill_newlines(const uint8_t* in, uint64_t n, const unsigned long long* block_base, uint64_t* nl, int n_blocks, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    k_fill_newlines<<<n_blocks, 256, 0, st>>>(in, n, block_base, nl); return cudaGetLastError();
}

// exclusive scan of u64 data in place: per-block reduce -> single-block scan of block sums -> per-block apply
constexpr int SCAN_ELEMS = 2048;   // per block (256 threads x 8)
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* s_w, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = warp_scan_u64(v, lane);
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long woff = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { if (w < warp) woff += s_w[w]; tot += s_w[w]; }
    __syncthreads();
    *total = tot;
    return woff + inc - v;
}
__global__ void __launch_bounds__(256) k_scan_reduce(const unsigned long long* __restrict__ d, uint64_t n, unsigned long long* __restrict__ sums) {
    const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_ELEMS;
    unsigned long long a = 0;
    for (int k = 0; k < 8; k++) { const uint64_t i = b0 + (uint64_t)threadIdx.x * 8 + k; if (i < n) a += d[i]; }
    __shared__ unsigned long long s[8];
    a = warp_sum_u64(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < 8; w++) t += s[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_top(unsigned long long* __restrict__ sums, uint64_t nb, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nb; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < nb ? sums[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_excl_scan(v, s_w, &tot);
        const unsigned long long c = carry;
        if (i < nb) sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_apply(unsigned long long* __restrict__ d, uint64_t n, const unsigned long long* __restrict__ sums) {
    const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_ELEMS;
    unsigned long long v[8], a = 0;
    for (int k = 0; k < 8; k++) { const uint64_t i = b0 + (uint64_t)threadIdx.x * 8 + k; v[k] = i < n ? d[i] : 0ull; a += v[k]; }
    __shared__ unsigned long long s_w[8];
    unsigned long long tot;
    unsigned long long ex = block_excl_scan(a, s_w, &tot) + sums[blockIdx.x];
    for (int k = 0; k < 8; k++) { const uint64_t i = b0 + (uint64_t)threadIdx.x * 8 + k; if (i < n) d[i] = ex; ex += v[k]; }
}
cudaError_t launch_exclusive_scan(unsigned long long* data, uint64_t n, unsigned long long* scratch, unsigned long long* total, cudaStream_t st) {
    const uint64_t nb = (n + SCAN_ELEMS - 1) / SCAN_ELEMS;
    if (n == 0) { return cudaMemsetAsync(total, 0, sizeof(unsigned long long), st); }
    k_scan_reduce<<<(unsigned)nb, 256, 0, st>>>(data, n, scratch);
    k_scan_top<<<1, 1024, 0, st>>>(scratch, nb, total);
    k_scan_apply<<<(unsigned)nb, 256, 0, st>>>(data, n, scratch);
    return cudaGetLastError();
}

// record r of a 2-line stream: id line [ids, ide), sequence line [ss, se); se is the position of its newline
struct RecPos { uint64_t ids, ide, ss, se; };
__device__ __forceinline__ RecPos rec_pos(const RecStream& R, uint64_t r) {
    RecPos p;
    p.ids = r == 0 ? 0ull : R.nl[2 * r - 1] + 1ull;
    p.ide = R.nl[2 * r];
    p.ss = p.ide + 1ull;
    p.se = (2 * r + 1 < R.n_nl) ? R.nl[2 * r + 1] : R.n;
    return p;
}

// ASCII record in global memory (already upper-cased) as the sequence the event draws work on
struct SeqAscii {
    uint8_t* p;
    __device__ __forceinline__ int code(int i) const { return ascii2code(p[i]); }
    __device__ __forceinline__ void put(int i, uint32_t oldb, uint32_t newb) { if (newb != oldb) p[i] = code2ascii(newb); }
};
// the event-drawn models (-matfile, -damage) on one record; PH as in matrix_events / briggs_events
template <int PH, class SeqT>
__device__ __forceinline__ void deam_events(const DamageDev& D, const BriggsHdr& B, const Key& key, uint64_t id, int L, SeqT& S) {
    if (D.mode == 1) matrix_events<PH, false, false>(key, D, D.term, D.acc, D.geoM, D.guideA, S, L, id);
    else if (D.mode == 2) briggs_events<PH, false>(key, D, B, D.geoS, D.guideA, D.geoD, D.guideB, S, L, id);
}

// stand-alone deamSim main loop (deamSim.cpp:1300-2042): a warp takes 32 records; the lanes copy them (upper-casing the sequence,
// deamSim.cpp:1319-1321, and applying the models that are drawn per base), then every lane draws the damage events of one record
__global__ void __launch_bounds__(256) k_deam_records(const RecStream R, uint8_t* __restrict__ out, const DamageDev D, int has_damage,
                                                      Key key, uint64_t first_id, unsigned long long* stats) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const bool perbase = has_damage && D.mode >= 3, events = has_damage && (D.mode == 1 || D.mode == 2);
    for (uint64_t r0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32ull; r0 < R.n_rec; r0 += nw * 32ull) {
        const int nr = (int)min((uint64_t)32, R.n_rec - r0);
        for (int rr = 0; rr < nr; rr++) {
            const uint64_t r = r0 + (uint64_t)rr;
            const RecPos p = rec_pos(R, r);
            const uint64_t id = first_id + r;
            if (lane == 0 && (p.ide <= p.ids || R.in[p.ids] != '>')) atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
            for (uint64_t i = p.ids + lane; i < p.ss; i += 32) out[i] = R.in[i];          // ID line + '\n' unchanged (deamSim.cpp:2038)
            const int L = (int)(p.se - p.ss);
            BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
            if (perbase && D.mode == 5) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
            for (int i = lane; i < L; i += 32) {
                const uint8_t raw = R.in[p.ss + i];
                uint8_t c = upper(raw);                                                    // toupper, deamSim.cpp:1319-1321
                const int code = ascii2code(c);
                if (perbase && D.mode == 6) {                                              // methylation: no toupper; a resolved base with a row comes out upper-case (deamSim.cpp:1630-1787)
                    c = raw;
                    if (code >= 0) { const int nb = dmg_meth_base(D, i, L, code | (raw != upper(raw) ? 4 : 0), rng_word(key, id, ST_DEAM, (uint32_t)i)); if (nb >= 0) c = code2ascii((uint32_t)nb); }
                } else if (perbase && code >= 0) {                                         // unresolved bases are never touched (deamSim.cpp:1795)
                    uint32_t nb = (uint32_t)code;
                    if (D.mode == 5) nb = dmg_briggs_ss_base(D, B, i, L, nb, rng_word(key, id, ST_DEAM, 4u + (uint32_t)i));
                    else nb = dmg_sd_base(D, i, L, nb, rng_word(key, id, ST_DEAM, (uint32_t)i), rng_word(key, id, ST_DEAM2, (uint32_t)i));
                    c = code2ascii(nb);
                }
                out[p.ss + i] = c;
            }
            if (lane == 0) out[p.se] = '\n';
        }
        if (events) {
            __syncwarp();                                                                  // the copies above are visible to the lane that owns the record
            if (lane < nr) {
                const uint64_t r = r0 + (uint64_t)lane;
                const RecPos p = rec_pos(R, r);
                const uint64_t id = first_id + r;
                const int L = (int)(p.se - p.ss);
                BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
                if (D.mode == 2) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
                SeqAscii S; S.p = out + p.ss;
                deam_events<7>(D, B, key, id, L, S);
            }
        }
    }
}
// ======================================================================================= stand-alone deamSim, single pass
// deamSim's main loop (deamSim.cpp:1300-2042) over a 2-line FASTA record stream in ONE kernel: persistent, autonomous warps take
// 4 KB tiles of the input by ticket.  A warp copies its tile (plus 1.5 KB of the next one: the record that starts in a tile may end
// behind it) into shared memory with 16-byte loads, finds the newlines, learns the number of lines in front of the tile from a
// decoupled look-back over per-tile line counts (that gives the line parity and the record index = the Philox fragment id), then
// every lane takes one record that STARTS in the tile: upper-cases the sequence (deamSim.cpp:1319-1321), draws its damage events
// and the warp stores the bytes of its records with 16-byte stores at the input's own offsets (the stage does not change the record
// sizes without -name).  No record index in global memory, no host round trip.  Anything unusual (more than DS_NLMAX lines in a
// tile, a record that does not end inside the buffer) raises ERRF_FALLBACK and the host takes the general multi-kernel path.
constexpr int DS_TILE = 4096, DS_TAIL = 1536, DS_BUF = DS_TILE + DS_TAIL, DS_NLMAX = 320;      // at most 255 lines per tile on this path, plus the tail's
constexpr int DS_WARP_BYTES = DS_BUF + 16 + 2 * DS_NLMAX + 32;
constexpr int DS_WARPS = 32;
constexpr int DS_TAB_BYTES = 16384;      // CTA-shared copy of the matrix tables (acceptance rows, then terminal rows) when they fit
constexpr int DS_GEO_N = 1024;           // ... and of the first entries of the skip tables of the event draws with their guides (records up to that length draw their skips from shared memory)
constexpr int DS_GEO_BYTES = 2 * (4 * DS_GEO_N + 512);

// exact per-byte "== '\n'" flags of a word: 0x80 in every matching byte
__device__ __forceinline__ uint32_t nl_flags(uint32_t x) {
    const uint32_t t = x ^ 0x0A0A0A0Au;
    return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);
}
// upper-case the ASCII letters of a word (toupper of deamSim.cpp:1319-1321 on four bytes)
__device__ __forceinline__ uint32_t upper4(uint32_t x) {
    const uint32_t y = x & 0x7F7F7F7Fu;
    const uint32_t m = (y + 0x1F1F1F1Fu) & ~(y + 0x05050505u) & ~x & 0x80808080u;      // 'a' <= byte <= 'z'
    return x ^ (m >> 2);
}
struct SeqAsciiSmem {                    // an upper-cased record in shared memory
    uint8_t* p;
    __device__ __forceinline__ int code(int i) const {
        const uint32_t c = p[i];
        const uint32_t v = (c >> 1) & 3u, k = v ^ (v >> 1);                 // A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 -> 3
        return ((0x54474341u >> (8 * k)) & 0xFFu) == c ? (int)k : -1;
    }
    __device__ __forceinline__ void put(int i, uint32_t oldb, uint32_t newb) { if (newb != oldb) p[i] = code2ascii(newb); }
};

__global__ void __launch_bounds__(DS_WARPS * 32, 1) k_deam_stream(const uint8_t* __restrict__ in, uint64_t n, uint8_t* __restrict__ out, const DamageDev D, int has_damage,
                                                                 const __grid_constant__ KeySched key, uint64_t first_id, unsigned long long* __restrict__ tile_state, unsigned int* __restrict__ ticket,
                                                                 unsigned long long* __restrict__ stats, uint32_t n_tiles) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* const buf = smem + (size_t)warp * DS_WARP_BYTES;               // tile bytes at their global alignment (tiles start at multiples of 4096)
    uint16_t* const nlpos = (uint16_t*)(buf + DS_BUF + 16);
    uint32_t my_err = 0;
    // matrix tables: global -> shared, once per CTA (L1 is small at this carve-out and the record bytes stream through it)
    const int n_acc = 4 * (D.rows5 + D.rows3 + 2), n_term = 4 * (D.n5t + D.n3t);
    const bool tab_sm = has_damage && D.mode == 1 && (n_acc + n_term) * 16 <= DS_TAB_BYTES;
    uint4* const s_tab = (uint4*)(smem + (size_t)DS_WARPS * DS_WARP_BYTES);
    if (tab_sm) {
        for (int q = threadIdx.x; q < n_acc; q += blockDim.x) s_tab[q] = __ldg(D.acc + q);
        for (int q = threadIdx.x; q < n_term; q += blockDim.x) s_tab[n_acc + q] = __ldg(D.term + q);
    }
    // skip tables of the event draws (matrix: geoM; Briggs: geoS, geoD): a skip is two dependent loads and a record draws several in a row
    uint32_t* const s_geo = (uint32_t*)(smem + (size_t)DS_WARPS * DS_WARP_BYTES + DS_TAB_BYTES);   // [geoA | guideA | geoB | guideB]
    if (has_damage && (D.mode == 1 || D.mode == 2)) {
        const int ng = min(DS_GEO_N, D.n_geo);
        for (int q = threadIdx.x; q < ng; q += blockDim.x) { s_geo[q] = __ldg(D.geoM + q); if (D.mode == 2) s_geo[DS_GEO_N + 128 + q] = __ldg(D.geoD + q); }
        for (int q = threadIdx.x; q < 128; q += blockDim.x) { s_geo[DS_GEO_N + q] = __ldg((const uint32_t*)D.guideA + q); if (D.mode == 2) s_geo[2 * DS_GEO_N + 128 + q] = __ldg((const uint32_t*)D.guideB + q); }
    }
    __syncthreads();
    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const uint64_t b0 = (uint64_t)tile * DS_TILE;
        const int tlen = (int)min((uint64_t)DS_TILE, n - b0);              // bytes of the tile proper
        const int blen = (int)min((uint64_t)DS_BUF, n - b0);               // bytes in the buffer (tile + tail)
        // the tile this warp will most likely take next (tickets sweep the stream in order): pull it into L2 now, one line per lane
        { const uint64_t pf = b0 + (uint64_t)gridDim.x * DS_WARPS * DS_TILE + 128ull * lane; if (pf < n) asm volatile("prefetch.global.L2 [%0];" :: "l"(in + pf)); }
        const uint8_t prev_byte = tile == 0 ? (uint8_t)'\n' : __ldg(in + b0 - 1);               // does a line start at the tile's first byte?  (loaded with the tile)
        // ---- load: 16-byte chunks, chunk c = lane + 32 k; newline flags of the chunks are kept for the counting below
        uint32_t cnt_lo = 0, cnt_hi = 0;                                    // newline counts of my chunks k = 0..3 / 4..7 (tile proper), 8 bits each
        uint32_t msk[11];                                                   // newline masks of my chunks: bit 8 j + q = byte j of word q, i.e. position 4 q + j of the chunk
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const int c = lane + 32 * k;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (16 * c < blen) v = __ldg((const uint4*)(in + b0) + c);      // (the input buffer has 16 bytes of slack behind n)
            *(uint4*)(buf + 16 * c) = v;
            uint32_t f0 = nl_flags(v.x), f1 = nl_flags(v.y), f2 = nl_flags(v.z), f3 = nl_flags(v.w);   // 0x80 per newline byte
            if (16 * c + 16 > blen) {                                       // the stream's last chunk: bytes behind the end are not data
                const int valid = max(0, blen - 16 * c);
                f0 &= valid >= 4 ? 0xFFFFFFFFu : (1u << (8 * max(valid, 0))) - 1u;
                f1 &= valid >= 8 ? 0xFFFFFFFFu : valid <= 4 ? 0u : (1u << (8 * (valid - 4))) - 1u;
                f2 &= valid >= 12 ? 0xFFFFFFFFu : valid <= 8 ? 0u : (1u << (8 * (valid - 8))) - 1u;
                f3 &= valid >= 16 ? 0xFFFFFFFFu : valid <= 12 ? 0u : (1u << (8 * (valid - 12))) - 1u;
            }
            const uint32_t m = (f0 >> 7) | (f1 >> 6) | (f2 >> 5) | (f3 >> 4);
            msk[k] = m;
            if (k < 8) { const uint32_t pc = (uint32_t)__popc(m); if (k < 4) cnt_lo |= pc << (8 * k); else cnt_hi |= pc << (8 * (k - 4)); }   // (behind the last tile's end the chunks are empty: tile proper == valid bytes)
        }
        // ---- ranks: newlines are ordered by position = by (k, lane); exclusive prefix of the counts in that order.  The 8-bit fields of
        // the packed scans cannot overflow when the tile has at most 255 newlines; a tile with more falls back.
        uint32_t my_nl = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) my_nl += ((cnt_lo >> (8 * k)) & 0xFFu) + ((cnt_hi >> (8 * k)) & 0xFFu);
        const uint32_t tot_nl = __reduce_add_sync(0xffffffffu, my_nl);
        uint32_t inc_lo = cnt_lo, inc_hi = cnt_hi;                          // per-field inclusive scans over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, inc_lo, o), b = __shfl_up_sync(0xffffffffu, inc_hi, o);
            if (lane >= o) { inc_lo += a; inc_hi += b; }
        }
        const uint32_t tot_lo = __shfl_sync(0xffffffffu, inc_lo, 31), tot_hi = __shfl_sync(0xffffffffu, inc_hi, 31);
        uint32_t kbase[9]; kbase[0] = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) kbase[k + 1] = kbase[k] + (((k < 4 ? tot_lo : tot_hi) >> (8 * (k & 3))) & 0xFFu);
        const uint32_t nl_tile = kbase[8];                                  // newlines in the tile proper
        bool fallback = tot_nl > 255u;
        if (lane == 0) st_volatile_u64(tile_state + tile, ((tile == 0 ? 2ull : 1ull) << 62) | (unsigned long long)tot_nl);
        // first look-back probe (consumed below)
        unsigned long long lb = 2ull << 62;
        if (tile > 0 && (long long)tile - 1 - lane >= 0) lb = ld_volatile_u64(tile_state + (tile - 1 - lane));
        // ---- newline positions of the tile proper, in order, then those of the tail
        if (!fallback) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int c = lane + 32 * k;
                uint32_t m = msk[k];
                uint32_t r = kbase[k] + (((k < 4 ? inc_lo : inc_hi) >> (8 * (k & 3))) & 0xFFu) - (uint32_t)__popc(m);
                // byte j of word q sits at bit t = 8 j + q, i.e. position 4 (t & 7) + (t >> 3) of the chunk.  A 16-byte chunk of a record stream
                // rarely holds two newlines: one newline needs no loop; otherwise walk q = 0..3 and, inside, j upwards (ascending positions)
                if (m) {
                    if ((m & (m - 1u)) == 0u) { const int t = __ffs(m) - 1; nlpos[r] = (uint16_t)(16 * c + 4 * (t & 7) + (t >> 3)); }
                    else {
#pragma unroll 1
                        for (int q = 0; q < 4; q++) {
                            uint32_t mq = (m >> q) & 0x01010101u;
                            while (mq) { const int b = __ffs(mq) - 1; nlpos[r++] = (uint16_t)(16 * c + 4 * q + (b >> 3)); mq &= mq - 1u; }
                        }
                    }
                }
            }
        }
        // tail (chunks 256..351): at most 48 positions are kept (two are needed per record that crosses the tile's end)
        uint32_t nl_all = nl_tile;
        {
            uint32_t tc = 0;
#pragma unroll
            for (int k = 8; k < 11; k++) tc |= (uint32_t)__popc(msk[k]) << (8 * (k - 8));
            uint32_t ti = tc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += a; }
            const uint32_t tt = __shfl_sync(0xffffffffu, ti, 31);
            uint32_t tb = 0;
            if (!fallback) {
#pragma unroll
                for (int k = 8; k < 11; k++) {
                    const uint32_t m = msk[k];
                    uint32_t r = nl_tile + tb + ((ti >> (8 * (k - 8))) & 0xFFu) - (uint32_t)__popc(m);
                    if (m) {
                        if ((m & (m - 1u)) == 0u) { const int t = __ffs(m) - 1; if (r < (uint32_t)DS_NLMAX) nlpos[r] = (uint16_t)(16 * (lane + 32 * k) + 4 * (t & 7) + (t >> 3)); }
                        else {
#pragma unroll 1
                            for (int q = 0; q < 4; q++) {
                                uint32_t mq = (m >> q) & 0x01010101u;
                                while (mq) { const int b = __ffs(mq) - 1; if (r < (uint32_t)DS_NLMAX) nlpos[r] = (uint16_t)(16 * (lane + 32 * k) + 4 * q + (b >> 3)); r++; mq &= mq - 1u; }
                            }
                        }
                    }
                    tb += (tt >> (8 * (k - 8))) & 0xFFu;
                }
            }
            nl_all += (tt & 0xFFu) + ((tt >> 8) & 0xFFu) + ((tt >> 16) & 0xFFu);
        }
        nl_all = min(nl_all, (uint32_t)DS_NLMAX);
        // ---- look-back: lines in front of this tile
        unsigned long long base = 0;
        if (tile > 0) {
            long long look = (long long)tile - 1;
            while (true) {
                const long long i = look - lane;
                while (__any_sync(0xffffffffu, (lb >> 62) == 0)) { if ((lb >> 62) == 0) lb = ld_volatile_u64(tile_state + i); }
                const uint32_t inc = __ballot_sync(0xffffffffu, (lb >> 62) == 2ull);
                const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                base += warp_sum_u64(lane <= first_inc ? (lb & 0x3FFFFFFFFFFFFFFFull) : 0ull);
                if (inc) break;
                look -= 32;
                lb = (look - lane >= 0) ? ld_volatile_u64(tile_state + (look - lane)) : (2ull << 62);
            }
            if (lane == 0) st_volatile_u64(tile_state + tile, (2ull << 62) | (base + (unsigned long long)tot_nl));
        }
        if (tile == n_tiles - 1 && lane == 0) {                             // lines of the whole stream: a last line without '\n' counts (deamSim reads it)
            const bool open_end = buf[tlen - 1] != '\n';
            stats[ST_TOTAL] = base + nl_tile + (open_end ? 1ull : 0ull);
            stats[ST_ATTEMPTS] = open_end ? 1ull : 0ull;                    // (the record gets its newline: one more output byte)
        }
        __syncwarp();
        // ---- the records that start in this tile.  The line that starts behind newline i of the tile has global index base + i + 1 (the
        // line at the tile's first byte, when one starts there, has index base); ID lines are the even ones.
        const bool starts_here = prev_byte == '\n';
        int i0 = (base & 1ull) ? 0 : -1;                                    // first newline index i with (base + i + 1) even
        if (i0 == -1 && !starts_here) i0 = 1;
        // records q = 0,1,..: newline index i = i0 + 2q, start s = nlpos[i] + 1 (or 0), ID line ends at newline i+1, sequence at newline i+2
        int n_own = 0;
        if ((int)nl_tile - 1 >= i0) n_own = ((int)nl_tile - 1 - i0) / 2 + 1;                     // records with i <= nl_tile - 1 ...
        if (n_own > 0 && i0 + 2 * (n_own - 1) >= 0 && (int)nlpos[i0 + 2 * (n_own - 1)] + 1 >= tlen) n_own--;   // ... whose first byte is still inside the tile proper
        int own_lo = DS_BUF, own_hi = 0;                                    // byte range of my records in the buffer
        if (!fallback)
        for (int q = lane; q < n_own; q += 32) {
            const int i = i0 + 2 * q;
            const int s = i < 0 ? 0 : (int)nlpos[i] + 1;
            int e1, e2; bool closed = true;
            if ((uint32_t)(i + 1) < nl_all) e1 = nlpos[i + 1]; else { e1 = blen; closed = false; }
            if ((uint32_t)(i + 2) < nl_all) e2 = nlpos[i + 2];
            else if (closed && b0 + (uint64_t)blen == n) e2 = blen;                              // the stream's last line has no newline
            else { e2 = blen; closed = false; }
            if (!closed) { fallback = true; break; }
            if (e2 >= DS_BUF && !(b0 + (uint64_t)blen == n)) { fallback = true; break; }
            if (buf[s] != '>' || e1 <= s) my_err |= ERRF_PARSE;                                  // FastQParser: a record starts with '>' (deamSim.cpp:1288)
            const int ss = e1 + 1, L = e2 - ss;
            // upper-case the sequence (deamSim.cpp:1319-1321).  A, C, G, T, N have bit 5 clear: records without any byte that has it set
            // (everything fragSim writes) are left alone after one OR over their words
            uint32_t any;
            {
                const int pa = ss & ~3, pe = e2 & ~3;                                            // the words that hold the first base / the byte behind the last
                const uint32_t head = 0xFFFFFFFFu << (8 * (ss & 3)), tail = (1u << (8 * (e2 & 3))) - 1u;   // bytes of those words that belong to the sequence
                if (pa == pe) any = *(const uint32_t*)(buf + pa) & head & tail;
                else {
                    any = *(const uint32_t*)(buf + pa) & head;
#pragma unroll 2
                    for (int p = pa + 4; p < pe; p += 4) any |= *(const uint32_t*)(buf + p);
                    any |= *(const uint32_t*)(buf + pe) & tail;
                }
            }
            if (any & 0x20202020u) {
                int p = ss;
                for (; p < e2 && (p & 3); p++) buf[p] = upper(buf[p]);
                for (; p + 4 <= e2; p += 4) *(uint32_t*)(buf + p) = upper4(*(const uint32_t*)(buf + p));
                for (; p < e2; p++) buf[p] = upper(buf[p]);
            }
            if (e2 < DS_BUF + 16) buf[e2] = '\n';
            if (has_damage && L > 0) {
                const uint64_t id = first_id + ((base + (unsigned long long)(i + 1)) >> 1);
                SeqAsciiSmem S; S.p = buf + ss;
                // skips: thresholds from the shared copy when the record is short enough for it (generic pointers, plain loads), guides always
                const bool gsm = L <= DS_GEO_N;
                const uint32_t* const gA = gsm ? s_geo : D.geoM; const uint8_t* const uA = (const uint8_t*)(s_geo + DS_GEO_N);
                if (D.mode == 1) {
                    if (tab_sm) matrix_events<7, true, true>(key, D, s_tab + n_acc, s_tab, gA, uA, S, L, id);
                    else matrix_events<7, false, true>(key, D, D.term, D.acc, gA, uA, S, L, id);
                } else {
                    const BriggsHdr B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
                    const uint32_t* const gB = gsm ? s_geo + DS_GEO_N + 128 : D.geoD; const uint8_t* const uB = (const uint8_t*)(s_geo + 2 * DS_GEO_N + 128);
                    briggs_events<7, true>(key, D, B, gA, uA, gB, uB, S, L, id);
                }
            }
            own_lo = min(own_lo, s); own_hi = max(own_hi, e2 + 1);
        }
        fallback = __any_sync(0xffffffffu, fallback);
        if (fallback) { my_err |= ERRF_FALLBACK; continue; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { own_lo = min(own_lo, __shfl_xor_sync(0xffffffffu, own_lo, o)); own_hi = max(own_hi, __shfl_xor_sync(0xffffffffu, own_hi, o)); }
        __syncwarp();
        // ---- store [own_lo, own_hi) at the input's own offsets (16-byte stores for the aligned body)
        if (own_hi > own_lo) {
            uint8_t* const gal = out + b0;
            const int c_lo = (own_lo + 15) >> 4, c_hi = own_hi >> 4;
            int c = c_lo + lane;
#pragma unroll 1
            for (; c + 32 < c_hi; c += 64) {                                   // two independent 16-byte copies per trip
                const uint4 v0 = *(const uint4*)(buf + 16 * c), v1 = *(const uint4*)(buf + 16 * c + 512);
                st_global_v4(gal + 16 * c, v0); st_global_v4(gal + 16 * c + 512, v1);
            }
            if (c < c_hi) st_global_v4(gal + 16 * c, *(const uint4*)(buf + 16 * c));
            const int t = lane < 16 ? own_lo + lane : 16 * max(c_hi, c_lo) + (lane - 16);
            if (lane < 16 ? t < min(16 * c_lo, own_hi) : t < own_hi) gal[t] = buf[t];
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_err |= __shfl_xor_sync(0xffffffffu, my_err, o);
    if (lane == 0 && my_err) atomicOr(stats + ST_ERR, (unsigned long long)my_err);
}
cudaError_t launch_deam_stream(const uint8_t* in, uint64_t n, uint8_t* out, const DamageDev& D, int has_damage, Key key, uint64_t first_id,
                               unsigned long long* tile_state, unsigned int* ticket, unsigned long long* stats, uint32_t n_tiles, int sm_count, cudaStream_t st) {
    if (n_tiles == 0) return cudaSuccess;
    const int smem_bytes = DS_WARPS * DS_WARP_BYTES + DS_TAB_BYTES + DS_GEO_BYTES;
    cudaError_t e = cudaFuncSetAttribute(k_deam_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)min((uint64_t)sm_count, ((uint64_t)n_tiles + DS_WARPS - 1) / DS_WARPS);
    k_deam_stream<<<grid, DS_WARPS * 32, smem_bytes, st>>>(in, n, out, D, has_damage, make_key_sched(key.k0, key.k1), first_id, tile_state, ticket, stats, n_tiles);
    return cudaGetLastError();
}
int deam_stream_tile_bytes() { return DS_TILE; }

// ---- deamSim -name: "_DEAM:p1,p2,..." appended to the ID of every damaged record (deamSim.cpp:2029-2037).
// Position lists follow the reference's push order: matrix: i ascending, +(i+1) in the 5' half, -(len-i) in the 3' half
// (:1857,1919); Briggs: +(i+1) (:1489-1589); single/double: all pass-1 hits +(i+1), then all pass-2 hits -(len-i) (:1949-1994).
struct DeamEvt { uint32_t nb; int v0, v1; bool keep; };     // new base; list-0 value (0 = none); list-1 value (0 = none); keep = the input byte stays as it is
// the models drawn per base (-damagess, -mat single|double, -matfilemeth/-matfilenonmeth); lower = the input byte was a lower-case letter
__device__ __forceinline__ DeamEvt deam_event(const DamageDev& D, const BriggsHdr& B, const Key& key, uint64_t id, int i, int L, int code, bool lower) {
    DeamEvt e; e.nb = (uint32_t)code; e.v0 = 0; e.v1 = 0; e.keep = false;
    const uint32_t b = (uint32_t)code;
    if (D.mode == 6) {                                                                     // deamSim.cpp:1709-1711,1770-1773: one list, in position order
        const int nb = dmg_meth_base(D, i, L, code | (lower ? 4 : 0), rng_word(key, id, ST_DEAM, (uint32_t)i));
        if (nb < 0) e.keep = true;
        else { e.nb = (uint32_t)nb; if (e.nb != b) e.v0 = i < (L >> 1) ? i + 1 : -(L - i); }
    } else if (D.mode == 5) {
        e.nb = dmg_briggs_ss_base(D, B, i, L, b, rng_word(key, id, ST_DEAM, 4u + (uint32_t)i));
        if (e.nb != b) e.v0 = i + 1;
    } else {
        const uint32_t x1 = rng_word(key, id, ST_DEAM, (uint32_t)i), x2 = rng_word(key, id, ST_DEAM2, (uint32_t)i);
        const int r5 = min(i, D.rows5 - 1), r3 = min(L - 1 - i, D.rows3 - 1);
        uint32_t c = b;
        if (c == 1u && (x1 >> 1) < __ldg(D.sd5 + r5)) { c = 3u; e.v0 = i + 1; }
        if (D.mode == 3) { if (c == 1u && (x2 >> 1) < __ldg(D.sd3 + r3)) { c = 3u; e.v1 = -(L - i); } }
        else             { if (c == 2u && (x2 >> 1) < __ldg(D.sd3 + r3)) { c = 0u; e.v1 = -(L - i); } }
        e.nb = c;
    }
    return e;
}
__device__ __forceinline__ int evt_text_len(int v) { return v == 0 ? 0 : 1 + (v < 0 ? 1 : 0) + ndigits((uint64_t)(v < 0 ? -v : v)); }   // separator + sign + digits
// the event-drawn models under -name: the phases of matrix_events / briggs_events run in position order, so the list comes out in the
// reference's order.  Sizing pass: in = the input sequence (upper-cased on the fly), txt = nullptr; writing pass: the output record.
struct SeqAsciiNamed {
    const uint8_t* in; uint8_t* outp; uint8_t* txt; uint32_t cur; int L, half_or_neg;     // half_or_neg: matrix = L/2 (3' half listed as -(L-i)), Briggs = -1
    __device__ __forceinline__ int code(int i) const { return ascii2code(upper(in[i])); }
    __device__ __forceinline__ void put(int i, uint32_t oldb, uint32_t newb) {
        if (newb == oldb) return;
        const int v = (half_or_neg < 0 || i < half_or_neg) ? i + 1 : -(L - i);
        if (txt) {
            outp[i] = code2ascii(newb);
            uint8_t* q = txt + cur;
            *q++ = cur == 5u ? ':' : ',';                                                  // the first event follows "_DEAM"
            if (v < 0) *q++ = '-';
            uint32_t a = (uint32_t)(v < 0 ? -v : v); const int nd = ndigits((uint64_t)a);
            for (int k = nd - 1; k >= 0; k--) { q[k] = (uint8_t)('0' + a % 10u); a /= 10u; }
        }
        cur += (uint32_t)evt_text_len(v);
    }
};

// pass 1: output bytes per record
__global__ void __launch_bounds__(256) k_deam_name_sizes(const RecStream R, const DamageDev D, Key key, uint64_t first_id,
                                                         unsigned long long* __restrict__ sizes, unsigned long long* stats) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (D.mode == 1 || D.mode == 2) {                                                      // event-drawn models: thread per record
        for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R.n_rec; r += (uint64_t)gridDim.x * blockDim.x) {
            const RecPos p = rec_pos(R, r);
            const uint64_t id = first_id + r;
            if (p.ide <= p.ids || R.in[p.ids] != '>') atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
            const int L = (int)(p.se - p.ss);
            BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
            if (D.mode == 2) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
            SeqAsciiNamed S; S.in = R.in + p.ss; S.outp = nullptr; S.txt = nullptr; S.cur = 5u; S.L = L; S.half_or_neg = D.mode == 1 ? (L >> 1) : -1;
            deam_events<1>(D, B, key, id, L, S); deam_events<2>(D, B, key, id, L, S); deam_events<4>(D, B, key, id, L, S);
            sizes[r] = (p.ide - p.ids) + (S.cur > 5u ? (unsigned long long)S.cur : 0ull) + 1ull + (unsigned long long)L + 1ull;
        }
        return;
    }
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R.n_rec; r += nw) {
        const RecPos p = rec_pos(R, r);
        const uint64_t id = first_id + r;
        if (lane == 0 && (p.ide <= p.ids || R.in[p.ids] != '>')) atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
        const int L = (int)(p.se - p.ss);
        BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
        if (D.mode == 5) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
        unsigned long long txt = 0;
        for (int i = lane; i < L; i += 32) {
            const uint8_t raw = R.in[p.ss + i];
            const int code = ascii2code(upper(raw));
            if (code >= 0) { const DeamEvt e = deam_event(D, B, key, id, i, L, code, raw != upper(raw)); txt += (unsigned long long)(evt_text_len(e.v0) + evt_text_len(e.v1)); }
        }
        txt = warp_sum_u64(txt);
        if (lane == 0) sizes[r] = (p.ide - p.ids) + (txt ? 5ull + txt : 0ull) + 1ull + (unsigned long long)L + 1ull;
    }
}
// pass 2: write "ID[_DEAM:list]\nSEQ\n"
__global__ void __launch_bounds__(256) k_deam_name_records(const RecStream R, const unsigned long long* __restrict__ out_off, uint8_t* __restrict__ out,
                                                           const DamageDev D, Key key, uint64_t first_id) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    if (D.mode == 1 || D.mode == 2) {                                                      // event-drawn models: the warp copies 32 records, then a lane per record
        for (uint64_t r0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32ull; r0 < R.n_rec; r0 += nw * 32ull) {
            const int nr = (int)min((uint64_t)32, R.n_rec - r0);
            for (int rr = 0; rr < nr; rr++) {
                const uint64_t r = r0 + (uint64_t)rr;
                const RecPos p = rec_pos(R, r);
                const int L = (int)(p.se - p.ss);
                const uint64_t idlen = p.ide - p.ids, total = out_off[r + 1] - out_off[r];
                uint8_t* o = out + out_off[r];
                uint8_t* seqo = o + (total - (uint64_t)L - 1);
                for (uint64_t i = lane; i < idlen; i += 32) o[i] = R.in[p.ids + i];
                if (total > idlen + 1 + (uint64_t)L + 1 && lane < 5) o[idlen + lane] = (uint8_t)"_DEAM"[lane];
                for (int i = lane; i < L; i += 32) seqo[i] = upper(R.in[p.ss + i]);
                if (lane == 0) { seqo[-1] = '\n'; seqo[L] = '\n'; }
            }
            __syncwarp();
            if (lane < nr) {
                const uint64_t r = r0 + (uint64_t)lane;
                const RecPos p = rec_pos(R, r);
                const uint64_t id = first_id + r;
                const int L = (int)(p.se - p.ss);
                const uint64_t idlen = p.ide - p.ids, total = out_off[r + 1] - out_off[r];
                uint8_t* o = out + out_off[r];
                BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
                if (D.mode == 2) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
                SeqAsciiNamed S; S.in = R.in + p.ss; S.outp = o + (total - (uint64_t)L - 1); S.txt = o + idlen; S.cur = 5u; S.L = L; S.half_or_neg = D.mode == 1 ? (L >> 1) : -1;
                deam_events<1>(D, B, key, id, L, S); deam_events<2>(D, B, key, id, L, S); deam_events<4>(D, B, key, id, L, S);
            }
            __syncwarp();
        }
        return;
    }
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R.n_rec; r += nw) {
        const RecPos p = rec_pos(R, r);
        const uint64_t id = first_id + r;
        const int L = (int)(p.se - p.ss);
        const uint64_t idlen = p.ide - p.ids;
        uint8_t* o = out + out_off[r];
        for (uint64_t i = lane; i < idlen; i += 32) o[i] = R.in[p.ids + i];
        BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
        if (D.mode == 5) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
        const uint64_t total = out_off[r + 1] - out_off[r];                       // sizes array has a sentinel entry
        uint8_t* seqo = o + (total - (uint64_t)L - 1);                            // sequence start
        const bool any = total > idlen + 1 + (uint64_t)L + 1;
        if (any && lane < 5) o[idlen + lane] = (uint8_t)"_DEAM"[lane];
        uint32_t cur = (uint32_t)idlen + 5;                                        // next text offset within the record
        for (int list = 0; list < 2; list++) {
            if (list == 1 && D.mode != 3 && D.mode != 4) break;
            for (int i0 = 0; i0 < L; i0 += 32) {
                const int i = i0 + lane;
                int v = 0; uint32_t nb = 0; int code = -1; uint8_t ch = 0;
                if (i < L) {
                    const uint8_t raw = R.in[p.ss + i];
                    ch = upper(raw); code = ascii2code(ch);
                    bool keep = code < 0;
                    if (code >= 0) { const DeamEvt e = deam_event(D, B, key, id, i, L, code, raw != ch); v = list == 0 ? e.v0 : e.v1; nb = e.nb; keep = e.keep; }
                    if (list == 0) seqo[i] = !keep ? code2ascii(nb) : (D.mode == 6 ? raw : ch);   // sequence bytes written once (methylation mode does not upper-case what it leaves alone)
                }
                const int len = evt_text_len(v);
                int inc = len;
                for (int s2 = 1; s2 < 32; s2 <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, s2); if (lane >= s2) inc += t; }
                const int tot = __shfl_sync(0xffffffffu, inc, 31);
                if (len) {
                    uint8_t* q = o + cur + (uint32_t)(inc - len);
                    *q++ = (cur + (uint32_t)(inc - len) == (uint32_t)idlen + 5) ? ':' : ',';   // first event overall follows "_DEAM"
                    if (v < 0) *q++ = '-';
                    uint32_t a = (uint32_t)(v < 0 ? -v : v); const int nd = ndigits((uint64_t)a);
                    for (int k = nd - 1; k >= 0; k--) { q[k] = (uint8_t)('0' + a % 10u); a /= 10u; }
                }
                cur += (uint32_t)tot;
            }
        }
        if (lane == 0) { seqo[-1] = '\n'; seqo[L] = '\n'; }
    }
}
cudaError_t launch_deam_name_sizes(const RecStream& R, const DamageDev& D, Key key, uint64_t first_id, unsigned long long* sizes, unsigned long long* stats, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    k_deam_name_sizes<<<(unsigned)(want < 148ull * 16 ? want : 148ull * 16), 256, 0, st>>>(R, D, key, first_id, sizes, stats);
    return cudaGetLastError();
}
cudaError_t launch_deam_name_records(const RecStream& R, const unsigned long long* out_off, uint8_t* out, const DamageDev& D, Key key, uint64_t first_id, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    k_deam_name_records<<<(unsigned)(want < 148ull * 16 ? want : 148ull * 16), 256, 0, st>>>(R, out_off, out, D, key, first_id);
    return cudaGetLastError();
}
cudaError_t launch_deam_records(const RecStream& R, uint8_t* out, const DamageDev& D, int has_damage, Key key,
                                uint64_t first_id, unsigned long long* stats, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    const unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    k_deam_records<<<grid, 256, 0, st>>>(R, out, D, has_damage, key, first_id, stats);
    return cudaGetLastError();
}

// "_adpt" if the forward read got adapter bases, "_rand" if it also got random bases (flags come from the forward read
// only, adptSim.cpp:294-310,378-386), then the -tag string
__device__ __forceinline__ int adpt_name_extra(const AdptNaming& nm, const Adapters& ad, int L) {
    int x = nm.tag_len;
    if (nm.add_in_name) { if (L < ad.read_len) x += 5; if (L + ad.lenF < ad.read_len) x += 5; }
    return x;
}
__device__ __forceinline__ uint64_t adpt_rec_bytes(int emit, uint64_t idlen, int RL) {
    switch (emit) {
        case 2: return (idlen + 1 + RL + 1) + (idlen + 2 + 1 + RL + 1);     // STDOUT4
        case 3: case 5: return idlen + 1 + RL + 1;                            // ARTS, FR
        case 6: return idlen + 2 + 1 + RL + 1;                                // RR
        case 4: return idlen + 1 + 2 * (uint64_t)RL + 1;                      // ARTP
        default: return 0;
    }
}
__global__ void __launch_bounds__(256) k_adpt_sizes(const RecStream R, int emit, Adapters ad, AdptNaming nm, unsigned long long* __restrict__ sizes, unsigned long long* stats) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n_rec) return;
    const RecPos p = rec_pos(R, r);
    if (p.ide <= p.ids || R.in[p.ids] != '>') atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
    sizes[r] = adpt_rec_bytes(emit, (p.ide - p.ids) + (uint64_t)adpt_name_extra(nm, ad, (int)(p.se - p.ss)), ad.read_len);
}
cudaError_t launch_adpt_sizes(const RecStream& R, int emit, Adapters ad, AdptNaming nm, unsigned long long* sizes, unsigned long long* stats, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    k_adpt_sizes<<<(unsigned)((R.n_rec + 255) / 256), 256, 0, st>>>(R, emit, ad, nm, sizes, stats);
    return cudaGetLastError();
}

// stand-alone adptSim main loop (adptSim.cpp:263-420): warp per record, ASCII in/out
__global__ void __launch_bounds__(256) k_adpt_records(const RecStream R, const unsigned long long* __restrict__ out_off, uint8_t* __restrict__ out,
                                                      int emit, Adapters ad, const uint8_t* __restrict__ adF, const uint8_t* __restrict__ adS,
                                                      AdptNaming nm, Key key, uint64_t first_id) {
    const int lane = threadIdx.x & 31;
    const int RL = ad.read_len;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R.n_rec; r += nw) {
        const RecPos p = rec_pos(R, r);
        const uint64_t id = first_id + r;
        const int L = (int)(p.se - p.ss);
        const uint64_t idlen = p.ide - p.ids;
        const uint8_t* seq = R.in + p.ss;
        uint8_t* o = out + out_off[r];
        const int nlines = emit == 2 ? 2 : 1;
        for (int ln = 0; ln < nlines; ln++) {
            const bool second = (emit == 6) || (emit == 2 && ln == 1);
            for (uint64_t i = lane; i < idlen; i += 32) o[i] = R.in[p.ids + i];
            o += idlen;
            if (second) { if (lane == 0) { o[0] = '_'; o[1] = 'r'; } o += 2; }       // namer = namef + "_r", adptSim.cpp:286
            if (nm.add_in_name && L < RL) { if (lane < 5) o[lane] = (uint8_t)"_adpt"[lane]; o += 5; }                    // adptSim.cpp:378-379,383-384
            if (nm.add_in_name && L + ad.lenF < RL) { if (lane < 5) o[lane] = (uint8_t)"_rand"[lane]; o += 5; }          // adptSim.cpp:380-381,385-386
            if (lane < nm.tag_len) o[lane] = (uint8_t)nm.tag[lane];                                                      // adptSim.cpp:389-392
            o += nm.tag_len;
            if (lane == 0) o[0] = '\n';
            o += 1;
            const int S = emit == 4 ? 2 * RL : RL;
            for (int j = lane; j < S; j += 32) {
                uint8_t c;
                // S1[j]  (adptSim.cpp:298-310)
                auto S1 = [&](int jj) -> uint8_t {
                    if (jj < L) return upper(seq[jj]);
                    if (jj - L < ad.lenF) return adF[jj - L];
                    const int q = jj - L - ad.lenF;
                    return code2ascii((rng_word(key, id, ST_ADPT_F, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u);
                };
                // SR[q]  (adptSim.cpp:289,313-322)
                auto SR = [&](int qq) -> uint8_t {
                    if (qq < L) return complement_ascii(upper(seq[L - 1 - qq]));
                    if (qq - L < ad.lenS) return adS[qq - L];
                    const int q = qq - L - ad.lenS;
                    return code2ascii((rng_word(key, id, ST_ADPT_R, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u);
                };
                if (second) c = SR(j);
                else if (emit == 4 && j >= RL) c = complement_ascii(SR(2 * RL - 1 - j));   // revcomp(seqr), adptSim.cpp:404
                else c = S1(j);
                o[j] = c;
            }
            o += S;
            if (lane == 0) o[0] = '\n';
            o += 1;
        }
    }
}
cudaError_t launch_adpt_records(const RecStream& R, const unsigned long long* out_off, uint8_t* out, int emit, Adapters ad,
                                const uint8_t* adF, const uint8_t* adS, AdptNaming nm, Key key, uint64_t first_id, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    const unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    k_adpt_records<<<grid, 256, 0, st>>>(R, out_off, out, emit, ad, adF, adS, nm, key, first_id);
    return cudaGetLastError();
}

// small results (byte totals, error flags, histograms) reach the host through SM stores into mapped pinned memory, not through the
// copy engine: a 16-byte D2H memcpy queues behind whatever bulk D2H the egress stream has in flight and would serialise the pipeline
__global__ void k_publish(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst_mapped, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst_mapped[i] = src[i];
    __threadfence_system();
}
cudaError_t launch_publish(const unsigned long long* src, unsigned long long* dst_mapped, int n, cudaStream_t st) {
    k_publish<<<1, 64, 0, st>>>(src, dst_mapped, n); return cudaGetLastError();
}
__global__ void __launch_bounds__(256) k_l2_flush(uint4* __restrict__ buf, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4((uint32_t)i, 0u, 1u, 2u);
}
cudaError_t launch_l2_flush(uint4* buf, uint64_t n16, cudaStream_t st) {
    k_l2_flush<<<148 * 8, 256, 0, st>>>(buf, n16); return cudaGetLastError();
}


// ======================================================================================= device BGZF
// Blocked gzip of a byte stream already resident in HBM (the host gzip hops are the cliff after the kernel).
// Huffman-only deflate: one dynamic code per stream (built on the host from a sampled histogram), every BGZF member = one
// final deflate block of literals.  On simulated reads this matches zlib level 1 (random bases give LZ77 nothing to find).
// One CTA per 65280-byte member: stage the input in shared memory, thread t owns bytes [68t, 68t+68): pass 1 = code-length sum
// + CRC-32 of the run, block scan -> bit offsets, CRC runs combined by polynomial shifts, decoupled look-back over members ->
// byte offset in the output, pass 2 = bit packing into a shared-memory image pre-shifted to the output's 16-byte phase,
// 16-byte stores.  A member that would not shrink is written as a stored block.
__global__ void __launch_bounds__(256) k_byte_hist(const uint8_t* __restrict__ in, uint64_t n, uint64_t group_stride, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int h[8][256];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint64_t n_groups = n >> 4;
    for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * group_stride; g < n_groups; g += (uint64_t)gridDim.x * blockDim.x * group_stride) {
        const uint4 v = __ldg((const uint4*)in + g);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&h[warp][w[k] & 0xFFu], 1u); atomicAdd(&h[warp][(w[k] >> 8) & 0xFFu], 1u);
            atomicAdd(&h[warp][(w[k] >> 16) & 0xFFu], 1u); atomicAdd(&h[warp][w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0) for (uint64_t i = (n_groups << 4) + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&h[warp][in[i]], 1u);   // tail bytes
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        unsigned long long c = 0;
        for (int k = 0; k < 8; k++) c += h[k][i];
        if (c) atomicAdd(hist + i, c);
    }
}
cudaError_t launch_byte_hist(const uint8_t* in, uint64_t n, uint64_t group_stride, unsigned long long* hist, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_byte_hist<<<148 * 4, 256, 0, st>>>(in, n, group_stride ? group_stride : 1, hist);
    return cudaGetLastError();
}

__device__ __forceinline__ uint32_t crc_mul(uint32_t a, uint32_t b) {      // polynomial product mod P, reflected (bit 31 = x^0)
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        p ^= b & (uint32_t)((int32_t)a >> 31);                              // coefficient of x^i in a
        a <<= 1;
        b = (b >> 1) ^ (0xEDB88320u & (0u - (b & 1u)));
    }
    return p;
}
__device__ __forceinline__ void put_bits_atomic(uint32_t* w, uint32_t bitpos, uint32_t v, int nbits) {    // nbits <= 16
    if (nbits <= 0) return;
    const uint32_t sh = bitpos & 31u;
    atomicOr(w + (bitpos >> 5), v << sh);
    if (sh + (uint32_t)nbits > 32u) atomicOr(w + (bitpos >> 5) + 1, v >> (32u - sh));
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void bgzf_prefetch(const BgzfParams& P, uint32_t b, uint8_t* s_dst, int tid) {
    if (b >= P.n_blocks) return;
    const uint64_t base = (uint64_t)b * BGZF_BLOCK;
    const int nb = (int)min((uint64_t)BGZF_BLOCK, P.n - base);
    for (int g = tid; g < BGZF_BLOCK / 16; g += BGZF_THREADS) {
        const int valid = max(0, min(16, nb - 16 * g));                      // bytes past the stream are zero-filled
        cp_async16_zfill(s_dst + 16 * g, P.in + base + (valid ? 16 * g : 0), valid);
    }
}

__global__ void __launch_bounds__(BGZF_THREADS, 1) k_bgzf_deflate(const __grid_constant__ BgzfParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* const s_out = (uint32_t*)(smem + 2 * BGZF_BLOCK);             // 65536 + 64 bytes: the member's payload image, bit 0 = first payload bit
    uint32_t* const s_tab = s_out + (65536 + 64) / 4;                       // BGZF_TAB_WORDS
    uint32_t* const s_scan = s_tab + BGZF_TAB_WORDS;                        // 32 + 32 + 16
    const uint32_t* const s_sym = s_tab + BGZF_TAB_SYM;
    const uint32_t* const s_crc = s_tab + BGZF_TAB_CRC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < BGZF_TAB_WORDS; i += BGZF_THREADS) s_tab[i] = __ldg(P.tab + i);
    const uint32_t eob = __ldg(P.tab + 256);
    // members are handed out in order (look-back never waits on an unstarted one); the next member's input is prefetched with
    // cp.async into the other input buffer while this one is encoded
    if (tid == 0) s_scan[64] = atomicAdd(P.ticket, 1u);
    __syncthreads();
    uint32_t b = s_scan[64];
    int cur = 0;
    bgzf_prefetch(P, b, smem, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    while (b < P.n_blocks) {
        uint8_t* const s_in = smem + cur * BGZF_BLOCK;
        const uint64_t base = (uint64_t)b * BGZF_BLOCK;
        const int nb = (int)min((uint64_t)BGZF_BLOCK, P.n - base);
        if (tid == 0) s_scan[71] = atomicAdd(P.ticket, 1u);
        for (int g = tid; g < (65536 + 64) / 16; g += BGZF_THREADS) ((uint4*)s_out)[g] = make_uint4(0, 0, 0, 0);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        const uint32_t b_next = s_scan[71];
        bgzf_prefetch(P, b_next, smem + (cur ^ 1) * BGZF_BLOCK, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- pass 1: bits and CRC of my run
        const int r0 = tid * BGZF_RUN;
        const int len = max(0, min(BGZF_RUN, nb - r0));
        const uint32_t* const wp = (const uint32_t*)(s_in + r0);
        uint32_t bits = 0, crc = 0xFFFFFFFFu;
        const int nw = len >> 2;                                             // whole words; only the stream's last run has a tail
#pragma unroll 2
        for (int w = 0; w < nw; w++) {
            const uint32_t x = wp[w];
            bits += (s_sym[x & 0xFFu] >> 16) + (s_sym[(x >> 8) & 0xFFu] >> 16) + (s_sym[(x >> 16) & 0xFFu] >> 16) + (s_sym[x >> 24] >> 16);
            crc ^= x;
            crc = s_crc[768 + (crc & 0xFFu)] ^ s_crc[512 + ((crc >> 8) & 0xFFu)] ^ s_crc[256 + ((crc >> 16) & 0xFFu)] ^ s_crc[crc >> 24];
        }
        for (int k = 4 * nw; k < len; k++) {
            const uint32_t c = s_in[r0 + k];
            bits += s_sym[c] >> 16;
            crc = s_crc[(crc ^ c) & 0xFFu] ^ (crc >> 8);
        }
        crc = len ? ~crc : 0u;
        // my run's CRC shifted over the bytes that follow it in the member (crc(A||B) = crc(A) * x^(8|B|) + crc(B))
        if (len) {
            uint32_t sh;
            if (nb == BGZF_BLOCK) sh = s_tab[BGZF_TAB_SHIFT + tid];
            else {
                sh = 0x80000000u;
                uint32_t after = (uint32_t)(nb - r0 - len);
                for (int k = 0; after; k++, after >>= 1) if (after & 1u) sh = crc_mul(sh, s_tab[BGZF_TAB_X8 + k]);
            }
            crc = crc_mul(crc, sh);
        }
        // ---- block scan of bits, XOR-reduction of the shifted CRCs
        uint32_t inc = bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
        if (lane == 31) s_scan[warp] = inc;
        if (lane == 0) s_scan[32 + warp] = crc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = lane < BGZF_THREADS / 32 ? s_scan[lane] : 0u;
            uint32_t wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            uint32_t c = lane < BGZF_THREADS / 32 ? s_scan[32 + lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
            const uint32_t data_bits = __shfl_sync(0xffffffffu, wi, 31);
            s_scan[lane] = wi - v;                                           // exclusive warp offsets
            // member size: published at once so that successors can look back over it
            const uint32_t total_bits = (uint32_t)P.hdr_nbits + data_bits + (eob >> 16);
            uint32_t payload = (total_bits + 7u) >> 3;
            const bool stored = payload >= (uint32_t)nb + 5u;
            if (stored) payload = (uint32_t)nb + 5u;
            if (lane == 0) {
                st_volatile_u64(P.state + b, ((b == 0 ? 2ull : 1ull) << 62) | (18ull + payload + 8ull));
                s_scan[67] = payload; s_scan[68] = stored ? 1u : 0u; s_scan[69] = c; s_scan[70] = data_bits;
            }
        }
        __syncthreads();
        const uint32_t payload = s_scan[67]; const bool stored = s_scan[68] != 0; const uint32_t crc_all = s_scan[69], data_bits = s_scan[70];
        if (!stored) {
            // ---- pass 2: pack my run's codes at my bit offset (first and last word shared with the neighbours: atomicOr)
            const uint32_t bitpos = (uint32_t)P.hdr_nbits + s_scan[warp] + (inc - bits);
            uint32_t* wout = s_out + (bitpos >> 5);
            // pending bits live in lo (bits [0,fill) of the current word, fill < 32 between steps) and hi (the spill of one step);
            // two codes (<= 30 bits) are merged in 32-bit arithmetic and enter with one shift + one funnel shift
            uint32_t lo = 0, hi = 0, fill = bitpos & 31u; bool first = true;
#define BGZF_PUT(p, n) { hi = __funnelshift_l((p), 0u, fill); lo |= (p) << fill; fill += (n); \
                if (fill >= 32u) { if (first) { atomicOr(wout, lo); first = false; } else *wout = lo; wout++; lo = hi; fill -= 32u; } }
#pragma unroll 2
            for (int w = 0; w < nw; w++) {
                const uint32_t x = wp[w];
                const uint32_t s0 = s_sym[x & 0xFFu], s1 = s_sym[(x >> 8) & 0xFFu], s2 = s_sym[(x >> 16) & 0xFFu], s3 = s_sym[x >> 24];
                const uint32_t l0 = s0 >> 16, l2 = s2 >> 16;
                const uint32_t p01 = (s0 & 0xFFFFu) | ((s1 & 0xFFFFu) << l0), p23 = (s2 & 0xFFFFu) | ((s3 & 0xFFFFu) << l2);
                BGZF_PUT(p01, l0 + (s1 >> 16));
                BGZF_PUT(p23, l2 + (s3 >> 16));
            }
            for (int k = 4 * nw; k < len; k++) {
                const uint32_t sy = s_sym[s_in[r0 + k]];
                BGZF_PUT(sy & 0xFFFFu, sy >> 16);
            }
#undef BGZF_PUT
            if (fill > (first ? (bitpos & 31u) : 0u)) atomicOr(wout, lo);
            // the code description in front and the end-of-block code behind
            if (tid >= 32 && tid < 32 + BGZF_HDR_WORDS * 2) {
                const int piece = tid - 32;                                 // 16 bits each
                const int nbp = min(16, P.hdr_nbits - 16 * piece);
                if (nbp > 0) put_bits_atomic(s_out, 16u * (uint32_t)piece, (s_tab[BGZF_TAB_HDR + (piece >> 1)] >> (16 * (piece & 1))) & (0xFFFFu >> (16 - nbp)), nbp);
            }
            if (tid == BGZF_THREADS - 1) put_bits_atomic(s_out, (uint32_t)P.hdr_nbits + data_bits, eob & 0xFFFFu, (int)(eob >> 16));
        }
        // ---- the member's place in the output: decoupled look-back over the members before it (by now they have published)
        if (warp == 0) {
            const unsigned long long S = 18ull + payload + 8ull;
            unsigned long long gofs = 0;
            if (b > 0) {
                long long look = (long long)b - 1;
                while (true) {
                    const long long i = look - lane;
                    unsigned long long lb = 2ull << 62;
                    if (i >= 0) { do { lb = ld_volatile_u64(P.state + i); } while ((lb >> 62) == 0); }
                    const uint32_t incl = __ballot_sync(0xffffffffu, (lb >> 62) == 2ull);
                    const int first_inc = incl ? (__ffs(incl) - 1) : 32;
                    gofs += warp_sum_u64(lane <= first_inc ? (lb & 0x3FFFFFFFFFFFFFFFull) : 0ull);
                    if (incl) break;
                    look -= 32;
                }
                if (lane == 0) st_volatile_u64(P.state + b, (2ull << 62) | (gofs + S));
            }
            if (lane == 0) {
                if (b == P.n_blocks - 1) *P.total = gofs + S;
                s_scan[65] = (uint32_t)gofs; s_scan[66] = (uint32_t)(gofs >> 32);
            }
        }
        __syncthreads();
        const unsigned long long gofs = (unsigned long long)s_scan[65] | ((unsigned long long)s_scan[66] << 32);
        if (gofs + 18ull + payload + 8ull <= P.out_cap) {
            // ---- member header (RFC 1952 + the BGZF "BC" extra field), payload, CRC-32 and ISIZE
            uint8_t* const dst = P.out + gofs;
            if (tid < 18) {
                const uint32_t bsize1 = 18u + payload + 8u - 1u;
                const uint8_t h[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)(bsize1 & 0xFFu), (uint8_t)(bsize1 >> 8)};
                dst[tid] = h[tid];
            } else if (tid >= 32 && tid < 40) {
                const int k = tid - 32;
                dst[18 + payload + k] = (uint8_t)((k < 4 ? crc_all >> (8 * k) : (uint32_t)nb >> (8 * (k - 4))) & 0xFFu);
            }
            uint8_t* const pay = dst + 18;
            if (!stored) {
                // 16-byte stores at the output's own alignment: group g holds payload bytes [16g - A, 16g - A + 16), A = pay & 15,
                // assembled from the image with byte funnel-shifts (the image starts at payload byte 0)
                const uint32_t A = (uint32_t)((uintptr_t)pay & 15u);
                const uint32_t end = A + payload;
                const uint8_t* img = (const uint8_t*)s_out;
                for (uint32_t g = tid; 16u * g < end; g += BGZF_THREADS) {
                    if (16u * g >= A && 16u * g + 16u <= end) {
                        const uint32_t s0 = 16u * g - A;                    // first payload byte of the group
                        const uint32_t* wsrc = s_out + (s0 >> 2);
                        const uint32_t sh = 8u * (s0 & 3u);
                        uint4 v;
                        v.x = __funnelshift_r(wsrc[0], wsrc[1], sh); v.y = __funnelshift_r(wsrc[1], wsrc[2], sh);
                        v.z = __funnelshift_r(wsrc[2], wsrc[3], sh); v.w = __funnelshift_r(wsrc[3], wsrc[4], sh);
                        *(uint4*)(pay - A + 16u * g) = v;
                    } else for (uint32_t k = max(16u * g, A); k < min(16u * g + 16u, end); k++) pay[k - A] = img[k - A];
                }
            } else {
                if (tid < 5) { const uint32_t L = (uint32_t)nb; const uint8_t sh[5] = {1, (uint8_t)(L & 0xFFu), (uint8_t)(L >> 8), (uint8_t)(~L & 0xFFu), (uint8_t)((~L >> 8) & 0xFFu)}; pay[tid] = sh[tid]; }
                for (int k = tid; k < nb; k += BGZF_THREADS) pay[5 + k] = s_in[k];
            }
        } else if (tid == 0) atomicOr(P.total + 1, 1ull);                   // capacity error flag
        __syncthreads();
        b = b_next; cur ^= 1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}
cudaError_t bgzf_prepare(size_t* smem_bytes) {
    const size_t sm = 2 * (size_t)BGZF_BLOCK + 65536 + 64 + (size_t)BGZF_TAB_WORDS * 4 + 80 * 4;
    *smem_bytes = sm;
    return cudaFuncSetAttribute(k_bgzf_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
}
cudaError_t launch_bgzf_deflate(const BgzfParams& P, int grid, cudaStream_t st) {
    size_t sm = 0;
    cudaError_t e = bgzf_prepare(&sm);
    if (e != cudaSuccess) return e;
    k_bgzf_deflate<<<grid, BGZF_THREADS, sm, st>>>(P);
    return cudaGetLastError();
}

// ======================================================================================= device BGZF with LZ77 matches
// The same member framing as k_bgzf_deflate, but a member's deflate block carries matches: one warp parses one 2176-byte run of the member
// in steps of 32 positions (hash of the next 6 bytes -> 512-entry table of the most recent earlier positions, greedy selection, byte-wise
// extension; gg_lz.h is the specification and ggh::bgzf_lz_model its host model: the output is the model's byte for byte).  A warp
// appends its codes to a private region of shared memory; once the bit counts of the 30 runs are scanned, the member's payload is
// gathered from the header bits, the regions and the end-of-block code with funnel shifts, straight into 16-byte global stores.
// COUNT: the sampling pass in front (every stride-th member): the same parse, symbol counts instead of codes.
constexpr int BGZL_REGION_WORDS = gglz::REGION / 4;
constexpr int BGZL_PROV_BYTES = 30 * gglz::REGION + 16;
constexpr int BGZL_SCAN_WORDS = 160;
// 6 bytes at byte offset p of a word-aligned shared-memory buffer: low four, then the next two
__device__ __forceinline__ void lz_load6(const uint8_t* buf, int p, uint32_t& lo, uint32_t& hi16) {
    const uint32_t* wp = (const uint32_t*)(buf + (p & ~3));
    const uint32_t sh = 8u * ((uint32_t)p & 3u);
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    lo = __funnelshift_r(w0, w1, sh); hi16 = __funnelshift_r(w1, w2, sh) & 0xFFFFu;
}
// segment k of a member's payload: 0 = code description, 1..30 = the runs' regions, 31 = end-of-block code
__device__ __forceinline__ const uint32_t* lz_seg_src(int k, const uint32_t* s_tab, const uint32_t* s_prov, const uint32_t* s_scan) {
    return k == 0 ? s_tab + BGZL_TAB_HDR : k == 31 ? s_scan + 120 : s_prov + (k - 1) * BGZL_REGION_WORDS;
}
// 32 bits of the payload from bit t on (zeros behind its end); S[k] = first bit of segment k, S[32] = total; k: the caller's running segment index
__device__ __forceinline__ uint32_t lz_gather32(uint32_t t, int& k, const uint32_t* S, const uint32_t* s_tab, const uint32_t* s_prov, const uint32_t* s_scan) {
    while (k < 32 && S[k + 1] <= t) k++;
    uint32_t out = 0, got = 0, tt = t; int kk = k;
    while (got < 32u && kk < 32) {
        const uint32_t o = tt - S[kk], avail = S[kk + 1] - tt;
        const uint32_t* src = lz_seg_src(kk, s_tab, s_prov, s_scan) + (o >> 5);
        uint32_t v = __funnelshift_r(src[0], src[1], o & 31u);
        const uint32_t take = min(avail, 32u - got);
        if (take < 32u) v &= (1u << take) - 1u;
        out |= v << got; got += take; tt += take;
        kk++; while (kk < 32 && S[kk + 1] <= tt) kk++;
    }
    return out;
}

template <bool COUNT>
__global__ void __launch_bounds__(BGZF_THREADS, 1) k_bgzf_lz(const __grid_constant__ BgzfLzParams P) {
    using namespace gglz;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* const s_prov = (uint32_t*)(smem + 2 * BGZF_BLOCK);            // 30 regions of REGION bytes: the runs' codes, bit 0 = first bit of the run
    uint32_t* const s_tab = s_prov + BGZL_PROV_BYTES / 4;
    uint16_t* const s_hash = (uint16_t*)(s_tab + BGZL_TAB_WORDS);           // 30 tables of TABLE positions
    uint32_t* const s_scan = (uint32_t*)(s_hash + 30 * TABLE);              // BGZL_SCAN_WORDS: [0,30) run bits, [32,62) run CRCs, [64,97) segment starts S, misc from 100, [120,122) end-of-block code
    uint32_t* const s_len = s_scan + BGZL_SCAN_WORDS;                       // 256: match length - 3 -> length code with its extra bits | total bits << 24
    uint32_t* const s_hist = s_prov;                                        // COUNT: 316 symbol counters
    const uint32_t* const s_sym = s_tab + BGZL_TAB_SYM;
    const uint32_t* const s_crc = s_tab + BGZL_TAB_CRC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!COUNT) {
        for (int i = tid; i < BGZL_TAB_WORDS; i += BGZF_THREADS) s_tab[i] = __ldg(P.tab + i);
    } else {
        for (int i = tid; i < N_SYM; i += BGZF_THREADS) s_hist[i] = 0u;
    }
    if (tid == 0) s_scan[100] = atomicAdd(P.ticket, 1u);
    __syncthreads();
    const uint32_t eob = COUNT ? 0u : s_sym[256];
    if (tid == 0) { s_scan[120] = eob & 0xFFFFu; s_scan[121] = 0u; }
    if (!COUNT && tid < 256) {
        int ls, lnx; uint32_t lxv;
        len_symbol(tid + 3, 31 - __clz(tid | 1), &ls, &lnx, &lxv);
        const uint32_t sy = s_sym[ls], n1 = sy >> 16;
        s_len[tid] = (sy & 0xFFFFu) | (lxv << n1) | ((n1 + (uint32_t)lnx) << 24);
    }
    __syncthreads();
    uint32_t v = s_scan[100];                                                // visit number; member index = visit * stride
    int cur = 0;
    auto prefetch = [&](uint32_t visit, uint8_t* s_dst) {
        if (visit >= P.n_blocks) return;
        const uint64_t base = (uint64_t)visit * P.stride * BGZF_BLOCK;
        const int nbp = (int)min((uint64_t)BGZF_BLOCK, P.n - base);
        for (int g = tid; g < BGZF_BLOCK / 16; g += BGZF_THREADS) {
            const int valid = max(0, min(16, nbp - 16 * g));
            cp_async16_zfill(s_dst + 16 * g, P.in + base + (valid ? 16 * g : 0), valid);
        }
    };
    prefetch(v, smem);
    asm volatile("cp.async.commit_group;" ::: "memory");

    while (v < P.n_blocks) {
        uint8_t* const s_in = smem + cur * BGZF_BLOCK;
        const uint64_t b = (uint64_t)v * P.stride;
        const uint64_t base = b * BGZF_BLOCK;
        const int nb = (int)min((uint64_t)BGZF_BLOCK, P.n - base);
        if (tid == 0) s_scan[101] = atomicAdd(P.ticket, 1u);
        if (!COUNT) for (int g = tid; g < BGZL_PROV_BYTES / 16; g += BGZF_THREADS) ((uint4*)s_prov)[g] = make_uint4(0, 0, 0, 0);
        {   // this warp's table: empty
            uint4* T4 = (uint4*)(s_hash + warp * TABLE);
            for (int g = lane; g < TABLE * 2 / 16; g += 32) T4[g] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        const uint32_t v_next = s_scan[101];
        prefetch(v_next, smem + (cur ^ 1) * BGZF_BLOCK);
        asm volatile("cp.async.commit_group;" ::: "memory");

        // ---- CRC-32 of my 68-byte run, shifted over the bytes behind it (as in k_bgzf_deflate)
        uint32_t crc = 0;
        if (!COUNT) {
            const int r0 = tid * BGZF_RUN;
            const int len = max(0, min(BGZF_RUN, nb - r0));
            const uint32_t* const wp = (const uint32_t*)(s_in + r0);
            crc = 0xFFFFFFFFu;
            const int nw = len >> 2;
#pragma unroll 2
            for (int w = 0; w < nw; w++) {
                crc ^= wp[w];
                crc = s_crc[768 + (crc & 0xFFu)] ^ s_crc[512 + ((crc >> 8) & 0xFFu)] ^ s_crc[256 + ((crc >> 16) & 0xFFu)] ^ s_crc[crc >> 24];
            }
            for (int k = 4 * nw; k < len; k++) crc = s_crc[(crc ^ s_in[r0 + k]) & 0xFFu] ^ (crc >> 8);
            crc = len ? ~crc : 0u;
            if (len) {
                uint32_t sh;
                if (nb == BGZF_BLOCK) sh = s_tab[BGZL_TAB_SHIFT + tid];
                else {
                    sh = 0x80000000u;
                    uint32_t after = (uint32_t)(nb - r0 - len);
                    for (int k = 0; after; k++, after >>= 1) if (after & 1u) sh = crc_mul(sh, s_tab[BGZL_TAB_X8 + k]);
                }
                crc = crc_mul(crc, sh);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
        }

        // ---- parse of my warp's run (gg_lz.h), codes appended to the warp's region
        uint32_t cursor = 0; bool overflow = false;
        {
            const int start = warp * RUN, end = min(nb, start + RUN);
            uint16_t* const T = s_hash + warp * TABLE;
            uint32_t* const reg = s_prov + warp * BGZL_REGION_WORDS;
            __syncwarp();
            if (start < end) {
                if (warp > 0) {                                              // the 160 positions in front of the run: inserted, not looked up
#pragma unroll 1
                    for (int bs = start - WARM; bs < start; bs += 32) {
                        const int p = bs + lane;
                        const bool ins = p + MIN_MATCH <= nb;
                        uint32_t lo, hi; lz_load6(s_in, p, lo, hi);
                        const uint32_t h = lz_hash(lo, hi);
                        if (ins) T[h] = (uint16_t)p;
                        __syncwarp();
                        bool again = ins && (int)T[h] < p;                   // of equal hashes the highest position stays
                        while (__any_sync(0xffffffffu, again)) { if (again) T[h] = (uint16_t)p; __syncwarp(); again = ins && (int)T[h] < p; }
                        __syncwarp();
                    }
                }
                int cover = start;
                // a step's look-ups are issued one step ahead: behind the previous step's inserts (the table state the specification asks for),
                // in front of its selection and code assembly, whose shared-memory round trips they overlap
                int p = start + lane;
                bool active = p < end, hv = p + MIN_MATCH <= end;
                uint32_t lo = 0, hi = 0;
                if (active) lz_load6(s_in, p, lo, hi);
                uint32_t h = lz_hash(lo, hi);
                uint32_t c = hv ? (uint32_t)T[h] : 0xFFFFu;
#pragma unroll 1
                for (int bs = start; bs < end; bs += 32) {
                    // inserts of this step: of equal hashes the highest position stays (store, read back, the larger ones store again)
                    __syncwarp();
                    if (hv) T[h] = (uint16_t)p;
                    __syncwarp();
                    {
                        bool again = hv && (int)T[h] < p;
                        while (__any_sync(0xffffffffu, again)) { if (again) T[h] = (uint16_t)p; __syncwarp(); again = hv && (int)T[h] < p; }
                    }
                    const int np = p + 32;
                    const bool n_active = np < end, n_hv = np + MIN_MATCH <= end;
                    uint32_t n_lo = 0, n_hi = 0;
                    if (n_active) lz_load6(s_in, np, n_lo, n_hi);
                    const uint32_t n_h = lz_hash(n_lo, n_hi);
                    const uint32_t n_c = n_hv ? (uint32_t)T[n_h] : 0xFFFFu;
                    if (cover < min(bs + 32, end)) {                         // (a step inside an earlier match emits nothing)
                    bool m = false; int dist = 0;
                    if (c != 0xFFFFu && p >= cover && p - (int)c <= 32768) {
                        uint32_t clo, chi; lz_load6(s_in, (int)c, clo, chi);
                        m = clo == lo && chi == hi; dist = p - (int)c;
                    }
                    bool is_lit = false, is_match = false; int mylen = 0;
                    while (true) {
                        const uint32_t mm = __ballot_sync(0xffffffffu, m && p >= cover);
                        if (!mm) break;
                        const int L = __ffs((int)mm) - 1, q = bs + L;
                        const int d = __shfl_sync(0xffffffffu, dist, L);
                        const int maxlen = min(MAX_MATCH, end - q);
                        int len = maxlen;
#pragma unroll 1
                        for (int k0 = MIN_MATCH; k0 < maxlen; k0 += 32) {
                            const int k = k0 + lane;
                            const bool eq = k < maxlen && s_in[q + k] == s_in[q + k - d];
                            const uint32_t ne = __ballot_sync(0xffffffffu, !eq);
                            if (ne) { len = k0 + __ffs((int)ne) - 1; break; }
                        }
                        is_lit |= p >= cover && p < q;
                        if (lane == L) { is_match = true; mylen = len; }
                        cover = q + len;
                    }
                    is_lit |= active && p >= cover;
                    // codes of this step
                    unsigned long long tb = 0; uint32_t tn = 0;
                    if (is_lit) {
                        if (COUNT) atomicAdd(&s_hist[lo & 0xFFu], 1u);
                        else { const uint32_t sy = s_sym[lo & 0xFFu]; tb = sy & 0xFFFFu; tn = sy >> 16; }
                    } else if (is_match) {
                        int ds, dnx; uint32_t dxv;
                        dist_symbol(dist, 31 - __clz((dist - 1) | 1), &ds, &dnx, &dxv);
                        if (COUNT) {
                            int ls, lnx; uint32_t lxv;
                            len_symbol(mylen, 31 - __clz(mylen - 3), &ls, &lnx, &lxv);
                            atomicAdd(&s_hist[ls], 1u); atomicAdd(&s_hist[N_LITLEN + ds], 1u);
                        } else {
                            const uint32_t s1 = s_len[mylen - 3], s2 = s_sym[N_LITLEN + ds];      // length code with its extra bits | bits << 24
                            const uint32_t n2 = s2 >> 16;
                            tn = s1 >> 24;
                            const uint32_t dbits = (s2 & 0xFFFFu) | (dxv << n2);                  // <= 28 bits
                            tb = (unsigned long long)(s1 & 0xFFFFFFu) | ((unsigned long long)dbits << tn);
                            tn += n2 + (uint32_t)dnx;
                        }
                    }
                    if (!COUNT) {
                        uint32_t inc = tn;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
                        if (cursor + tot > (uint32_t)(REGION * 8)) { overflow = true; break; }       // (warp-uniform)
                        if (tn) {
                            const uint32_t pos = cursor + inc - tn, sh = pos & 31u;
                            uint32_t* w = reg + (pos >> 5);
                            const unsigned long long lo64 = tb << sh;
                            atomicOr(w, (uint32_t)lo64);
                            if (sh + tn > 32u) atomicOr(w + 1, (uint32_t)(lo64 >> 32));
                            if (sh + tn > 64u) atomicOr(w + 2, (uint32_t)(tb >> (64u - sh)));
                        }
                        cursor += tot;
                    }
                    }
                    p = np; active = n_active; hv = n_hv; lo = n_lo; hi = n_hi; h = n_h; c = n_c;
                }
            }
        }
        if (COUNT) {
            if (tid == 0) atomicAdd(&s_hist[256], 1u);
            __syncthreads();
            v = v_next; cur ^= 1;
            continue;
        }
        if (lane == 0) { s_scan[warp] = overflow ? 0x80000000u : cursor; s_scan[32 + warp] = crc; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t raw = lane < 30 ? s_scan[lane] : 0u;
            const uint32_t ovf = __ballot_sync(0xffffffffu, (raw & 0x80000000u) != 0u);
            const uint32_t bits = raw & 0x7FFFFFFFu;
            uint32_t wi = bits;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            uint32_t c = lane < 30 ? s_scan[32 + lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
            const uint32_t data_bits = __shfl_sync(0xffffffffu, wi, 31);
            // segment starts: S[0] = 0 (description), S[1 + w] = first bit of run w, S[31] = end-of-block code, S[32] = total
            if (lane < 30) s_scan[64 + 1 + lane] = (uint32_t)P.hdr_nbits + wi - bits;
            const uint32_t total_bits = (uint32_t)P.hdr_nbits + data_bits + (eob >> 16);
            uint32_t payload = (total_bits + 7u) >> 3;
            const bool stored = ovf != 0u || payload >= (uint32_t)nb + 5u;
            if (stored) payload = (uint32_t)nb + 5u;
            if (lane == 0) {
                s_scan[64] = 0u; s_scan[64 + 31] = (uint32_t)P.hdr_nbits + data_bits; s_scan[64 + 32] = total_bits;
                st_volatile_u64(P.state + b, ((b == 0 ? 2ull : 1ull) << 62) | (18ull + payload + 8ull));
                s_scan[102] = payload; s_scan[103] = stored ? 1u : 0u; s_scan[104] = c;
            }
            // ---- the member's place in the output: decoupled look-back over the members before it
            const unsigned long long Sz = 18ull + payload + 8ull;
            unsigned long long gofs = 0;
            if (b > 0) {
                long long look = (long long)b - 1;
                while (true) {
                    const long long i = look - lane;
                    unsigned long long lb = 2ull << 62;
                    if (i >= 0) { do { lb = ld_volatile_u64(P.state + i); } while ((lb >> 62) == 0); }
                    const uint32_t incl = __ballot_sync(0xffffffffu, (lb >> 62) == 2ull);
                    const int first_inc = incl ? (__ffs(incl) - 1) : 32;
                    gofs += warp_sum_u64(lane <= first_inc ? (lb & 0x3FFFFFFFFFFFFFFFull) : 0ull);
                    if (incl) break;
                    look -= 32;
                }
                if (lane == 0) st_volatile_u64(P.state + b, (2ull << 62) | (gofs + Sz));
            }
            if (lane == 0) {
                if (b == (uint64_t)P.n_blocks - 1) *P.total = gofs + Sz;
                s_scan[105] = (uint32_t)gofs; s_scan[106] = (uint32_t)(gofs >> 32);
            }
        }
        __syncthreads();
        const uint32_t payload = s_scan[102]; const bool stored = s_scan[103] != 0; const uint32_t crc_all = s_scan[104];
        const unsigned long long gofs = (unsigned long long)s_scan[105] | ((unsigned long long)s_scan[106] << 32);
        if (gofs + 18ull + payload + 8ull <= P.out_cap) {
            uint8_t* const dst = P.out + gofs;
            if (tid < 18) {
                const uint32_t bsize1 = 18u + payload + 8u - 1u;
                const uint8_t h[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)(bsize1 & 0xFFu), (uint8_t)(bsize1 >> 8)};
                dst[tid] = h[tid];
            } else if (tid >= 32 && tid < 40) {
                const int k = tid - 32;
                dst[18 + payload + k] = (uint8_t)((k < 4 ? crc_all >> (8 * k) : (uint32_t)nb >> (8 * (k - 4))) & 0xFFu);
            }
            uint8_t* const pay = dst + 18;
            if (!stored) {
                // 16-byte stores at the output's own alignment: group g holds payload bytes [16g - A, 16g - A + 16), A = pay & 15
                const uint32_t* const S = s_scan + 64;
                const uint32_t A = (uint32_t)((uintptr_t)pay & 15u);
                const uint32_t endb = A + payload;
                int k = 0;
                for (uint32_t g = tid; 16u * g < endb; g += BGZF_THREADS) {
                    if (16u * g >= A && 16u * g + 16u <= endb) {
                        const uint32_t t0 = 8u * (16u * g - A);
                        uint4 o;
                        o.x = lz_gather32(t0, k, S, s_tab, s_prov, s_scan); o.y = lz_gather32(t0 + 32u, k, S, s_tab, s_prov, s_scan);
                        o.z = lz_gather32(t0 + 64u, k, S, s_tab, s_prov, s_scan); o.w = lz_gather32(t0 + 96u, k, S, s_tab, s_prov, s_scan);
                        *(uint4*)(pay - A + 16u * g) = o;
                    } else {
                        for (uint32_t i = max(16u * g, A); i < min(16u * g + 16u, endb); i++) {
                            int k2 = 0;
                            pay[i - A] = (uint8_t)lz_gather32(8u * (i - A), k2, S, s_tab, s_prov, s_scan);
                        }
                    }
                }
            } else {
                if (tid < 5) { const uint32_t L = (uint32_t)nb; const uint8_t sh[5] = {1, (uint8_t)(L & 0xFFu), (uint8_t)(L >> 8), (uint8_t)(~L & 0xFFu), (uint8_t)((~L >> 8) & 0xFFu)}; pay[tid] = sh[tid]; }
                for (int k = tid; k < nb; k += BGZF_THREADS) pay[5 + k] = s_in[k];
            }
        } else if (tid == 0) atomicOr(P.total + 1, 1ull);
        __syncthreads();
        v = v_next; cur ^= 1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (COUNT) {
        __syncthreads();
        for (int i = tid; i < N_SYM; i += BGZF_THREADS) if (s_hist[i]) atomicAdd(P.hist + i, (unsigned long long)s_hist[i]);
    }
}
cudaError_t launch_bgzf_lz(const BgzfLzParams& P, bool count_only, int grid, cudaStream_t st) {
    const size_t sm = 2 * (size_t)BGZF_BLOCK + BGZL_PROV_BYTES + (size_t)BGZL_TAB_WORDS * 4 + 30 * gglz::TABLE * 2 + BGZL_SCAN_WORDS * 4 + 256 * 4;
    cudaError_t e = cudaFuncSetAttribute(count_only ? k_bgzf_lz<true> : k_bgzf_lz<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    if (count_only) k_bgzf_lz<true><<<grid, BGZF_THREADS, sm, st>>>(P); else k_bgzf_lz<false><<<grid, BGZF_THREADS, sm, st>>>(P);
    return cudaGetLastError();
}

} // namespace ggk
