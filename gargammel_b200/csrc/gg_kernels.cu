// gg_kernels.cu -- hand-written sm_100a kernels of the gargammel hot path.
//
//   k_pack_genome   FASTA bytes -> 2-bit codes + non-ACGT mask (replaces mmap + per-char fetchSeq, fragSim.cpp:305-329)
//   k_fused         fragSim -> deamSim -> adptSim for a tile of fragments per CTA iteration, ONE kernel:
//                   thread-per-fragment rejection sampling (fragSim.cpp:1361-1507), thread-per-16-base-chunk damage
//                   (deamSim.cpp:1477-1613,1794-2002), thread-per-16-byte-group adapter/emit (adptSim.cpp:281-411), records
//                   staged in shared memory, global byte offsets from a decoupled look-back scan over tiles, and
//                   16-byte vector stores of the final FASTA byte stream.
//   k_deam_records / k_adpt_records   the stand-alone stages over a 2-line FASTA record stream (for the drop-in CLIs)
//
// HBM-bound integer/byte work: no tensor cores.  Layout and byte accounting: DESIGN.md.
#include "gg_kernels.cuh"

namespace ggk {
using namespace ggd;

// ======================================================================================= pack
// one thread per 32 bases of the packed base space (2 seq2 words + 1 mask word)
__global__ void __launch_bounds__(256) k_pack_genome(const uint8_t* __restrict__ fasta, uint64_t n,
        const uint64_t* __restrict__ c_off, const uint32_t* __restrict__ c_len,
        const int32_t* __restrict__ c_blen, const int32_t* __restrict__ c_llen,
        const uint64_t* __restrict__ c_gbase, int n_contigs, uint64_t n_units,
        uint32_t* __restrict__ seq2, uint32_t* __restrict__ nmask) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
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
            uint64_t line = local / blen, col = local % blen;
            uint64_t pos = c_off[c] + line * llen + col;     // fetchSeq, fragSim.cpp:312-314
            m = 0;
            for (int t = 0; t < 32; t++) {
                uint32_t code = 0, bad = 1;
                if (local + t < len) {
                    const uint8_t ch = pos < n ? upper(fasta[pos]) : 0;   // toupper, fragSim.cpp:315
                    const int cd = ascii2code(ch);
                    if (cd >= 0) { code = (uint32_t)cd; bad = 0; }
                    pos++; col++;
                    if (col == blen) { col = 0; pos += llen - blen; }
                }
                if (t < 16) w0 |= code << (2 * t); else w1 |= code << (2 * (t - 16));
                m |= bad << t;
            }
        }
    }
    seq2[2 * u] = w0; seq2[2 * u + 1] = w1; nmask[u] = m;
}

cudaError_t launch_pack_genome(const uint8_t* fasta, uint64_t n, const uint64_t* c_off, const uint32_t* c_len,
                               const int32_t* c_blen, const int32_t* c_llen, const uint64_t* c_gbase,
                               const uint64_t*, int n_contigs, uint64_t n_units,
                               uint32_t* seq2, uint32_t* nmask, cudaStream_t st) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t blocks = (n_units + 255) / 256;
    k_pack_genome<<<(unsigned)blocks, 256, 0, st>>>(fasta, n, c_off, c_len, c_blen, c_llen, c_gbase, n_contigs, n_units, seq2, nmask);
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

// positions p..p+15 of a packed smem source; positions < 0 read as garbage-free zeros
__device__ __forceinline__ uint32_t fetch_run_smem(const uint32_t* src, int p) {
    return p >= 0 ? fetch16_smem(src, p) : (fetch16_smem(src, 0) << (2 * (-p)));
}
__device__ __forceinline__ uint32_t fetch_run_glob(const uint32_t* __restrict__ src, int p) {
    return p >= 0 ? fetch16(src, p) : (fetch16(src, 0) << (2 * (-p)));
}
// group-relative mask of [lo,hi) clipped to [0,16)
__device__ __forceinline__ uint32_t run_mask(int lo, int hi) {
    lo = min(max(lo, 0), 16); hi = min(max(hi, 0), 16);
    return lowfields(hi) & ~lowfields(lo);          // empty when hi <= lo
}

// 16 output codes for positions j0..j0+15 of sequence line `line` of one fragment
// (adptSim.cpp:288-322,396-411 restated as runs over 2-bit sources; D = damaged fragment in smem)
__device__ __forceinline__ uint32_t line_codes(const FusedParams& P, const uint32_t* D, int L, int line, int j0,
                                               uint64_t id) {
    const int emit = P.emit, RL = P.ad.read_len;
    if (emit <= 1) return fetch_run_smem(D, j0);                           // FRAG / DEAM: the fragment itself
    uint32_t out = 0;
    const bool second = (emit == 6) || (emit == 2 && line == 1);           // the reverse read line (RR, STDOUT4 line 1)
    if (!second) {
        // forward read S1 = (frag + adapterF)[0:RL] + pad      adptSim.cpp:298-310
        const int nD = min(L, RL);
        uint32_t m = run_mask(0 - j0, nD - j0);
        if (m) out |= fetch_run_smem(D, j0) & m;
        const int nA = min(max(RL - L, 0), P.ad.lenF);
        m = run_mask(L - j0, L + nA - j0);
        if (m) out |= fetch_run_glob(P.ad.adF2, j0 - L) & m;
        const int padStart = L + nA;                                        // random tail, adptSim.cpp:307-310
        if (padStart < RL) {
            for (int t = max(padStart - j0, 0); t < min(RL - j0, 16); t++) {
                const int q = j0 + t - padStart;
                const uint32_t c = (rng_word(P.key, id, ST_ADPT_F, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u;
                out |= c << (2 * t);
            }
        }
        if (emit == 4) {
            // ARTP second half = revcomp(SR), positions RL..2RL-1      adptSim.cpp:404-405
            const int mAd = max(RL - L, 0);                 // adapter+pad bases of SR
            const int nS = min(mAd, P.ad.lenS);             // adapter bases
            const int nP = mAd - nS;                        // pad bases
            // [RL, RL+nP): complement of padR reversed
            if (nP > 0)
            for (int t = max(RL - j0, 0); t < min(RL + nP - j0, 16); t++) {
                const int q = nP - 1 - (j0 + t - RL);
                const uint32_t c = (rng_word(P.key, id, ST_ADPT_R, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u;
                out |= (3u - c) << (2 * t);
            }
            // [RL+nP, RL+mAd): suffix of revcomp(adapterS)
            m = run_mask(RL + nP - j0, RL + mAd - j0);
            if (m) out |= fetch_run_glob(P.ad.rcS2, j0 - (RL + nP) + (P.ad.lenS - nS)) & m;
            // [RL+mAd, 2RL): fragment forward from max(L-RL,0)
            m = run_mask(RL + mAd - j0, 2 * RL - j0);
            if (m) out |= fetch_run_smem(D, j0 - (RL + mAd) + max(L - RL, 0)) & m;
        }
    } else {
        // reverse read SR = (revcomp(frag) + adapterS)[0:RL] + pad      adptSim.cpp:313-322
        const int nD = min(L, RL);
        uint32_t m = run_mask(0 - j0, nD - j0);
        if (m) out |= revcomp16(fetch_run_smem(D, L - 16 - j0)) & m;
        const int nA = min(max(RL - L, 0), P.ad.lenS);
        m = run_mask(L - j0, L + nA - j0);
        if (m) out |= fetch_run_glob(P.ad.adS2, j0 - L) & m;
        const int padStart = L + nA;
        if (padStart < RL) {
            for (int t = max(padStart - j0, 0); t < min(RL - j0, 16); t++) {
                const int q = j0 + t - padStart;
                const uint32_t c = (rng_word(P.key, id, ST_ADPT_R, (uint32_t)(q >> 4)) >> (2 * (q & 15))) & 3u;
                out |= c << (2 * t);
            }
        }
    }
    return out;
}

__device__ __forceinline__ uint8_t* put_decimal(uint8_t* p, uint64_t v, int nd) {
    int k = nd - 1;
    for (; v > 0xFFFFFFFFull; k--) { const uint64_t q = v / 10ull; p[k] = (uint8_t)('0' + (uint32_t)(v - q * 10ull)); v = q; }
    uint32_t w = (uint32_t)v;
    for (; k >= 0; k--) { const uint32_t q = w / 10u; p[k] = (uint8_t)('0' + (w - q * 10u)); w = q; }
    return p + nd;
}

__global__ void __launch_bounds__(FUSED_THREADS, 3) k_fused(const __grid_constant__ FusedParams P) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t*  const s_stage  = smem + P.sm.stage;
    uint64_t* const s_gpos   = (uint64_t*)(smem + P.sm.gpos);
    uint32_t* const s_meta   = (uint32_t*)(smem + P.sm.meta);
    uint32_t* const s_recoff = (uint32_t*)(smem + P.sm.recoff);
    uint16_t* const s_hdr    = (uint16_t*)(smem + P.sm.hdr);
    uint16_t* const s_cstart = (uint16_t*)(smem + P.sm.cstart);
    uint16_t* const s_gstart = (uint16_t*)(smem + P.sm.gstart);
    uint32_t* const s_dwords = (uint32_t*)(smem + P.sm.dwords);
    uint8_t*  const s_cmap   = smem + P.sm.cmap;
    uint8_t*  const s_gmap   = smem + P.sm.gmap;
    uint32_t* const s_lut    = (uint32_t*)(smem + P.sm.lut);     // 256 x 4 ASCII bytes
    __shared__ unsigned long long s_warp_tot[FUSED_THREADS / 32];
    __shared__ unsigned long long s_excl;
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int emit = P.emit, RL = P.ad.read_len;
    const int nlines = emit_lines(emit);

    // 4 bases (one byte of codes) -> 4 ASCII chars
    for (int v = tid; v < 256; v += FUSED_THREADS) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) w |= (uint32_t)code2ascii((v >> (2 * k)) & 3) << (8 * k);
        s_lut[v] = w;
    }
    unsigned long long my_attempts = 0, my_gather = 0;
    uint32_t my_err = 0;

    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd(P.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.n_tiles) break;
        const uint64_t tile_first = (uint64_t)tile * (uint64_t)P.F;
        const int nf = (int)min((uint64_t)P.F, P.count - tile_first);

        // ------------------------------------------------------------ pass A: sample (thread per fragment)
        uint64_t idx = 0; int L = 0, ci = 0, hdr = 0, tag_len = 0; uint32_t plus = 1; const char* tag = nullptr;
        unsigned long long pk = 0;
        if (tid < nf) {
            const uint64_t id = P.first + tile_first + (uint64_t)tid;
            int lo = 0, hi = P.n_ranges - 1;                    // last range with first_id <= id
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (P.ranges[mid].first_id <= id) lo = mid; else hi = mid - 1; }
            const RangeDev* rg = P.ranges + lo;
            const GenomeDev G = P.genomes[rg->genome];
            const uint64_t* cum = P.ctab.cum_end + G.first_contig;
            uint64_t gpos = 0; bool ok = false;
            for (uint32_t attempt = 0; attempt < 65536u; attempt++) {
                const uint4 x = rng_block(P.key, id, attempt, ST_FRAG);
                my_attempts++;
                // selectFragmentLength, fragSim.cpp:137-187
                if (P.size_mode == 0) L = P.fixed_len;
                else if (P.size_mode == 1) L = __ldg(P.size_len + __umulhi(x.x, (uint32_t)P.n_sizes));
                else {
                    int k = first_below(P.size_thr, P.n_sizes, x.x >> 1);
                    if (k == P.n_sizes) k = (int)(((uint64_t)(x.y >> 1) * (uint64_t)P.n_sizes) >> 31);   // fragSim.cpp:176
                    L = __ldg(P.size_len + k);
                }
                if (L < 0 || L < P.minlen || L > P.maxlen) continue;                 // fragSim.cpp:1375-1381
                const uint64_t r64 = ((uint64_t)x.w << 32) | x.z;
                const uint64_t coord = __umul64hi(r64, G.G);                           // randGenomicCoord, fragSim.cpp:80-93
                if (coord == 0) continue;                                              // never matches a chromosome (starts are 1-based)
                const uint64_t cp = coord - 1;
                int a = 0, b = G.n_contigs - 1;                                        // first contig with cp <= cum_end (fragSim.cpp:1397-1403)
                while (a < b) { int mid = (a + b) >> 1; if (cp <= __ldg(cum + mid)) b = mid; else a = mid + 1; }
                ci = G.first_contig + a;
                const uint64_t clen = __ldg(P.ctab.len + ci);
                idx = cp - (__ldg(cum + a) - clen);                                    // coord - startIndexChr, fragSim.cpp:1404
                if (idx + (uint64_t)L > clen) continue;                                // run-off past the contig end is always rejected (fragSim.cpp:1472)
                gpos = __ldg(P.ctab.gbase + ci) + idx;
                if (L > 0) {                                                           // isResolvedDNAstring, fragSim.cpp:333-339,1472
                    const uint64_t e = gpos + (uint64_t)L - 1;
                    const uint64_t w0 = gpos >> 5, w1 = e >> 5;
                    uint32_t any = 0;
                    for (uint64_t w = w0; w <= w1; w++) {
                        uint32_t m = __ldg(G.nmask + w);
                        if (w == w0) m &= 0xFFFFFFFFu << (gpos & 31);
                        if (w == w1) m &= 0xFFFFFFFFu >> (31 - (e & 31));
                        any |= m;
                    }
                    if (any) continue;
                }
                plus = P.norev ? 1u : (x.y & 1u);                                      // randomBool, fragSim.cpp:1462-1467
                ok = true; break;
            }
            if (!ok) { my_err |= ERRF_GIVEUP; L = 0; idx = 0; }
            my_gather += (unsigned long long)((L + 3) / 4 + (L + 7) / 8);
            tag = rg->tag; tag_len = rg->tag_len;
            const int nlen = __ldg(P.ctab.name_len + ci);
            hdr = 1 + nlen + 3 + ndigits(idx) + 1 + ndigits(idx + (uint64_t)L) + 1 + ndigits((uint64_t)L) + tag_len;
            const int S = emit_seqlen(emit, L, RL);
            uint32_t bytes = 0;
            for (int ln = 0; ln < nlines; ln++) bytes += hdr + emit_suffix(emit, ln) + 1 + S + 1;
            const uint32_t chunks = (L + 15) >> 4;
            const uint32_t groups = nlines * (((S + 15) >> 4) + 1);
            pk = PK(bytes, chunks, groups);
            s_gpos[tid] = gpos;
            s_meta[tid] = META(L, plus, rg->slot, rg->genome);
            s_hdr[tid] = (uint16_t)hdr;
        }
        // block exclusive scan of pk
        unsigned long long incl = warp_scan_u64(pk, lane);
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        unsigned long long woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < FUSED_THREADS / 32; w++) { const unsigned long long t = s_warp_tot[w]; if (w < warp) woff += t; tot += t; }
        const unsigned long long excl = woff + incl - pk;
        const uint32_t T = PK_BYTES(tot), n_chunks = PK_CHUNKS(tot), n_groups = PK_GROUPS(tot);
        if (tid == 0) {   // publish this tile's byte count for the look-back
            st_volatile_u64(P.tile_state + tile, ((tile == 0 ? 2ull : 1ull) << 62) | (unsigned long long)T);
            if (T > (uint32_t)P.sm.stage_bytes) my_err |= ERRF_OVERFLOW;
        }
        const bool overflow = T > (uint32_t)P.sm.stage_bytes;    // host sizing bug guard: never write past staging

        // ------------------------------------------------------------ pass B: headers + work maps (owner thread)
        if (tid < nf && !overflow) {
            const uint32_t roff = PK_BYTES(excl), c0 = PK_CHUNKS(excl), g0 = PK_GROUPS(excl);
            s_recoff[tid] = roff; s_cstart[tid] = (uint16_t)c0; s_gstart[tid] = (uint16_t)g0;
            const uint32_t nch = PK_CHUNKS(pk), ngr = PK_GROUPS(pk);
            for (uint32_t k = 0; k < nch; k++) s_cmap[c0 + k] = (uint8_t)tid;
            for (uint32_t k = 0; k < ngr; k++) s_gmap[g0 + k] = (uint8_t)tid;
            const int S = emit_seqlen(emit, L, RL);
            const char* nm = P.ctab.names + __ldg(P.ctab.name_off + ci);
            const int nlen = __ldg(P.ctab.name_len + ci);
            uint8_t* p = s_stage + roff;
            for (int ln = 0; ln < nlines; ln++) {
                // ">chr:+:start:end:len" + tag   (fragSim.cpp:1489,1498,1505-1507,1603)
                *p++ = '>';
                for (int k = 0; k < nlen; k++) *p++ = (uint8_t)__ldg(nm + k);
                *p++ = ':'; *p++ = plus ? '+' : '-'; *p++ = ':';
                p = put_decimal(p, idx, ndigits(idx)); *p++ = ':';
                p = put_decimal(p, idx + (uint64_t)L, ndigits(idx + (uint64_t)L)); *p++ = ':';
                p = put_decimal(p, (uint64_t)L, ndigits((uint64_t)L));
                for (int k = 0; k < tag_len; k++) *p++ = (uint8_t)tag[k];
                if (emit_suffix(emit, ln)) { *p++ = '_'; *p++ = 'r'; }
                *p++ = '\n';
                p += S;
                *p++ = '\n';
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ pass C: gather + damage (thread per 16-base chunk)
        if (!overflow)
        for (uint32_t e = tid; e < n_chunks; e += FUSED_THREADS) {
            const int f = s_cmap[e];
            const int k = (int)e - (int)s_cstart[f];
            const uint32_t meta = s_meta[f];
            const int fl = META_L(meta), slot = META_SLOT(meta);
            const uint64_t gp = s_gpos[f];
            const uint32_t* seq2 = P.genomes[META_GEN(meta)].seq2;
            const int n = min(16, fl - 16 * k);
            uint32_t W = META_PLUS(meta) ? fetch16(seq2, (int64_t)(gp + 16ull * k))
                                         : revcomp16(fetch16(seq2, (int64_t)(gp + (uint64_t)fl) - 16ll * k - 16ll));   // reverseComplement, fragSim.cpp:1484
            W &= lowfields(n);
            if (slot >= 0 && emit != 0) {
                const DamageDev& D = P.dmg[slot];
                const uint64_t id = P.first + tile_first + (uint64_t)f;
                if (D.mode == 1) {                                   // -matfile: one uniform per base (deamSim.cpp:1847,1911)
                    uint32_t Wn = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (4 * q < n) {
                            const uint4 x = rng_block(P.key, id, (uint32_t)(4 * k + q), ST_DEAM);
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                const int j = 4 * q + jj;
                                if (j < n) Wn |= dmg_matrix_base(D, 16 * k + j, fl, (W >> (2 * j)) & 3u, pick(x, jj)) << (2 * j);
                            }
                        }
                    }
                    W = Wn;
                } else if (D.mode == 2) {                            // Briggs: header block 0, per-base word 4+i
                    const BriggsHdr B = dmg_briggs_hdr(D, rng_block(P.key, id, 0u, ST_DEAM));
                    uint32_t Wn = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (4 * q < n) {
                            const uint4 x = rng_block(P.key, id, (uint32_t)(4 * k + q + 1), ST_DEAM);
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                const int j = 4 * q + jj;
                                if (j < n) Wn |= dmg_briggs_base(D, B, 16 * k + j, fl, (W >> (2 * j)) & 3u, pick(x, jj)) << (2 * j);
                            }
                        }
                    }
                    W = Wn;
                } else if (D.mode == 3 || D.mode == 4) {             // single / double: two passes, two streams
                    uint32_t Wn = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (4 * q < n) {
                            const uint4 x1 = rng_block(P.key, id, (uint32_t)(4 * k + q), ST_DEAM);
                            const uint4 x2 = rng_block(P.key, id, (uint32_t)(4 * k + q), ST_DEAM2);
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                const int j = 4 * q + jj;
                                if (j < n) Wn |= dmg_sd_base(D, 16 * k + j, fl, (W >> (2 * j)) & 3u, pick(x1, jj), pick(x2, jj)) << (2 * j);
                            }
                        }
                    }
                    W = Wn;
                }
            }
            s_dwords[e] = W;
        }
        __syncthreads();

        // ------------------------------------------------------------ pass D: adapters + ASCII (thread per 16-byte staging group)
        if (!overflow)
        for (uint32_t e = tid; e < n_groups; e += FUSED_THREADS) {
            const int f = s_gmap[e];
            const uint32_t meta = s_meta[f];
            const int fl = META_L(meta);
            const int S = emit_seqlen(emit, fl, RL);
            const int gpl = ((S + 15) >> 4) + 1;                    // group slots per line
            int g = (int)e - (int)s_gstart[f];
            const int line = g >= gpl ? 1 : 0;
            g -= line * gpl;
            const int h = s_hdr[f];
            // staging offset of the first sequence byte of this line
            const int ls = (int)s_recoff[f] + line * (h + 1 + S + 1) + h + emit_suffix(emit, line) + 1;
            const int sigma = ls & 15;
            const int j0 = 16 * g - sigma;
            const int jlo = max(j0, 0), jhi = min(j0 + 16, S);
            if (jhi <= jlo) continue;
            const uint32_t codes = line_codes(P, s_dwords + s_cstart[f], fl, line, j0, P.first + tile_first + (uint64_t)f);
            uint4 a;
            a.x = s_lut[codes & 0xFFu]; a.y = s_lut[(codes >> 8) & 0xFFu];
            a.z = s_lut[(codes >> 16) & 0xFFu]; a.w = s_lut[codes >> 24];
            uint8_t* dst = s_stage + (ls - sigma) + 16 * g;         // 16-byte aligned
            if (jlo == j0 && jhi == j0 + 16) {
                *(uint4*)dst = a;
            } else {
                // boundary group: merge under a byte mask.  Headers were written before the barrier; two sequence lines are
                // always >= 12 bytes apart (newline + a header), so no other thread touches these 4-byte words in this pass.
                const int t0 = jlo - j0, t1 = jhi - j0;                  // valid bytes [t0,t1) of the group
                uint32_t* d4 = (uint32_t*)dst;
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const int lo = min(max(t0 - 4 * w, 0), 4), hi = min(max(t1 - 4 * w, 0), 4);
                    uint32_t mh, ml;
                    asm("shl.b32 %0, 1, %1;" : "=r"(mh) : "r"(8 * hi));
                    asm("shl.b32 %0, 1, %1;" : "=r"(ml) : "r"(8 * lo));
                    const uint32_t m = (mh - 1u) & ~(ml - 1u);
                    if (m == 0xFFFFFFFFu) d4[w] = aw[w];
                    else if (m) d4[w] = (d4[w] & ~m) | (aw[w] & m);
                }
            }
        }

        // ------------------------------------------------------------ look-back: global byte offset of this tile
        if (warp == 0) {
            unsigned long long ex = 0;
            if (tile > 0) {
                int look = tile - 1;
                while (true) {
                    const int i = look - lane;
                    unsigned long long v;
                    do { v = i >= 0 ? ld_volatile_u64(P.tile_state + i) : (2ull << 62); } while (__any_sync(0xffffffffu, (v >> 62) == 0));
                    const uint32_t inc = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                    const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                    ex += warp_sum_u64(lane <= first_inc ? (v & 0x3FFFFFFFFFFFFFFFull) : 0ull);
                    if (inc) break;
                    look -= 32;
                }
                if (lane == 0) st_volatile_u64(P.tile_state + tile, (2ull << 62) | (ex + (unsigned long long)T));
            }
            if (lane == 0) {
                s_excl = ex;
                if (tile == P.n_tiles - 1) P.stats[ST_TOTAL] = ex + (unsigned long long)T;
            }
        }
        __syncthreads();

        // ------------------------------------------------------------ copy-out: staging -> global, 16-byte aligned vector stores
        if (!overflow) {
            const unsigned long long B = s_excl;
            if (B + T > P.out_cap) { my_err |= ERRF_OVERFLOW; }
            else {
                uint8_t* const gout = P.out + B;
                const int pad = (int)((uintptr_t)gout & 15);            // misalignment of the tile's first byte
                // aligned global chunk c covers staging bytes [16c - pad, 16c - pad + 16)
                const int nchunks = (int)((T + pad + 15) >> 4);
                const int sh = (16 - pad) & 15;                          // staging byte offset within its aligned 16 B
                for (int c = tid; c < nchunks; c += FUSED_THREADS) {
                    const int s0 = 16 * c - pad;
                    uint8_t* gdst = gout + s0;
                    if (s0 >= 0 && s0 + 16 <= (int)T) {
                        uint4 v;
                        if (sh == 0) v = *(const uint4*)(s_stage + s0);
                        else {
                            const uint4 lo4 = *(const uint4*)(s_stage + s0 - sh), hi4 = *(const uint4*)(s_stage + s0 - sh + 16);
                            const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                            const int ws = sh >> 2; const uint32_t bs = (uint32_t)(sh & 3) * 8u;
                            uint32_t r[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                uint32_t a0 = 0, a1 = 0;
#pragma unroll
                                for (int q = 0; q < 4; q++) if (ws == q) { a0 = w[q + k]; a1 = w[q + k + 1 < 8 ? q + k + 1 : 7]; }
                                r[k] = __funnelshift_r(a0, a1, bs);
                            }
                            v = make_uint4(r[0], r[1], r[2], r[3]);
                        }
                        st_global_v4(gdst, v);
                    } else {
                        for (int t = max(s0, 0); t < min(s0 + 16, (int)T); t++) gout[t] = s_stage[t];
                    }
                }
            }
        }
        __syncthreads();
    }

    // per-CTA statistics
    my_attempts = warp_sum_u64(my_attempts); my_gather = warp_sum_u64(my_gather);
    uint32_t e = my_err;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e |= __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) {
        if (my_attempts) atomicAdd(P.stats + ST_ATTEMPTS, my_attempts);
        if (my_gather) atomicAdd(P.stats + ST_GATHER, my_gather);
        if (e) atomicOr(P.stats + ST_ERR, (unsigned long long)e);
    }
}

cudaError_t fused_occupancy(int smem_bytes, int* blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_fused, FUSED_THREADS, smem_bytes);
}
cudaError_t launch_fused(const FusedParams& P, int grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, P.sm.total);
    if (e != cudaSuccess) return e;
    k_fused<<<grid, FUSED_THREADS, P.sm.total, st>>>(P);
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

// stand-alone deamSim main loop (deamSim.cpp:1300-2042): warp per record
__global__ void __launch_bounds__(256) k_deam_records(const RecStream R, uint8_t* __restrict__ out, const DamageDev D, int has_damage,
                                                      Key key, uint64_t first_id, unsigned long long* stats) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < R.n_rec; r += nw) {
        const RecPos p = rec_pos(R, r);
        const uint64_t id = first_id + r;
        if (lane == 0 && (p.ide <= p.ids || R.in[p.ids] != '>')) atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
        for (uint64_t i = p.ids + lane; i < p.ss; i += 32) out[i] = R.in[i];          // ID line + '\n' unchanged (deamSim.cpp:2038)
        const int L = (int)(p.se - p.ss);
        BriggsHdr B; B.o5 = B.o3 = B.nick = 0;
        if (has_damage && D.mode == 2) B = dmg_briggs_hdr(D, rng_block(key, id, 0u, ST_DEAM));
        for (int i = lane; i < L; i += 32) {
            uint8_t c = upper(R.in[p.ss + i]);                                         // toupper, deamSim.cpp:1319-1321
            const int code = ascii2code(c);
            if (has_damage && code >= 0) {                                             // unresolved bases are never touched (deamSim.cpp:1795)
                uint32_t nb = (uint32_t)code;
                if (D.mode == 1) nb = dmg_matrix_base(D, i, L, nb, rng_word(key, id, ST_DEAM, (uint32_t)i));
                else if (D.mode == 2) nb = dmg_briggs_base(D, B, i, L, nb, rng_word(key, id, ST_DEAM, 4u + (uint32_t)i));
                else if (D.mode == 3 || D.mode == 4) nb = dmg_sd_base(D, i, L, nb, rng_word(key, id, ST_DEAM, (uint32_t)i), rng_word(key, id, ST_DEAM2, (uint32_t)i));
                c = code2ascii(nb);
            }
            out[p.ss + i] = c;
        }
        if (lane == 0) out[p.se] = '\n';
    }
}
cudaError_t launch_deam_records(const RecStream& R, uint8_t* out, const DamageDev& D, int has_damage, Key key,
                                uint64_t first_id, unsigned long long* stats, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    const unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    k_deam_records<<<grid, 256, 0, st>>>(R, out, D, has_damage, key, first_id, stats);
    return cudaGetLastError();
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
__global__ void __launch_bounds__(256) k_adpt_sizes(const RecStream R, int emit, int RL, unsigned long long* __restrict__ sizes, unsigned long long* stats) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n_rec) return;
    const RecPos p = rec_pos(R, r);
    if (p.ide <= p.ids || R.in[p.ids] != '>') atomicOr(stats + ST_ERR, (unsigned long long)ERRF_PARSE);
    sizes[r] = adpt_rec_bytes(emit, p.ide - p.ids, RL);
}
cudaError_t launch_adpt_sizes(const RecStream& R, int emit, int read_len, unsigned long long* sizes, unsigned long long* stats, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    k_adpt_sizes<<<(unsigned)((R.n_rec + 255) / 256), 256, 0, st>>>(R, emit, read_len, sizes, stats);
    return cudaGetLastError();
}

// stand-alone adptSim main loop (adptSim.cpp:263-420): warp per record, ASCII in/out
__global__ void __launch_bounds__(256) k_adpt_records(const RecStream R, const unsigned long long* __restrict__ out_off, uint8_t* __restrict__ out,
                                                      int emit, Adapters ad, const uint8_t* __restrict__ adF, const uint8_t* __restrict__ adS,
                                                      Key key, uint64_t first_id) {
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
                                const uint8_t* adF, const uint8_t* adS, Key key, uint64_t first_id, cudaStream_t st) {
    if (R.n_rec == 0) return cudaSuccess;
    const uint64_t want = (R.n_rec + 7) / 8;
    const unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    k_adpt_records<<<grid, 256, 0, st>>>(R, out_off, out, emit, ad, adF, adS, key, first_id);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_l2_flush(uint4* __restrict__ buf, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4((uint32_t)i, 0u, 1u, 2u);
}
cudaError_t launch_l2_flush(uint4* buf, uint64_t n16, cudaStream_t st) {
    k_l2_flush<<<148 * 8, 256, 0, st>>>(buf, n16); return cudaGetLastError();
}

} // namespace ggk
